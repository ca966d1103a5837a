#!/bin/bash
# ncu launch list of one bench step (per-launch gpu__time_duration, cold-cache and serialised: shares, not absolutes).
# usage: tools/launch_list.sh <out.csv> [bench args...]
out=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:vxb_ -c 400 --csv --log-file "$out" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-config4 "$@" > "${out%.csv}.log" 2>&1
python - "$out" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
out = [(r[ki][:70], float(r[vi].replace(',', '')) * (1e-3 if r[ui] == 'ns' else 1.0)) for r in rows[1:]]
# the last step = everything after the last vxb_scan_kernel launch
last = max(i for i, (k, _) in enumerate(out) if 'vxb_scan_kernel' in k)
for k, v in out[last:]:
    print("%-70s %9.1f us" % (k, v))
print("sum of the step: %.1f us" % sum(v for _, v in out[last:]))
PY
