#!/bin/bash
# `ncu --set full` of the vxb_* kernels of ONE bench step (the last one of `--steps 1 --warmup 1`), with sources, plus the
# exports the profiles/ summaries are made from.  usage: tools/profile_full.sh <out_prefix> [extra bench args]
# (a number printed by a run under ncu is never a bench value)
out=$1; shift
mkdir -p "$(dirname "$out")"
export VXB200_NO_GRAPH=1   # plain launches: one ncu record per kernel, in launch order
ncu --set full --clock-control none --import-source on --kernel-name regex:vxb_ --launch-skip 21 --launch-count 19 -f -o "$out" \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-config4 "$@" > "$out.log" 2>&1
ncu -i "$out.ncu-rep" --page raw --csv > "$out.raw.csv" 2>/dev/null
ncu -i "$out.ncu-rep" --page details > "$out.details.txt" 2>/dev/null
for k in vxb_scan_kernel vxb_block_kernel vxb_transition_kernel vxb_vertex_kernel vxb_triangle_kernel vxb_transition_vertex_kernel; do
  ncu -i "$out.ncu-rep" --page source --csv --print-source cuda,sass --kernel-name regex:$k > "$out.src.$k.csv" 2>/dev/null
done
python tools/ncu_kernels.py "$out.raw.csv" "$out" "ncu --set full --clock-control none, one step of bench.py (1024^3 terrain, plain launches)"
