#!/usr/bin/env python
"""prints selected fields of the last JSON line of a file (tools output may be preceded by library banners)"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "ms_per_step", d.get("ms_per_step"), d.get("device_ms_per_rank"), d.get("blocks_per_rank"))
for k, v in (d.get("kernel_ms_per_rank") or {}).items():
    print("   %-45s %s" % (k, v))
if "verify" in d:
    print("   verify", d["verify"])
