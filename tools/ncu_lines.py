#!/usr/bin/env python
"""Summarises `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line:
stall samples, executed instructions and the dominant stall reasons.  usage: ncu_lines.py file.csv [top]"""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]; continue
        if r and r[0] == "Line No":
            hdr = r; continue
        if hdr is None or len(r) != len(hdr) or not r[0]:
            continue
        lines.append((cur_file, r))
    si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(float(r[si] or 0) for _, r in lines)
    total_inst = sum(float(r[ii] or 0) for _, r in lines)
    print("total samples %d, total warp-instructions %d" % (total, total_inst))
    agg = {}
    for i, h in stall_cols:
        agg[h] = sum(float(r[i] or 0) for _, r in lines)
    print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(total, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for f, r in sorted(lines, key=lambda fr: -float(fr[1][si] or 0))[:top]:
        st = sorted(((float(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:3]
        print("%-16s %4s %5.1f%% inst %5.1f%% [%s] | %s" % (f, r[0], 100 * float(r[si] or 0) / max(total, 1), 100 * float(r[ii] or 0) / max(total_inst, 1),
                                                        " ".join("%s:%d" % (n, v) for v, n in st if v), r[1].strip()[:95]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
