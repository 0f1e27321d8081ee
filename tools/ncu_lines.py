#!/usr/bin/env python
"""Summarise an ncu source-page CSV (`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`) per CUDA source line:
share of warp-stall samples and of executed warp instructions.  Usage: ncu_lines.py file.csv [top]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hdr = None
    lines = []
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r or r[0] == "":
            continue
        if len(r) < len(hdr) or not r[0].isdigit():
            continue
        d = dict(zip(hdr[4:], r[4:]))
        try:
            lines.append((int(r[0]), r[1].strip(), int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), d))
        except (ValueError, KeyError):
            pass
    ts = sum(l[2] for l in lines) or 1
    ti = sum(l[3] for l in lines) or 1
    print("total samples %d, warp instructions %d" % (ts, ti))
    stall_keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    for ln, src, s, i, d in sorted(lines, key=lambda x: -x[2])[:top]:
        st = sorted(((int(d.get(k) or 0), k[6:]) for k in stall_keys), reverse=True)[:3]
        print("%5d  smp %5.1f%%  inst %5.1f%%  %-28s %s" % (ln, 100.0 * s / ts, 100.0 * i / ti, ",".join("%s:%d" % (k, v) for v, k in st if v), src[:100]))


if __name__ == "__main__":
    main()
