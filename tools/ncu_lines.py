#!/usr/bin/env python
"""Aggregate warp-stall samples of an ncu `--page source --csv --print-source cuda,sass` export by source file:line.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python tools/ncu_lines.py x.csv [top]
Lines in the export are grouped per source file; SASS rows follow the CUDA line they belong to.
"""
import csv, sys, collections

def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    per = collections.Counter()
    text = {}
    cur_file, cur_line, hdr = None, None, None
    for row in csv.reader(open(path, newline="")):
        if not row:
            continue
        if row[0] in ("File Name", "File Path"):
            cur_file = row[1].split("/")[-1]
            continue
        if row[0] == "Line No":
            hdr = row
            i_samp = hdr.index("# Samples")
            continue
        if hdr is None:
            continue
        if row[0] != "":
            if not row[0].isdigit():
                continue
            cur_line = int(row[0])
            text[(cur_file, cur_line)] = row[1]
        try:
            s = int(row[i_samp] or 0)
        except (ValueError, IndexError):
            s = 0
        if row[0] == "" or True:
            # SASS rows carry the samples; CUDA rows carry the per-line sum -> count only the CUDA rows
            pass
        if row[0] != "":
            per[(cur_file, cur_line)] += s
    tot = sum(per.values())
    print("total samples", tot)
    for (f, l), s in per.most_common(top):
        print(f"{100.0*s/tot:6.2f}%  {f}:{l:<5d} {text[(f,l)].strip()[:110]}")
    # cumulative by file in line order, 5% buckets
    print("--- per-file running share (line ranges holding >=4%)")
    for f in sorted({k[0] for k in per}):
        acc, start = 0, None
        for (ff, l) in sorted(k for k in per if k[0] == f):
            s = per[(ff, l)]
            if s == 0:
                continue
            if start is None:
                start = l
            acc += s
            if acc >= 0.04 * tot:
                print(f"{100.0*acc/tot:6.2f}%  {f}:{start}-{l}")
                acc, start = 0, None
        if acc:
            print(f"{100.0*acc/tot:6.2f}%  {f}:{start}-end")

if __name__ == "__main__":
    main()
