#!/usr/bin/env python3
"""tools/ncu_hot_lines.py <report.ncu-rep> [topN] [kernel-regex] [ins] -- source lines ranked by warp-stall samples (or by executed instructions with "ins")
(needs -lineinfo at compile time and --import-source on at capture time)."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    kern = sys.argv[3] if len(sys.argv) > 3 else None
    by_ins = len(sys.argv) > 4 and sys.argv[4] == "ins"
    out = subprocess.run(["ncu", "-i", rep] + (["-k", "regex:" + kern] if kern else []) +
                         ["--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, acc = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name" or hdr is None:
            continue
        if r[0] != "":  # a source line row (its SASS rows follow with empty first columns)
            try:
                si = hdr.index("# Samples")
                ii = hdr.index("Instructions Executed")
                n, ins = int(r[si]), int(r[ii])
            except (ValueError, IndexError):
                continue
            key = (cur_file, r[0])
            a = acc.setdefault(key, [0, 0, r[1].strip()[:100]])
            a[0] += n
            a[1] += ins
    tot = sum(a[0] for a in acc.values()) or 1
    toti = sum(a[1] for a in acc.values()) or 1
    print("total samples %d, warp instructions %d" % (tot, toti))
    for (f, ln), a in sorted(acc.items(), key=lambda kv: -kv[1][1 if by_ins else 0])[:top]:
        print("%5.1f%% smp %5.1f%% ins  %s:%s  %s" % (100.0 * a[0] / tot, 100.0 * a[1] / toti, f, ln, a[2]))


if __name__ == "__main__":
    main()
