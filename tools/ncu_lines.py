#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "sass,cuda", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    agg = defaultdict(lambda: [0, 0, ""])
    cur_file = ""
    for r in rows:
        if r and r[0] == "File Name":
            cur_file = r[1].split("/")[-1]
        if "Instructions Executed" in r and "# Samples" in r:
            hdr = r
            ii, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= ii or not r[ii].isdigit():
            continue
        if not r[0]:
            continue
        key = (cur_file, r[0])
        agg[key][0] += int(r[ii])
        agg[key][1] += int(r[sm]) if r[sm].isdigit() else 0
        if r[0]:
            agg[key][2] = r[1]
    tot = sum(v[0] for v in agg.values()) or 1
    smp = sum(v[1] for v in agg.values()) or 1
    print("total warp instructions %d, samples %d" % (tot, smp))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%10d %5.1f%%  smp %5.1f%%  %s:%s  %s" % (v[0], 100.0 * v[0] / tot, 100.0 * v[1] / smp, k[0], k[1], v[2].strip()[:100]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
