#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an ncu report captured with --import-source on.
usage: ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
fname = None
cur = None
src = ""
agg = collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        if r[0] != "":
            cur = int(r[0])
            src = r[1]
        if r[2].startswith("0x"):
            a = agg[(fname, cur)]
            a[0] += int(r[7] or 0)
            a[1] += int(r[6] or 0)
            a[2] = src
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print("total warp instructions %d, stall samples %d" % (tot, tots))
byfile = collections.Counter()
for (f, l), a in agg.items():
    byfile[f] += a[0]
print({k: "%.1f%%" % (100 * v / tot) for k, v in byfile.most_common()})
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %4d  %5.2f%% inst  %5.2f%% smp  %s" % (f, l, 100 * a[0] / tot, 100 * a[1] / tots, a[2].strip()[:100]))
