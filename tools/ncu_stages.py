"""Time (stall samples) and instruction shares per FUNCTION of the csrc headers, from an .ncu-rep captured with
--import-source on.  python tools/ncu_stages.py report.ncu-rep"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, agg = None, None, {}
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10 or r[0] == "":
        continue
    try:
        ln = int(r[0]); ins = int(r[hdr.index("Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
    except Exception:
        continue
    agg[(cur, ln)] = (ins, samp)
funcs = {}
for path in set(p for p, _ in agg):
    starts = []
    if os.path.exists(path) and path.startswith(ROOT):
        for i, line in enumerate(open(path), 1):
            m = re.match(r"^(?:template.*\n)?(?:SSFM_HD(?:_NOINLINE)?|__device__|__global__|inline|static)[^;]*?\b(\w+)\s*\(", line)
            if m and not line.startswith(" "):
                starts.append((i, m.group(1)))
    funcs[path] = starts
def owner(path, ln):
    name = os.path.basename(path)
    for i, fn in reversed(funcs.get(path, [])):
        if ln >= i:
            return name + ":" + fn
    return name
bi, bs = collections.Counter(), collections.Counter()
for (p, ln), (ins, samp) in agg.items():
    o = owner(p, ln); bi[o] += ins; bs[o] += samp
ti, ts = sum(bi.values()), sum(bs.values())
print("total warp instructions %d, stall samples %d" % (ti, ts))
for b, v in bs.most_common(25):
    print("%-52s time %5.1f%%   instructions %5.1f%%" % (b, 100 * v / ts, 100 * bi[b] / ti))
