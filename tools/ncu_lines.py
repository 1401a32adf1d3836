"""Per-source-line instruction shares of an .ncu-rep captured with --import-source on (needs -lineinfo).
python tools/ncu_lines.py report.ncu-rep [top N]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, agg, tot = None, None, {}, 0
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10 or r[0] == "":
        continue
    try:
        ln = int(r[0])
        ins = int(r[hdr.index("Instructions Executed")])
        samp = int(r[hdr.index("# Samples")])
        thr = int(r[hdr.index("Thread Instructions Executed")])
    except Exception:
        continue
    agg[(cur, ln)] = (ins, samp, thr, r[1][:100])
    tot += ins
print("total warp instructions", tot)
for (f, ln), (ins, samp, thr, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print("%-24s %4d %5.2f%% samples %6d lanes %4.1f  %s" % (f, ln, 100 * ins / tot, samp, thr / max(ins, 1), src))
byfile = collections.Counter()
for (f, ln), (ins, samp, thr, src) in agg.items():
    byfile[f] += ins
print({k: round(100 * v / tot, 1) for k, v in byfile.items()})
