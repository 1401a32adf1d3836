"""ncu launch list (csv) -> markdown table of per-kernel shares."""
import csv, sys
src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = {}
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
own = {k: v for k, v in agg.items() if (k.startswith("ssfm::") or k.startswith("k_")) and "k_fma_peak" not in k}
tot = sum(v[1] for v in own.values())
with open(out, "w") as f:
    f.write("# %s\n\n" % title)
    f.write("Raw CSV: %s.  Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n" % src)
    f.write("Shares among the engine's own kernels (torch data-generation kernels and the FFMA peak probe excluded):\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, (c, t) in sorted(own.items(), key=lambda x: -x[1][1]):
        f.write("| %s | %d | %.3f | %.1f%% |\n" % (k, c, t / 1e6, 100 * t / tot))
print(open(out).read())
