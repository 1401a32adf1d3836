"""GPU-box probe: ssfm_match_pairs throughput.  python tools/match_probe.py [images] [descriptors per image] [pairs]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import spherical_sfm_b200 as S  # noqa: E402

ni = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
P = int(sys.argv[3]) if len(sys.argv) > 3 else 2016
rng = np.random.default_rng(0)
base = np.minimum(rng.gamma(0.6, 30.0, (n, 128)), 255).astype(np.int32)
descs = []
for i in range(ni):  # consecutive images share ~40 % of their descriptors (perturbed), like overlapping views
    d = np.minimum(rng.gamma(0.6, 30.0, (n, 128)), 255).astype(np.int32)
    keep = rng.random(n) < 0.4
    d[keep] = np.clip(base[keep] + rng.integers(-2, 3, (int(keep.sum()), 128)), 0, 255)
    descs.append(d.astype(np.float32))
allrows = np.concatenate(descs)
offs = (np.arange(ni + 1) * n).astype(np.int64)
pairs = np.array([(i, j) for i in range(ni) for j in range(i + 1, ni)], np.int32)[:P]
eng = S.Engine(0)
eng.match_pairs(allrows, offs, pairs[:8])
t0 = time.perf_counter()
mo, mm = eng.match_pairs(allrows, offs, pairs)
dt = time.perf_counter() - t0
flop = 2.0 * n * n * 128 * len(pairs)
print("pairs %d x (%d x %d): %.3f s  %.1f pairs/s  %.1f TFLOP/s (fp16 tensor)  %.3e distance evaluations/s  matches/pair %.0f" % (
    len(pairs), n, n, dt, len(pairs) / dt, flop / dt / 1e12, float(n) * n * len(pairs) / dt, mo[-1] / len(pairs)))
ms = eng.match_stats()
print("  stages: pack+H2D %.2f ms, k_match_2nn %.2f ms (%.1f TFLOP/s algorithmic, %.1f executed), compaction+D2H %.2f ms, call %.2f ms; tiles %d, ctas %d" % (
    ms.pack_ms, ms.knn_ms, flop / (ms.knn_ms * 1e-3) / 1e12, ms.mma_tiles * 2.0 * 256 * 128 * 144 / (ms.knn_ms * 1e-3) / 1e12,
    ms.compact_ms, ms.total_ms, ms.mma_tiles, ms.ctas))
import match_oracle as MO  # noqa: E402
t0 = time.perf_counter()
om = MO.match(descs[pairs[0][0]], descs[pairs[0][1]])
dc = time.perf_counter() - t0
a, b = mo[0], mo[1]
print("oracle (numpy, 1 core) one pair: %.2f s; identical: %s" % (dc, bool(len(om) == b - a and (mm[a:b] == om).all())))
