"""The FP32 pre-filter's arithmetic on the CPU: k_score_rounds evaluates the Sampson error of a unit-z correspondence through
expanded bilinear forms (csrc/ssfm_kernels.cuh, sampson2_unitz_bilinear).  Restated here in numpy float32 (without fused
multiply-adds, i.e. with slightly more rounding than the device), against the float64 oracle on config-C3-shaped pairs:
the MSAC cost of every minimal-sample model must agree far inside the pre-filter's 2e-4 margin, and no worse than the plain
sum-of-squares form it replaced."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from conftest import THR2, E_of  # noqa: E402


def _costs_f32(p, x, y, z, w, a, b, st, thr):
    f = np.float32
    two = f(2)
    A = p[0] * p[0] + p[1] * p[1]
    Bu, Cu = two * (p[0] * p[2] + p[1] * p[3]), two * (p[1] * p[2] - p[0] * p[3])
    Bt, Ct = two * (p[0] * p[4] + p[1] * p[5]), two * (p[1] * p[4] - p[0] * p[5])
    D = (p[2] * p[2] + p[3] * p[3]) + (p[4] * p[4] + p[5] * p[5])
    d = p[0] * a + (p[1] * b + (p[2] * z + (p[3] * w + (p[4] * x + p[5] * y))))
    den = A * st + (Bu * x + (Cu * y + (Bt * z + (Ct * w + D))))
    expanded = np.minimum(d * d / den, thr).astype(np.float64).sum()
    Eu0, Eu1, Eu2 = p[0] * x + p[1] * y + p[2], p[1] * x - p[0] * y + p[3], p[4] * x + p[5] * y
    Et0, Et1 = p[0] * z + p[1] * w + p[4], p[1] * z - p[0] * w + p[5]
    dp = z * Eu0 + w * Eu1 + Eu2
    plain = np.minimum(dp * dp / (Eu0 * Eu0 + Eu1 * Eu1 + Et0 * Et0 + Et1 * Et1), thr).astype(np.float64).sum()
    return expanded, plain


def test_expanded_fp32_scoring_form_is_far_inside_the_prefilter_margin(S, orc):
    f = np.float32
    worst_expanded, worst_plain, models = 0.0, 0.0, 0
    for seed in range(4):
        pr = S.problems.make_problem(S.problems.make_rng(77, seed), 1500, False, None, 1 / 600, 1050, 20.0)  # C3: 70 % outliers
        rays = pr.rays
        assert (rays[:, 2] == 1.0).all() and (rays[:, 5] == 1.0).all()
        x, y, z, w = (rays[:, k].astype(f) for k in (0, 1, 3, 4))
        a, b, st = z * x - w * y, z * y + w * x, x * x + y * y + z * z + w * w  # the per-correspondence products of a tile
        for it in range(50):
            idx = [int(v) for v in np.random.default_rng(100 * seed + it).integers(0, 1500, 3)]
            if len(set(idx)) < 3:
                continue
            _, mo = orc.solve(rays, idx, 0)
            for m in mo:
                if not np.isfinite(m).all():
                    continue
                s64 = np.minimum(orc.sampson(E_of(m), rays), THR2).sum()
                se, sp = _costs_f32([f(v) for v in m], x, y, z, w, a, b, st, f(THR2))
                worst_expanded = max(worst_expanded, abs(se - s64) / s64)
                worst_plain = max(worst_plain, abs(sp - s64) / s64)
                models += 1
    assert models > 500
    assert worst_expanded < 2e-5, worst_expanded          # the pre-filter margin is 2e-4 (Params::cand_margin)
    assert worst_expanded < 4 * worst_plain + 1e-7, (worst_expanded, worst_plain)
