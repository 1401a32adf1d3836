"""Synthetic spherical two-view problems with ground truth.

Mirror of evaluation/problem_generator/problem_generator.cpp:14-65 (conventions) and
problem_generator.h:17-39 (error metrics), plus outlier injection (the reference never injects
outliers; BASELINE.json's configs do).  The reference draws from an unseeded libstdc++
std::default_random_engine (random.h:8); here the stream is numpy's Philox so host, tests and
bench agree everywhere.
"""
import numpy as np


def so3exp(r):
    """src/so3.cpp:16-23"""
    r = np.asarray(r, np.float64)
    th = np.linalg.norm(r)
    if th < 1e-10:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def skew3(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], np.float64)


def make_spherical_E(R, inward=False):
    """src/spherical_utils.cpp:9-14: t = R e3 - e3 (negated if inward), E = [t]x R."""
    t = R[:, 2] - np.array([0, 0, 1.0])
    if inward:
        t = -t
    return skew3(t) @ R, t


class Problem:
    __slots__ = ("rays", "E", "R", "t", "inlier_mask", "inward")


def make_problem(rng, num_corr, inward=False, rotation_deg=None, noise=0.0, num_outliers=0, max_angle_deg=180.0):
    """One problem.  rays: (N, 6) float64 = [u.xyz, v.xyz] per row == RayPair memory layout.

    rotation_deg None -> angle = U(-1,1)*max_angle (the reference uses max_angle = 180 deg,
    problem_generator.cpp:22); whole problems are redrawn until every point has positive depth
    in the second view (:50,61).  Noise is added to both images' xy (:54-55).  Outliers replace
    v.xy of `num_outliers` randomly chosen rows by fresh N(0,1) draws.
    """
    while True:
        angle = (rng.uniform(-1, 1) * np.deg2rad(max_angle_deg)) if rotation_deg is None else np.deg2rad(rotation_deg)
        axis = rng.standard_normal(3)
        axis /= np.linalg.norm(axis)
        R = so3exp(axis * angle)
        E, t = make_spherical_E(R, inward)
        u = np.ones((num_corr, 3))
        u[:, :2] = rng.standard_normal((num_corr, 2))
        depth = rng.uniform(-1, 1, num_corr) * (0.25 if inward else 2.0) + (0.5 if inward else 6.0)
        X = u * depth[:, None]
        P2 = X @ R.T + t
        good = np.all(P2[:, 2] >= 0)
        v = np.ones((num_corr, 3))
        with np.errstate(divide="ignore", invalid="ignore"):
            v[:, :2] = P2[:, :2] / P2[:, 2:3]
        nu = rng.standard_normal((num_corr, 2))
        nv = rng.standard_normal((num_corr, 2))
        u[:, :2] += noise * nu
        v[:, :2] += noise * nv
        if good:
            break
    mask = np.ones(num_corr, bool)
    if num_outliers > 0:
        idx = rng.permutation(num_corr)[:num_outliers]
        v[idx, :2] = rng.standard_normal((num_outliers, 2))
        mask[idx] = False
    p = Problem()
    p.rays = np.ascontiguousarray(np.concatenate([u, v], axis=1))
    p.E, p.R, p.t, p.inlier_mask, p.inward = E, R, t, mask, inward
    return p


def make_rng(seed, stream=0):
    return np.random.Generator(np.random.Philox(key=[seed, stream]))


def make_batch(seed, num_pairs, num_corr, inward=False, rotation_deg=None, noise=1.0 / 600, outlier_frac=0.0,
               max_angle_deg=20.0, first_pair=0):
    """CSR batch: rays (sum N, 6), offsets (P+1,), list of Problems.  Pair p uses stream first_pair+p."""
    probs = []
    n_out = int(round(outlier_frac * num_corr))
    for p in range(num_pairs):
        rng = make_rng(seed, first_pair + p)
        probs.append(make_problem(rng, num_corr, inward, rotation_deg, noise, n_out, max_angle_deg))
    rays = np.concatenate([q.rays for q in probs], axis=0)
    offsets = np.arange(num_pairs + 1, dtype=np.int64) * num_corr
    return rays, offsets, probs


# ---- error metrics (evaluation/problem_generator/problem_generator.h:17-39) ----
def frob_error(E_gt, E):
    a = E_gt / np.linalg.norm(E_gt)
    b = np.asarray(E).reshape(3, 3)
    b = b / np.linalg.norm(b)
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b))


def rot_error(R_gt, R):
    c = (np.trace(R @ R_gt.T) - 1) / 2
    return float(np.arccos(np.clip(c, -1, 1)))


def trans_error(t_gt, t):
    a = t_gt / np.linalg.norm(t_gt)
    b = t / np.linalg.norm(t)
    d = np.clip(a @ b, -1, 1)
    return float(min(np.arccos(d), np.arccos(-d)))


def make_sixpt_batch(seed, num_pairs, num_corr, outlier_frac=0.5, noise_px=0.5, focal_range=(400.0, 1200.0), max_angle_deg=20.0):
    """Config C4 inputs: general (non-spherical) two-view problems with an unknown shared focal; rays are
    (x - cx, y - cy, 1) in pixel units.  Returns rays (P*N, 6), offsets, focal (P,), R (P,3,3), t (P,3)."""
    rng = np.random.default_rng(seed)
    P, N = num_pairs, num_corr
    f = rng.uniform(focal_range[0], focal_range[1], P)
    ax = rng.normal(size=(P, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = np.deg2rad(rng.uniform(-max_angle_deg, max_angle_deg, P))
    K = np.zeros((P, 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -ax[:, 2], ax[:, 1], ax[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ax[:, 0], -ax[:, 1], ax[:, 0]
    R = np.eye(3)[None] + np.sin(ang)[:, None, None] * K + (1 - np.cos(ang))[:, None, None] * (K @ K)
    t = rng.normal(size=(P, 3))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    X = np.stack([rng.uniform(-2, 2, (P, N)), rng.uniform(-2, 2, (P, N)), rng.uniform(4, 8, (P, N))], axis=2)
    Y = np.einsum("pij,pnj->pni", R, X) + t[:, None, :]
    rays = np.ones((P, N, 6))
    rays[:, :, 0:2] = f[:, None, None] * X[:, :, :2] / X[:, :, 2:3] + rng.normal(size=(P, N, 2)) * noise_px
    rays[:, :, 3:5] = f[:, None, None] * Y[:, :, :2] / Y[:, :, 2:3] + rng.normal(size=(P, N, 2)) * noise_px
    out = rng.random((P, N)) < outlier_frac
    v = rays[:, :, 3:5]
    v[out] = rng.uniform(-0.5, 0.5, (int(out.sum()), 2)) * np.repeat(f[:, None], N, 1)[out][:, None]
    rays[:, :, 3:5] = v
    return rays.reshape(P * N, 6), (np.arange(P + 1) * N).astype(np.int64), f, R, t


def make_tracks(seed, num_cameras, num_points, obs_range=(3, 24), focal=600.0, noise_px=0.5, outlier_frac=0.2):
    """Inputs of SfM::Retriangulate (src/sfm.cpp:156-192): cameras on the unit sphere looking outward (spherical
    motion: t = R e3 - e3 composed along a circle), 3-D points in front of them, ragged tracks with pixel noise and
    gross outliers.  Returns camera_tr (C,6: t, r), obs_offsets (P+1), obs_camera, obs_xy, focal, X_true (P,3)."""
    rng = np.random.default_rng(seed)
    ang = np.linspace(0, 2 * np.pi, num_cameras, endpoint=False)
    cam = np.zeros((num_cameras, 6))
    for i, a in enumerate(ang):
        r = np.array([0.0, a, 0.0]) + rng.normal(size=3) * 0.01
        R = so3exp(r)
        cam[i, :3] = R[:, 2] - np.array([0, 0, 1.0])  # t = R e3 - e3 (src/spherical_utils.cpp:9-14)
        cam[i, 3:] = r
    offs = [0]
    oc, oxy, Xs = [], [], []
    for p in range(num_points):
        n = int(rng.integers(obs_range[0], obs_range[1] + 1))
        c0 = int(rng.integers(0, num_cameras))
        cams = (c0 + np.arange(n)) % num_cameras
        R0 = so3exp(cam[c0 + 0 if n == 0 else cams[n // 2], 3:])
        t0 = cam[cams[n // 2], :3]
        Xc = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(4, 8)])  # in the middle camera's frame
        X = R0.T @ (Xc - t0)
        for c in cams:
            R = so3exp(cam[c, 3:])
            PX = R @ X + cam[c, :3]
            if PX[2] < 1.0 or np.abs(PX[:2] / PX[2]).max() > 1.5:  # not visible from this camera
                continue
            xy = focal * PX[:2] / PX[2] + rng.normal(size=2) * noise_px
            if rng.random() < outlier_frac:
                xy = rng.uniform(-300, 300, 2)
            oc.append(c)
            oxy.append(xy)
        offs.append(len(oc))
        Xs.append(X)
    return (cam, np.array(offs, np.int64), np.array(oc, np.int32), np.array(oxy, np.float64).reshape(-1, 2), focal,
            np.array(Xs))
