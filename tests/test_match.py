"""Descriptor matching: ssfm_match_pairs (tcgen05 kernel) against oracle/match_oracle.py, the
restatement of match() + cv::BFMatcher::knnMatch (examples/spherical_sfm_tools.cpp:235-251, 575-600).  Bit-exact index pairs."""
import os
import sys

import numpy as np
import pytest

from conftest import THR2

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import match_oracle as MO  # noqa: E402


def sift_like(rng, n, dup_from=None, ndup=0):
    """Integer descriptors 0..255 with SIFT-like sparsity; optionally exact duplicates (ties) of rows of another image."""
    d = np.minimum(rng.gamma(0.6, 30.0, (n, 128)), 255).astype(np.int32).astype(np.float32)
    if dup_from is not None and ndup > 0 and n > 0 and len(dup_from) > 0:
        src = rng.integers(0, len(dup_from), ndup)
        dst = rng.integers(0, n, ndup)
        d[dst] = dup_from[src]
    return d


def test_oracle_known_answers():
    """Hand-checkable cases of the restated cv::BFMatcher semantics: ratio test, overwrite order, ties -> lower train index."""
    e = np.zeros((6, 128), np.float32)
    for k in range(6):
        e[k, k] = 10.0 * (k + 1)
    train = e[:4].copy()
    query = np.stack([e[1], e[1], e[3] + e[0] * 0.1, e[5]])
    query = np.rint(query).astype(np.float32)
    m = MO.match(train, query, 0.75)
    # queries 0 and 1 equal train 1 (distance 0 < 0.75 * d1): train 1 -> query 1 (the later query overwrites query 0);
    # query 2 is train 3 + (1,0,..): nearest train 3 at distance 1; query 3 is far from everything: fails the ratio test
    assert m.tolist() == [[1, 1], [3, 2]]
    # exact tie between two train rows: the lower index is kept as nearest, and d0 == d1 fails the ratio test
    train2 = np.stack([e[0], e[2], e[2]])
    idx, dist = MO.knn2(train2, e[2:3])
    assert idx[0].tolist() == [1, 2] and dist[0, 0] == 0 and dist[0, 1] == 0
    assert len(MO.match(train2, e[2:3], 0.75)) == 0
    assert len(MO.match(train[:1], query, 0.75)) == 0  # fewer than two train descriptors


@pytest.mark.gpu
def test_match_pairs_bit_exact_with_oracle(S, engine):
    rng = np.random.default_rng(3)
    sizes = [300, 257, 1000, 5, 1, 0, 2, 700, 128, 256]
    descs = []
    for i, n in enumerate(sizes):
        descs.append(sift_like(rng, n, descs[0] if i > 0 else None, ndup=n // 3))
    # near-duplicates: rows of image 0 with +-1 on a few bins (close float distances, collisions after sqrtf)
    nd = descs[2]
    nd[:200] = descs[0][rng.integers(0, 300, 200)]
    nd[:200, :8] = np.clip(nd[:200, :8] + rng.integers(-1, 2, (200, 8)), 0, 255)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    allrows = np.concatenate(descs)
    pairs = [(0, 1), (1, 0), (0, 2), (2, 0), (2, 7), (3, 0), (0, 3), (4, 0), (0, 4), (5, 0), (0, 5), (6, 1), (1, 6), (8, 9), (9, 8), (0, 0),
             (7, 2), (2, 2)]
    mo, mm = engine.match_pairs(allrows, offs, np.array(pairs, np.int32), 0.75)
    oo, om = MO.match_exhaustive(descs, pairs, 0.75)
    assert mo.tolist() == oo.tolist()
    assert (mm == om).all()
    assert mo[-1] > 500  # the duplicates really produce matches
    # another ratio, same data
    mo2, mm2 = engine.match_pairs(allrows, offs, np.array(pairs[:6], np.int32), 0.9)
    oo2, om2 = MO.match_exhaustive(descs, pairs[:6], 0.9)
    assert mo2.tolist() == oo2.tolist() and (mm2 == om2).all()


@pytest.mark.gpu
def test_match_pairs_full_size_images(S, engine):
    """4000 x 4000 descriptors per pair (the reference keeps at most 4000 keypoints per image, :186): several train tiles and
    query blocks per pair, padded tails on both sides."""
    rng = np.random.default_rng(4)
    a = sift_like(rng, 4000)
    b = sift_like(rng, 3777, a, ndup=1500)
    b[rng.integers(0, 3777, 600), :4] += 1  # perturb some of the copies
    b = np.clip(b, 0, 255)
    offs = np.array([0, 4000, 7777], np.int64)
    mo, mm = engine.match_pairs(np.concatenate([a, b]), offs, np.array([[0, 1], [1, 0]], np.int32), 0.75)
    oo, om = MO.match_exhaustive([a, b], [(0, 1), (1, 0)], 0.75)
    assert mo.tolist() == oo.tolist() and (mm == om).all()
    assert mo[1] > 800


def _far_tie_images(rng, nq=60, nt=500):
    """Train rows at d^2 ~ 7.8e6 from every query, differing from each other by a few units: above 2^22 sqrtf maps
    neighbouring integers to the same float, and cv::BFMatcher compares the floats (first index wins a float tie)."""
    base_t = np.full(128, 255, np.int32)
    base_t[120:] = 7
    train = np.tile(base_t, (nt, 1))
    train[:, 120:] += (rng.random((nt, 8)) < 0.3).astype(np.int32)  # d^2 = D + number of bumped bins
    train[:, :4] -= rng.integers(0, 2, (nt, 4))
    query = np.zeros((nq, 128), np.int32)
    query[:, 120:] = 7
    query[:, 4:12] += rng.integers(0, 3, (nq, 8))
    return train.astype(np.float32), query.astype(np.float32)


def test_far_regime_data_separates_float_from_integer_ordering():
    """The far-regime case below has teeth: ordering by the exact integer d^2 gives a different 2-NN than ordering by
    sqrtf(d^2) in float, which is what the restated cv::BFMatcher (and the kernel) must do."""
    train, query = _far_tie_images(np.random.default_rng(12))
    d2 = ((query[:, None, :].astype(np.int64) - train[None, :, :].astype(np.int64)) ** 2).sum(2)
    assert d2.min() > 2 ** 22
    idx, _ = MO.knn2(train, query)
    by_integer = np.argsort(d2, axis=1, kind="stable")[:, :2]
    assert (by_integer != idx).any(1).sum() > 10


@pytest.mark.gpu
def test_match_far_regime_float_ties(S, engine):
    """Squared distances above 2^22 (see _far_tie_images), mixed with near rows so both regimes and the switch between
    them are walked; several ratios, including > 1 so that float-tied neighbours pass the test and show up in the output."""
    rng = np.random.default_rng(12)
    descs, pairs = [], []
    for k in range(12):
        train, query = _far_tie_images(rng, nq=40 + 7 * k, nt=300 + 31 * k)
        if k % 3 == 1:  # a few near rows in the middle of the train image: the second neighbour drops below 2^22 there
            train[150:153] = query[:3]
        if k % 3 == 2:  # exactly one near row: the nearest is near, the second stays far
            train[200] = query[5]
        descs += [train, query]
        pairs += [(2 * k, 2 * k + 1), (2 * k + 1, 2 * k)]
    offs = np.concatenate([[0], np.cumsum([len(d) for d in descs])]).astype(np.int64)
    for ratio in (0.75, 1.0, 1.5):
        mo, mm = engine.match_pairs(np.concatenate(descs), offs, np.array(pairs, np.int32), ratio)
        oo, om = MO.match_exhaustive(descs, pairs, ratio)
        assert mo.tolist() == oo.tolist() and (mm == om).all()
    assert oo[-1] > 20  # the float-tied neighbours that pass at ratio 1.5 do show up in the output


@pytest.mark.gpu
def test_match_pairs_rejects_non_integer_descriptors(S, engine):
    rng = np.random.default_rng(5)
    d = sift_like(rng, 64)
    d[3, 7] += 0.5
    with pytest.raises(S.SsfmError) as e:
        engine.match_pairs(d, np.array([0, 32, 64], np.int64), np.array([[0, 1]], np.int32))
    assert e.value.code == S.SSFM_ERR_INVALID


@pytest.mark.gpu
def test_matches_feed_the_pose_engine(S, engine):
    """ssfm_match_pairs output is exactly SsfmMatchBatch.matches: descriptors -> matches -> poses without reshaping."""
    rng = np.random.default_rng(6)
    n = 600
    pr = S.problems.make_problem(S.problems.make_rng(81, 0), n, False, None, 1 / 600, 0, 20.0)
    f = 600.0
    kp0 = (pr.rays[:, 0:2] * f).astype(np.float32)
    kp1 = (pr.rays[:, 3:5] * f).astype(np.float32)
    d0 = sift_like(rng, n)
    perm = rng.permutation(n)
    d1 = d0[perm].copy()  # keypoint perm[k] of image 0 reappears as keypoint k of image 1
    kp1 = kp1[perm]
    offs = np.array([0, n, 2 * n], np.int64)
    mo, mm = engine.match_pairs(np.concatenate([d0, d1]), offs, np.array([[0, 1]], np.int32), 0.75)
    assert mo[-1] > 0.9 * n and (perm[mm[:, 1]] == mm[:, 0]).all()
    Kinv = np.array([[1 / f, 0, 0], [0, 1 / f, 0], [0, 0, 1.0]])
    res, flags = engine.estimate_pairs_from_matches(np.concatenate([kp0, kp1]), offs, np.array([[0, 1]], np.int32), mm, mo, Kinv,
                                                    S.pipeline_options(THR2))
    assert res["status"][0] == 0 and res["best_num_inliers"][0] > 0.9 * mo[-1]
    assert np.rad2deg(S.problems.rot_error(pr.R, S.problems.so3exp(res["r"][0]))) < 0.2
