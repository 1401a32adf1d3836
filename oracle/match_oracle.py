"""CPU ORACLE for descriptor matching (test infrastructure; only tests/, smoke() and bench.py's CPU legs may import it).

Restates match() of the reference (examples/spherical_sfm_tools.cpp:235-251):

    cv::BFMatcher matcher;                                   // NORM_L2, crossCheck = false
    matcher.knnMatch(features1.descs, features0.descs, matches, 2);   // query = image 1, train = image 0
    for i: if (matches[i][0].distance < ratio * matches[i][1].distance) m01[matches[i][0].trainIdx] = matches[i][0].queryIdx;

and the part of OpenCV 4.10 (docker/Dockerfile pins 4.10.0; OpenCV is not installed here) that decides the result:
cv::batchDistance with NORM_L2 / CV_32F computes dist = std::sqrt(normL2Sqr(q, t)) in float (modules/core/src/batch_distance.cpp:
batchDistL2_32f), starts from dist = FLT_MAX, idx = -1, and inserts train index j in increasing j with
`if (d < dist[K-1]) { shift while dist[k] > d; ... }` -- strict comparisons, so on equal float distances the LOWER train index
stays in front.  That is a stable sort by (float distance, index), taken to two places.

Parity status: pinned by construction on integer-valued descriptors (cv::SIFT: integers 0..255 stored as float), where the
float accumulation of normL2Sqr is exact whatever the SIMD summation order; the reference has no test or golden vector for
this function and OpenCV cannot be run here, so beyond that it is "parity unpinned".
"""
import numpy as np


def knn2(desc_train, desc_query):
    """(idx[n_query, 2], dist[n_query, 2]) float32 distances as cv::BFMatcher::knnMatch(query, train, k=2) returns them."""
    t = np.asarray(desc_train, np.float32)
    q = np.asarray(desc_query, np.float32)
    idx = np.full((len(q), 2), -1, np.int64)
    dist = np.full((len(q), 2), np.finfo(np.float32).max, np.float32)
    if len(t) == 0:
        return idx, dist
    integer = bool((t == np.rint(t)).all() and (q == np.rint(q)).all() and np.abs(t).max(initial=0) < 4096 and np.abs(q).max(initial=0) < 4096)
    for i0 in range(0, len(q), 256):
        qq = q[i0:i0 + 256]
        if integer:  # exact integer arithmetic == float accumulation for these values
            ti, qi = t.astype(np.int64), qq.astype(np.int64)
            d2 = ((qi * qi).sum(1)[:, None] + (ti * ti).sum(1)[None, :] - 2 * (qi @ ti.T)).astype(np.float32)
        else:  # sequential float accumulation (OpenCV's SIMD order may differ in the last bit here)
            d2 = np.zeros((len(qq), len(t)), np.float32)
            for k in range(t.shape[1]):
                df = qq[:, k][:, None] - t[:, k][None, :]
                d2 += df * df
        d = np.sqrt(d2)  # float32
        order = np.argsort(d, axis=1, kind="stable")[:, :2]
        for c in range(order.shape[1]):
            idx[i0:i0 + len(qq), c] = order[:, c]
            dist[i0:i0 + len(qq), c] = np.take_along_axis(d, order[:, c:c + 1], 1)[:, 0]
    return idx, dist


def match(desc0, desc1, ratio=0.75):
    """match(features0, features1, m01, ratio): the Matches map as an array of (index in image 0, index in image 1) in the
    map's iteration order (sorted by the first)."""
    if len(desc0) < 2:  # knnMatch(k=2) returns fewer than two neighbours: the reference would read past the vector
        return np.zeros((0, 2), np.int32)
    idx, dist = knn2(desc0, desc1)
    m01 = {}
    for i in range(len(desc1)):
        if float(dist[i, 0]) < ratio * float(dist[i, 1]):
            m01[int(idx[i, 0])] = i  # later queries overwrite earlier ones
    return np.array(sorted(m01.items()), np.int32).reshape(-1, 2)


def match_exhaustive(descs, pairs, ratio=0.75):
    """match_exhaustive (examples/spherical_sfm_tools.cpp:575-600) for the listed pairs: (offsets, matches)."""
    out, offs = [], [0]
    for i0, i1 in pairs:
        m = match(descs[i0], descs[i1], ratio)
        out.append(m)
        offs.append(offs[-1] + len(m))
    return np.array(offs, np.int64), (np.concatenate(out) if out else np.zeros((0, 2), np.int32))
