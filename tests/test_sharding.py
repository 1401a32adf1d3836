"""Multi-GPU plumbing on CPU: pair sharding + the all-gather of per-pair result records, world_size 2,
gloo backend.  (The records are produced here by the oracle -- as the checker's stand-in for the GPU
engine -- because this test runs without a GPU; the NCCL path is the same code with tensors on cuda.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_all_pairs(S):
    offs = np.concatenate([[0], np.cumsum(np.random.default_rng(0).integers(0, 3000, 1001))])
    for w in (1, 2, 3, 8):
        b = S.sharding.partition_pairs(offs, w)
        assert b[0] == 0 and b[-1] == 1001 and all(b[i] <= b[i + 1] for i in range(w))
        loads = [offs[b[i + 1]] - offs[b[i]] for i in range(w)]
        assert max(loads) - min(loads) <= 2 * 3000
    # degenerate: fewer pairs than ranks, empty batch
    b = S.sharding.partition_pairs(np.array([0, 10, 20]), 8)
    assert b[0] == 0 and b[-1] == 2 and all(b[i] <= b[i + 1] for i in range(8))
    assert S.partition_pairs(np.array([0]), 4) == [0, 0, 0, 0, 0]
    # it is the library's C function (ssfm_partition_pairs, the one ssfm_estimate_pairs_multi shards with)
    assert S.sharding.partition_pairs(offs, 5) == S.partition_pairs(offs, 5)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import oracle as O
    import spherical_sfm_b200 as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    thr2 = (2.0 / 600.0) ** 2
    rays, offsets, _ = S.problems.make_batch(77, 7, 200, noise=1 / 600, outlier_frac=0.4)
    bounds = S.sharding.partition_pairs(offsets, world)
    my_rays, my_offs, p0 = S.sharding.shard(rays, offsets, rank, world)
    orc = O.load()
    opt = O.pipeline_options(thr2)
    rec = np.zeros(len(my_offs) - 1, S.RESULT_DTYPE)
    for i in range(len(rec)):
        r, _ = orc.estimate_pair(my_rays[my_offs[i]:my_offs[i + 1]], opt, p0 + i)
        rec["E"][i] = r.E
        rec["num_iterations"][i] = r.num_iterations
        rec["best_num_inliers"][i] = r.best_num_inliers
    counts = [bounds[i + 1] - bounds[i] for i in range(world)]
    table = S.sharding.allgather_results(rec, counts, dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, table["num_iterations"].tolist(), table["best_num_inliers"].tolist(), table["E"].tolist()))


def test_two_rank_gloo_allgather_matches_single_process(S, O, orc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rays, offsets, _ = S.problems.make_batch(77, 7, 200, noise=1 / 600, outlier_frac=0.4)
    opt = O.pipeline_options((2.0 / 600.0) ** 2)
    want_it, want_inl, want_E = [], [], []
    for p in range(7):
        r, _ = orc.estimate_pair(rays[offsets[p]:offsets[p + 1]], opt, p)
        want_it.append(r.num_iterations)
        want_inl.append(r.best_num_inliers)
        want_E.append(list(r.E))
    for rank, it, inl, E in outs:
        assert it == want_it and inl == want_inl and E == want_E
