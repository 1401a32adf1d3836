"""Multi-GPU plumbing: image pairs are independent (no cross-pair state anywhere in
estimate_pairwise, examples/spherical_sfm_tools.cpp:332-420), so pairs are sharded across ranks with
no data-path collective; only the fixed-size per-pair result records are all-gathered
(torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np


def partition_pairs(offsets, world_size):
    """Contiguous block partition of the pair list balanced by correspondence count
    (cost ~ N_pair x iterations; N_pair is the a-priori proxy).  Returns world_size+1 pair bounds.
    One implementation: the library's ssfm_partition_pairs (a host function; ssfm_estimate_pairs_multi uses the same)."""
    from . import partition_pairs as _pp
    return _pp(offsets, world_size)


def shard(rays, offsets, rank, world_size):
    """This rank's CSR slice: (rays_slice, offsets_slice (rebased to 0), first_pair)."""
    offsets = np.asarray(offsets, np.int64)
    b = partition_pairs(offsets, world_size)
    p0, p1 = b[rank], b[rank + 1]
    c0, c1 = int(offsets[p0]), int(offsets[p1])
    return rays[c0:c1], offsets[p0:p1 + 1] - c0, p0


def allgather_results(local_records, counts, dist, device=None):
    """All-gather per-pair records (a structured numpy array, see RESULT_DTYPE) from every rank.
    counts: pairs per rank.  Returns the concatenated table in global pair order."""
    import torch

    world = dist.get_world_size()
    itemsize = local_records.dtype.itemsize
    maxn = int(max(counts)) if len(counts) else 0
    buf = np.zeros(maxn * itemsize, np.uint8)
    raw = local_records.view(np.uint8).reshape(-1)
    buf[:raw.size] = raw
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = []
    for r in range(world):
        a = outs[r].cpu().numpy()[:int(counts[r]) * itemsize]
        parts.append(a.view(local_records.dtype))
    return np.concatenate(parts) if parts else local_records[:0]
