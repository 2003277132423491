"""Row-sharding of the featurise + score path over the GPUs of one box.

One process per GPU (``torch.distributed``, backend ``nccl``; ``gloo`` in the CPU tests).
Every contig is independent given the bin tables, so the data path has exactly two
exchange steps: rank 0 broadcasts the bin centroids, and every rank all-gathers the
per-contig ``(bin, weight)`` so that the sequential marker-gene pass of
label_prop_utils.assign_to_bins (label_prop_utils.py:348-391) can run anywhere.
Nothing else crosses ranks; the contig bases, profiles and coverages live only on the
owning GPU.
"""
import numpy as np
import torch
import torch.distributed as dist


def partition_by_bases(offsets, world_size):
    """Contiguous contig ranges with (nearly) equal total bases.

    Returns ``bounds`` int64[world_size+1]; rank r owns contigs bounds[r]:bounds[r+1].
    Balancing by bases rather than by contig count is what keeps the scan even when
    contig lengths span 1 kb .. 5 Mb (BASELINE.json config 4).
    """
    offsets = np.asarray(offsets, dtype=np.int64)
    n = offsets.shape[0] - 1
    total = int(offsets[-1] - offsets[0])
    bounds = np.zeros(world_size + 1, dtype=np.int64)
    bounds[world_size] = n
    for r in range(1, world_size):
        target = offsets[0] + (total * r) // world_size
        bounds[r] = int(np.searchsorted(offsets, target, side="left"))
    bounds[1:world_size] = np.clip(bounds[1:world_size], 0, n)
    return np.maximum.accumulate(bounds)


def shard(offsets, blob, coverage, rank, world_size):
    """The slice of (offsets, blob, coverage) that ``rank`` owns, rebased to offset 0."""
    bounds = partition_by_bases(offsets, world_size)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    base = int(offsets[lo])
    sub_off = np.asarray(offsets[lo:hi + 1], dtype=np.int64) - base
    sub_blob = blob[base:int(offsets[hi])]
    sub_cov = None if coverage is None else coverage[lo:hi]
    return lo, hi, sub_off, sub_blob, sub_cov


def broadcast_centroids(cen_prof, cen_cov, src=0):
    """In-place broadcast of the bin tables ([B,136] and [B,S] float64 tensors) from
    ``src``; one fused buffer so that it is a single collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return cen_prof, cen_cov
    b = cen_prof.shape[0]
    fused = torch.cat([cen_prof.reshape(b, -1), cen_cov.reshape(b, -1)], dim=1).contiguous()
    dist.broadcast(fused, src=src)
    k = cen_prof.shape[1]
    cen_prof.copy_(fused[:, :k])
    cen_cov.copy_(fused[:, k:])
    return cen_prof, cen_cov


def gather_assignments(argmin, wmin, counts=None):
    """All-gather the per-contig results of every rank, in rank (= contig) order.

    argmin int32[n_local], wmin float64[n_local]; shards may have different sizes
    (``counts`` = rows per rank, gathered if not given).  Returns (argmin, wmin) over all
    contigs, on the same device as the inputs.
    """
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return argmin, wmin
    world = dist.get_world_size()
    if counts is None:
        mine = torch.tensor([argmin.shape[0]], dtype=torch.int64, device=argmin.device)
        all_counts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(all_counts, mine)
        counts = [int(c.item()) for c in all_counts]
    width = max(counts)
    if all(c == width for c in counts):
        out_a = torch.empty(world * width, dtype=argmin.dtype, device=argmin.device)
        out_w = torch.empty(world * width, dtype=wmin.dtype, device=wmin.device)
        dist.all_gather_into_tensor(out_a, argmin.contiguous())
        dist.all_gather_into_tensor(out_w, wmin.contiguous())
        return out_a, out_w
    pad_a = torch.full((width,), -1, dtype=argmin.dtype, device=argmin.device)
    pad_w = torch.zeros(width, dtype=wmin.dtype, device=wmin.device)
    pad_a[:argmin.shape[0]] = argmin
    pad_w[:wmin.shape[0]] = wmin
    out_a = torch.empty(world * width, dtype=argmin.dtype, device=argmin.device)
    out_w = torch.empty(world * width, dtype=wmin.dtype, device=wmin.device)
    dist.all_gather_into_tensor(out_a, pad_a)
    dist.all_gather_into_tensor(out_w, pad_w)
    keep = torch.cat([torch.arange(r * width, r * width + c, device=argmin.device)
                      for r, c in enumerate(counts)])
    return out_a[keep], out_w[keep]
