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


# ---- label propagation: candidate generation sharded over the ranks ---------------
# The sequential part of label_prop / final_label_prop (heap, marker-gene rules) runs replicated
# and deterministically on every rank; what is sharded is the batch work in front of it -- the
# bounded BFS from every unbinned long contig and the rule-B scores of its hits
# (label_prop_utils.py:24-100).  Every rank takes a contiguous slice of the start nodes and the
# (node, labelled node, bin, depth, score) rows are all-gathered, so that all ranks hold the
# same candidate sets, bit for bit, and stay in lock step.
MIN_SHARDED_NODES = 256     # below this a collective costs more than the work it saves
stats = {"sharded_calls": 0, "rows_gathered": 0}


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_slice(n, rank=None, world_size=None):
    """[lo, hi) of ``n`` items for this rank, contiguous, sizes differing by at most one."""
    world_size = world() if world_size is None else world_size
    rank = (dist.get_rank() if world_size > 1 else 0) if rank is None else rank
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _collective_device():
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_rows(ints, floats):
    """All-gather ragged row blocks: ``ints`` int64[k, a] and ``floats`` float64[k, b] of this
    rank -> the concatenation over ranks, in rank order (numpy in, numpy out)."""
    dev = _collective_device()
    ints = np.ascontiguousarray(ints, dtype=np.int64)
    floats = np.ascontiguousarray(floats, dtype=np.float64)
    k = ints.shape[0]
    w = dist.get_world_size()
    mine = torch.tensor([k], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(mine) for _ in range(w)]
    dist.all_gather(counts, mine)
    counts = [int(c.item()) for c in counts]
    width = max(counts + [1])
    pad_i = torch.zeros((width, ints.shape[1]), dtype=torch.int64, device=dev)
    pad_f = torch.zeros((width, floats.shape[1]), dtype=torch.float64, device=dev)
    if k:
        pad_i[:k] = torch.from_numpy(ints).to(dev)
        pad_f[:k] = torch.from_numpy(floats).to(dev)
    out_i = torch.empty((w * width, ints.shape[1]), dtype=torch.int64, device=dev)
    out_f = torch.empty((w * width, floats.shape[1]), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out_i, pad_i)
    dist.all_gather_into_tensor(out_f, pad_f)
    out_i, out_f = out_i.cpu().numpy(), out_f.cpu().numpy()
    keep = np.concatenate([np.arange(r * width, r * width + c) for r, c in enumerate(counts)])
    stats["sharded_calls"] += 1
    stats["rows_gathered"] += int(keep.shape[0])
    return out_i[keep], out_f[keep]


class PeerExchange:
    """The two exchange steps of the path as stores into peer memory (NVLink / NVSwitch)
    instead of NCCL collectives: every rank owns one symmetric buffer (torch's symmetric-memory
    allocator: CUDA VMM handles exchanged at rendezvous, every peer's buffer mapped into every
    process) holding, per step parity, the gathered assignments of all ranks and the bin
    tables.

    Ordering rules (all enforced here, not left to the caller):
      * ``push_*`` and ``barrier`` run on ONE stream: the ``stream`` argument, or torch's current
        stream when it is None -- never the library's private stream, which nothing would order
        against the signal.  ``barrier(stream=...)`` enters that stream before it signals.
      * shards are ragged (``rows_per_rank`` may be a list: partition_by_bases gives every rank a
        different row count); ``push_assignments`` checks the size of what it is given against
        this rank's row count.
      * a view returned by ``assignments(par)`` / ``centroids(par)`` is overwritten by the
        peers' pushes of the NEXT use of that parity.  A reader on another stream must call
        ``reader_done(par, stream)`` when its last read is queued; the next ``push_*`` of that
        parity from THIS rank waits for it, and the peers' pushes are behind the barrier this
        rank only passes after that wait (every exchange step ends with ``barrier``).

    Raises when symmetric memory cannot be set up (no peer access between the GPUs): callers
    fall back to ``broadcast_centroids`` / ``gather_assignments``.
    """

    def __init__(self, rows_per_rank, n_bins, n_samples, device):
        import torch.distributed._symmetric_memory as symm
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        if isinstance(rows_per_rank, (int, np.integer)):
            counts = [int(rows_per_rank)] * self.world
        else:
            counts = [int(c) for c in rows_per_rank]
        if len(counts) != self.world or min(counts) < 0:
            raise ValueError("rows_per_rank must give one row count per rank")
        self.counts = counts
        self.first = [0]
        for c in counts:
            self.first.append(self.first[-1] + c)
        self.rows = counts[self.rank]
        self.n_bins, self.n_samples = int(n_bins), int(n_samples)

        def up(x):
            return (x + 255) // 256 * 256
        total = self.first[-1]
        self.total = total
        self.off_arg = 0
        self.off_w = up(4 * total)
        self.off_prof = self.off_w + up(8 * total)
        self.off_cov = self.off_prof + up(8 * 136 * self.n_bins)
        self.slot = self.off_cov + up(8 * self.n_samples * self.n_bins)
        self.buf = symm.empty(2 * self.slot, dtype=torch.uint8, device=device)
        self.handle = symm.rendezvous(self.buf, dist.group.WORLD)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(self.ptrs) != self.world or not all(self.ptrs):
            raise RuntimeError("symmetric memory rendezvous gave no peer pointers")
        self.buf.zero_()
        self._read = [None, None]
        torch.cuda.synchronize()
        self.handle.barrier(channel=0)

    def _view(self, par, off, count, dtype):
        start = par * self.slot + off
        return self.buf[start:start + count * dtype.itemsize].view(dtype)

    @staticmethod
    def _stream(stream):
        return stream if stream is not None else torch.cuda.current_stream()

    def assignments(self, par):
        """(argmin int32[total rows], wmin float64[total rows]) of this rank's buffer, ranks in
        order; valid until the next exchange of the same parity (see the class notes)."""
        return (self._view(par, self.off_arg, self.total, torch.int32),
                self._view(par, self.off_w, self.total, torch.float64))

    def centroids(self, par):
        return (self._view(par, self.off_prof, 136 * self.n_bins, torch.float64).view(self.n_bins, 136),
                self._view(par, self.off_cov, self.n_samples * self.n_bins,
                           torch.float64).view(self.n_bins, self.n_samples))

    def reader_done(self, par, stream=None):
        """The last read of the parity-``par`` views is queued on ``stream``."""
        ev = torch.cuda.Event()
        ev.record(self._stream(stream))
        self._read[par] = ev

    def _before_push(self, par, st):
        ev, self._read[par] = self._read[par], None
        if ev is not None:
            st.wait_event(ev)

    def push_assignments(self, ctx, par, d_argmin, d_wmin, stream=None):
        """This rank's rows into everyone's gathered arrays (an all-gather by stores)."""
        if d_argmin.numel() != self.rows or d_wmin.numel() != self.rows:
            raise ValueError("rank %d owns %d rows, got %d / %d" % (
                self.rank, self.rows, d_argmin.numel(), d_wmin.numel()))
        if d_argmin.dtype != torch.int32 or d_wmin.dtype != torch.float64:
            raise ValueError("assignments must be int32 / float64")
        st = self._stream(stream)
        self._before_push(par, st)
        if self.rows == 0:
            return
        base = par * self.slot
        mine = self.first[self.rank]
        ctx.push_to_peers(d_argmin.data_ptr(), 4 * self.rows,
                          [p + base + self.off_arg + 4 * mine for p in self.ptrs], st.cuda_stream)
        ctx.push_to_peers(d_wmin.data_ptr(), 8 * self.rows,
                          [p + base + self.off_w + 8 * mine for p in self.ptrs], st.cuda_stream)

    def push_centroids(self, ctx, par, cen_prof, cen_cov, stream=None):
        """The source rank's bin tables into everyone's buffer (a broadcast by stores)."""
        if cen_prof.numel() != 136 * self.n_bins or cen_cov.numel() != self.n_samples * self.n_bins:
            raise ValueError("centroid tables must be [%d,136] and [%d,%d]" % (
                self.n_bins, self.n_bins, self.n_samples))
        st = self._stream(stream)
        self._before_push(par, st)
        base = par * self.slot
        ctx.push_to_peers(cen_prof.data_ptr(), cen_prof.numel() * 8,
                          [p + base + self.off_prof for p in self.ptrs], st.cuda_stream)
        ctx.push_to_peers(cen_cov.data_ptr(), cen_cov.numel() * 8,
                          [p + base + self.off_cov for p in self.ptrs], st.cuda_stream)

    def barrier(self, channel=0, stream=None):
        """All ranks meet on ``stream`` (default: torch's current stream): everything pushed on
        it before has landed everywhere when the barrier is passed."""
        with torch.cuda.stream(self._stream(stream)):
            self.handle.barrier(channel=channel)
