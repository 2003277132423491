"""Worker of tests/test_peers_multi.py: one process per GPU (torch.distributed.run), the
peer-memory exchange of DESIGN.md section 5b against the NCCL collectives it replaces.

Ragged shards (a different row count per rank, one of them empty when world > 2), both step
parities, pushes and barrier on a side stream as bench.py runs them, the reader_done ordering
rule, and the argument checks of push_assignments.  Prints PEER-OK on rank 0 when every rank
agrees."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from metacoag_b200 import _lib, parallel  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = _lib.Context(local)
    n_bins, n_samples = 37, 5
    counts = [1000 + 517 * r for r in range(world)]
    if world > 2:
        counts[1] = 0                      # an empty shard
    ok = True
    try:
        peer = parallel.PeerExchange(counts, n_bins, n_samples, dev)
    except Exception as err:               # no peer access between these GPUs
        if rank == 0:
            print("PEER-UNAVAILABLE %s" % err)
        dist.destroy_process_group()
        return
    comm = torch.cuda.Stream(device=dev)
    rng = np.random.default_rng(100 + rank)
    for step in range(4):
        par = step & 1
        arg = torch.from_numpy(rng.integers(-1, n_bins, counts[rank]).astype(np.int32)).to(dev)
        w = torch.from_numpy(rng.random(counts[rank])).to(dev)
        # every rank holds different tables: only rank 0's may come out
        cen_prof = torch.from_numpy(np.random.default_rng([step, rank]).random((n_bins, 136))).to(dev)
        cen_cov = torch.from_numpy(np.random.default_rng([step, rank, 7]).random((n_bins, n_samples))).to(dev)
        torch.cuda.synchronize()
        with torch.cuda.stream(comm):
            if rank == 0:
                peer.push_centroids(ctx, par, cen_prof, cen_cov, stream=comm)
            peer.barrier(channel=0, stream=comm)
            peer.push_assignments(ctx, par, arg, w, stream=comm)
            peer.barrier(channel=1, stream=comm)
            got_a, got_w = peer.assignments(par)
            got_p, got_c = peer.centroids(par)
            got = [t.clone() for t in (got_a, got_w, got_p, got_c)]
            peer.reader_done(par, comm)
        comm.synchronize()
        want_a, want_w = parallel.gather_assignments(arg, w, counts=counts)
        want_p, want_c = cen_prof.clone(), cen_cov.clone()
        parallel.broadcast_centroids(want_p, want_c, src=0)
        torch.cuda.synchronize()
        ok = ok and torch.equal(got[0], want_a) and torch.equal(got[1], want_w)
        ok = ok and torch.equal(got[2], want_p) and torch.equal(got[3], want_c)
    # a shard of the wrong size is refused, not read out of bounds
    try:
        peer.push_assignments(ctx, 0, torch.zeros(counts[rank] + 1, dtype=torch.int32, device=dev),
                              torch.zeros(counts[rank] + 1, dtype=torch.float64, device=dev))
        ok = False
    except ValueError:
        pass
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PEER-OK" if int(flag.item()) == 1 else "PEER-MISMATCH")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
