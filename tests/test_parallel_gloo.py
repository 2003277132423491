"""The N > 1 path of metacoag_b200.parallel on CPU: world_size 2, gloo backend.

Checks the base-balanced partition, the centroid broadcast and the ragged all-gather of
assignments; the scoring itself is replaced by a rank-independent function of the contig id,
so the gathered result must equal the single-process one.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metacoag_b200 import parallel


def test_partition_by_bases_balances_bases_not_counts():
    rng = np.random.default_rng(0)
    lens = np.concatenate([rng.integers(1000, 3000, 900), rng.integers(1_000_000, 5_000_000, 20)])
    rng.shuffle(lens)
    offsets = np.concatenate([[0], np.cumsum(lens)])
    for world in (1, 2, 4, 8):
        bounds = parallel.partition_by_bases(offsets, world)
        assert bounds[0] == 0 and bounds[-1] == len(lens) and np.all(np.diff(bounds) >= 0)
        per_rank = np.array([offsets[bounds[r + 1]] - offsets[bounds[r]] for r in range(world)])
        assert per_rank.sum() == offsets[-1]
        # no shard is off the mean by more than the longest contig
        assert np.abs(per_rank - offsets[-1] / world).max() <= lens.max()
    lo, hi, sub_off, sub_blob, sub_cov = parallel.shard(offsets, np.zeros(int(offsets[-1]), np.uint8),
                                                         None, 1, 2)
    assert sub_off[0] == 0 and sub_off[-1] == sub_blob.shape[0] and hi - lo == len(sub_off) - 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_contigs, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        lens = rng.integers(1000, 50000, n_contigs)
        offsets = np.concatenate([[0], np.cumsum(lens)])
        bounds = parallel.partition_by_bases(offsets, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        # rank 0 owns the bin tables; the others start from garbage
        cen_prof = torch.arange(5 * 136, dtype=torch.float64).reshape(5, 136) if rank == 0 \
            else torch.full((5, 136), -1.0, dtype=torch.float64)
        cen_cov = torch.arange(5 * 3, dtype=torch.float64).reshape(5, 3) if rank == 0 \
            else torch.full((5, 3), -1.0, dtype=torch.float64)
        parallel.broadcast_centroids(cen_prof, cen_cov, src=0)
        assert torch.equal(cen_prof, torch.arange(5 * 136, dtype=torch.float64).reshape(5, 136))
        assert torch.equal(cen_cov, torch.arange(5 * 3, dtype=torch.float64).reshape(5, 3))
        # "score" the local shard, then gather (ragged shards)
        ids = torch.arange(lo, hi)
        argmin = (ids % 5).to(torch.int32)
        wmin = ids.to(torch.float64) * 0.5 + cen_prof[0, 1]
        all_arg, all_w = parallel.gather_assignments(argmin, wmin)
        want_ids = torch.arange(0, n_contigs)
        assert torch.equal(all_arg, (want_ids % 5).to(torch.int32))
        assert torch.equal(all_w, want_ids.to(torch.float64) * 0.5 + 1.0)
        # equal shards take the single-collective path
        eq_arg, eq_w = parallel.gather_assignments(torch.full((4,), rank, dtype=torch.int32),
                                                   torch.full((4,), float(rank), dtype=torch.float64),
                                                   counts=[4] * world)
        assert eq_arg.tolist() == [r for r in range(world) for _ in range(4)]
        open(os.path.join(result_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_broadcast_and_gather_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 1001, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


# ---- label propagation with the candidate generation sharded over two ranks ---------------------
def _label_prop_worker(rank, world, port, golden_dir, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import json
        import sys
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path.insert(0, here)
        sys.path.insert(0, os.path.join(here, "golden"))
        import test_pipeline
        from oracle_backend import OracleBackend
        from metacoag_b200 import _backend
        parallel.MIN_SHARDED_NODES = 2          # the golden assembly is small
        _backend.set_backend(OracleBackend())
        states, want = test_pipeline.run(golden_dir, os.path.join(result_dir, "work%d" % rank))
        test_pipeline.compare(states, want, exact_floats=True)
        assert parallel.stats["sharded_calls"] > 0 and parallel.stats["rows_gathered"] > 0
        json.dump({"final": {str(k): v for k, v in states["final_label_prop"].items()},
                   "sharded_calls": parallel.stats["sharded_calls"]},
                  open(os.path.join(result_dir, "rank%d.json" % rank), "w"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_label_prop_sharded_candidates_world2(golden_dir, tmp_path):
    """Both ranks shard the BFS + rule-B work of label_prop / final_label_prop, all-gather the
    candidate rows and run the sequential part replicated: same bins as the reference after
    every stage, on both ranks, with the same number of collectives."""
    import json
    world = 2
    for r in range(world):
        os.makedirs(tmp_path / ("work%d" % r))
    mp.spawn(_label_prop_worker, args=(world, _free_port(), golden_dir, str(tmp_path)),
             nprocs=world, join=True)
    res = [json.load(open(tmp_path / ("rank%d.json" % r))) for r in range(world)]
    assert res[0]["final"] == res[1]["final"]
    assert res[0]["sharded_calls"] == res[1]["sharded_calls"] > 0
