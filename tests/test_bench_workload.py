"""The synthetic workload of bench.py: shards of a multi-GPU run are different
contigs of ONE community -- the bins rank 0 sends out are the bins of every shard."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_shards_share_the_community():
    a = bench.workload_tables(1001, 500, 4, 20)
    b = bench.workload_tables(1001, 500, 4, 20, shard=3)
    again = bench.workload_tables(1001, 500, 4, 20, shard=3)
    assert np.array_equal(a["trans"], b["trans"])                 # same genomes
    assert not np.array_equal(a["genome_of"], b["genome_of"])     # other contigs
    assert not np.array_equal(a["lengths"], b["lengths"])
    assert np.array_equal(b["genome_of"], again["genome_of"])     # seeded
    # coverage of a contig follows its genome's abundance in both shards
    for t in (a, b):
        g = t["genome_of"]
        per_genome = np.array([np.median(t["coverage"][g == k], axis=0) for k in range(20)
                               if (g == k).sum() > 5])
        assert per_genome.std(axis=0).mean() > 1.0
    ga = np.array([np.median(a["coverage"][a["genome_of"] == k], axis=0) for k in range(20)])
    gb = np.array([np.median(b["coverage"][b["genome_of"] == k], axis=0) for k in range(20)])
    assert np.nanmedian(np.abs(ga - gb) / ga) < 0.15
    # shard 0 is the single-GPU workload, unchanged
    assert np.array_equal(bench.workload_tables(1001, 500, 4, 20, shard=0)["genome_of"], a["genome_of"])
