"""Rule C at scale is a branch and bound (include/mcg_b200.h, mcg_set_prune): bins whose
tetramer distance alone puts their weight above the contig's current minimum are not scored.
It must not change a bit of the result: argmin (first index of the minimum,
label_prop_utils.py:340) and minimum weight with pruning == without == the weight matrix's
row minima == the oracle, on sets built to stress the bound -- duplicated centroids (exact
ties: the lower index has to win even when the later copy would be pruned first), crowded
centroids (many bins inside the bound), contigs far from every bin (nothing to prune with),
coverage mismatches that make the nearest bin lose, S on both sides of the guard (30 / 31+),
one and several bin groups."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from metacoag_b200 import _lib  # noqa: E402
from oracle import oracle  # noqa: E402

pytestmark = pytest.mark.gpu
MAXW = _lib.MAX_WEIGHT


def make_set(seed, n, B, S, spread):
    rng = np.random.default_rng(seed)
    base = rng.dirichlet(np.full(136, 40.0), size=B)
    # crowd: a third of the centroids are small steps away from centroid 0 .. 9
    crowd = rng.choice(B, B // 3, replace=False)
    base[crowd] = base[crowd % 10] + rng.normal(0, spread, (crowd.shape[0], 136))
    base = np.abs(base)
    base /= base.sum(axis=1, keepdims=True)
    cen_cov = rng.lognormal(np.log(20), 1.0, (B, S))
    # exact duplicates: bins 3 and B-2 are the same bin twice; 7 and 8 share the profile only
    base[B - 2] = base[3]
    cen_cov[B - 2] = cen_cov[3]
    base[8] = base[7]
    owner = rng.integers(0, B, n)
    prof = np.abs(base[owner] + rng.normal(0, 4e-4, (n, 136)))
    prof /= prof.sum(axis=1, keepdims=True)
    cov = cen_cov[owner] * rng.normal(1.0, 0.1, (n, S)).clip(0.05)
    # a tenth of the contigs sit on their centroid exactly (d = 0), a tenth far from everything,
    # a tenth carry the coverage of another bin (the nearest bin loses)
    k = n // 10
    prof[:k] = base[owner[:k]]
    far = rng.dirichlet(np.full(136, 3.0), size=k)
    prof[k:2 * k] = far
    cov[2 * k:3 * k] = cen_cov[rng.integers(0, B, k)]
    cov[::97] = 1e-4
    return prof, np.maximum(cov, 1e-4), base, np.maximum(cen_cov, 1e-4)


@pytest.mark.parametrize("n,B,S,spread", [
    (4096, 200, 10, 2e-4),
    (3000, 97, 30, 1e-3),
    (2500, 500, 31, 2e-4),
    (2304, 1000, 50, 5e-4),
    (2100, 330, 100, 2e-4),
])
def test_pruned_argmin_is_the_argmin(n, B, S, spread):
    prof, cov, cen_prof, cen_cov = make_set(n + B + S, n, B, S, spread)
    ctx = _lib.Context(0)
    try:
        ctx.set_profiles(prof)
        ctx.load_coverage(cov)
        ctx.set_prune(True)
        arg_p, w_p, _ = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov)
        redo_p = ctx.last_redo_count() if S > 30 else 0
        ctx.set_prune(False)
        arg_f, w_f, _ = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov)
        arg_m, w_m, weights = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov,
                                                  want_weights=True)
        ctx.set_prune(True)
    finally:
        ctx.close()
    assert np.array_equal(arg_p, arg_f)
    assert np.array_equal(w_p, w_f)                     # bit for bit
    assert redo_p >= 0
    # the duplicated bin never wins under its higher index
    assert not (arg_p == B - 2).any() and (arg_p == 3).any()
    # the weight matrix (every pair scored, exact kernel for S > 30) agrees
    row_min = weights.min(axis=1)
    decided = row_min != MAXW
    assert np.array_equal(arg_p == -1, ~decided)
    assert np.array_equal(arg_p[decided], weights.argmin(axis=1)[decided])
    rel = np.abs(w_p[decided] - row_min[decided]) / row_min[decided]
    assert rel.max() <= 1e-9
    # and so does the oracle on a subset
    pick = np.arange(0, n, max(1, n // 300))
    want_arg, want_w, _ = oracle.score_centroids(prof[pick], cov[pick], cen_prof, cen_cov,
                                                 nthreads=0)
    assert np.array_equal(arg_p[pick], want_arg)
    both = want_w != MAXW
    assert np.array_equal(w_p[pick][~both], want_w[~both])
    assert (np.abs(w_p[pick][both] - want_w[both]) / want_w[both]).max() <= 1e-9


def test_candidate_list_overflow_falls_back_to_all_pairs():
    """Every centroid is the same bin: every bin is a candidate of every contig, the list would
    hold all pairs (more than its capacity, a quarter of them) and the device-side gate hands
    the batch to prob_kernel.  First index of the minimum: bin 0."""
    n, B, S = 2304, 96, 10
    rng = np.random.default_rng(5)
    one = rng.dirichlet(np.full(136, 40.0))
    cen_prof = np.tile(one, (B, 1))
    cen_cov = np.tile(rng.lognormal(np.log(20), 1.0, S), (B, 1))
    prof = np.abs(one + rng.normal(0, 3e-4, (n, 136)))
    prof /= prof.sum(axis=1, keepdims=True)
    cov = np.maximum(cen_cov[0] * rng.normal(1.0, 0.1, (n, S)), 1e-4)
    ctx = _lib.Context(0)
    try:
        ctx.set_profiles(prof)
        ctx.load_coverage(cov)
        arg_p, w_p, _ = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov)
        stats = ctx.prune_stats()
        ctx.set_prune(False)
        arg_f, w_f, _ = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov)
        assert ctx.prune_stats()["used"] is False
    finally:
        ctx.close()
    assert stats["candidates"] == n * B and stats["candidates"] > stats["capacity"]
    assert stats["used"] is False
    assert np.array_equal(arg_p, arg_f) and np.array_equal(w_p, w_f)
    assert (arg_p[w_p != MAXW] == 0).all() and (w_p != MAXW).mean() > 0.9


def test_pruned_equals_all_pairs_on_random_shapes():
    """Seeded sweep over bin / sample counts around every boundary of the candidate path (one bin,
    tile and slab edges of 32 / 64 bins, the 2048-bin limit of the candidate list, the S = 30 / 31
    guard, more than one bin group): pruned == all pairs, bit for bit."""
    rng = np.random.default_rng(2024)
    shapes = [(1, 1), (2, 3), (31, 10), (32, 30), (33, 31), (63, 7), (64, 33), (65, 2), (129, 12),
              (700, 40), (2048, 5), (2049, 5)]
    ctx = _lib.Context(0)
    try:
        for B, S in shapes:
            n = int(rng.integers(2048, 2600))
            prof, cov, cen_prof, cen_cov = make_set(int(rng.integers(1 << 30)), n, max(B, 12), S,
                                                    float(rng.choice([2e-4, 1e-3, 5e-3])))
            cen_prof, cen_cov = cen_prof[:B], cen_cov[:B]
            ctx.set_profiles(prof)
            ctx.load_coverage(cov)
            ctx.set_prune(True)
            arg_p, w_p, _ = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov)
            used = ctx.prune_stats()
            ctx.set_prune(False)
            arg_f, w_f, _ = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov)
            assert np.array_equal(arg_p, arg_f), (B, S)
            assert np.array_equal(w_p, w_f), (B, S)
            assert (used["pairs"] == n * B) == (B <= 2048), (B, S, used)
            # a gathered query list (rows through an index) takes the same path
            q = rng.permutation(n)[:2100].astype(np.int64)
            ctx.set_prune(True)
            arg_q, w_q, _ = ctx.score_centroids(queries=q)
            assert np.array_equal(arg_q, arg_f[q]) and np.array_equal(w_q, w_f[q]), (B, S)
    finally:
        ctx.set_prune(True)
        ctx.close()
