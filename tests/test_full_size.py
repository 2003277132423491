"""The CUDA path at BASELINE.json's full per-GPU sizes, through size-independent properties and
oracle checks on seeded subsets (the oracle finishes those in seconds):

  config 2   100 000 contigs of 1-50 kb (1.25 Gbp), 10 samples, 200 bins
  config 3   one 8-GPU shard of the 1 M-contig set: 125 000 contigs, 50 samples, 500 bins
  config 5   one 8-GPU shard of the 5 M-contig set: 625 000 contigs, 100 samples, 1000 bins
  config 4   Flye-style contigs of 10 kb-5 Mb (scaled to ~1.6 Gbp)

Properties: every count row sums to len - 3; a contig's profile and its assignment do not
depend on which batch (and therefore which kernel path: the two-kernel path for >= 2048 queries,
the fused pair kernel below) it was scored in; the host/GPU split of the upload does not matter.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (the workload generator of the benchmark)
from metacoag_b200 import _lib  # noqa: E402
from oracle import oracle  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 1e-9          # BASELINE.json: probabilities / weights within 1e-9 relative
MAXW = _lib.MAX_WEIGHT


def build(ctx, kind, n, S, B, seed):
    old = bench.LENGTH_KIND
    bench.LENGTH_KIND = kind
    try:
        tabs = bench.workload_tables(seed, n, S, B)
    finally:
        bench.LENGTH_KIND = old
    blob = ctx.synth_bases(tabs["offsets"], tabs["genome_of"], tabs["trans"], seed)
    return tabs, blob


def close(got, want):
    got, want = np.asarray(got), np.asarray(want)
    same = (want == MAXW) | (got == MAXW)
    assert np.array_equal(got[same], want[same])
    err = np.abs(got[~same] - want[~same]) / np.abs(want[~same])
    assert err.size == 0 or err.max() <= RTOL, err.max()


def check_scoring(ctx, tabs, blob, n_oracle, seed):
    """Whole set on the two-kernel path; a seeded subset again on its own (fused kernel) and
    through the oracle."""
    offsets, cov = tabs["offsets"], tabs["coverage"]
    n = offsets.shape[0] - 1
    ctx.load_contigs(blob, offsets)
    ctx.scan()
    counts, freqs = ctx.get_profiles(want_counts=True)
    lens = np.diff(offsets)
    assert np.array_equal(counts.sum(axis=1, dtype=np.int64), np.maximum(lens - 3, 0))
    ctx.load_coverage(cov)
    cen_prof, cen_cov = ctx.bin_profiles(tabs["members"], tabs["seg"])
    arg, w, _ = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov)
    ctx.set_prune(False)      # every pair scored: not a bit may differ (tests/test_prune.py)
    arg_all, w_all, _ = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov)
    ctx.set_prune(True)
    assert np.array_equal(arg, arg_all) and np.array_equal(w, w_all)
    assert arg.shape == (n,) and (arg >= -1).all() and (arg < cen_prof.shape[0]).all()
    assert np.array_equal(w == MAXW, arg == -1)

    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(n, n_oracle, replace=False))
    # the same contigs as a small batch: different kernel, same answer
    arg_s, w_s, _ = ctx.score_centroids(queries=pick)
    close(w_s, w[pick])
    decided = w[pick] != MAXW
    assert np.array_equal(arg_s[decided], arg[pick][decided])
    # and through the oracle, from the bases
    sub_off = np.zeros(n_oracle + 1, dtype=np.int64)
    np.cumsum(lens[pick], out=sub_off[1:])
    sub_blob = np.concatenate([blob[offsets[i]:offsets[i + 1]] for i in pick])
    want_counts, want_prof = oracle.count_kmers_batch(sub_blob, sub_off, nthreads=0)
    assert np.array_equal(counts[pick].astype(np.float64), want_counts)
    assert np.array_equal(freqs[pick], want_prof)
    want_arg, want_w, _ = oracle.score_centroids(want_prof, cov[pick], cen_prof, cen_cov, nthreads=0)
    close(w[pick], want_w)
    assert np.array_equal(arg[pick], want_arg)
    return arg, w


def test_config2_full_size():
    ctx = _lib.Context(0)
    try:
        tabs, blob = build(ctx, "spades", 100000, 10, 200, 1002)
        assert blob.shape[0] > 1.2e9
        arg, w = check_scoring(ctx, tabs, blob, 400, 1)
        # most contigs go back to the genome they were drawn from
        assert (arg == tabs["genome_of"]).mean() > 0.5
        # the plug-in call from host buffers scans and scores contig ranges while the rest of
        # the blob is still uploading: same answer, whatever the split and the wave boundaries
        ctx.load_coverage(tabs["coverage"])
        cen_prof, cen_cov = ctx.bin_profiles(tabs["members"], tabs["seg"])
        for threads in (-1, 3, 0):
            ctx.set_host_pack_threads(threads)
            arg3, w3 = ctx.featurise_score(blob, tabs["offsets"], tabs["coverage"], cen_prof, cen_cov)
            assert np.array_equal(arg3, arg) and np.array_equal(w3, w)
            counts3 = ctx.get_profiles(want_counts=True)[0]
            assert np.array_equal(counts3.sum(axis=1, dtype=np.int64),
                                  np.maximum(np.diff(tabs["offsets"]) - 3, 0))
        # the upload split between copy engine and host pack threads changes nothing
        ctx.set_host_pack_threads(0)
        ctx.load_contigs(blob, tabs["offsets"])
        ctx.step_resident()
        arg2, w2 = ctx.get_assignments()
        assert np.array_equal(arg2, arg) and np.array_equal(w2, w)
        # the two halves of the step (multi-GPU callers put a broadcast between them), with SMs
        # kept out of the persistent scan's grid, and every pair scored: same bits
        ctx.set_reserved_sms(4)
        ctx.set_prune(False)
        ctx.step_featurise()
        ctx.step_score()
        ctx.set_prune(True)
        ctx.set_reserved_sms(0)
        arg4, w4 = ctx.get_assignments()
        assert np.array_equal(arg4, arg) and np.array_equal(w4, w)
        with pytest.raises(_lib.McgError):
            ctx.set_reserved_sms(100000)
    finally:
        ctx.close()


def test_config3_shard_full_size():
    ctx = _lib.Context(0)
    try:
        tabs, blob = build(ctx, "megahit", 125000, 50, 500, 1003)
        check_scoring(ctx, tabs, blob, 150, 2)
        assert 0 <= ctx.last_redo_count() <= 125000
    finally:
        ctx.close()


def test_config5_shard_full_size():
    """One 8-GPU shard of configs[4]: 625 000 contigs x 1000 bins x 100 samples (the guarded
    S > 30 path, 16 bin tiles, chunked rows), pruned == all pairs bit for bit, 120 seeded contigs
    through the oracle."""
    ctx = _lib.Context(0)
    try:
        tabs, blob = build(ctx, "megahit", 625000, 100, 1000, 1005)
        check_scoring(ctx, tabs, blob, 120, 5)
        assert 0 <= ctx.last_redo_count() <= 625000
    finally:
        ctx.close()


def test_config4_flye_lengths(monkeypatch):
    ctx = _lib.Context(0)
    try:
        tabs, blob = build(ctx, "flye", 2000, 10, 200, 1004)
        assert np.diff(tabs["offsets"]).max() > 3_000_000
        arg, w = check_scoring(ctx, tabs, blob, 12, 3)
        # fewer than 2048 contigs but > 64 MB of bases: the plug-in call runs in waves and must
        # still take the kernels (and give the bits) of the staged call, for any thread count
        # and wave size
        cen_prof, cen_cov = ctx.bin_profiles(tabs["members"], tabs["seg"])
        rng = np.random.default_rng(4)
        for _ in range(10):
            ctx.set_host_pack_threads(int(rng.choice([0, 1, 3, 7, 15, 24])))
            monkeypatch.setenv("MCG_WAVE_DIV", str(int(rng.choice([2, 5, 12, 30]))))
            arg3, w3 = ctx.featurise_score(blob, tabs["offsets"], tabs["coverage"], cen_prof, cen_cov)
            assert np.array_equal(arg3, arg) and np.array_equal(w3, w)
    finally:
        ctx.close()
