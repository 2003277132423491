"""The tensor-core distance bound (csrc/mcg_tc.cu: tcgen05.mma on an fp16 hi/lo split, TMA
operands, fp32 accumulators in tensor memory) against float64: every approximate squared
distance lies within the margin the library states, with room to spare, on real profiles and
on adversarial rows.  Reference arithmetic: matching_utils.py:28-29 (get_tetramer_distance).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from metacoag_b200 import _lib, synth  # noqa: E402
from oracle import oracle  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def _rows(seed, n):
    asm = synth.make_assembly(seed=seed, n_genomes=9, n_contigs=n, n_samples=2,
                              length_kind="mixed", with_graph=False, with_markers=False)
    _, prof = oracle.count_kmers_batch(asm.blob, asm.offsets, nthreads=0)
    return asm, prof


def _check(ctx, prof, cen, typical=True):
    n, B = prof.shape[0], cen.shape[0]
    ctx.set_profiles(prof)
    ctx.load_coverage(np.ones((n, 2)))
    ctx.set_centroids(cen, np.ones((B, 2)))
    d2a, margin = ctx.tc_distances(nq=n)
    d2 = np.zeros((n, B))
    for lo in range(0, n, 512):
        diff = prof[lo:lo + 512, None, :] - cen[None, :, :]
        d2[lo:lo + 512] = np.einsum("qbk,qbk->qb", diff, diff)
    err = np.abs(d2a.astype(np.float64) - d2)
    ratio = err / margin[:, None].astype(np.float64)
    assert ratio.max() <= 0.5, (ratio.max(), np.unravel_index(ratio.argmax(), ratio.shape))
    # the margin itself is small against the distances the bound has to separate
    # (d^2 ~ 1e-4..4e-2); rows far from every genome (a lower-cased contig counts as poly-A:
    # |x| ~ 1) have a larger one, and are far from every bin by much more than that
    assert margin.max() < 1e-4 and (not typical or np.median(margin) < 2e-6)
    print("max |d2a - d2| / margin = %.3f, median margin %.2e, max %.2e" % (
        ratio.max(), np.median(margin), margin.max()))
    return ratio.max()


@pytest.mark.parametrize("n,B", [(1000, 200), (300, 7), (129, 130), (4000, 513), (128, 128), (1, 1)])
def test_bound_holds_on_contig_profiles(ctx, n, B):
    asm, prof = _rows(100 + B, max(n, B + 5))
    rng = np.random.default_rng(n)
    cen = np.stack([prof[rng.choice(prof.shape[0], 5)].mean(axis=0) for _ in range(B)])
    _check(ctx, prof[:n], cen)


def test_bound_holds_on_adversarial_rows(ctx):
    """Rows equal to a centroid, all windows in one slot (poly-A contigs, lower-cased contigs:
    everything counts as A), uniform profiles, rows a hair apart, zeros (contigs shorter than 4)."""
    _, prof = _rows(7, 600)
    rng = np.random.default_rng(3)
    cen = np.stack([prof[rng.choice(600, 4)].mean(axis=0) for _ in range(150)])
    one_hot = np.zeros((136, 136))
    one_hot[np.arange(136), np.arange(136)] = 1.0
    codes = np.bincount(_lib.kmer_table(), minlength=136) / 256.0
    near = cen[:50] * (1.0 + 1e-7 * rng.normal(size=(50, 136)))
    rows = np.concatenate([cen[:60], one_hot, codes[None, :], np.zeros((3, 136)), near,
                           prof[:100] * 0.5 + 0.5 * one_hot[0][None, :]])
    cen2 = np.concatenate([cen, one_hot[:3], codes[None, :]])
    _check(ctx, rows, cen2, typical=False)
    d2a, margin = ctx.tc_distances(queries=np.arange(60))
    # a row that IS a centroid: |d2a| itself is within the margin of 0
    assert np.all(np.abs(d2a[np.arange(60), np.arange(60)]) <= margin)
