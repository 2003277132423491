"""The CPU oracle against the golden vectors taken from the live reference.

This is what "pins" oracle/: every oracle function is compared with values the
unmodified reference produced (tests/golden/make_golden.py).  Integer work and the
libm-identical float work are bit-exact; nothing here needs a GPU.
"""
import hashlib
import math
import os
import sys

import numpy as np
import pytest

from oracle import oracle

MAXW = sys.float_info.max


@pytest.fixture(scope="module")
def scan(golden_dir):
    return np.load(os.path.join(golden_dir, "scan.npz"))


@pytest.fixture(scope="module")
def score(golden_dir):
    return np.load(os.path.join(golden_dir, "score.npz"))


def _stripped(scan):
    blob, off = scan["raw_blob"], scan["raw_offsets"]
    parts = [oracle.strip_ascii(blob[off[i]:off[i + 1]].tobytes()) for i in range(len(off) - 1)]
    offsets = np.zeros(len(parts) + 1, dtype=np.int64)
    np.cumsum([len(p) for p in parts], out=offsets[1:])
    return np.frombuffer(b"".join(parts), dtype=np.uint8), offsets


def test_kmer_table_known_answer(scan):
    table, n = oracle.kmer_table()
    assert n == 136
    assert np.array_equal(table, scan["table"])
    # SURVEY.md section 8a row 2 known answer
    assert hashlib.sha256(table.tobytes()).hexdigest() == \
        "9a57f94a839828cfdd42650170aa4db6e761932f7724c8056ab4ea818d0bc72f"
    assert list(table[:32]) == list(range(31)) + [11]
    assert list(table[-16:]) == [135, 121, 91, 45, 133, 115, 81, 31, 130, 108, 70, 16, 126, 100,
                                 58, 0]


def test_count_kmers_bit_exact(scan):
    blob, offsets = _stripped(scan)
    counts, freqs = oracle.count_kmers_batch(blob, offsets)
    assert np.array_equal(counts, scan["counts"])
    assert np.array_equal(freqs, scan["freqs"])
    lens = np.diff(offsets)
    assert np.array_equal(counts.sum(axis=1), np.maximum(0, lens - 3))
    # threads give the same answer
    counts4, freqs4 = oracle.count_kmers_batch(blob, offsets, nthreads=4)
    assert np.array_equal(counts4, counts) and np.array_equal(freqs4, freqs)


def test_lgamma_matches_cpython():
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(1e-4, 6, 4000) + 1.0, np.exp(rng.uniform(0, 11, 4000)),
                         np.arange(1, 60, dtype=float)])
    for x in xs:
        assert oracle.lib().orc_lgamma(float(x)) == math.lgamma(float(x))


def test_gaussians_bit_exact(score):
    for d, gi, ge, pc in zip(score["dist"], score["pdf_intra"], score["pdf_inter"],
                             score["comp_prob"]):
        assert oracle.normpdf(d, 0.0, 0.01037897 / 2) == gi
        assert oracle.normpdf(d, 0.0676654, 0.03419337) == ge
        assert oracle.comp_prob(d) == pc
    # probed values of SURVEY.md section 8a row 7
    assert abs(oracle.comp_prob(0.0) - 0.979) < 1e-3
    assert oracle.comp_prob(0.21) == 0.0
    assert math.isnan(oracle.comp_prob(1.40))


@pytest.mark.parametrize("S", [1, 3, 10, 33, 50, 100])
def test_cov_probability_bit_exact(score, S):
    a, b, want = score["covA_%d" % S], score["covB_%d" % S], score["covprob_%d" % S]
    got = np.array([oracle.cov_prob(x, y) for x, y in zip(a, b)])
    assert np.array_equal(got, want)
    if S == 100:
        assert (want == 0.0).any()  # floor^S underflows


def test_pairs_bit_exact(score):
    prof, cov = score["prof"], score["cov"]
    n = prof.shape[0]
    for i in range(n):
        for j in range(n):
            ok, lp, pc, pv = oracle.pair_lp(prof[i], prof[j], cov[i], cov[j])
            assert oracle.euclidean(prof[i], prof[j]) == score["pair_dist"][i, j]
            assert pc == score["pair_pc"][i, j]
            assert pv == score["pair_pv"][i, j]
            assert lp == score["pair_lp"][i, j]
            assert ok == (lp != MAXW)
    assert (score["pair_lp"] == MAXW).any() and (score["pair_lp"] != MAXW).any()


def test_segment_mean_bit_exact(score):
    got_p = oracle.segment_mean(score["prof"], score["members"], score["seg_offsets"])
    got_c = oracle.segment_mean(score["cov"], score["members"], score["seg_offsets"])
    assert np.array_equal(got_p, score["cen_prof"])
    assert np.array_equal(got_c, score["cen_cov"])


def test_rule_c_matches_assign_long(score):
    arg, w, weights = oracle.score_centroids(score["prof"], score["cov"], score["cen_prof"],
                                             score["cen_cov"], want_weights=True)
    assert np.array_equal(arg, score["ruleC_arg"])
    assert np.array_equal(w, score["ruleC_w"])
    assert weights.shape == (score["prof"].shape[0], score["cen_prof"].shape[0])


def test_rule_b_matches_run_bfs_long(score):
    want = score["ruleB_w"]
    queries = np.flatnonzero(~np.isnan(want[:, 0]))
    got = oracle.score_members(score["prof"], score["cov"], queries, score["members"],
                               score["seg_offsets"], rule=1)
    assert np.array_equal(got, want[queries])


def test_rule_a_matches_match_contigs(score):
    want = score["ruleA_w"]
    queries = np.flatnonzero(~np.isnan(want[:, 0]))
    assert queries.size > 0
    got = oracle.score_members(score["prof"], score["cov"], queries, score["members"],
                               score["seg_offsets"], rule=0, nthreads=2)
    assert np.array_equal(got, want[queries])


def test_cov_terms(score):
    logc, lgam = oracle.cov_terms(score["cov"])
    flat = score["cov"].reshape(-1)
    assert np.array_equal(logc.reshape(-1), np.array([math.log(c) for c in flat]))
    assert np.array_equal(lgam.reshape(-1), np.array([math.lgamma(c + 1.0) for c in flat]))


def test_bfs_long_matches_the_reference_walk(golden_dir):
    """orc_bfs_long against run_bfs_long of the live reference (tests/golden/bfs.json, written by
    tests/golden/make_bfs.py): the set of (labelled node, bin, depth) per start node and depth
    limit, including the depths the overwrite behaviour (:93-98) produces."""
    import json
    from metacoag_b200 import synth
    cases = json.load(open(os.path.join(golden_dir, "bfs.json")))
    for name, case in cases.items():
        indptr, indices = synth.csr_from_edges(case["n"], np.array(case["edges"]))
        labels = np.array(case["labels"], dtype=np.int32)
        for limit, per_start in case["expect"].items():
            for s, want in zip(case["starts"], per_start):
                got = oracle.bfs_long(indptr, indices, labels, s, int(limit))
                assert sorted(set(got)) == [tuple(t) for t in want], (name, limit, s)


def test_synth_generator_restated():
    """orc_synth_bases draws the bytes of the library's generator (same splitmix64 streams);
    here: determinism, alphabet and a prefix being independent of what follows it.  The
    byte-for-byte comparison with mcg_synth_bases needs a GPU (tests/test_resident.py)."""
    rng = np.random.default_rng(0)
    lengths = rng.integers(100, 5000, 50)
    offsets = np.zeros(51, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    genome_of = rng.integers(0, 3, 50).astype(np.int32)
    trans = rng.dirichlet(np.ones(4), size=12).astype(np.float32).reshape(3, 4, 4)
    a = oracle.synth_bases(offsets, genome_of, trans, 7, n_frac=0.01, lower_frac=0.1)
    b = oracle.synth_bases(offsets, genome_of, trans, 7, n_frac=0.01, lower_frac=0.1, n_contigs=20)
    assert a.shape[0] == offsets[-1] and np.array_equal(a[:offsets[20]], b)
    assert set(np.unique(a).tolist()) <= set(b"ACGTNacgtn")
    assert (a == ord("N")).any() and (a >= ord("a")).any()
