"""GPU parity: the CUDA path (through the C ABI) against the golden vectors taken from the
live reference and against the pinned CPU oracle on seeded synthetic inputs.

Bars (BASELINE.json north_star): tetramer counts, frequencies, coverage clamp and bin
profiles bit-exact; probabilities / weights within 1e-9 relative; argmins identical.
"""
import math
import os
import sys

import numpy as np
import pytest

from metacoag_b200 import _lib, synth
from oracle import oracle

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

MAXW = sys.float_info.max
RTOL = 1e-9   # north_star: probabilities within 1e-9 relative


@pytest.fixture(scope="module")
def ctx():
    return _lib.Context(0)


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


def pack_2bit(blob, offsets):
    """Host-side MCG_ENC_2BIT packing as include/mcg_b200.h defines it."""
    lut = np.zeros(256, dtype=np.uint32)
    lut[ord("C")], lut[ord("G")], lut[ord("T")] = 1, 2, 3
    lens = np.diff(offsets)
    words_per = 4 * ((lens + 63) // 64)
    word_off = np.concatenate([[0], np.cumsum(words_per)])
    out = np.zeros(int(word_off[-1]), dtype=np.uint32)
    for c in range(len(lens)):
        codes = lut[blob[offsets[c]:offsets[c + 1]]]
        pad = np.zeros(int(words_per[c]) * 16, dtype=np.uint32)
        pad[:codes.shape[0]] = codes
        shifts = (2 * (np.arange(pad.shape[0]) % 16)).astype(np.uint32)
        out[word_off[c]:word_off[c + 1]] = (pad << shifts).reshape(-1, 16).sum(
            axis=1, dtype=np.uint64).astype(np.uint32)
    return out.view(np.uint8)


def assert_close(got, want, rtol=RTOL):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    special = (want == MAXW) | ~np.isfinite(want) | (want == 0.0)
    assert np.array_equal(got[special], want[special])
    # denormal results carry the reference's own rounding of a denormal intermediate;
    # they are compared through their logarithm (what the weights use)
    tiny = ~special & (np.abs(want) < 1e-290)
    if tiny.any():
        assert np.all(got[tiny] > 0)
        assert np.allclose(np.log(got[tiny]), np.log(want[tiny]), rtol=1e-6, atol=0)
    rest = ~special & ~tiny
    err = np.abs(got[rest] - want[rest]) / np.abs(want[rest])
    assert err.size == 0 or err.max() <= rtol, "max rel err %.3e" % err.max()


# ---- scan -------------------------------------------------------------------------------
def test_scan_golden_ascii(ctx, scan):
    blob, offsets = _stripped(scan)
    counts, freqs = ctx.tetramer_scan(blob, offsets)
    assert np.array_equal(counts.astype(np.float64), scan["counts"])
    assert np.array_equal(freqs, scan["freqs"])


def test_scan_golden_2bit(ctx, scan):
    blob, offsets = _stripped(scan)
    packed = pack_2bit(blob, offsets)
    counts, freqs = ctx.tetramer_scan(packed, offsets, encoding=_lib.ENC_2BIT)
    assert np.array_equal(counts.astype(np.float64), scan["counts"])
    assert np.array_equal(freqs, scan["freqs"])


@pytest.mark.parametrize("kind,n", [("mixed", 700), ("spades", 300), ("megahit", 1500)])
def test_scan_vs_oracle(ctx, kind, n):
    asm = synth.make_assembly(seed=11, n_genomes=6, n_contigs=n, n_samples=2, length_kind=kind,
                              n_frac=2e-3, lower_frac=0.05, with_graph=False, with_markers=False)
    want_counts, want_freqs = oracle.count_kmers_batch(asm.blob, asm.offsets, nthreads=0)
    counts, freqs = ctx.tetramer_scan(asm.blob, asm.offsets)
    assert np.array_equal(counts.astype(np.float64), want_counts)
    assert np.array_equal(freqs, want_freqs)
    assert np.array_equal(counts.sum(axis=1), np.maximum(0, asm.lengths - 3))


def test_scan_edge_lengths(ctx):
    """Empty and ragged inputs, lengths around every tiling boundary, one long contig."""
    rng = np.random.default_rng(3)
    lens = [0, 1, 3, 4, 5, 15, 16, 17, 18, 19, 20, 63, 64, 65, 66, 67, 255, 256, 257, 258, 259,
            260, 261, 511, 515, 8191, 8192, 8195, 8196, 0, 70000, 2, 300001, 4]
    offsets = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    blob = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(offsets[-1]))].copy()
    blob[rng.random(blob.shape[0]) < 0.01] = ord("N")
    want_counts, want_freqs = oracle.count_kmers_batch(blob, offsets, nthreads=0)
    counts, freqs = ctx.tetramer_scan(blob, offsets)
    assert np.array_equal(counts.astype(np.float64), want_counts)
    assert np.array_equal(freqs, want_freqs)


def test_scan_counter_saturation(ctx):
    """Homopolymer and N runs put every window of a segment on one counter: the byte
    counters must not wrap (N counts as A, feature_utils.py:21)."""
    seqs = [b"A" * 5000, b"N" * 777, b"T" * 259, b"ACGT" + b"N" * 1024 + b"ACGT", b"G" * 256 * 33,
            b"C" * 258 + b"A"]
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offsets[1:])
    blob = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    want_counts, want_freqs = oracle.count_kmers_batch(blob, offsets)
    counts, freqs = ctx.tetramer_scan(blob, offsets)
    assert np.array_equal(counts.astype(np.float64), want_counts)
    assert np.array_equal(freqs, want_freqs)


def test_scan_properties_large(ctx):
    """Size-independent properties at a size the oracle would take long on: every row sums
    to len-3, and the profile of a contig does not depend on its neighbours in the batch."""
    rng = np.random.default_rng(5)
    lens = np.exp(rng.uniform(np.log(1000), np.log(50000), 4000)).astype(np.int64)
    offsets = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    blob = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(offsets[-1]))]
    counts, freqs = ctx.tetramer_scan(blob, offsets)
    assert np.array_equal(counts.sum(axis=1, dtype=np.int64), lens - 3)
    assert np.allclose(freqs.sum(axis=1), 1.0, rtol=0, atol=1e-12)
    pick = rng.choice(len(lens), 64, replace=False)
    sub_off = np.zeros(65, dtype=np.int64)
    np.cumsum(lens[pick], out=sub_off[1:])
    sub_blob = np.concatenate([blob[offsets[i]:offsets[i + 1]] for i in pick])
    sub_counts, _ = ctx.tetramer_scan(sub_blob, sub_off)
    assert np.array_equal(sub_counts, counts[pick])
    want, _ = oracle.count_kmers_batch(sub_blob, sub_off, nthreads=0)
    assert np.array_equal(sub_counts.astype(np.float64), want)


# ---- coverage features and bin profiles ------------------------------------------------------
def test_coverage_terms(ctx, score):
    raw = score["cov"].copy()
    raw[0, 0] = 0.0          # exercises the clamp of feature_utils.py:152-153
    raw[1, 1] = 5e-5
    ctx.load_coverage(raw)
    cov, logc, lgam = ctx.coverage_terms()
    want = raw.copy()
    want[want < 1e-4] = 1e-4
    assert np.array_equal(cov, want)                      # coverage features: bit-exact
    want_log, want_lg = oracle.cov_terms(want)
    assert_close(logc, want_log, rtol=1e-15)
    # lgamma(c+1) is ~0 near c = 0 and c = 1: absolute tolerance there
    assert np.allclose(lgam, want_lg, rtol=1e-14, atol=1e-15)


def test_lgamma_is_cpythons(ctx):
    rng = np.random.default_rng(1)
    c = np.concatenate([rng.uniform(1e-4, 6, 3000), np.exp(rng.uniform(0, 11, 3000)),
                        np.arange(0, 40, dtype=float)]).reshape(-1, 1)
    c[c < 1e-4] = 1e-4
    ctx.load_coverage(c)
    _, _, lgam = ctx.coverage_terms()
    want = np.array([math.lgamma(x + 1.0) for x in c[:, 0]])
    assert np.allclose(lgam[:, 0], want, rtol=5e-15, atol=5e-15)


def test_bin_profiles_bit_exact(ctx, score):
    ctx.set_profiles(score["prof"])
    ctx.load_coverage(score["cov"])
    cen_prof, cen_cov = ctx.bin_profiles(score["members"], score["seg_offsets"])
    assert np.array_equal(cen_prof, score["cen_prof"])
    assert np.array_equal(cen_cov, score["cen_cov"])


# ---- scoring primitives -----------------------------------------------------------------------
def test_gaussians(ctx, score):
    d = score["dist"]
    assert_close(ctx.normpdf(d, 0.0, 0.01037897 / 2), score["pdf_intra"])
    assert_close(ctx.normpdf(d, 0.0676654, 0.03419337), score["pdf_inter"])
    assert_close(ctx.comp_probability(d), score["comp_prob"])
    with pytest.raises(ZeroDivisionError):    # matching_utils.py:39 at d >~ 1.39
        ctx.comp_probability(np.array([0.1, 1.40]))


@pytest.mark.parametrize("S", [1, 3, 10, 33, 50, 100])
def test_cov_probability(ctx, score, S):
    a, b, want = score["covA_%d" % S], score["covB_%d" % S], score["covprob_%d" % S]
    assert_close(ctx.cov_probability(a, b), want)


def test_pair_matrix(ctx, score):
    prof, cov = score["prof"], score["cov"]
    n = prof.shape[0]
    ctx.set_profiles(prof)
    ctx.load_coverage(cov)
    ids = np.arange(n)
    got = ctx.pair_scores(ids, ids)
    assert_close(got["dist"], score["pair_dist"], rtol=1e-12)
    assert_close(got["pc"], score["pair_pc"])
    assert_close(got["pv"], score["pair_pv"])
    assert_close(got["lp"], score["pair_lp"])
    u = prof[np.arange(n)]
    v = prof[(np.arange(n) * 7 + 3) % n]
    want = np.array([oracle.euclidean(x, y) for x, y in zip(u, v)])
    assert_close(ctx.tetramer_distance(u, v), want, rtol=1e-12)


def test_rule_c_golden(ctx, score):
    ctx.set_profiles(score["prof"])
    ctx.load_coverage(score["cov"])
    arg, w, weights = ctx.score_centroids(cen_prof=score["cen_prof"], cen_cov=score["cen_cov"],
                                          want_weights=True)
    assert np.array_equal(arg, score["ruleC_arg"])
    assert_close(w, score["ruleC_w"])
    _, _, want = oracle.score_centroids(score["prof"], score["cov"], score["cen_prof"],
                                        score["cen_cov"], want_weights=True)
    assert_close(weights, want)
    # resident centroids from mcg_bin_profiles give the same answer
    ctx.bin_profiles(score["members"], score["seg_offsets"])
    arg2, w2, _ = ctx.score_centroids()
    assert np.array_equal(arg2, arg) and np.array_equal(w2, w)
    # a subset of queries
    q = np.array([5, 0, 17, 17, 40])
    arg3, w3, _ = ctx.score_centroids(queries=q)
    assert np.array_equal(arg3, arg[q]) and np.array_equal(w3, w[q])


def test_rules_a_b_golden(ctx, score):
    ctx.set_profiles(score["prof"])
    ctx.load_coverage(score["cov"])
    for rule, key in ((_lib.RULE_A, "ruleA_w"), (_lib.RULE_B, "ruleB_w")):
        want = score[key]
        queries = np.flatnonzero(~np.isnan(want[:, 0]))
        got = ctx.score_members(queries, score["members"], score["seg_offsets"], rule)
        assert_close(got, want[queries])


@pytest.mark.parametrize("S,B,n", [(1, 7, 300), (10, 200, 1000), (30, 40, 300), (31, 40, 300),
                                   (50, 70, 333), (100, 33, 150),
                                   # >= 2048 queries take the two-kernel rule-C path; 300 bins
                                   # need two shared-memory groups there
                                   (3, 40, 2100), (10, 300, 2500), (30, 70, 2200)])
def test_rule_c_vs_oracle(ctx, S, B, n):
    """Seeded synthetic contig tables at sizes the oracle finishes in seconds, including
    S > 32 where floor^S underflows and the MAX_WEIGHT / None paths trigger."""
    asm = synth.make_assembly(seed=100 + S, n_genomes=B, n_contigs=n, n_samples=S,
                              length_kind="mixed", with_graph=False, with_markers=False)
    _, prof = oracle.count_kmers_batch(asm.blob, asm.offsets, nthreads=0)
    cov = asm.coverage
    members, seg = [], [0]
    for g in range(B):
        ids = np.flatnonzero(asm.genome_of == g)[:6]
        if ids.size == 0:
            ids = np.array([g % n])
        members.extend(ids.tolist())
        seg.append(len(members))
    cen_prof = oracle.segment_mean(prof, members, seg)
    cen_cov = oracle.segment_mean(cov, members, seg)
    want_arg, want_w, want_weights = oracle.score_centroids(prof, cov, cen_prof, cen_cov,
                                                            want_weights=True, nthreads=0)
    ctx.set_profiles(prof)
    ctx.load_coverage(cov)
    arg, w, weights = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov, want_weights=True)
    assert_close(weights, want_weights)
    assert np.array_equal(arg, want_arg)
    assert_close(w, want_w)
    for rule in (_lib.RULE_A, _lib.RULE_B):
        q = np.arange(0, n, 3)
        want = oracle.score_members(prof, cov, q, members, seg, rule, nthreads=0)
        assert_close(ctx.score_members(q, members, seg, rule), want)


@pytest.mark.parametrize("S,B,n,orphans", [(33, 70, 2100, False), (50, 300, 2300, False),
                                           (50, 300, 2300, True), (100, 40, 2100, True),
                                           (64, 520, 2050, True)])
def test_rule_c_guarded_path_vs_oracle(ctx, S, B, n, orphans):
    """S > 30 on the two-kernel path: log-domain scoring with deferred pairs, and the redo of
    the contigs that have no weight below 295 (label_prop_utils.py:308-345 must come out the
    same: first argmin, None, weights).  `orphans` gives every third contig coverages that
    match no bin, which forces the redo path."""
    asm = synth.make_assembly(seed=300 + S, n_genomes=B, n_contigs=n, n_samples=S,
                              length_kind="mixed", with_graph=False, with_markers=False)
    _, prof = oracle.count_kmers_batch(asm.blob, asm.offsets, nthreads=0)
    cov = asm.coverage.copy()
    if orphans:
        rng = np.random.default_rng(S)
        rows = np.arange(0, n, 3)
        cov[rows] = np.maximum(np.round(rng.lognormal(3.0, 1.5, (rows.size, S)), 4), 1e-4)
        cov[rows[::4], : S // 2] = 1e-4      # long runs of clamped zeros: terms ~ 1, "sticking"
    members, seg = [], [0]
    for g in range(B):
        ids = np.flatnonzero(asm.genome_of == g)[:6]
        if ids.size == 0:
            ids = np.array([g % n])
        members.extend(ids.tolist())
        seg.append(len(members))
    cen_prof = oracle.segment_mean(prof, members, seg)
    cen_cov = oracle.segment_mean(cov, members, seg)
    want_arg, want_w, _ = oracle.score_centroids(prof, cov, cen_prof, cen_cov, nthreads=0)
    ctx.set_profiles(prof)
    ctx.load_coverage(cov)
    arg, w = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov)[:2]
    redo = ctx.last_redo_count()
    assert np.array_equal(arg, want_arg)
    assert_close(w, want_w)
    assert 0 <= redo <= n
    if orphans:
        assert redo > 0          # the in-order path ran ...
    assert redo < n              # ... and the log-domain path decided the rest


def test_fused_call_matches_staged(ctx):
    asm = synth.make_assembly(seed=9, n_genomes=12, n_contigs=500, n_samples=5,
                              with_graph=False, with_markers=False)
    _, prof = oracle.count_kmers_batch(asm.blob, asm.offsets, nthreads=0)
    members, seg = [], [0]
    for g in range(12):
        members.extend(np.flatnonzero(asm.genome_of == g)[:4].tolist())
        seg.append(len(members))
    cen_prof = oracle.segment_mean(prof, members, seg)
    cen_cov = oracle.segment_mean(asm.coverage, members, seg)
    want_arg, want_w, _ = oracle.score_centroids(prof, asm.coverage, cen_prof, cen_cov,
                                                 nthreads=0)
    arg, w = ctx.featurise_score(asm.blob, asm.offsets, asm.coverage, cen_prof, cen_cov)
    assert np.array_equal(arg, want_arg)
    assert_close(w, want_w)
    assert np.array_equal(ctx.get_profiles(), prof)
    # the resident step recomputes the same thing from the packed store
    ctx.step_resident()
    arg2, w2 = ctx.get_assignments()
    assert np.array_equal(arg2, arg) and np.array_equal(w2, w)


# ---- frontier ----------------------------------------------------------------------------------
def _ref_bfs(start, limit, labels, indptr, indices):
    """label_prop_utils.py:24-100 restated on CSR (traversal only), pure Python."""
    queue, visited, depth, hits = [start], set(), {start: 0}, []
    while queue:
        v = queue.pop(0)
        visited.add(v)
        if labels[v] >= 0 and len(visited) > 1:
            hits.append((v, int(labels[v]), depth[v]))
        else:
            for nb in indices[indptr[v]:indptr[v + 1]]:
                nb = int(nb)
                if nb not in visited:
                    depth[nb] = depth[v] + 1
                    if depth[nb] > limit:
                        continue
                    queue.append(nb)
    return hits


@pytest.mark.parametrize("limit", [1, 3, 10])
def test_bfs_candidates(ctx, limit):
    asm = synth.make_assembly(seed=21, n_genomes=5, n_contigs=400, n_samples=1,
                              length_kind="mixed", with_markers=False)
    n = asm.n_contigs
    indptr, indices = synth.csr_from_edges(n, asm.edges)
    rng = np.random.default_rng(2)
    labels = np.full(n, -1, dtype=np.int32)
    lab = rng.choice(n, 60, replace=False)
    labels[lab] = rng.integers(0, 5, 60)
    starts = np.concatenate([np.flatnonzero(labels < 0)[:150], lab[:5]])
    node, bin_, depth, n_hits = ctx.bfs_candidates(indptr, indices, labels, starts, limit)
    for i, s in enumerate(starts):
        want = _ref_bfs(int(s), limit, labels, indptr, indices)
        got = [(int(node[i, h]), int(bin_[i, h]), int(depth[i, h])) for h in range(n_hits[i])]
        assert got == want, (i, s)


def test_bfs_depth_overwrite_quirk(ctx):
    # triangle 0-1-2 with 2 labelled: SURVEY.md section 7 hard part 4
    indptr, indices = synth.csr_from_edges(3, np.array([[0, 1], [0, 2], [1, 2]]))
    labels = np.array([-1, -1, 0], dtype=np.int32)
    node, bin_, depth, n_hits = ctx.bfs_candidates(indptr, indices, labels, [0], 10)
    got = [(int(node[0, h]), int(depth[0, h])) for h in range(n_hits[0])]
    assert got == [(v, d) for v, _, d in _ref_bfs(0, 10, labels, indptr, indices)]
    assert got[0] == (2, 2)      # distance-1 node reported at depth 2


# ---- shortest-path lengths (graph gate of match_contigs) -------------------------------------
def _csr(n, edges):
    adj = [set() for _ in range(n)]
    for a, b in edges:
        if a != b:
            adj[a].add(b)
            adj[b].add(a)
    indptr = np.zeros(n + 1, dtype=np.int64)
    for v in range(n):
        indptr[v + 1] = indptr[v] + len(adj[v])
    indices = np.array([u for v in range(n) for u in sorted(adj[v])], dtype=np.int32)
    return indptr, indices


@pytest.mark.parametrize("kind", ["sparse", "chain", "star"])
def test_path_lengths_vs_bfs(ctx, kind):
    """matching_utils.py:191-197: len(get_shortest_paths(v, to=w)[0]) for all pairs -- more than
    32 sources (two passes), repeated sources and targets, unreachable components, v == w."""
    from oracle_backend import OracleBackend
    rng = np.random.default_rng({"sparse": 1, "chain": 2, "star": 3}[kind])
    if kind == "sparse":
        n = 3000
        edges = [(i, i + 1) for i in range(0, 2000)] + \
            [tuple(rng.integers(0, 2400, 2)) for _ in range(700)]       # 2400.. are isolated
    elif kind == "chain":
        n = 20000
        edges = [(i, i + 1) for i in range(n - 1)]                      # 20 000 levels
    else:
        n = 5000
        edges = [(0, i) for i in range(1, 4000)] + [(i, i + 1) for i in range(4000, 4990)]
    indptr, indices = _csr(n, edges)
    sources = rng.integers(0, n, 70)
    sources[5] = sources[6]
    targets = np.concatenate([rng.integers(0, n, 300), sources[:10], [0, n - 1]])
    targets[3] = targets[4]
    got = ctx.path_lengths(indptr, indices, sources, targets)
    want = OracleBackend().path_lengths(indptr, indices, sources, targets)
    assert np.array_equal(got, want)
    assert (got == 0).any() or kind == "chain"
