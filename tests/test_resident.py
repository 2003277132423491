"""Round-2 entry points through the C ABI: the resident assembly graph (mcg_graph_load /
mcg_labels_set / mcg_bfs_resident), rules A / B for listed items (mcg_score_member_items), the
device pool (mcg_pool_*) and the 2-bit ("packed at parse time") input of the fused call.

Reference behaviour: label_prop_utils.py:24-100 (walks, rule B), :308-345 (rule C); the checker
is the pinned oracle / the Python restatement of the walk in tests/oracle_backend.py.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from metacoag_b200 import _lib, synth  # noqa: E402
from oracle import oracle  # noqa: E402
from oracle_backend import ref_bfs  # noqa: E402

MAX = sys.float_info.max


def test_parallel_host_packer_matches_the_serial_one():
    """mcg_pack_2bit_host_mt == mcg_pack_2bit_host (no device needed)."""
    rng = np.random.default_rng(5)
    lens = np.concatenate([rng.integers(0, 200, 300), rng.integers(60000, 900000, 30), [0, 1, 63, 64, 65]])
    rng.shuffle(lens)
    offsets = np.zeros(lens.shape[0] + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    blob = rng.choice(np.frombuffer(b"ACGTNacgt", dtype=np.uint8), int(offsets[-1]))
    want = _lib.pack_2bit_host(blob, offsets)
    for threads in (0, 1, 3, 16):
        assert np.array_equal(_lib.pack_2bit_host(blob, offsets, threads=threads), want)


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def _graph(seed, n=3000, labelled=400, n_bins=7):
    asm = synth.make_assembly(seed=seed, n_genomes=5, n_contigs=n, n_samples=1,
                              length_kind="mixed", with_markers=False)
    indptr, indices = synth.csr_from_edges(n, asm.edges)
    rng = np.random.default_rng(seed)
    labels = np.full(n, -1, dtype=np.int32)
    lab = rng.choice(n, labelled, replace=False)
    labels[lab] = rng.integers(0, n_bins, labelled)
    return indptr, indices, labels


def _check_walks(first, count, hits, starts, limit, labels, indptr, indices):
    assert hits.shape[0] == int(count.sum())
    for i, s in enumerate(starts):
        want = ref_bfs(int(s), limit, labels, indptr, indices)
        got = [tuple(r) for r in hits[first[i]:first[i] + count[i]].tolist()]
        assert got == want, (i, int(s))


@pytest.mark.gpu
@pytest.mark.parametrize("limit", [1, 3, 10])
def test_bfs_resident_matches_the_reference_walk(ctx, limit):
    indptr, indices, labels = _graph(31)
    ctx.graph_load(indptr, indices)
    ctx.labels_load(labels)
    starts = np.concatenate([np.flatnonzero(labels < 0)[:700], np.flatnonzero(labels >= 0)[:9]])
    first, count, hits = ctx.bfs_resident(starts, limit)
    _check_walks(first, count, hits, starts, limit, labels, indptr, indices)


@pytest.mark.gpu
def test_labels_set_patches_single_entries_and_batches(ctx):
    indptr, indices, labels = _graph(32)
    ctx.graph_load(indptr, indices)
    ctx.labels_load(labels)
    rng = np.random.default_rng(0)
    starts = np.flatnonzero(labels < 0)[:300]
    for step in range(6):
        free = np.flatnonzero(labels < 0)
        k = 1 if step < 4 else 40          # kernel-argument patch, then the bulk path
        ids = rng.choice(free, k, replace=False)
        bins = rng.integers(0, 7, k).astype(np.int32)
        labels[ids] = bins
        ctx.labels_set(ids, bins)
        first, count, hits = ctx.bfs_resident(starts, 4)
        _check_walks(first, count, hits, starts, 4, labels, indptr, indices)
    with pytest.raises(_lib.McgError):
        ctx.labels_set([labels.shape[0]], [0])


@pytest.mark.gpu
def test_bfs_resident_large_walks_take_the_global_table_pass(ctx):
    """Few labels and depth 10 on a random 4-regular-ish graph: walks touch thousands of
    vertices, overflow the 128-slot shared table and are repeated over global memory."""
    n = 20000
    rng = np.random.default_rng(7)
    edges = np.concatenate([np.stack([np.arange(n - 1), np.arange(1, n)], 1),
                            rng.integers(0, n, (n, 2))])
    indptr, indices = synth.csr_from_edges(n, edges)
    labels = np.full(n, -1, dtype=np.int32)
    labels[rng.choice(n, 40, replace=False)] = rng.integers(0, 3, 40)
    ctx.graph_load(indptr, indices)
    ctx.labels_load(labels)
    starts = rng.choice(np.flatnonzero(labels < 0), 64, replace=False)
    first, count, hits = ctx.bfs_resident(starts, 7)
    _check_walks(first, count, hits, starts, 7, labels, indptr, indices)
    # and the legacy dense call agrees
    node, bin_, depth, n_hits = ctx.bfs_candidates(indptr, indices, labels, starts, 7)
    for i in range(starts.shape[0]):
        dense = [(int(node[i, h]), int(bin_[i, h]), int(depth[i, h])) for h in range(n_hits[i])]
        assert dense == [tuple(r) for r in hits[first[i]:first[i] + count[i]].tolist()]


@pytest.mark.gpu
def test_bfs_resident_triangle_quirk_and_empty(ctx):
    indptr, indices = synth.csr_from_edges(3, np.array([[0, 1], [0, 2], [1, 2]]))
    ctx.graph_load(indptr, indices)
    ctx.labels_load(np.array([-1, -1, 0], dtype=np.int32))
    first, count, hits = ctx.bfs_resident([0], 10)
    # vertex 2 is queued from 0 (depth 1) and again from 1, which overwrites its depth with 2
    # (:93-98): both pops report depth 2 (the reference's set keeps one of the equal tuples)
    assert hits.tolist() == [[2, 0, 2], [2, 0, 2]]
    assert [tuple(r) for r in hits.tolist()] == ref_bfs(0, 10, np.array([-1, -1, 0]), indptr, indices)
    first, count, hits = ctx.bfs_resident([], 10)
    assert hits.shape == (0, 3) and count.shape == (0,)


def _features(seed, n, S, n_genomes=6):
    asm = synth.make_assembly(seed=seed, n_genomes=n_genomes, n_contigs=n, n_samples=S,
                              with_graph=False, with_markers=False)
    _, prof = oracle.count_kmers_batch(asm.blob, asm.offsets, nthreads=0)
    return asm, prof


@pytest.mark.gpu
@pytest.mark.parametrize("S", [1, 10, 30, 31, 50, 100])
@pytest.mark.parametrize("rule", [_lib.RULE_A, _lib.RULE_B])
def test_member_items_match_the_oracle(ctx, S, rule):
    asm, prof = _features(40 + S, 260, S)
    ctx.set_profiles(prof)
    ctx.load_coverage(asm.coverage)
    rng = np.random.default_rng(S)
    members, seg = [], [0]
    for g in range(6):
        ids = np.flatnonzero(asm.genome_of == g)[:0 if g == 2 else rng.integers(1, 6)]   # one empty bin
        members.extend(ids.tolist())
        seg.append(len(members))
    members, seg = np.array(members, dtype=np.int64), np.array(seg, dtype=np.int64)
    queries = rng.integers(0, 260, 500)
    cols = rng.integers(0, 6, 500).astype(np.int32)
    got = ctx.score_member_items(queries, cols, members, seg, rule)
    uniq = np.unique(queries)
    dense = oracle.score_members(prof, asm.coverage, uniq, members, seg, rule, nthreads=0)
    want = dense[np.searchsorted(uniq, queries), cols]
    # rule A over an empty bin is 0 / 0 (the reference raises ZeroDivisionError there,
    # matching_utils.py:152): NaN in the oracle, in the dense call and here
    nan = np.isnan(want)
    assert np.isnan(got[nan]).all() and not np.isnan(got[~nan]).any()
    assert nan.any() == (rule == _lib.RULE_A)
    top = (want == MAX) & ~nan
    fin = ~top & ~nan
    assert np.array_equal(got[top], want[top])
    assert np.all(np.abs(got[fin] - want[fin]) <= 1e-9 * np.abs(want[fin]))
    # and the dense call of the library itself (norm form of the distance): same to 1e-11
    lib_dense = ctx.score_members(uniq, members, seg, rule)[np.searchsorted(uniq, queries), cols]
    assert np.array_equal(lib_dense == MAX, top) and np.array_equal(np.isnan(lib_dense), nan)
    assert np.all(np.abs(got[fin] - lib_dense[fin]) <= 1e-11 * np.abs(want[fin]))


def _centroids(asm, prof, per_bin=4):
    members, seg = [], [0]
    for g in range(int(asm.genome_of.max()) + 1):
        members.extend(np.flatnonzero(asm.genome_of == g)[:per_bin].tolist())
        seg.append(len(members))
    return (oracle.segment_mean(prof, members, seg), oracle.segment_mean(asm.coverage, members, seg),
            np.array(members, dtype=np.int64), np.array(seg, dtype=np.int64))


@pytest.mark.gpu
def test_two_bit_input_of_the_fused_call(ctx):
    """mcg_featurise_score with MCG_ENC_2BIT input (packed at parse time), small (staged path)
    and above 64 Mbp (wave path): same results as the ASCII call, bit for bit."""
    for n, kind in ((400, "spades"), (9000, "spades")):
        asm = synth.make_assembly(seed=50 + n, n_genomes=10, n_contigs=n, n_samples=6,
                                  length_kind=kind, with_graph=False, with_markers=False)
        prof = ctx.tetramer_scan(asm.blob, asm.offsets)[1]
        cen_prof, cen_cov, _, _ = _centroids(asm, prof)
        arg, w = ctx.featurise_score(asm.blob, asm.offsets, asm.coverage, cen_prof, cen_cov)
        arg, w = arg.copy(), w.copy()
        packed = _lib.pack_2bit_host(asm.blob, asm.offsets, threads=0)
        arg2, w2 = ctx.featurise_score(packed, asm.offsets, asm.coverage, cen_prof, cen_cov,
                                       encoding=_lib.ENC_2BIT)
        assert np.array_equal(arg, arg2) and np.array_equal(w, w2)
        assert np.array_equal(ctx.get_profiles(), prof)
        if n == 9000:
            assert int(asm.offsets[-1]) >= 64 << 20      # the wave path was taken
            assert ctx.ingest_stats()["h2d_bytes"] == packed.shape[0]


def _pools():
    n = _lib.device_count()
    sizes = [1] + ([2] if n >= 2 else []) + ([n] if n > 2 else [])
    return sizes


@pytest.mark.gpu
def test_pool_results_do_not_depend_on_the_device_count():
    """Every pooled call against the single-context call, bit for bit: 1 device always, 2 and
    all devices when the box has them (skips nothing on a 1-GPU box but the N > 1 legs)."""
    asm = synth.make_assembly(seed=77, n_genomes=12, n_contigs=6000, n_samples=8,
                              length_kind="megahit", with_graph=False, with_markers=False)
    one = _lib.Context(0)
    counts1, prof1 = one.tetramer_scan(asm.blob, asm.offsets)
    cen_prof, cen_cov, members, seg = _centroids(asm, prof1)
    arg1, w1 = one.featurise_score(asm.blob, asm.offsets, asm.coverage, cen_prof, cen_cov)
    arg1, w1 = arg1.copy(), w1.copy()
    one.set_profiles(prof1)
    one.load_coverage(asm.coverage)
    rng = np.random.default_rng(1)
    queries = rng.permutation(6000)[:5000].astype(np.int64)
    qa1, qw1, _ = one.score_centroids(queries=queries, cen_prof=cen_prof, cen_cov=cen_cov)
    small = queries[:300]
    sa1, sw1, sW1 = one.score_centroids(queries=small, cen_prof=cen_prof, cen_cov=cen_cov,
                                        want_weights=True)
    mem1 = one.score_members(small, members, seg, _lib.RULE_A)
    items_q = rng.integers(0, 6000, 9000)
    items_b = rng.integers(0, 12, 9000).astype(np.int32)
    it1 = one.score_member_items(items_q, items_b, members, seg, _lib.RULE_B)
    bp1 = one.bin_profiles(members, seg)
    indptr, indices, labels = _graph(33, n=6000, labelled=900)
    one.graph_load(indptr, indices)
    one.labels_load(labels)
    starts = np.flatnonzero(labels < 0)[:5000]
    f1, c1, h1 = one.bfs_resident(starts, 5)
    one.close()
    packed = _lib.pack_2bit_host(asm.blob, asm.offsets, threads=0)
    for size in _pools():
        pool = _lib.Pool(list(range(size)))
        try:
            assert pool.size == size
            counts, prof = pool.tetramer_scan(asm.blob, asm.offsets)
            assert np.array_equal(counts, counts1) and np.array_equal(prof, prof1)
            b = pool.bounds()
            assert b[0] == 0 and b[-1] == 6000 and np.all(np.diff(b) >= 0)
            for enc, bases in ((_lib.ENC_ASCII, asm.blob), (_lib.ENC_2BIT, packed)):
                arg, w = pool.featurise_score(bases, asm.offsets, asm.coverage, cen_prof, cen_cov,
                                              encoding=enc)
                assert np.array_equal(arg, arg1) and np.array_equal(w, w1), (size, enc)
            pool.set_features(prof=prof1, cov=asm.coverage)
            qa, qw, _ = pool.score_centroids(queries=queries, cen_prof=cen_prof, cen_cov=cen_cov)
            assert np.array_equal(qa, qa1) and np.array_equal(qw, qw1)
            sa, sw, sW = pool.score_centroids(queries=small, cen_prof=cen_prof, cen_cov=cen_cov,
                                              want_weights=True)
            assert np.array_equal(sa, sa1) and np.array_equal(sw, sw1) and np.array_equal(sW, sW1)
            assert np.array_equal(pool.score_members(small, members, seg, _lib.RULE_A), mem1)
            assert np.array_equal(
                pool.score_member_items(items_q, items_b, members, seg, _lib.RULE_B), it1)
            bp = pool.bin_profiles(members, seg)
            assert np.array_equal(bp[0], bp1[0]) and np.array_equal(bp[1], bp1[1])
            pool.graph_load(indptr, indices)
            pool.labels_load(labels)
            f, c, h = pool.bfs_resident(starts, 5)
            assert np.array_equal(c, c1) and h.shape == h1.shape
            # (where a walk's hits sit in the flat list depends on which walk ended first)
            for i in range(starts.shape[0]):
                assert np.array_equal(h[f[i]:f[i] + c[i]], h1[f1[i]:f1[i] + c1[i]])
            lab2 = labels.copy()
            ids = np.flatnonzero(labels < 0)[-3:]
            lab2[ids] = [1, 2, 3]
            pool.labels_set(ids, np.array([1, 2, 3], dtype=np.int32))
            f, c, h = pool.bfs_resident(starts, 5)
            _check_walks(f, c, h, starts[:400], 5, lab2, indptr, indices)
        finally:
            pool.close()


@pytest.mark.gpu
def test_bfs_resident_against_the_reference_golden(ctx, golden_dir):
    """mcg_bfs_resident against run_bfs_long of the live reference (tests/golden/bfs.json): the
    reference returns a set, so the triples of a start node are compared as sets."""
    import json
    cases = json.load(open(os.path.join(golden_dir, "bfs.json")))
    for name, case in cases.items():
        indptr, indices = synth.csr_from_edges(case["n"], np.array(case["edges"]))
        ctx.graph_load(indptr, indices)
        ctx.labels_load(np.array(case["labels"], dtype=np.int32))
        for limit, per_start in case["expect"].items():
            first, count, hits = ctx.bfs_resident(case["starts"], int(limit))
            for i, want in enumerate(per_start):
                got = sorted(set(tuple(r) for r in hits[first[i]:first[i] + count[i]].tolist()))
                assert got == [tuple(t) for t in want], (name, limit, case["starts"][i])


@pytest.mark.gpu
def test_oracle_generator_writes_the_librarys_bytes(ctx):
    """bench.py --impl reference draws its bases with oracle.synth_bases (no CUDA library in
    that process): same bytes as mcg_synth_bases, also for a prefix of the contigs."""
    rng = np.random.default_rng(3)
    lengths = rng.integers(200, 40000, 400)
    offsets = np.zeros(401, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    genome_of = rng.integers(0, 5, 400).astype(np.int32)
    trans = rng.dirichlet(np.ones(4), size=20).astype(np.float32).reshape(5, 4, 4)
    want = ctx.synth_bases(offsets, genome_of, trans, 1001)
    assert np.array_equal(oracle.synth_bases(offsets, genome_of, trans, 1001), want)
    assert np.array_equal(oracle.synth_bases(offsets, genome_of, trans, 1001, n_contigs=37),
                          want[:offsets[37]])
