#!/usr/bin/env python3
"""Measurements of the rows next to the hot path: the graph gate of
match_contigs, the bin x bin merge matrix, the FASTA reader.  Prints one JSON object.

    python tests/bench_aux.py            (needs a B200; the CPU side is timed in the same run)

Lives under tests/ because it checks every GPU number against the CPU oracle (test
infrastructure, only tests may use it); pytest does not collect it.

CPU comparisons restate what the reference does per call: one BFS per (source, target) pair
(igraph get_shortest_paths, here a plain-Python BFS over the same CSR -- igraph itself is not
installed), one get_*_probability triple per bin pair (oracle port), a Python FASTA parser.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from metacoag_b200 import _lib, feature_utils, synth  # noqa: E402


def graph(n, seed):
    """Per-genome chains + chords + a few cross edges, as CSR."""
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 200, n)
    order = np.argsort(genome, kind="stable")
    a = order[:-1]
    b = order[1:]
    same = genome[a] == genome[b]
    edges = [np.stack([a[same], b[same]], 1)]
    chords = rng.integers(0, n, (n // 4, 2))
    chords = chords[genome[chords[:, 0]] == genome[chords[:, 1]]]
    cross = rng.integers(0, n, (n // 50, 2))
    e = np.concatenate(edges + [chords, cross])
    e = e[e[:, 0] != e[:, 1]]
    e = np.unique(np.concatenate([e, e[:, ::-1]]), axis=0)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, e[:, 0] + 1, 1)
    np.cumsum(indptr, out=indptr)
    return indptr, e[:, 1].astype(np.int32)


def bfs_pair(indptr, indices, s, t):
    """One (source, target) query the way the reference asks igraph: BFS until t is found."""
    if s == t:
        return 1
    dist = {s: 1}
    frontier = [s]
    while frontier:
        nxt = []
        for v in frontier:
            for u in indices[indptr[v]:indptr[v + 1]]:
                u = int(u)
                if u not in dist:
                    dist[u] = dist[v] + 1
                    if u == t:
                        return dist[u]
                    nxt.append(u)
        frontier = nxt
    return 0


def main():
    out = {}
    ctx = _lib.Context(0)

    # ---- graph gate ------------------------------------------------------------------------
    n = 100000
    indptr, indices = graph(n, 5)
    rng = np.random.default_rng(6)
    sources = rng.choice(n, 200, replace=False)
    targets = rng.choice(n, 2000, replace=False)
    ctx.path_lengths(indptr, indices, sources[:4], targets[:4])
    gpu_s = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        got = ctx.path_lengths(indptr, indices, sources, targets)
        gpu_s = min(gpu_s, time.perf_counter() - t0)
    sample = [(int(rng.integers(0, 200)), int(rng.integers(0, 2000))) for _ in range(60)]
    t0 = time.perf_counter()
    for i, j in sample:
        assert bfs_pair(indptr, indices, int(sources[i]), int(targets[j])) == got[i, j]
    cpu_s = (time.perf_counter() - t0) / len(sample)
    out["path_lengths"] = {"vertices": n, "edges": int(indices.shape[0]) // 2, "sources": 200,
                           "targets": 2000, "gpu_ms_total_incl_copies": gpu_s * 1e3,
                           "gpu_us_per_pair": gpu_s * 1e6 / 4e5,
                           "cpu_ms_per_pair_python_bfs": cpu_s * 1e3,
                           "levels": int(got.max()), "unreachable_pairs": int((got == 0).sum())}

    # ---- connected components --------------------------------------------------------------
    ctx.connected_components(indptr, indices)
    t0 = time.perf_counter()
    labels = ctx.connected_components(indptr, indices)
    gpu_s = time.perf_counter() - t0
    from oracle_backend import OracleBackend
    t0 = time.perf_counter()
    want = OracleBackend().connected_components(indptr, indices)
    cpu_s = time.perf_counter() - t0
    assert np.array_equal(labels, want)
    # the reference's own sweep (graph_utils.py:407-434), restated, for ONE start vertex of
    # the largest component; it runs it for every binned contig
    big = int(np.bincount(labels).argmax())
    t0 = time.perf_counter()
    comp = [big]
    length = 0
    while length != len(comp) and len(comp) < 4000:      # capped: the sweep is quadratic
        length = len(comp)
        for j in list(comp):
            for u in indices[indptr[j]:indptr[j + 1]]:
                if int(u) not in comp:
                    comp.append(int(u))
    ref_s = time.perf_counter() - t0
    out["connected_components"] = {"vertices": n, "components": int(np.unique(labels).shape[0]),
                                   "largest": int(np.bincount(labels).max()),
                                   "gpu_ms_total_incl_copies": gpu_s * 1e3,
                                   "cpu_ms_python_dfs_labelling": cpu_s * 1e3,
                                   "reference_style_sweep_s_for_one_start_capped_at_4000_vertices": ref_s}

    # ---- label propagation candidates: bounded BFS + rule B ----------
    # one stage of label_prop: every unlabelled long contig walks to the labelled contigs within
    # `depth` (label_prop_utils.py:24-53, 92-100) and is scored against the seed members of every
    # bin it reached (:54-86).  CPU side: the reference's BFS restated in Python (ref_bfs) and the
    # C port of rule B, both on a sample.
    from oracle_backend import ref_bfs
    from oracle import oracle
    B, S = 200, 10
    import bench
    bench.LENGTH_KIND = "megahit"
    tabs = bench.workload_tables(11, n, S, B)
    lp_blob = ctx.synth_bases(tabs["offsets"], tabs["genome_of"], tabs["trans"], 11)
    rngl = np.random.default_rng(12)
    genome = np.asarray(tabs["genome_of"])
    labels = np.where(rngl.random(n) < 0.3, genome, -1).astype(np.int32)
    starts = np.flatnonzero(labels < 0)[:40000].astype(np.int64)
    ctx.bfs_candidates(indptr, indices, labels, starts[:64], 10)
    t0 = time.perf_counter()
    node, bin_, depth, n_hits = ctx.bfs_candidates(indptr, indices, labels, starts, 10)
    bfs_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    k = 0
    for i in range(0, starts.shape[0], starts.shape[0] // 200):
        want = ref_bfs(int(starts[i]), 10, labels, indptr, indices)
        got = [(int(node[i, h]), int(bin_[i, h]), int(depth[i, h])) for h in range(int(n_hits[i]))]
        assert got == want, (i, got[:3], want[:3])
        k += 1
    bfs_cpu = (time.perf_counter() - t0) / k
    # rule B of every start node against 20 seed members per bin
    ctx.load_contigs(lp_blob, tabs["offsets"])
    ctx.scan()
    cov = np.maximum(tabs["coverage"], 1e-4)
    ctx.load_coverage(cov)
    order = np.argsort(genome, kind="stable")
    first = np.searchsorted(genome[order], np.arange(B + 1))
    members = np.concatenate([order[first[b]:first[b + 1]][:20] for b in range(B)]).astype(np.int64)
    seg = np.zeros(B + 1, dtype=np.int64)
    np.cumsum([min(20, first[b + 1] - first[b]) for b in range(B)], out=seg[1:])
    keep = np.flatnonzero(np.diff(seg) > 0)
    seg = np.concatenate([[0], seg[1:][keep]]).astype(np.int64)
    ctx.score_members(starts, members, seg, _lib.RULE_B)      # sizes the device buffers
    t0 = time.perf_counter()
    wB = ctx.score_members(starts, members, seg, _lib.RULE_B)
    ruleb_s = time.perf_counter() - t0
    prof_h = ctx.get_profiles()
    pick = starts[:: starts.shape[0] // 40]
    t0 = time.perf_counter()
    want = oracle.score_members(prof_h, cov, pick, members, seg, _lib.RULE_B, nthreads=1)
    ruleb_cpu = (time.perf_counter() - t0) / (pick.shape[0] * members.shape[0])
    got = wB[:: starts.shape[0] // 40]
    same = (want == _lib.MAX_WEIGHT) | (got == _lib.MAX_WEIGHT)
    assert np.array_equal(want[same], got[same])
    assert (np.abs(got[~same] - want[~same]) / np.abs(want[~same])).max() <= 1e-9
    out["label_prop_candidates"] = {
        "vertices": n, "labelled": int((labels >= 0).sum()), "start_nodes": int(starts.shape[0]),
        "depth": 10, "hits": int(n_hits.sum()),
        "bfs_gpu_ms_total_incl_copies": bfs_s * 1e3, "bfs_gpu_us_per_start": bfs_s * 1e6 / starts.shape[0],
        "bfs_cpu_us_per_start_python": bfs_cpu * 1e6,
        "rule_b_pairs": int(starts.shape[0] * members.shape[0]),
        "rule_b_gpu_ms_total_incl_copies": ruleb_s * 1e3,
        "rule_b_gpu_ns_per_pair": ruleb_s * 1e9 / (starts.shape[0] * members.shape[0]),
        "rule_b_cpu_ns_per_pair_c_port_1_thread": ruleb_cpu * 1e9}

    # ---- bin x bin merge matrix --------------------------------------------------------------------
    for B, S in ((200, 10), (1000, 100)):
        asm = synth.make_assembly(seed=B, n_genomes=B, n_contigs=B, n_samples=S,
                                  length_kind="spades", with_graph=False, with_markers=False)
        _, prof = oracle.count_kmers_batch(asm.blob, asm.offsets, nthreads=0)
        cov = np.maximum(asm.coverage, 1e-4)
        ctx.set_profiles(prof)
        ctx.load_coverage(cov)
        ids = np.arange(B, dtype=np.int64)
        ctx.pair_scores(ids[:8], ids[:8], want=("lp",))
        t0 = time.perf_counter()
        lp = ctx.pair_scores(ids, ids, want=("pc", "lp"))["lp"]
        gpu_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        k = 0
        for i in range(0, B, max(1, B // 20)):
            for j in range(0, B, max(1, B // 20)):
                want = oracle.pair_lp(prof[i], prof[j], cov[i], cov[j])[1]
                assert want == lp[i, j] or abs(want - lp[i, j]) <= 1e-9 * abs(want)
                k += 1
        cpu_s = (time.perf_counter() - t0) / k
        out["merge_matrix_B%d_S%d" % (B, S)] = {
            "pairs": B * B, "gpu_ms_total_incl_copies": gpu_s * 1e3,
            "gpu_ns_per_pair": gpu_s * 1e9 / (B * B), "cpu_us_per_pair_c_port": cpu_s * 1e6}

    # ---- FASTA reader ---------------------------------------------------------------------------------
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "contigs.fasta")
        asm = synth.make_assembly(seed=9, n_genomes=50, n_contigs=20000, n_samples=2,
                                  length_kind="spades", with_graph=False, with_markers=False)
        with open(path, "wb") as fh:
            for i in range(asm.n_contigs):
                s = asm.blob[asm.offsets[i]:asm.offsets[i + 1]]
                fh.write(b">NODE_%d_length_%d\n" % (i + 1, s.shape[0]))
                fh.write(b"\n".join(s[k:k + 60].tobytes() for k in range(0, s.shape[0], 60)) + b"\n")
        size = os.path.getsize(path)
        _lib.read_fasta(path, pinned=True)
        t0 = time.perf_counter()
        names, blob, offsets = _lib.read_fasta(path, pinned=True)
        nat = time.perf_counter() - t0
        t0 = time.perf_counter()
        recs = list(feature_utils._fasta_records(path))
        b2, o2 = feature_utils._ascii_blob([r[1] for r in recs])
        py = time.perf_counter() - t0
        assert np.array_equal(blob, b2) and np.array_equal(offsets, o2)
        out["fasta_reader"] = {"file_mb": size / 1e6, "native_s": nat, "native_gb_s": size / nat / 1e9,
                               "python_parse_and_join_s": py, "speedup": py / nat}

        # ---- per-bin FASTA writer (metacoag:1231-1257) ------------------------------------------------
        n_files = 200
        dest = (np.arange(len(names)) * 7919 % (n_files + 20)).astype(np.int32)
        dest[dest >= n_files] = -1
        paths = [os.path.join(tmp, "bin_%d.fasta" % i) for i in range(n_files)]
        t0 = time.perf_counter()
        _lib.write_bin_fastas(names, blob, offsets, dest, paths)
        nat = time.perf_counter() - t0
        written = sum(os.path.getsize(p) for p in paths)
        t0 = time.perf_counter()                         # the reference's way: second parse + write()
        files = [open(os.path.join(tmp, "ref_%d.fasta" % i), "w+") for i in range(n_files)]
        for k, (name, seq) in enumerate(feature_utils._fasta_records(path)):
            if dest[k] >= 0:
                files[dest[k]].write(f">{name}\n{seq}\n")
        for fh in files:
            fh.close()
        py = time.perf_counter() - t0
        for i in range(0, n_files, 17):
            assert open(paths[i], "rb").read() == open(os.path.join(tmp, "ref_%d.fasta" % i), "rb").read()
        out["bin_fasta_writer"] = {"files": n_files, "mb_written": written / 1e6, "native_s": nat,
                                   "native_gb_s": written / nat / 1e9,
                                   "python_second_parse_and_write_s": py, "speedup": py / nat}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
