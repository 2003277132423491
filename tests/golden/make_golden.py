#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the LIVE reference (build container only).

    python tests/golden/make_golden.py            # writes scan.npz, score.npz, pipeline.*

Every expected value stored here is produced by calling the unmodified reference
functions imported from /root/reference (see _refload.py).  Inputs are stored next
to the outputs so that the GPU-box tests never need the reference or this script.
"""
import hashlib
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refload import ShimGraph, load_reference  # noqa: E402

from metacoag_b200 import synth  # noqa: E402

fu, mu, lpu = load_reference()
MAX_WEIGHT = sys.float_info.max


def scan_cases():
    rng = np.random.default_rng(20240601)
    seqs = ["", "A", "ACG", "ACGT", "ACGTN", "NNNN", "acgtacgt", "ACGTNNacgtRYKMSWBDHV*-",
            "  ACGTACGTTTGACCA\n", "\tGGGGCCCCAAAATTTT  ", "T" * 257, "ACGT" * 100,
            "GATTACA" * 33 + "N" * 17 + "gattaca" * 5, "AC GT"]
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    for length in (5, 31, 32, 33, 63, 64, 65, 100, 511, 512, 513, 1000, 2047, 4099, 8192, 20011):
        gc = rng.uniform(0.25, 0.75)
        p = [(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2]
        s = alphabet[rng.choice(4, size=length, p=p)].copy()
        if length >= 100:
            for _ in range(3):
                a = int(rng.integers(0, length - 10))
                s[a:a + int(rng.integers(1, 10))] = ord("N")
        if length in (513, 2047):
            s |= 0x20
        seqs.append(s.tobytes().decode())
    return seqs


def make_scan():
    table, n_slots = fu.compute_kmer_inds(4)
    tab = np.array([table[c] for c in range(256)], dtype=np.uint8)
    assert n_slots == 136
    sha = hashlib.sha256(tab.tobytes()).hexdigest()
    assert sha == "9a57f94a839828cfdd42650170aa4db6e761932f7724c8056ab4ea818d0bc72f", sha
    seqs = scan_cases()
    counts = np.zeros((len(seqs), 136))
    freqs = np.zeros((len(seqs), 136))
    for i, s in enumerate(seqs):
        c, f = fu.count_kmers((s, 4, table, n_slots))
        counts[i], freqs[i] = c, f
    raw = [s.encode("ascii") for s in seqs]
    offsets = np.zeros(len(raw) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in raw], out=offsets[1:])
    blob = np.frombuffer(b"".join(raw), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "scan.npz"), table=tab, raw_blob=blob,
                        raw_offsets=offsets, counts=counts, freqs=freqs)
    print("scan.npz:", len(seqs), "sequences,", blob.shape[0], "bytes")


def cov_pairs(rng, n, S):
    a = np.round(rng.lognormal(np.log(20.0), 1.2, (n, S)), 4)
    b = np.round(a * np.clip(rng.normal(1.0, 0.25, (n, S)), 0.0, None), 4)
    swap = rng.random((n, S)) < 0.3
    b[swap] = np.round(rng.lognormal(np.log(20.0), 1.2, int(swap.sum())), 4)
    a[rng.random((n, S)) < 0.03] = 0.0
    b[rng.random((n, S)) < 0.03] = 0.0
    if n > 4:
        a[0, :] = 2500.0 + np.arange(S)          # large coverage: cancellation stress
        b[0, :] = 2510.5 + np.arange(S)
        a[1, :] = b[1, :]                        # identical vectors
        a[2, :] = 0.0                            # all clamped
        b[3, :] = 1.0                            # lgamma(2) = 0 special case
        a[3, :] = 1.0
    a[a < 1e-4] = 1e-4
    b[b < 1e-4] = 1e-4
    return a, b


def make_score():
    rng = np.random.default_rng(77)
    out = {}
    dist = np.concatenate([np.linspace(0.0, 0.25, 126), np.geomspace(1e-9, 1.35, 60),
                           [0.2, 0.21, 0.3, 1.0, 1.38]])
    out["dist"] = dist
    out["pdf_intra"] = np.array([mu.normpdf(d, mu.MU_INTRA, mu.SIGMA_INTRA) for d in dist])
    out["pdf_inter"] = np.array([mu.normpdf(d, mu.MU_INTER, mu.SIGMA_INTER) for d in dist])
    out["comp_prob"] = np.array([mu.get_comp_probability(d) for d in dist])
    for S in (1, 3, 10, 33, 50, 100):
        a, b = cov_pairs(rng, 64, S)
        out["covA_%d" % S] = a
        out["covB_%d" % S] = b
        out["covprob_%d" % S] = np.array([mu.get_cov_probability(list(x), list(y))
                                          for x, y in zip(a, b)])

    # a small contig table scored three ways
    asm = synth.make_assembly(seed=5, n_genomes=4, n_contigs=56, n_samples=4,
                              lengths=np.exp(np.random.default_rng(5).uniform(
                                  np.log(2500), np.log(60000), 56)).astype(np.int64))
    table, n_slots = fu.compute_kmer_inds(4)
    n = asm.n_contigs
    prof = np.stack([fu.count_kmers((asm.sequence(i), 4, table, n_slots))[1] for i in range(n)])
    cov = asm.coverage
    profiles = {i: prof[i] for i in range(n)}
    coverages = {i: list(cov[i]) for i in range(n)}
    out["prof"], out["cov"], out["genome_of"] = prof, cov, asm.genome_of

    pd = np.zeros((n, n)); pc = np.zeros((n, n)); pv = np.zeros((n, n)); lp = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            d = mu.get_tetramer_distance(profiles[i], profiles[j])
            c = mu.get_comp_probability(d)
            v = mu.get_cov_probability(coverages[i], coverages[j])
            pd[i, j], pc[i, j], pv[i, j] = d, c, v
            lp[i, j] = -(math.log(c, 10) + math.log(v, 10)) if c * v > 0.0 else MAX_WEIGHT
    out["pair_dist"], out["pair_pc"], out["pair_pv"], out["pair_lp"] = pd, pc, pv, lp

    # bins = first 5 contigs of every genome; rule C through assign_long itself
    bins = {g: [int(i) for i in np.flatnonzero(asm.genome_of == g)[:5]] for g in range(4)}
    bins[4] = [int(np.flatnonzero(asm.genome_of == 0)[5])]           # singleton bin
    members = np.array([c for b in bins for c in bins[b]], dtype=np.int64)
    seg = np.zeros(len(bins) + 1, dtype=np.int64)
    np.cumsum([len(bins[b]) for b in bins], out=seg[1:])
    out["members"], out["seg_offsets"] = members, seg
    cen_prof, cen_cov = fu.get_bin_profiles(bins, coverages, profiles)
    out["cen_prof"] = np.stack([cen_prof[b] for b in bins])
    out["cen_cov"] = np.stack([cen_cov[b] for b in bins])
    arg = np.full(n, -1, dtype=np.int32)
    wmin = np.full(n, MAX_WEIGHT)
    for i in range(n):
        r = lpu.assign_long(i, coverages, profiles, cen_prof, cen_cov)
        if r is not None:
            assert r[0] == i
            arg[i], wmin[i] = r[1], r[2]
    out["ruleC_arg"], out["ruleC_w"] = arg, wmin

    # rule B through run_bfs_long on a star graph: node -> one member of every bin
    rule_b = np.full((n, len(bins)), np.nan)
    smg_counts = [len(bins[b]) for b in bins]
    bin_of = {c: b for b in bins for c in bins[b]}
    for q in range(n):
        if q in bin_of:
            continue
        g = ShimGraph()
        g.add_vertices(n)
        g.add_edges([(q, bins[b][0]) for b in bins])
        g.simplify()
        for (node, hit, b, depth, w) in lpu.run_bfs_long(q, 1, list(bin_of.keys()), bin_of, bins,
                                                         smg_counts, g, profiles, coverages):
            assert node == q and depth == 1
            rule_b[q, b] = w
    out["ruleB_w"] = rule_b

    # rule A: capture the edge weights match_contigs hands to networkx
    captured = {}
    real_graph = mu.nx.Graph

    class Spy(real_graph):
        def add_edge(self, u, v, **kw):
            captured[(u, v)] = kw["weight"]
            super().add_edge(u, v, **kw)

    mu.nx.Graph = Spy
    try:
        unbinned = [i for i in range(n) if i not in bin_of][:12]
        g = ShimGraph(); g.add_vertices(n); g.simplify()
        bins_copy = {b: list(v) for b, v in bins.items()}
        mu.match_contigs({0: [bins[b][0] for b in bins], 1: unbinned}, bins_copy, len(bins),
                         dict(bin_of), [c for c in bin_of], {b: [] for b in bins},
                         {i: [] for i in range(n)}, {i: int(asm.lengths[i]) for i in range(n)},
                         {}, profiles, coverages, g, -1.0, MAX_WEIGHT, 20)
    finally:
        mu.nx.Graph = real_graph
    rule_a = np.full((n, len(bins)), np.nan)
    for (first_member, q), w in captured.items():
        b = [k for k in bins if bins[k][0] == first_member][0]
        rule_a[q, b] = w
    out["ruleA_w"] = rule_a
    assert np.isfinite(rule_a[unbinned]).all()
    np.savez_compressed(os.path.join(HERE, "score.npz"), **out)
    print("score.npz: n=%d contigs, %d bins; assigned %d" % (n, len(bins), int((arg >= 0).sum())))


if __name__ == "__main__":
    make_scan()
    make_score()
    if os.path.exists(os.path.join(HERE, "make_pipeline.py")):
        import make_pipeline
        make_pipeline.main()
