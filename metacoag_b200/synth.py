"""Synthetic metagenome assemblies for parity tests and benchmarks.

Host-side (numpy) generator.  A data set is G genomes, each with its own
dinucleotide composition and per-sample abundance; contigs are drawn from the
genomes, a fraction of bases become ``N`` runs, a fraction of contigs are
lower-cased (both count as ``A`` in the reference, feature_utils.py:21,31-35), and
coverage is abundance times N(1, 0.1^2) printed with four decimals with a share
forced to 0.0000 (exercises the 1e-4 clamp of feature_utils.py:150-153).

The big benchmark configurations generate their bases on the device instead
(``mcg_synth_bases`` in the C-ABI) because a 1.25 Gbp string takes minutes in
numpy; the statistical recipe is the same.
"""
from dataclasses import dataclass, field

import numpy as np

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
N_MARKER_GENES = 108


@dataclass
class SynthAssembly:
    blob: np.ndarray            # uint8 ASCII bases of all contigs, concatenated
    offsets: np.ndarray         # int64 [n+1]
    lengths: np.ndarray         # int64 [n]
    genome_of: np.ndarray       # int32 [n]
    coverage: np.ndarray        # float64 [n, S], already clamped like the reference
    coverage_text: list         # the "%.4f" strings before clamping, per contig
    edges: np.ndarray           # int64 [E, 2], undirected, simplified
    contig_markers: dict = field(default_factory=dict)   # contig -> [gene name]
    n_samples: int = 0

    @property
    def n_contigs(self):
        return int(self.lengths.shape[0])

    def sequence(self, i):
        return self.blob[self.offsets[i]:self.offsets[i + 1]].tobytes().decode("ascii")

    def sequences(self):
        return [self.sequence(i) for i in range(self.n_contigs)]


def sample_lengths(rng, n, kind):
    """Contig length distributions named in BASELINE.json's configs."""
    if kind == "spades":        # log-uniform 1-50 kb
        return np.exp(rng.uniform(np.log(1000), np.log(50000), n)).astype(np.int64)
    if kind == "megahit":       # lognormal(ln 1800, 0.9) clipped to [1 kb, 300 kb]
        return np.clip(rng.lognormal(np.log(1800), 0.9, n), 1000, 300000).astype(np.int64)
    if kind == "flye":          # log-uniform 10 kb - 5 Mb
        return np.exp(rng.uniform(np.log(10000), np.log(5000000), n)).astype(np.int64)
    if kind == "mixed":         # small test mix incl. contigs below min_length
        return np.exp(rng.uniform(np.log(300), np.log(40000), n)).astype(np.int64)
    raise ValueError(kind)


def genome_tables(rng, n_genomes, n_samples):
    """Per-genome dinucleotide distribution [G,16] and abundance [G,S]."""
    gc = rng.uniform(0.30, 0.70, n_genomes)
    base = np.stack([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2], axis=1)   # A C G T
    dinuc = np.empty((n_genomes, 16))
    for g in range(n_genomes):
        rows = np.stack([rng.dirichlet(8.0 * base[g] + 0.5) for _ in range(4)])
        dinuc[g] = (base[g][:, None] * rows).reshape(16)
    dinuc /= dinuc.sum(axis=1, keepdims=True)
    abundance = rng.lognormal(np.log(20.0), 1.0, (n_genomes, n_samples))
    return dinuc, abundance


def _draw_bases(rng, dinuc_row, length):
    pairs = rng.choice(16, size=(length + 1) // 2, p=dinuc_row)
    codes = np.empty(2 * pairs.shape[0], dtype=np.uint8)
    codes[0::2] = pairs >> 2
    codes[1::2] = pairs & 3
    return _BASES[codes[:length]]


def make_assembly(seed, n_genomes, n_contigs, n_samples, length_kind="spades",
                  n_frac=1e-4, lower_frac=0.01, zero_cov_frac=0.01, with_graph=True,
                  with_markers=True, lengths=None):
    rng = np.random.default_rng(seed)
    dinuc, abundance = genome_tables(rng, n_genomes, n_samples)
    if lengths is None:
        lengths = sample_lengths(rng, n_contigs, length_kind)
    lengths = np.asarray(lengths, dtype=np.int64)
    n_contigs = lengths.shape[0]
    genome_of = rng.integers(0, n_genomes, n_contigs).astype(np.int32)
    offsets = np.zeros(n_contigs + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    blob = np.empty(int(offsets[-1]), dtype=np.uint8)
    for i in range(n_contigs):
        seq = _draw_bases(rng, dinuc[genome_of[i]], int(lengths[i]))
        if rng.random() < lower_frac:
            seq = seq | 0x20                       # lower case
        blob[offsets[i]:offsets[i + 1]] = seq
    # N runs: n_frac of all bases, in runs of 1..50
    total = blob.shape[0]
    n_target = int(total * n_frac)
    placed = 0
    while placed < n_target and total > 0:
        start = int(rng.integers(0, total))
        run = int(rng.integers(1, 51))
        blob[start:start + run] = ord("N")
        placed += run

    noise = np.clip(rng.normal(1.0, 0.1, (n_contigs, n_samples)), 0.0, None)
    raw = abundance[genome_of] * noise
    raw[rng.random((n_contigs, n_samples)) < zero_cov_frac] = 0.0
    text = [["%.4f" % v for v in row] for row in raw]
    coverage = np.array([[float(t) for t in row] for row in text], dtype=np.float64)
    coverage = coverage.reshape(n_contigs, n_samples)
    coverage[coverage < 1e-4] = 1e-4

    edges = np.zeros((0, 2), dtype=np.int64)
    if with_graph:
        edges = _make_graph(rng, genome_of, n_genomes)
    markers = {}
    if with_markers:
        markers = _make_markers(rng, genome_of, lengths, n_genomes)
    return SynthAssembly(blob=blob, offsets=offsets, lengths=lengths, genome_of=genome_of,
                         coverage=coverage, coverage_text=text, edges=edges,
                         contig_markers=markers, n_samples=n_samples)


def _make_graph(rng, genome_of, n_genomes):
    """Per-genome chain + 0.25 n random chords + 2 % cross-genome edges."""
    n = genome_of.shape[0]
    pairs = set()
    for g in range(n_genomes):
        members = np.flatnonzero(genome_of == g)
        rng.shuffle(members)
        for a, b in zip(members[:-1], members[1:]):
            pairs.add((int(min(a, b)), int(max(a, b))))
        if members.shape[0] >= 3:
            for _ in range(int(0.25 * members.shape[0])):
                a, b = rng.choice(members, 2, replace=False)
                pairs.add((int(min(a, b)), int(max(a, b))))
    for _ in range(max(1, int(0.02 * n))) if n >= 2 else []:
        a, b = rng.choice(n, 2, replace=False)
        pairs.add((int(min(a, b)), int(max(a, b))))
    if not pairs:
        return np.zeros((0, 2), dtype=np.int64)
    return np.array(sorted(pairs), dtype=np.int64)


def _make_markers(rng, genome_of, lengths, n_genomes, min_marker_len=3000):
    """Each genome places each of 108 single-copy marker genes (present w.p. 0.9)
    on one of its 1..40 marker contigs (contigs >= 3 kb)."""
    markers = {}
    for g in range(n_genomes):
        eligible = np.flatnonzero((genome_of == g) & (lengths >= min_marker_len))
        if eligible.shape[0] == 0:
            continue
        k = int(min(eligible.shape[0], rng.integers(1, 41)))
        hosts = rng.choice(eligible, k, replace=False)
        for gene in range(N_MARKER_GENES):
            if rng.random() < 0.9:
                c = int(hosts[rng.integers(0, k)])
                markers.setdefault(c, []).append("MG%03d" % gene)
    return markers


def csr_from_edges(n, edges):
    """Undirected simplified adjacency in CSR with ascending neighbour ids (the order
    igraph's neighbors(v, mode="ALL") gives after simplify())."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    keep = edges[:, 0] != edges[:, 1]
    edges = edges[keep]
    both = np.concatenate([edges, edges[:, ::-1]], axis=0)
    if both.shape[0]:
        both = np.unique(both, axis=0)
    indptr = np.zeros(n + 1, dtype=np.int64)
    if both.shape[0]:
        np.add.at(indptr, both[:, 0] + 1, 1)
    np.cumsum(indptr, out=indptr)
    indices = both[:, 1].astype(np.int32) if both.shape[0] else np.zeros(0, dtype=np.int32)
    return indptr, indices


def graph_csr_large(seed, genome_of, chord_frac=0.25, cross_frac=0.02):
    """The graph recipe of ``_make_graph`` (per-genome chain + 0.25 n chords + 2 % cross-genome
    edges) without Python loops, for the million-vertex benchmark configurations; returns the
    simplified undirected CSR (indptr int64, indices int32, ascending neighbour ids)."""
    rng = np.random.default_rng([seed, 4242])
    genome_of = np.asarray(genome_of)
    n = genome_of.shape[0]
    # contigs of a genome in a random order, genomes one after the other
    order = np.lexsort((rng.random(n), genome_of))
    g_sorted = genome_of[order]
    same = g_sorted[1:] == g_sorted[:-1]
    chain = np.stack([order[:-1][same], order[1:][same]], axis=1)
    # chords: a random partner inside the same genome's run of `order`
    starts = np.flatnonzero(np.concatenate([[True], ~same]))
    sizes = np.diff(np.concatenate([starts, [n]]))
    run_of = np.repeat(np.arange(starts.shape[0]), sizes)
    k = int(chord_frac * n)
    a_pos = rng.integers(0, n, k)
    b_pos = starts[run_of[a_pos]] + (rng.random(k) * sizes[run_of[a_pos]]).astype(np.int64)
    chords = np.stack([order[a_pos], order[b_pos]], axis=1)
    cross = rng.integers(0, n, (max(1, int(cross_frac * n)), 2))
    edges = np.concatenate([chain, chords, cross]).astype(np.int64)
    edges = edges[edges[:, 0] != edges[:, 1]]
    both = np.concatenate([edges, edges[:, ::-1]])
    key = np.unique(both[:, 0] * n + both[:, 1])
    src, dst = key // n, key % n
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=indptr[1:])
    return indptr, dst.astype(np.int32)
