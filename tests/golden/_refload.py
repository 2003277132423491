"""Import the UNMODIFIED reference modules from /root/reference (build container only).

Used by tests/golden/make_golden.py to generate the committed fixtures.  Nothing
that runs on the GPU box imports this file: /root/reference does not exist there.

The reference needs two packages that are not installed here:
``Bio`` (only ``SeqIO.parse(path, "fasta")`` is used) and ``igraph`` (only a handful
of ``Graph`` methods).  Both are replaced by the minimal stand-ins below; the
reference sources themselves are imported as they are.
"""
import importlib
import sys
import types

REFERENCE_ROOT = "/root/reference"


class _Record:
    def __init__(self, name, description, seq):
        self.id = name
        self.description = description
        self.seq = seq


def _parse_fasta(path, fmt):
    """Bio.SeqIO.parse(path, "fasta") as far as the reference uses it.  The rules are those of
    Bio.SeqIO.FastaIO.SimpleFastaParser (BioPython 1.7x/1.8x): text before the first '>' line is
    skipped; title = line[1:].rstrip(); id = first word of the title; the sequence is the
    right-stripped lines joined, with ' ' and '\\r' removed; the handle is text mode."""
    assert fmt == "fasta"
    with open(path) as fh:
        title = None
        for line in fh:
            if line[0] == ">":
                title = line[1:].rstrip()
                break
        if title is None:
            return
        lines = []
        for line in fh:
            if line[0] == ">":
                words = title.split(None, 1)
                yield _Record(words[0] if words else "", title,
                              "".join(lines).replace(" ", "").replace("\r", ""))
                lines = []
                title = line[1:].rstrip()
                continue
            lines.append(line.rstrip())
        words = title.split(None, 1)
        yield _Record(words[0] if words else "", title,
                      "".join(lines).replace(" ", "").replace("\r", ""))


class ShimGraph:
    """Duck-typed stand-in for igraph.Graph: undirected, neighbours in ascending id
    order once simplified (igraph's adjacency order for mode="ALL")."""

    class _VS:
        def __init__(self, g):
            self.g = g

        def __getitem__(self, i):
            return self.g._vattr.setdefault(i, {})

        def __len__(self):
            return self.g.n

    def __init__(self):
        self.n = 0
        self.adj = []
        self._vattr = {}
        self.vs = ShimGraph._VS(self)

    def add_vertices(self, n):
        self.adj.extend([] for _ in range(n))
        self.n += n

    def add_edges(self, edges):
        for a, b in edges:
            self.add_edge(a, b)

    def add_edge(self, a, b):
        self.adj[a].append(b)
        self.adj[b].append(a)

    def simplify(self, multiple=True, loops=True, combine_edges=None):
        for v in range(self.n):
            self.adj[v] = sorted(set(w for w in self.adj[v] if w != v))

    @property
    def es(self):
        return [(a, b) for a in range(self.n) for b in self.adj[a] if a < b]

    def neighbors(self, v, mode="ALL"):
        return list(self.adj[v])

    def get_shortest_paths(self, v, to=None):
        prev = {v: None}
        frontier = [v]
        while frontier and to not in prev:
            nxt = []
            for a in frontier:
                for b in self.adj[a]:
                    if b not in prev:
                        prev[b] = a
                        nxt.append(b)
            frontier = nxt
        if to not in prev:
            return [[]]
        path = [to]
        while prev[path[-1]] is not None:
            path.append(prev[path[-1]])
        return [path[::-1]]

    def maximal_cliques(self):
        import networkx as nx

        g = nx.Graph()
        g.add_nodes_from(range(self.n))
        g.add_edges_from(self.es)
        return [tuple(sorted(c)) for c in nx.find_cliques(g)]


def load_reference():
    """Returns the reference's (feature_utils, matching_utils, label_prop_utils)."""
    if "Bio" not in sys.modules:
        bio = types.ModuleType("Bio")
        seqio = types.ModuleType("Bio.SeqIO")
        seqio.parse = _parse_fasta
        bio.SeqIO = seqio
        sys.modules["Bio"] = bio
        sys.modules["Bio.SeqIO"] = seqio
    if "igraph" not in sys.modules:
        ig = types.ModuleType("igraph")
        ig.Graph = ShimGraph
        sys.modules["igraph"] = ig
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    fu = importlib.import_module("metacoag_utils.feature_utils")
    mu = importlib.import_module("metacoag_utils.matching_utils")
    lp = importlib.import_module("metacoag_utils.label_prop_utils")
    assert fu.__file__.startswith(REFERENCE_ROOT), fu.__file__
    return fu, mu, lp
