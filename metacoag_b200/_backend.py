"""Dict <-> dense staging and the scoring back end behind the drop-in modules.

The reference carries its features as Python dicts keyed by contig id
(``{contig: ndarray float64[136]}``, ``{contig: list[float]}``).  The
kernels want dense row-major tables resident in HBM, so the first scoring call of a
pipeline run stages both dicts once (row = contig id) and later calls only send ids.

``GpuBackend`` is the product path: every number comes from libmcg_b200.so.  The module
keeps one process-wide back end; tests replace it with a CPU checker through
``set_backend`` to exercise the host logic without a GPU -- nothing in this package
provides such a fallback itself.
"""
import threading

import numpy as np

from . import _lib

NPROF = _lib.NPROF
MAX_WEIGHT = _lib.MAX_WEIGHT


def dense_tables(normalized_tetramer_profiles, coverages):
    """Row-per-contig-id tables from the reference's dicts.

    Returns (prof [n,136], cov [n,S], have bool[n]); ``have`` marks ids present in both.
    Rows of absent ids hold zeros / ones and are never read: every id sent to a scorer is
    checked against ``have`` first (the reference would raise KeyError there).
    """
    ids = [c for c in normalized_tetramer_profiles]
    n = (max(ids) + 1) if ids else 0
    n_samples = 0
    for c in ids:
        if c in coverages:
            n_samples = len(coverages[c])
            break
    prof = np.zeros((n, NPROF), dtype=np.float64)
    cov = np.ones((n, max(n_samples, 1)), dtype=np.float64)
    have = np.zeros(n, dtype=bool)
    for c in ids:
        prof[c] = normalized_tetramer_profiles[c]
        if c in coverages and len(coverages[c]) == n_samples:
            cov[c] = coverages[c]
            have[c] = True
    return prof, cov, have


def csr_bins(bins, bin_order, counts=None):
    """members / seg_offsets CSR of ``bins`` in ``bin_order``; ``counts[b]`` limits a bin
    to its first members (the marker-seeded ones, label_prop_utils.py:58)."""
    members, seg = [], [0]
    for b in bin_order:
        m = bins[b] if counts is None else bins[b][:counts[b]]
        members.extend(int(x) for x in m)
        seg.append(len(members))
    return np.asarray(members, dtype=np.int64), np.asarray(seg, dtype=np.int64)


class GraphCache:
    """CSR copy of the duck-typed assembly graph (``neighbors(v, mode="ALL")``), built once
    per graph object; neighbour order is kept as the graph reports it."""

    def __init__(self):
        self.key = None
        self.indptr = self.indices = None

    def get(self, assembly_graph, n_vertices):
        key = (id(assembly_graph), n_vertices)
        if key != self.key:
            indptr = np.zeros(n_vertices + 1, dtype=np.int64)
            chunks = []
            for v in range(n_vertices):
                nb = assembly_graph.neighbors(v, mode="ALL")
                indptr[v + 1] = indptr[v] + len(nb)
                chunks.append(np.asarray(nb, dtype=np.int32))
            self.indices = (np.concatenate(chunks) if chunks and indptr[-1] > 0
                            else np.zeros(0, dtype=np.int32))
            self.indptr = indptr
            self.key = key
        return self.indptr, self.indices


graphs = GraphCache()


def vertex_count(assembly_graph, *dicts):
    if hasattr(assembly_graph, "vcount"):
        return int(assembly_graph.vcount())
    if hasattr(assembly_graph, "vs"):
        try:
            return len(assembly_graph.vs)
        except TypeError:
            pass
    n = 0
    for d in dicts:
        if len(d):
            n = max(n, max(d) + 1)
    return n


class GpuBackend:
    """Scoring through the C ABI; one resident contig table per (profiles, coverages) pair."""

    def __init__(self, device=None):
        self._ctx = None
        self._device = device
        self._key = None
        self._have = None
        self._lock = threading.RLock()

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.default_context(self._device)
        return self._ctx

    # -- staging ---------------------------------------------------------------------
    def stage(self, normalized_tetramer_profiles, coverages):
        key = (id(normalized_tetramer_profiles), len(normalized_tetramer_profiles),
               id(coverages), len(coverages))
        with self._lock:
            if key != self._key:
                prof, cov, have = dense_tables(normalized_tetramer_profiles, coverages)
                self.ctx.set_profiles(prof)
                self.ctx.load_coverage(cov)
                self._have = have
                self._key = key
        return self._have

    def _check(self, ids):
        ids = np.asarray(ids, dtype=np.int64)
        bad = (ids < 0) | (ids >= self._have.shape[0])
        if bad.any():
            raise KeyError(int(ids[bad][0]))
        missing = ~self._have[ids]
        if missing.any():
            raise KeyError(int(ids[missing][0]))
        return ids

    # -- featurisation -----------------------------------------------------------------
    def tetramer_scan(self, blob, offsets):
        with self._lock:
            self._key = None        # the scan replaces the resident table
            return self.ctx.tetramer_scan(blob, offsets)

    def bin_profiles(self, profiles, coverages, members, seg_offsets):
        with self._lock:
            self.stage(profiles, coverages)
            self._check(members)
            return self.ctx.bin_profiles(members, seg_offsets)

    # -- scoring ---------------------------------------------------------------------------
    def score_members(self, profiles, coverages, queries, members, seg_offsets, rule):
        """Rules A / B: weights [len(queries), n_bins]."""
        with self._lock:
            self.stage(profiles, coverages)
            return self.ctx.score_members(self._check(queries), self._check(members),
                                          seg_offsets, rule)

    def score_centroids(self, profiles, coverages, queries, cen_prof, cen_cov):
        """Rule C: (argmin with -1 = None, min weight)."""
        with self._lock:
            self.stage(profiles, coverages)
            arg, w, _ = self.ctx.score_centroids(queries=self._check(queries), cen_prof=cen_prof,
                                                 cen_cov=cen_cov)
            return arg, w

    def pair_matrix(self, prof, cov):
        """All-pairs (prob_comp, -log10 weight) of the rows of dense [n,136] / [n,S] tables:
        one launch for the bin x bin merge scoring (metacoag:1104-1131)."""
        with self._lock:
            self._key = None        # the tables replace the resident contig features
            self.ctx.set_profiles(prof)
            self.ctx.load_coverage(cov)
            ids = np.arange(prof.shape[0], dtype=np.int64)
            out = self.ctx.pair_scores(ids, ids, want=("pc", "lp"))
            return out["pc"], out["lp"]

    def bfs_candidates(self, indptr, indices, labels, starts, depth_limit):
        with self._lock:
            return self.ctx.bfs_candidates(indptr, indices, labels, starts, depth_limit)

    def connected_components(self, indptr, indices):
        with self._lock:
            return self.ctx.connected_components(indptr, indices)

    def path_lengths(self, indptr, indices, sources, targets):
        """[len(sources), len(targets)] vertices on a shortest path, 0 = unreachable."""
        with self._lock:
            return self.ctx.path_lengths(indptr, indices, sources, targets)

    # -- scalar primitives --------------------------------------------------------------------
    def normpdf(self, x, mean, sd):
        with self._lock:
            return self.ctx.normpdf(x, mean, sd)

    def tetramer_distance(self, u, v):
        with self._lock:
            return self.ctx.tetramer_distance(u, v)

    def comp_probability(self, dist):
        with self._lock:
            return self.ctx.comp_probability(dist)

    def cov_probability(self, cov1, cov2):
        with self._lock:
            return self.ctx.cov_probability(cov1, cov2)


_backend = None
_backend_lock = threading.Lock()


def backend():
    global _backend
    with _backend_lock:
        if _backend is None:
            _backend = GpuBackend()
        return _backend


def set_backend(obj):
    """Replace the process-wide back end (tests); returns the previous one."""
    global _backend
    with _backend_lock:
        old, _backend = _backend, obj
        return old
