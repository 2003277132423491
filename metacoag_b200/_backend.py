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


class ProfileTable(dict):
    """What feature_utils.get_tetramer_profiles returns: the reference's ``{contig: float64[136]}``
    dict (the values are rows of ``dense``, not copies) that also carries the dense table of ALL
    records it was cut from, so that staging it for the scorers needs no per-contig Python work
    (the boundary is dict-shaped, the kernels want rows)."""

    dense = None        # float64 [records, 136], row i = i-th FASTA record
    resident = None     # token of the back end whose device table still holds `dense`


def dense_tables(normalized_tetramer_profiles, coverages):
    """Row-per-contig-id tables from the reference's dicts.

    Returns (prof [n,136], cov [n,S], have bool[n]); ``have`` marks ids present in both.
    Rows of absent ids hold zeros / ones and are never read: every id sent to a scorer is
    checked against ``have`` first (the reference would raise KeyError there).
    No Python loop over contigs on the usual path: one ``fromiter`` over the keys, one stacked
    assignment per table.
    """
    count = len(normalized_tetramer_profiles)
    ids = np.fromiter(normalized_tetramer_profiles.keys(), dtype=np.int64, count=count)
    dense = getattr(normalized_tetramer_profiles, "dense", None)
    if dense is not None and (count == 0 or int(ids.max()) < dense.shape[0]):
        prof = dense
        n = dense.shape[0]
    else:
        n = int(ids.max()) + 1 if count else 0
        prof = np.zeros((n, NPROF), dtype=np.float64)
        if count:
            prof[ids] = np.asarray(list(normalized_tetramer_profiles.values()), dtype=np.float64)
    have = np.zeros(n, dtype=bool)
    keys = [c for c in ids.tolist() if c in coverages]
    n_samples = len(coverages[keys[0]]) if keys else 0
    cov = np.ones((n, max(n_samples, 1)), dtype=np.float64)
    if keys:
        rows = [coverages[c] for c in keys]
        try:
            block = np.asarray(rows, dtype=np.float64)
        except ValueError:
            block = None
        if block is not None and block.ndim == 2 and block.shape[1] == n_samples:
            kid = np.asarray(keys, dtype=np.int64)
            cov[kid] = block
            have[kid] = True
        else:   # ragged rows: only those with the sample count of the first one take part
            for c, row in zip(keys, rows):
                if len(row) == n_samples:
                    cov[c] = row
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


def _edge_count(assembly_graph):
    try:
        return int(assembly_graph.ecount())
    except (AttributeError, TypeError):
        return -1


class GraphCache:
    """CSR copy of the duck-typed assembly graph (``neighbors(v, mode="ALL")``), built once
    per graph object; neighbour order is kept as the graph reports it.

    The cache holds a strong reference to the graph it was built from and compares with
    ``is`` (an ``id()`` can be reused by a new object once the old one is freed); vertex and
    edge counts are part of the key so that a graph edited in place is rebuilt.  ``version``
    changes with every rebuild: the device-resident copy (GpuBackend) follows it."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.graph = None
        self.shape = None
        self.indptr = self.indices = None
        self.version = 0

    def get(self, assembly_graph, n_vertices):
        shape = (n_vertices, _edge_count(assembly_graph))
        if self.graph is not assembly_graph or shape != self.shape:
            indptr = np.zeros(n_vertices + 1, dtype=np.int64)
            chunks = []
            for v in range(n_vertices):
                nb = assembly_graph.neighbors(v, mode="ALL")
                indptr[v + 1] = indptr[v] + len(nb)
                chunks.append(np.asarray(nb, dtype=np.int32))
            self.indices = (np.concatenate(chunks) if chunks and indptr[-1] > 0
                            else np.zeros(0, dtype=np.int32))
            self.indptr = indptr
            self.graph = assembly_graph
            self.shape = shape
            self.version += 1
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
    """Scoring through the C ABI.  One device pool for the process (every GPU of the box behind
    this one caller, ``_lib.default_pool``), one staged contig table per (profiles, coverages)
    pair.  The caches hold strong references to the objects they were built from and compare
    with ``is``: an ``id()`` is only unique while its object lives."""

    def __init__(self, pool=None):
        self._pool = pool
        self._lock = threading.RLock()
        self.reset()

    def reset(self):
        """Forget every staged table (a new pipeline run in the same process)."""
        with self._lock:
            self._staged = None          # (profiles, len, coverages, len)
            self._have = None
            self._scan_token = None      # dense profile table still resident after a scan
            self._graph_version = None
            self._graph_arrays = None
            self._labels = None

    @property
    def ctx(self):
        if self._pool is None:
            self._pool = _lib.default_pool()
        return self._pool

    # -- staging ---------------------------------------------------------------------
    def stage(self, normalized_tetramer_profiles, coverages):
        with self._lock:
            st = self._staged
            if (st is None or st[0] is not normalized_tetramer_profiles or
                    st[1] != len(normalized_tetramer_profiles) or st[2] is not coverages or
                    st[3] != len(coverages)):
                prof, cov, have = dense_tables(normalized_tetramer_profiles, coverages)
                # the table the scan left on the device is this very array: only the coverages
                # still have to travel (single device; a pool scans row shards)
                if (self._scan_token is not None and prof is self._scan_token and
                        self.ctx.size == 1):
                    self.ctx.set_features(cov=cov)
                else:
                    self.ctx.set_features(prof=prof, cov=cov)
                    self._scan_token = None
                self._have = have
                self._staged = (normalized_tetramer_profiles, len(normalized_tetramer_profiles),
                                coverages, len(coverages))
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
    def tetramer_scan(self, blob, offsets, packed=None):
        """(counts uint32 [n,136], freqs float64 [n,136]); ``packed`` = the same contigs in the
        2-bit layout (packed at parse time): a quarter of the bytes on the link."""
        with self._lock:
            self._staged = None        # the scan replaces the resident table
            if packed is not None:
                counts, freqs = self.ctx.tetramer_scan(packed, offsets, encoding=_lib.ENC_2BIT)
            else:
                counts, freqs = self.ctx.tetramer_scan(blob, offsets)
            self._scan_token = freqs
            return counts, freqs

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

    def score_member_items(self, profiles, coverages, item_query, item_bin, members, seg_offsets,
                           rule):
        """Rules A / B for listed (contig, bin index) items: weights [n_items]."""
        with self._lock:
            self.stage(profiles, coverages)
            return self.ctx.score_member_items(self._check(item_query), item_bin,
                                               self._check(members), seg_offsets, rule)

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
            self._staged = None        # the tables replace the resident contig features
            self._scan_token = None
            self.ctx.set_features(prof=prof, cov=cov)
            ids = np.arange(prof.shape[0], dtype=np.int64)
            out = self.ctx.pair_scores(ids, ids, want=("pc", "lp"))
            return out["pc"], out["lp"]

    # -- label propagation: graph and labels resident -----------------------------------------
    def frontier(self, indptr, indices, version, labels):
        """Graph (uploaded once per ``version`` of the CSR arrays) and labels on the device."""
        with self._lock:
            if self._graph_version != version or self._graph_arrays is not indptr:
                self.ctx.graph_load(indptr, indices)
                self._graph_version, self._graph_arrays = version, indptr
            self.ctx.labels_load(labels)

    def frontier_patch(self, ids, bins):
        with self._lock:
            self.ctx.labels_set(ids, bins)

    def frontier_walks(self, starts, depth_limit):
        """(hit_start, hit_count, hits [total,3]) of the bounded walks from ``starts``."""
        with self._lock:
            return self.ctx.bfs_resident(starts, depth_limit)

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


def reset():
    """Drop every cache that is keyed on the caller's objects (staged tables, CSR graph,
    centroid scores): called where a pipeline run starts (feature_utils.get_cov_len*), so that a
    second run in the same process can never read the first one's tables."""
    graphs.reset()
    with _backend_lock:
        obj = _backend
    if obj is not None and hasattr(obj, "reset"):
        obj.reset()
    for hook in _reset_hooks:
        hook()


_reset_hooks = []


def set_backend(obj):
    """Replace the process-wide back end (tests); returns the previous one."""
    global _backend
    with _backend_lock:
        old, _backend = _backend, obj
        return old
