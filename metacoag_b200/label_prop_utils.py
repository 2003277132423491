"""Drop-in for ``metacoag_utils/label_prop_utils.py`` (MetaCoAG v1.1.5) on the B200 path.

The candidate search (bounded BFS to labelled contigs with the reference's depth bookkeeping)
and the rule-B / rule-C scoring run in libmcg_b200.so; the priority-queue loop and the
marker-gene rules stay on the host, in the reference's order.  Citations are
``label_prop_utils.py:line`` of the reference.
"""
import heapq
import logging
import sys

import numpy as np

from . import _backend
from ._lib import RULE_B

MAX_WEIGHT = sys.float_info.max

logger = logging.getLogger("MetaCoaAG 1.1.5")


class DataWrap:
    """Heap entry ordered by (depth, score) of its 5-tuple (:16-21)."""

    def __init__(self, data):
        self.data = data

    def __lt__(self, other):
        return (self.data[3], self.data[4]) < (other.data[3], other.data[4])


# ---- assembly graph as CSR (shared with matching_utils) ----------------------------------------
_graphs = _backend.graphs
_vertex_count = _backend.vertex_count


class _SeedScores:
    """Rule-B score of (node, bin): mean -log10 score against the first
    ``smg_bin_counts[bin]`` members of the bin (:54-86).  Those members never change during a
    stage, so a score is computed once; like the reference, only the bins a node's walk
    reached are scored (one launch per batch of new (node, bin) items)."""

    def __init__(self, bins, smg_bin_counts, profiles, coverages):
        self.order = list(bins)
        self.col = {b: i for i, b in enumerate(self.order)}
        self.members, self.seg = _backend.csr_bins(bins, self.order, smg_bin_counts)
        self.profiles, self.coverages = profiles, coverages
        self.cache = {}

    def ensure(self, items):
        todo = [it for it in dict.fromkeys(items) if it not in self.cache]
        if todo:
            query = np.fromiter((it[0] for it in todo), dtype=np.int64, count=len(todo))
            col = np.fromiter((self.col[it[1]] for it in todo), dtype=np.int32, count=len(todo))
            w = _backend.backend().score_member_items(self.profiles, self.coverages, query, col,
                                                      self.members, self.seg, RULE_B)
            self.cache.update(zip(todo, w.tolist()))

    def get(self, node, bin_):
        return self.cache[(node, bin_)]


class _Frontier:
    """The assembly graph and the labels of one propagation stage, resident in the back end:
    the CSR graph goes up once per graph, the labels once per stage, and every acceptance
    patches ONE label (:270-271, :472-473) instead of rebuilding and re-sending both."""

    def __init__(self, assembly_graph, n_vertices, bin_of_contig):
        self.indptr, self.indices = _graphs.get(assembly_graph, n_vertices)
        labels = np.full(n_vertices, -1, dtype=np.int32)
        if len(bin_of_contig):
            ids = np.fromiter(bin_of_contig.keys(), dtype=np.int64, count=len(bin_of_contig))
            labels[ids] = np.fromiter(bin_of_contig.values(), dtype=np.int32,
                                      count=len(bin_of_contig))
        self.labels = labels
        self.backend = _backend.backend()
        self.backend.frontier(self.indptr, self.indices, _graphs.version, labels)

    def patch(self, contig, bin_):
        self.labels[contig] = bin_
        self.backend.frontier_patch([contig], [bin_])

    def walks(self, nodes, threshold):
        """Per start node the (labelled node, bin, depth) triples of run_bfs_long, pop order."""
        starts = np.asarray(list(nodes), dtype=np.int64)
        if starts.shape[0] == 0:
            return []
        first, count, hits = self.backend.frontier_walks(starts, threshold)
        rows = hits.tolist()
        return [rows[a:a + c] for a, c in zip(first.tolist(), count.tolist())]


def _local_candidates(nodes, threshold, frontier, scores):
    """(node, labelled, bin, depth, score) rows of ``nodes``, hits in pop order."""
    hits = frontier.walks(nodes, threshold)
    scores.ensure((n, b) for n, node_hits in zip(nodes, hits) for _, b, _ in node_hits)
    rows = []
    for n, node_hits in zip(nodes, hits):
        for labelled, b, d in node_hits:
            rows.append((n, labelled, b, d, scores.get(n, b)))
    return rows


def _distributed_world():
    """World size when torch.distributed is up (the package does not import torch otherwise)."""
    import sys
    if "torch.distributed" not in sys.modules:
        return 1
    from . import parallel
    return parallel.world()


def _candidates(nodes, threshold, frontier, scores):
    """``[set of (node, labelled, bin, depth, score)]`` per node, the sets built by inserting
    the hits in pop order like the reference does (:88-90).  A single process spreads the start
    nodes over its GPUs inside the library (mcg_pool_bfs).  With several PROCESSES
    (torch.distributed initialised) large batches are sharded: every rank walks and scores a
    slice of the nodes and the rows are all-gathered (metacoag_b200/parallel.py)."""
    nodes = [int(n) for n in nodes]
    if not nodes:
        return []
    rows = None
    if _distributed_world() > 1:
        from . import parallel
        if len(nodes) >= parallel.MIN_SHARDED_NODES:
            lo, hi = parallel.shard_slice(len(nodes))
            mine = _local_candidates(nodes[lo:hi], threshold, frontier, scores)
            ints = np.array([r[:4] for r in mine], dtype=np.int64).reshape(-1, 4)
            floats = np.array([[r[4]] for r in mine], dtype=np.float64).reshape(-1, 1)
            ints, floats = parallel.all_gather_rows(ints, floats)
            rows = [(int(a), int(b), int(c), int(d), float(f[0])) for (a, b, c, d), f
                    in zip(ints.tolist(), floats.tolist())]
    if rows is None:
        rows = _local_candidates(nodes, threshold, frontier, scores)
    by_node = {}
    for row in rows:
        by_node.setdefault(row[0], set()).add(row)
    return [by_node.get(n, set()) for n in nodes]


def run_bfs_long(node, threhold, binned_contigs, bin_of_contig, bins, smg_bin_counts,
                 assembly_graph, normalized_tetramer_profiles, coverages):
    """Bounded BFS from ``node`` to labelled contigs (:24-100): a set of
    ``(node, labelled node, bin, depth, score)``."""
    n_vertices = _vertex_count(assembly_graph, bin_of_contig, normalized_tetramer_profiles,
                               {node: 0})
    scores = _SeedScores(bins, smg_bin_counts, normalized_tetramer_profiles, coverages)
    frontier = _Frontier(assembly_graph, n_vertices, bin_of_contig)
    return _candidates([node], threhold, frontier, scores)[0]


def run_bfs_short(node, threhold, binned_contigs, bin_of_contig, assembly_graph, coverages):
    """Never called by the pipeline; kept for the interface.
    Same traversal, score = Euclidean distance of the coverage vectors (:103-141)."""
    n_vertices = _vertex_count(assembly_graph, bin_of_contig, {node: 0})
    hits = _Frontier(assembly_graph, n_vertices, bin_of_contig).walks([node], threhold)[0]
    found = set()
    for labelled, b, d in hits:
        dist = float(_backend.backend().tetramer_distance(
            np.asarray(coverages[labelled], dtype=np.float64),
            np.asarray(coverages[node], dtype=np.float64))[0])
        found.add((node, labelled, b, d, dist))
    return found


def getClosestLongVertices(graph, node, binned_contigs, contig_lengths, min_length):
    """Level-order search from ``node``: the long unbinned contigs of the first level that has
    any (:144-172).  Levels are deduplicated and never revisit a vertex."""
    binned = binned_contigs if isinstance(binned_contigs, (set, frozenset, dict)) \
        else set(binned_contigs)
    visited = {node}
    level = list(graph.neighbors(node, mode="ALL"))
    while level:
        visited.update(level)
        unlabelled = [n for n in level if contig_lengths[n] >= min_length and n not in binned]
        if unlabelled:
            return unlabelled
        merged = []
        for n in level:
            merged = list(set(merged + list(graph.neighbors(n, mode="ALL"))))
        level = [n for n in merged if n not in visited]
    return []


def _closest_long(graph, node, bin_of_contig, contig_lengths, min_length):
    return [x for x in getClosestLongVertices(graph, node, bin_of_contig, contig_lengths,
                                              min_length)
            if contig_lengths[x] >= min_length]


def _propagate(bin_of_contig, bins, contig_markers, bin_markers, binned_contigs_with_markers,
               smg_bin_counts, seeds, contig_lengths, min_length, assembly_graph,
               normalized_tetramer_profiles, coverages, depth, accept):
    """The shared heap loop of label_prop (:196-305) and final_label_prop (:419-511)."""
    n_vertices = _vertex_count(assembly_graph, contig_lengths, bin_of_contig)
    scores = _SeedScores(bins, smg_bin_counts, normalized_tetramer_profiles, coverages)

    contigs_to_bin = set()
    for contig in seeds:
        contigs_to_bin.update(_closest_long(assembly_graph, contig, bin_of_contig, contig_lengths,
                                            min_length))

    if not contigs_to_bin:     # nothing to propagate to (label_prop as shipped)
        return bins, bin_of_contig, bin_markers, binned_contigs_with_markers
    frontier = _Frontier(assembly_graph, n_vertices, bin_of_contig)

    heap = []
    for found in _candidates(contigs_to_bin, depth, frontier, scores):
        for data in list(found):
            heapq.heappush(heap, DataWrap(data))

    while heap:
        to_bin, binned, bin_, dist, score = heapq.heappop(heap).data

        has_mg = False
        common_mgs = set()
        if to_bin in contig_markers:
            has_mg = True
            common_mgs = set(bin_markers[bin_]).intersection(set(contig_markers[to_bin]))
            if binned in contig_markers and dist == 1:
                neighbour_common = set(contig_markers[binned]).intersection(
                    set(contig_markers[to_bin]))
                if neighbour_common == common_mgs:
                    common_mgs = set()

        if not accept(to_bin, score, dist, common_mgs):
            continue

        bins[bin_].append(to_bin)
        bin_of_contig[to_bin] = bin_
        frontier.patch(to_bin, bin_)
        if has_mg:
            binned_contigs_with_markers.append(to_bin)
            bin_markers[bin_] = list(set(bin_markers[bin_] + contig_markers[to_bin]))

        # the contigs around the new member get fresh candidates
        around = set(_closest_long(assembly_graph, to_bin, bin_of_contig, contig_lengths,
                                   min_length))
        # (filter + heapify also when `around` is empty: heapify re-arranges entries with equal
        # keys, and the pop order of ties is the reference's only with the same call sequence)
        heap = [x for x in heap if x.data[0] not in around]
        heapq.heapify(heap)
        for found in _candidates(around, depth, frontier, scores):
            for data in list(found):
                heapq.heappush(heap, DataWrap(data))

    return bins, bin_of_contig, bin_markers, binned_contigs_with_markers


def label_prop(bin_of_contig, bins, contig_markers, bin_markers, binned_contigs_with_markers,
               smg_bin_counts, non_isolated, contig_lengths, min_length, assembly_graph,
               normalized_tetramer_profiles, coverages, depth, weight):
    """Label propagation from the binned long contigs listed in ``non_isolated`` (:175-305).
    A candidate is accepted when it is still unbinned, scores below ``weight``, lies within
    ``depth`` and shares no marker gene with the bin (or one, if longer than 100 kb)."""
    seeds = [c for c in bin_of_contig
             if c in non_isolated and contig_lengths[c] >= min_length]

    def accept(to_bin, score, dist, common_mgs):
        if to_bin in bin_of_contig or not (score < weight and dist <= depth):
            return False
        return len(common_mgs) == 0 or (len(common_mgs) <= 1 and contig_lengths[to_bin] > 100000)

    return _propagate(bin_of_contig, bins, contig_markers, bin_markers,
                      binned_contigs_with_markers, smg_bin_counts, seeds, contig_lengths,
                      min_length, assembly_graph, normalized_tetramer_profiles, coverages, depth,
                      accept)


# ---- rule C: contig against bin centroids ------------------------------------------------------
class _CentroidCache:
    """assign_long is called once per contig, from several threads (metacoag:959-999).  The
    first call of a stage scores every staged long contig against the centroids in one
    launch; the other calls are lookups.  The cache keeps the four dicts it was computed from
    alive and compares them with ``is`` (an id() can be reused once its object is freed)."""

    def __init__(self):
        import threading
        self.lock = threading.Lock()
        self.reset()

    def reset(self):
        self.key = None
        self.result = {}

    def _same(self, key):
        return (self.key is not None and all(a is b for a, b in zip(self.key[:4], key[:4]))
                and self.key[4:] == key[4:])

    def lookup(self, contigid, coverages, profiles, bin_tetramer_profiles,
               bin_coverage_profiles):
        key = (coverages, profiles, bin_tetramer_profiles, bin_coverage_profiles,
               len(bin_tetramer_profiles), len(profiles))
        with self.lock:
            if not self._same(key) or contigid not in self.result:
                order = list(bin_tetramer_profiles)
                cen_prof = np.stack([np.asarray(bin_tetramer_profiles[b], dtype=np.float64)
                                     for b in order])
                cen_cov = np.stack([np.asarray(bin_coverage_profiles[b], dtype=np.float64)
                                    for b in order])
                queries = sorted(c for c in profiles if c in coverages)
                if contigid not in profiles:
                    raise KeyError(contigid)
                arg, w = _backend.backend().score_centroids(
                    profiles, coverages, np.asarray(queries, dtype=np.int64), cen_prof, cen_cov)
                self.result = dict(zip(queries, zip(arg.tolist(), w.tolist())))
                self.key = key
            return self.result[contigid]


_centroids = _CentroidCache()
_backend._reset_hooks.append(_centroids.reset)


def assign_long(contigid, coverages, normalized_tetramer_profiles, bin_tetramer_profiles,
                bin_coverage_profiles):
    """``(contigid, index of the best bin, weight)`` or ``None`` when no bin scores (:308-345).
    The index is the position in ``bin_tetramer_profiles``' iteration order, first minimum."""
    best, weight = _centroids.lookup(contigid, coverages, normalized_tetramer_profiles,
                                     bin_tetramer_profiles, bin_coverage_profiles)
    if best < 0:
        return None
    return contigid, best, weight


def assign_to_bins(put_to_bins, bins, bin_of_contig, bin_markers, binned_contigs_with_markers,
                   contig_markers, contig_lengths):
    """Applies assign_long's results in the given order with the marker-gene rule (:348-391)."""
    for contig, contig_bin, bin_weight in put_to_bins:
        if contig_bin is None:
            continue
        has_mg = contig in contig_markers
        common_mgs = []
        if has_mg:
            common_mgs = list(set(bin_markers[contig_bin]).intersection(
                set(contig_markers[contig])))
        can_bin = False
        if contig not in bin_of_contig and bin_weight != MAX_WEIGHT:
            if len(common_mgs) == 0 or (len(common_mgs) <= 1 and contig_lengths[contig] > 100000):
                can_bin = True
        if can_bin:
            bins[contig_bin].append(contig)
            bin_of_contig[contig] = contig_bin
            if has_mg:
                binned_contigs_with_markers.append(contig)
                bin_markers[contig_bin] = list(set(bin_markers[contig_bin]
                                                   + contig_markers[contig]))
    return bins, bin_of_contig, bin_markers, binned_contigs_with_markers


def final_label_prop(bin_of_contig, bins, contig_markers, bin_markers,
                     binned_contigs_with_markers, smg_bin_counts, contig_lengths, min_length,
                     assembly_graph, normalized_tetramer_profiles, coverages, depth, weight):
    """Last propagation pass from every binned long contig (:394-513): a candidate is accepted
    when it is still unbinned and its score is not ``weight`` (MAX_WEIGHT in the pipeline);
    the marker-gene sets are computed but do not gate this pass."""
    seeds = [c for c in bin_of_contig if contig_lengths[c] >= min_length]

    def accept(to_bin, score, dist, common_mgs):
        return to_bin not in bin_of_contig and score != weight

    return _propagate(bin_of_contig, bins, contig_markers, bin_markers,
                      binned_contigs_with_markers, smg_bin_counts, seeds, contig_lengths,
                      min_length, assembly_graph, normalized_tetramer_profiles, coverages, depth,
                      accept)
