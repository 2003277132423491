"""Drop-in for ``metacoag_utils/matching_utils.py`` (MetaCoAG v1.1.5) on the B200 path.

The scoring primitives and the (contig x bin x member) scorer loops run in libmcg_b200.so; the
sequential control around them -- the networkx / SciPy assignment, the shortest-path gate,
the marker-gene rules -- stays on the host with the reference's semantics and iteration
orders.  Citations are ``matching_utils.py:line`` of the reference.
"""
import logging
import math
import sys

import networkx as nx
import numpy as np

from . import _backend
from ._lib import RULE_A

# Constants set from MaxBin 2.0 (:12-15)
MU_INTRA, SIGMA_INTRA = 0, 0.01037897 / 2
MU_INTER, SIGMA_INTER = 0.0676654, 0.03419337
VERY_SMALL_DOUBLE = 1e-10
MAX_WEIGHT = sys.float_info.max

logger = logging.getLogger("MetaCoaAG 1.1.5")


# ---- scoring primitives (:21-66), one value per call like the reference -------------------
def normpdf(x, mean, sd):
    return float(_backend.backend().normpdf([float(x)], float(mean), float(sd))[0])


def get_tetramer_distance(seq1, seq2):
    return float(_backend.backend().tetramer_distance(np.asarray(seq1, dtype=np.float64),
                                                      np.asarray(seq2, dtype=np.float64))[0])


def get_coverage_distance(cov1, cov2):
    return float(_backend.backend().tetramer_distance(np.asarray(cov1, dtype=np.float64),
                                                      np.asarray(cov2, dtype=np.float64))[0])


def get_comp_probability(tetramer_dist):
    """Raises ZeroDivisionError where the reference does (both Gaussians underflow)."""
    return float(_backend.backend().comp_probability([float(tetramer_dist)])[0])


def get_cov_probability(cov1, cov2):
    cov1 = np.asarray(cov1, dtype=np.float64)
    cov2 = np.asarray(cov2, dtype=np.float64)
    if (cov1 <= 0).any() or (cov2 <= 0).any():
        raise ValueError("math domain error")      # math.log(0) in the reference (:49, :53)
    return float(_backend.backend().cov_probability(cov1, cov2)[0])


# ---- rule A: (query x bin x member) mean of -log10 scores (:117-155, :368-398) ----------------
def _rule_a_weights(queries, bin_order, bins, normalized_tetramer_profiles, coverages):
    """weights[len(queries), len(bin_order)]: sum of the member scores / members, MAX_WEIGHT
    when the sum overflowed to inf."""
    members, seg = _backend.csr_bins(bins, bin_order)
    return _backend.backend().score_members(normalized_tetramer_profiles, coverages,
                                            np.asarray(queries, dtype=np.int64), members, seg,
                                            RULE_A)


def _path_len_sum(assembly_graph, contig, bin_members):
    """Sum over the bin of the vertex count of a shortest path (0 when unreachable),
    :191-197 and :285-292."""
    total = 0
    for member in bin_members:
        paths = assembly_graph.get_shortest_paths(contig, to=member)
        if len(paths) != 0:
            total += len(paths[0])
    return total


class _PathLengths:
    """The ``len(get_shortest_paths(v, to=w)[0])`` values one marker-gene iteration can ask
    for (:191-197, :285-292), from one batched multi-source BFS (mcg_path_lengths) over the
    CSR copy of the assembly graph: sources are the contigs to bin, targets every binned
    contig plus the contigs that may join a bin during the iteration.  A graph object without
    ``neighbors`` falls back to the reference's per-pair calls."""

    def __init__(self, assembly_graph, sources, bins, contig_lengths):
        self.graph = assembly_graph
        self.table = None
        if not hasattr(assembly_graph, "neighbors") or len(sources) == 0:
            return
        targets = []
        seen = set()
        for b in bins:
            for c in bins[b]:
                if c not in seen:
                    seen.add(c)
                    targets.append(c)
        for c in sources:
            if c not in seen:
                seen.add(c)
                targets.append(c)
        n_vertices = _backend.vertex_count(assembly_graph, contig_lengths)
        n_vertices = max([n_vertices] + [t + 1 for t in targets])
        indptr, indices = _backend.graphs.get(assembly_graph, n_vertices)
        self.row = {c: i for i, c in enumerate(sources)}
        self.col = {c: j for j, c in enumerate(targets)}
        self.table = _backend.backend().path_lengths(indptr, indices,
                                                     np.asarray(sources, dtype=np.int64),
                                                     np.asarray(targets, dtype=np.int64))

    def total(self, contig, bin_members):
        if self.table is None or contig not in self.row or any(m not in self.col for m in bin_members):
            return _path_len_sum(self.graph, contig, bin_members)
        row = self.table[self.row[contig]]
        return int(sum(int(row[self.col[m]]) for m in bin_members))


def match_contigs(smg_iteration, bins, n_bins, bin_of_contig, binned_contigs_with_markers,
                  bin_markers, contig_markers, contig_lengths, contig_names,
                  normalized_tetramer_profiles, coverages, assembly_graph, w_intra, w_inter,
                  d_limit):
    """Marker-gene seeded matching (:69-336).  For every marker-gene iteration after the first,
    the unbinned contigs carrying that gene are matched against the bins by a minimum-weight
    full bipartite matching on the rule-A weights; a match is accepted when the weight is at
    most ``w_intra``, the contig is close to the bin in the assembly graph and shares no
    marker gene with it; otherwise the best-supported rejected contig may open a new bin."""
    for i in range(len(smg_iteration)):
        logger.debug(f"Iteration {i}: {len(smg_iteration[i])} contig(s) with seed marker genes")
        if i == 0:
            continue

        already = set(binned_contigs_with_markers).intersection(set(smg_iteration[i]))
        to_bin = list(set(smg_iteration[i]) - already)
        logger.debug(f"{len(to_bin)} contig(s) to bin in the iteration")
        n_bins = len(bins)
        binned_count = 0

        if len(to_bin) != 0:
            bin_order = list(range(n_bins))
            weights = _rule_a_weights(to_bin, bin_order, bins, normalized_tetramer_profiles,
                                      coverages)

            # bins are represented by their first contig; contigs and bins share one id space
            graph = nx.Graph()
            seeds = []
            for b in bin_order:
                if bins[b][0] not in seeds:
                    seeds.append(bins[b][0])
            graph.add_nodes_from(to_bin, bipartite=0)
            graph.add_nodes_from(seeds, bipartite=1)
            edge_weights = {}
            for qi, contig in enumerate(to_bin):
                for b in bin_order:
                    w = float(weights[qi, b])
                    edge_weights[(bins[b][0], contig)] = w
                    graph.add_edge(bins[b][0], contig, weight=w)

            top_nodes = {n for n, d in graph.nodes(data=True) if d["bipartite"] == 0}
            if len(top_nodes) > 0:
                matching = nx.algorithms.bipartite.matching.minimum_weight_full_matching(
                    graph, top_nodes, "weight")
                paths = _PathLengths(assembly_graph, to_bin, bins, contig_lengths)

                rejected = {}     # contig -> (seed it was matched to, bin)
                for seed in matching:
                    if seed not in bin_of_contig:
                        continue
                    b = bin_of_contig[seed]
                    contig = matching[seed]
                    if contig in bins[b] or (seed, contig) not in edge_weights:
                        continue
                    avg_path_len = math.floor(paths.total(contig, bins[b]) / len(bins[b]))
                    accepted = False
                    if edge_weights[(seed, contig)] <= w_intra and avg_path_len <= d_limit:
                        if len(set(bin_markers[b]).intersection(set(contig_markers[contig]))) == 0:
                            accepted = True
                    if accepted:
                        bins[b].append(contig)
                        bin_of_contig[contig] = b
                        binned_contigs_with_markers.append(contig)
                        binned_count += 1
                        bin_markers[b] = list(set(bin_markers[b] + contig_markers[contig]))
                    else:
                        rejected[contig] = (seed, b)

                # among the rejected contigs that are far from their seed (weight > w_inter)
                # prefer the seed with the most marker genes, then the longest seed
                chosen, chosen_len, chosen_mgs = -1, -1, -1
                for contig in rejected:
                    seed = rejected[contig][0]
                    if edge_weights[(seed, contig)] > w_inter:
                        n_mgs = len(contig_markers[seed])
                        if chosen_mgs < n_mgs or (chosen_mgs == n_mgs
                                                  and chosen_len < contig_lengths[seed]):
                            chosen, chosen_mgs, chosen_len = contig, n_mgs, contig_lengths[seed]

                if chosen != -1:
                    old_bin = rejected[chosen][1]
                    path_len_sum = paths.total(chosen, bins[old_bin])
                    avg_path_len = path_len_sum / len(bins[old_bin])
                    if math.floor(avg_path_len) >= d_limit or path_len_sum == 0:
                        logger.debug("Creating new bin...")
                        logger.debug("New bin has contig " + str(chosen) + " to bin "
                                     + str(n_bins + 1) + " weight="
                                     + str(edge_weights[(rejected[chosen][0], chosen)]))
                        bins[n_bins] = [chosen]
                        bin_of_contig[chosen] = n_bins
                        binned_count += 1
                        bin_markers[n_bins] = contig_markers[chosen]
                        n_bins += 1
                        binned_contigs_with_markers.append(chosen)

        logger.debug(f"{binned_count} contig(s) binned in the iteration")

    return bins, bin_of_contig, n_bins, bin_markers, binned_contigs_with_markers


def further_match_contigs(unbinned_mg_contigs, min_length, bins, n_bins, bin_of_contig,
                          binned_contigs_with_markers, bin_markers, contig_markers,
                          normalized_tetramer_profiles, coverages, w_intra):
    """Leftover marker-gene contigs, longest first as the caller sorts them (:339-416): each
    goes to the marker-disjoint bin with the smallest rule-A weight if that is <= w_intra."""
    for contig in unbinned_mg_contigs:
        if contig[1] < min_length:
            continue
        contigid = contig[0]
        possible_bins = [b for b in bin_markers
                         if len(set(bin_markers[b]).intersection(set(contig_markers[contigid])))
                         == 0]
        if len(possible_bins) == 0:
            continue
        weights = _rule_a_weights([contigid], possible_bins, bins, normalized_tetramer_profiles,
                                  coverages)[0]
        best = int(np.argmin(weights))            # first index of the minimum, like min()
        if weights[best] <= w_intra:
            b = possible_bins[best]
            bins[b].append(contigid)
            bin_of_contig[contigid] = b
            binned_contigs_with_markers.append(contigid)
            bin_markers[b] = list(set(bin_markers[b] + contig_markers[contigid]))

    return bins, bin_of_contig, n_bins, bin_markers, binned_contigs_with_markers


# ---- bin x bin merge scoring (metacoag:1090-1143, inline in the reference's CLI) --------------
def bin_merge_candidates(bins, bin_markers, bin_seed_tetramer_profiles,
                         bin_seed_coverage_profiles, w_intra, n_mg, bin_mg_threshold):
    """The pair loop of the bin-merge step as one batched call.

    The reference scores, for every bin ``b`` and every bin ``pb`` whose marker genes do not
    overlap with ``b``'s, the seed centroids against each other (get_tetramer_distance,
    get_comp_probability, get_cov_probability, metacoag:1108-1121) and links ``b`` to the
    first ``pb`` with the smallest weight among those with weight <= w_intra (strict ``<``,
    metacoag:1138-1140); a bin with no candidate and fewer than ``n_mg * bin_mg_threshold``
    marker genes is marked for removal (metacoag:1146-1147).  Here all B x B weights come
    from one launch; the marker tests, the order of ``bin_markers`` and the strict-minimum
    rule stay as they are.

    Returns ``(edges, bins_to_rem)``: ``edges`` is the list of ``(b, min_pb)`` in the order
    the reference would call ``bins_graph.add_edge``.
    """
    order = list(bin_markers)
    pos = {b: i for i, b in enumerate(order)}
    missing = [b for b in bins if b not in pos]
    if missing:
        raise KeyError(missing[0])
    prof = np.array([np.asarray(bin_seed_tetramer_profiles[b], dtype=np.float64) for b in order])
    cov = np.array([np.asarray(bin_seed_coverage_profiles[b], dtype=np.float64) for b in order])
    prof = prof.reshape(len(order), -1)
    cov = cov.reshape(len(order), -1)
    if (cov <= 0).any():
        raise ValueError("math domain error")      # math.log(0), matching_utils.py:49
    pc, lp = _backend.backend().pair_matrix(prof, cov)
    marker_sets = {b: set(bin_markers[b]) for b in order}
    edges, bins_to_rem = [], []
    for b in bins:
        min_pb, min_pb_weight = -1, MAX_WEIGHT
        for pb in order:
            if marker_sets[pb] & marker_sets[b]:
                continue
            i, j = pos[b], pos[pb]
            if math.isnan(pc[i, j]):
                raise ZeroDivisionError("float division by zero")   # matching_utils.py:36-39
            log_prob = float(lp[i, j])
            # (the reference recomputes the same weight once more before accepting it)
            if log_prob <= w_intra and log_prob < min_pb_weight:
                min_pb_weight = log_prob
                min_pb = pb
        if min_pb != -1:
            edges.append((b, min_pb))
        elif len(bin_markers[b]) < n_mg * bin_mg_threshold:
            bins_to_rem.append(b)
    return edges, bins_to_rem
