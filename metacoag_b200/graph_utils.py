"""The two graph_utils functions of MetaCoAG v1.1.5 that sit next to the scoring path
(``metacoag_utils/graph_utils.py:392-443``): ``get_isolated`` and ``get_non_isolated``.

The graph parsers of that module (GFA / paths / assembly_info readers) are host-side file
handling and stay with the reference (out of scope).
"""
import numpy as np

from . import _backend


def get_isolated(node_count, assembly_graph):
    """Contigs without neighbours (:392-404)."""
    indptr, _ = _backend.graphs.get(assembly_graph, node_count)
    return [int(i) for i in np.flatnonzero(np.diff(indptr) == 0)]


def _components(node_count, assembly_graph):
    indptr, indices = _backend.graphs.get(assembly_graph, node_count)
    labels = _backend.backend().connected_components(indptr, indices)
    order = np.argsort(labels, kind="stable")
    starts = np.flatnonzero(np.diff(labels[order], prepend=-1))
    groups = np.split(order, starts[1:])
    return labels, {int(labels[g[0]]): [int(v) for v in g] for g in groups}


def get_connected_components(i, assembly_graph, binned_contigs):
    """The component of contig ``i`` when ``i`` is binned, else ``[]`` (:407-434).  Same
    vertices as the reference; the order is ascending (the reference's order comes from a
    ``list(set(...))`` round trip and nothing reads it)."""
    if i not in binned_contigs:
        return []
    n = _backend.vertex_count(assembly_graph, {i: 0})
    labels, groups = _components(n, assembly_graph)
    return groups[int(labels[i])]


def get_non_isolated(node_count, assembly_graph, binned_contigs, nthreads):
    """One list per contig (:437-443): the component of every binned contig, ``[]`` for the
    others -- the reference's list-of-lists shape, which makes ``contig in non_isolated``
    false for every contig id and is kept for that reason.  One
    labelling of the whole graph replaces the per-contig sweeps and the process pool;
    ``nthreads`` is accepted for the signature."""
    labels, groups = _components(node_count, assembly_graph)
    binned = binned_contigs if isinstance(binned_contigs, (set, frozenset, dict)) \
        else set(binned_contigs)
    return [groups[int(labels[i])] if i in binned else [] for i in range(node_count)]
