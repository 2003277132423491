"""graph_utils.get_isolated / get_connected_components / get_non_isolated against
tests/golden/components.json (tests/golden/make_components.py, the reference's own functions):
same vertices per contig, same list-of-lists shape."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

from _refload import ShimGraph  # noqa: E402  (plain-Python graph; nothing from the reference)

from metacoag_b200 import _backend, graph_utils  # noqa: E402


def load(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "components.json")))
    graph = ShimGraph()
    graph.add_vertices(g["n"])
    graph.add_edges([tuple(e) for e in g["edges"]])
    graph.simplify()
    return g, graph


def check(golden_dir):
    g, graph = load(golden_dir)
    assert graph_utils.get_isolated(g["n"], graph) == g["isolated"]
    got = graph_utils.get_non_isolated(g["n"], graph, g["binned"], 4)
    assert isinstance(got, list) and len(got) == g["n"]
    assert [sorted(c) for c in got] == g["non_isolated"]
    for i in (0, g["binned"][0], g["n"] - 3):
        assert sorted(graph_utils.get_connected_components(i, graph, g["binned"])) == g["non_isolated"][i]
    # the shape the reference hands to label_prop: a contig id is never "in" a list of lists
    assert all(c not in got for c in g["binned"])


def test_components_host_logic(golden_dir):
    from oracle_backend import OracleBackend
    old = _backend.set_backend(OracleBackend())
    try:
        check(golden_dir)
    finally:
        _backend.set_backend(old)


@pytest.mark.gpu
def test_components_gpu(golden_dir):
    _backend.set_backend(None)
    check(golden_dir)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["chain", "random", "empty"])
def test_component_labels_vs_bfs(kind):
    """mcg_connected_components on graphs that need many rounds of plain label propagation:
    a 200 000-vertex chain (FastSV finishes it in O(log n) rounds), a sparse random graph with
    thousands of components, a graph without edges."""
    from oracle_backend import OracleBackend
    from metacoag_b200 import _lib
    rng = np.random.default_rng(11)
    if kind == "chain":
        n = 200000
        perm = rng.permutation(n)
        e = np.stack([perm[:-1], perm[1:]], 1)
    elif kind == "random":
        n = 50000
        e = rng.integers(0, n, (30000, 2))
    else:
        n = 1000
        e = np.zeros((0, 2), dtype=np.int64)
    e = e[e[:, 0] != e[:, 1]]
    e = np.unique(np.concatenate([e, e[:, ::-1]]), axis=0) if e.shape[0] else e
    indptr = np.zeros(n + 1, dtype=np.int64)
    if e.shape[0]:
        np.add.at(indptr, e[:, 0] + 1, 1)
    np.cumsum(indptr, out=indptr)
    indices = e[:, 1].astype(np.int32) if e.shape[0] else np.zeros(0, np.int32)
    ctx = _lib.Context(0)
    try:
        got = ctx.connected_components(indptr, indices)
    finally:
        ctx.close()
    want = OracleBackend().connected_components(indptr, indices)
    assert np.array_equal(got, want)
