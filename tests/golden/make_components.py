#!/usr/bin/env python3
"""Generate tests/golden/components.json from the LIVE reference (build container):
graph_utils.get_isolated / get_connected_components / get_non_isolated
(metacoag_utils/graph_utils.py:392-443, unmodified) on the golden pipeline's assembly graph
plus a few extra isolated vertices and small components."""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refload import ShimGraph, load_reference  # noqa: E402


def main():
    load_reference()
    gu = importlib.import_module("metacoag_utils.graph_utils")
    assert gu.__file__.startswith("/root/reference")
    data = np.load(os.path.join(HERE, "pipeline.npz"))
    meta = json.load(open(os.path.join(HERE, "pipeline.json")))
    n0 = int(data["lengths"].shape[0])
    n = n0 + 12
    edges = [(int(a), int(b)) for a, b in data["edges"]]
    edges += [(n0, n0 + 1), (n0 + 1, n0 + 2), (n0 + 4, n0 + 5), (n0 + 7, n0 + 7)]   # + isolated ones
    g = ShimGraph()
    g.add_vertices(n)
    g.add_edges(edges)
    g.simplify()
    binned = sorted(int(c) for members in meta["states"]["match_contigs"].values() for c in members)
    binned += [n0 + 1, n0 + 9]
    isolated = gu.get_isolated(n, g)
    per_contig = [gu.get_connected_components(i, g, binned) for i in range(n)]
    pooled = gu.get_non_isolated(n, g, binned, 2)
    assert [sorted(x) for x in pooled] == [sorted(x) for x in per_contig]
    out = {"n": n, "edges": edges, "binned": binned, "isolated": [int(x) for x in isolated],
           "non_isolated": [sorted(int(v) for v in comp) for comp in pooled]}
    with open(os.path.join(HERE, "components.json"), "w") as fh:
        json.dump(out, fh)
    print("components golden: %d vertices, %d binned, %d isolated, %d non-empty lists, largest %d" % (
        n, len(binned), len(isolated), sum(1 for c in pooled if c), max(len(c) for c in pooled)))


if __name__ == "__main__":
    main()
