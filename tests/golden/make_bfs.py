#!/usr/bin/env python3
"""Generate tests/golden/bfs.json from the LIVE reference's run_bfs_long (build container).

The traversal of label_prop_utils.py:24-100 on real graphs and depth limits, with the scoring
switched off (``smg_bin_counts`` of 0 members per bin: the member loop of :58-80 does not run).
The reference returns a *set* of (node, labelled node, bin, depth, score); stored per start node:
the sorted (labelled node, bin, depth) triples.  Graphs: a synthetic assembly graph (per-genome
chains + chords + cross edges), a denser random graph where the depth-overwrite behaviour
(:93-98) changes many reported depths, and the triangle of SURVEY.md section 7.4.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refload import ShimGraph, load_reference  # noqa: E402
from metacoag_b200 import synth  # noqa: E402


def cases():
    asm = synth.make_assembly(seed=21, n_genomes=5, n_contigs=400, n_samples=1,
                              length_kind="mixed", with_markers=False)
    rng = np.random.default_rng(2)
    labels = np.full(400, -1, dtype=np.int64)
    lab = rng.choice(400, 60, replace=False)
    labels[lab] = rng.integers(0, 5, 60)
    starts = np.concatenate([np.flatnonzero(labels < 0)[:120], lab[:5]])
    yield "assembly", 400, asm.edges, labels, starts, (1, 3, 10)
    n = 300
    edges = np.concatenate([np.stack([np.arange(n - 1), np.arange(1, n)], 1),
                            rng.integers(0, n, (2 * n, 2))])
    labels = np.full(n, -1, dtype=np.int64)
    labels[rng.choice(n, 25, replace=False)] = rng.integers(0, 4, 25)
    yield "dense", n, edges, labels, np.flatnonzero(labels < 0)[:80], (2, 4, 6)
    yield "triangle", 3, np.array([[0, 1], [0, 2], [1, 2]]), np.array([-1, -1, 0]), np.array([0, 1]), (1, 10)


def main():
    _, _, lpu = load_reference()
    out = {}
    for name, n, edges, labels, starts, limits in cases():
        g = ShimGraph()
        g.add_vertices(n)
        g.add_edges([(int(a), int(b)) for a, b in edges if a != b])
        g.simplify()
        bin_of = {int(v): int(labels[v]) for v in np.flatnonzero(labels >= 0)}
        n_bins = int(labels.max()) + 1
        bins = {b: [] for b in range(n_bins)}
        expect = {}
        for limit in limits:
            per_start = []
            for s in starts:
                found = lpu.run_bfs_long(int(s), limit, list(bin_of), bin_of, bins, [0] * n_bins, g,
                                         {}, {})
                assert all(t[0] == int(s) for t in found)
                per_start.append(sorted([t[1], t[2], t[3]] for t in found))
            expect[str(limit)] = per_start
        out[name] = {"n": n, "edges": [[int(a), int(b)] for a, b in edges],
                     "labels": [int(x) for x in labels], "starts": [int(s) for s in starts],
                     "expect": expect}
        print(name, {k: sum(len(x) for x in v) for k, v in expect.items()})
    with open(os.path.join(HERE, "bfs.json"), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))


if __name__ == "__main__":
    main()
