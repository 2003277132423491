#!/usr/bin/env python3
"""Generate tests/golden/pipeline.npz + pipeline.json from the LIVE reference (build container).

A synthetic assembly is pushed through the binning stages of the reference
(tests/pipeline_driver.py with the reference's own feature_utils / matching_utils /
label_prop_utils) and every intermediate state is stored next to the inputs, so the
GPU-box tests need neither the reference nor this script.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refload import ShimGraph, load_reference  # noqa: E402

import pipeline_driver  # noqa: E402
from metacoag_b200 import synth  # noqa: E402


def build_inputs(seed=6, n_genomes=8, n_contigs=160, n_samples=3, dup_markers=8):
    rng = np.random.default_rng(seed)
    lengths = np.exp(rng.uniform(np.log(400), np.log(45000), n_contigs)).astype(np.int64)
    asm = synth.make_assembly(seed=seed, n_genomes=n_genomes, n_contigs=n_contigs,
                              n_samples=n_samples, lengths=lengths, n_frac=2e-4, lower_frac=0.02)
    markers = {int(c): list(v) for c, v in asm.contig_markers.items()}
    # a few multi-copy marker genes: a long marker-free contig gets a gene that another contig
    # of its genome already carries, so the marker-gene conflict rules (and with them the
    # leftovers final_label_prop works on) are exercised
    for _ in range(dup_markers):
        g = int(rng.integers(0, n_genomes))
        own = [c for c in markers if asm.genome_of[c] == g]
        free = [int(c) for c in np.flatnonzero((asm.genome_of == g) & (asm.lengths >= 3000))
                if int(c) not in markers]
        if own and free:
            donor = own[int(rng.integers(0, len(own)))]
            gene = markers[donor][int(rng.integers(0, len(markers[donor])))]
            markers.setdefault(free[int(rng.integers(0, len(free)))], []).append(gene)
    return asm, markers


def graph_from_edges(n, edges):
    g = ShimGraph()
    g.add_vertices(n)
    g.add_edges([(int(a), int(b)) for a, b in edges])
    g.simplify()
    return g


def main():
    fu, mu, lpu = load_reference()
    asm, markers = build_inputs()
    n = asm.n_contigs
    min_length = 1000
    sequences = asm.sequences()
    contig_lengths = {i: int(asm.lengths[i]) for i in range(n)}
    coverages = {i: [float(x) for x in asm.coverage[i]] for i in range(n)
                 if asm.lengths[i] >= min_length}
    graph = graph_from_edges(n, asm.edges)
    with tempfile.TemporaryDirectory() as tmp:
        states = pipeline_driver.run_pipeline(fu, mu, lpu, sequences, coverages, contig_lengths,
                                              markers, graph, asm.n_samples, tmp + "/",
                                              min_length=min_length)
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), blob=asm.blob, offsets=asm.offsets,
                        coverage=asm.coverage, edges=asm.edges, lengths=asm.lengths)
    with open(os.path.join(HERE, "pipeline.json"), "w") as fh:
        json.dump({"markers": {str(k): v for k, v in markers.items()}, "min_length": min_length,
                   "n_samples": asm.n_samples,
                   "states": {k: (v if not isinstance(v, dict) else {str(a): b for a, b in v.items()})
                              for k, v in states.items()}}, fh)
    sizes = {k: sorted(len(m) for m in states[k].values())
             for k in ("initial", "match_contigs", "further_match_contigs", "assign_to_bins",
                       "final_label_prop")}
    print("pipeline golden: %d contigs, %d long, %d with markers" % (
        n, sum(1 for i in range(n) if asm.lengths[i] >= min_length), len(markers)))
    for k, v in sizes.items():
        print("  %-24s %d bins, sizes %s" % (k, len(v), v))
    print("  assign_long results:", len(states["assign_long"]))


if __name__ == "__main__":
    main()
