#!/usr/bin/env python3
"""Generate tests/golden/merge.json from the LIVE reference (build container).

The bin-merge pair loop is inline in the reference's CLI (metacoag:1090-1147); it is restated
here around the reference's own matching_utils functions, imported unmodified from
/root/reference, on seed bins that contain near-duplicates (two bins cut from one genome),
marker-overlapping pairs and small bins, so that edges, ties and removals all occur.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refload import load_reference  # noqa: E402

from metacoag_b200 import synth  # noqa: E402

MAX_WEIGHT = sys.float_info.max


def reference_merge_loop(mu, bins, bin_markers, seed_prof, seed_cov, w_intra, n_mg, thr):
    import math
    edges, to_rem = [], []
    for b in bins:
        best, best_w = -1, MAX_WEIGHT
        for pb in bin_markers:
            if set(bin_markers[pb]) & set(bin_markers[b]):
                continue
            d = mu.get_tetramer_distance(seed_prof[b], seed_prof[pb])
            pc = mu.get_comp_probability(d)
            pv = mu.get_cov_probability(seed_cov[pb], seed_cov[b])
            w = -(math.log(pc, 10) + math.log(pv, 10)) if pc * pv > 0.0 else MAX_WEIGHT
            if w <= w_intra and w < best_w:
                best, best_w = pb, w
        if best != -1:
            edges.append([int(b), int(best)])
        elif len(bin_markers[b]) < n_mg * thr:
            to_rem.append(int(b))
    return edges, to_rem


def main():
    fu, mu, _ = load_reference()
    n_genomes, n_samples = 14, 4
    asm = synth.make_assembly(seed=77, n_genomes=n_genomes, n_contigs=700, n_samples=n_samples,
                              length_kind="mixed", with_graph=False, with_markers=False)
    rng = np.random.default_rng(78)
    lengths = {i: int(asm.lengths[i]) for i in range(asm.n_contigs)}
    with tempfile.TemporaryDirectory() as tmp:
        profiles = fu.get_tetramer_profiles(tmp + "/", asm.sequences(), "c.fasta", lengths, 1000, 2)
    coverages = {i: [max(float(x), 1e-4) for x in asm.coverage[i]] for i in profiles}
    genes = ["g%03d" % i for i in range(108)]
    bins, bin_markers = {}, {}
    for g in range(n_genomes):
        ids = [int(c) for c in np.flatnonzero(asm.genome_of == g) if int(c) in profiles]
        rng.shuffle(ids)
        halves = [ids[:len(ids) // 2][:8], ids[len(ids) // 2:][:8]] if g % 3 != 2 else [ids[:10]]
        pool = list(rng.permutation(genes))
        for h, members in enumerate(halves):
            if not members:
                continue
            b = len(bins)
            bins[b] = members
            if g % 3 == 0:      # the two halves carry disjoint markers: merge candidates
                mine = pool[:50] if h == 0 else pool[50:95]
            elif g % 3 == 1:    # overlapping markers: never scored
                mine = pool[:60] if h == 0 else pool[40:100]
            else:               # a small bin on its own: removal candidate
                mine = pool[:int(rng.integers(5, 60))]
            bin_markers[b] = [str(x) for x in mine]
    seed_prof, seed_cov = fu.get_bin_profiles(bins, coverages, profiles)
    w_intra = -(np.log10(0.1)) * (n_samples + 1)
    out = {"bins": {str(b): v for b, v in bins.items()},
           "bin_markers": {str(b): v for b, v in bin_markers.items()},
           "seed_prof": {str(b): [float(x) for x in seed_prof[b]] for b in seed_prof},
           "seed_cov": {str(b): [float(x) for x in seed_cov[b]] for b in seed_cov},
           "n_mg": 108, "bin_mg_threshold": 0.33333, "cases": []}
    for w in (float(w_intra), 2.0 * float(w_intra), 60.0, MAX_WEIGHT):
        edges, to_rem = reference_merge_loop(mu, bins, bin_markers, seed_prof, seed_cov, w, 108,
                                             0.33333)
        out["cases"].append({"w_intra": w, "edges": edges, "bins_to_rem": to_rem})
        print("w_intra %.4g: %d edges, %d removals" % (w, len(edges), len(to_rem)))
    with open(os.path.join(HERE, "merge.json"), "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
