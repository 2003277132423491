"""Soak test of the split upload + waves: the same call many times with changing thread counts
and wave sizes must give bit-identical assignments and counts every time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from metacoag_b200 import _lib
rng = np.random.default_rng(0)
for kind, n, S, B in (("spades", 100000, 10, 200), ("megahit", 125000, 50, 500), ("flye", 1500, 10, 200)):
    old = bench.LENGTH_KIND; bench.LENGTH_KIND = kind
    tabs = bench.workload_tables(1000 + n, n, S, B); bench.LENGTH_KIND = old
    offsets = tabs["offsets"]; nb = int(offsets[-1])
    ctx = _lib.Context(0)
    blob = _lib.pinned_empty(nb)
    ctx.synth_bases(offsets, tabs["genome_of"], tabs["trans"], 7, out=blob)
    cov = tabs["coverage"]
    ctx.set_host_pack_threads(0)
    ctx.load_contigs(blob, offsets); ctx.load_coverage(cov); ctx.scan()
    cen_prof, cen_cov = ctx.bin_profiles(tabs["members"], tabs["seg"])
    arg0, w0, _ = ctx.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov)
    counts0 = ctx.get_profiles(want_counts=True)[0]
    bad = 0
    t0 = time.time()
    reps = 60 if kind != "flye" else 30
    for i in range(reps):
        threads = int(rng.choice([0, 1, 2, 5, 9, 14, 15, 16, 24]))
        os.environ["MCG_WAVE_DIV"] = str(int(rng.choice([2, 5, 12, 30])))
        ctx.set_host_pack_threads(threads)
        arg, w = ctx.featurise_score(blob, offsets, cov, cen_prof, cen_cov)
        ok = np.array_equal(arg, arg0) and np.array_equal(w, w0)
        if i % 10 == 0:
            ok = ok and np.array_equal(ctx.get_profiles(want_counts=True)[0], counts0)
        if not ok:
            bad += 1
            print("MISMATCH", kind, i, threads, os.environ["MCG_WAVE_DIV"], int((arg != arg0).sum()), int((w != w0).sum()), flush=True)
    print("%s: %d calls, %d mismatches, %.1f s" % (kind, reps, bad, time.time() - t0), flush=True)
    ctx.close()
