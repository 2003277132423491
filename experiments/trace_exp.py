import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from metacoag_b200 import _lib
tabs = bench.workload_tables(1001, 100000, 10, 200)
offsets = tabs["offsets"]; n_bases = int(offsets[-1]); n = 100000
ctx = _lib.Context(0)
blob = _lib.pinned_empty(n_bases)
ctx.synth_bases(offsets, tabs["genome_of"], tabs["trans"], 1001, out=blob)
pin_off = _lib.pinned_empty(offsets.nbytes).view(np.int64); pin_off[:] = offsets
cov = tabs["coverage"]
pin_cov = _lib.pinned_empty(cov.nbytes).view(np.float64).reshape(cov.shape); pin_cov[:] = cov
ctx.load_contigs(blob, pin_off); ctx.load_coverage(pin_cov); ctx.scan()
cen_prof, cen_cov = ctx.bin_profiles(tabs["members"], tabs["seg"])
h_arg = _lib.pinned_empty(4 * n).view(np.int32); h_w = _lib.pinned_empty(8 * n).view(np.float64)
for i in range(4):
    print("---- call", i, file=sys.stderr, flush=True)
    t0 = time.perf_counter(); ctx.featurise_score(blob, pin_off, pin_cov, cen_prof, cen_cov, argmin=h_arg, wmin=h_w)
    print("python wall %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)
