import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from metacoag_b200 import _lib
tabs = bench.workload_tables(1001, 100000, 10, 200)
offsets = tabs["offsets"]; n_bases = int(offsets[-1])
ctx = _lib.Context(0)
blob = _lib.pinned_empty(n_bases)
ctx.synth_bases(offsets, tabs["genome_of"], tabs["trans"], 1001, out=blob)
pin_off = _lib.pinned_empty(offsets.nbytes).view(np.int64); pin_off[:] = offsets
for threads in [0, 4, 8, 12, 14, 15, 16, 20]:
    ctx.set_host_pack_threads(threads)
    ctx.load_contigs(blob, pin_off)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); ctx.load_contigs(blob, pin_off); ts.append(time.perf_counter() - t0)
    st = ctx.ingest_stats()
    print("threads %2d: load %.2f ms (min %.2f)  host %.0f MB  gpu %.0f MB  h2d %.0f MB" % (
        threads, 1e3 * np.median(ts), 1e3 * min(ts), st["host_packed_bases"] / 1e6, st["gpu_packed_bases"] / 1e6, st["h2d_bytes"] / 1e6), flush=True)
