"""Can the tetramer scan (L1TEX data pipe) and the rule-C scoring (FP64) of two different contig
sets share the SMs?  Two contexts on one device, one thread each; MCG_SCAN_WARPS shrinks the
scan's CTA so that a scoring CTA fits next to it."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from metacoag_b200 import _lib

tabs = bench.workload_tables(1001, 100000, 10, 200)
offsets = tabs["offsets"]
a = _lib.Context(0)
blob = a.synth_bases(offsets, tabs["genome_of"], tabs["trans"], 1001)
a.load_contigs(blob, offsets); a.load_coverage(tabs["coverage"]); a.scan()
cen_prof, cen_cov = a.bin_profiles(tabs["members"], tabs["seg"])
prof = a.get_profiles()
b = _lib.Context(0)
b.set_profiles(prof); b.load_coverage(tabs["coverage"])
b.score_centroids(cen_prof=cen_prof, cen_cov=cen_cov)
K = 40

def run_scan():
    for _ in range(K):
        a.scan()
    a.synchronize()

def run_score():
    for _ in range(K):
        b.score_centroids()
    b.synchronize()

for f in (run_scan, run_score):
    f()
t0 = time.perf_counter(); run_scan(); ts = time.perf_counter() - t0
t0 = time.perf_counter(); run_score(); tc = time.perf_counter() - t0
t0 = time.perf_counter()
th = [threading.Thread(target=run_scan), threading.Thread(target=run_score)]
[t.start() for t in th]; [t.join() for t in th]
tb = time.perf_counter() - t0
print("MCG_SCAN_WARPS=%s  scan %.3f ms  score %.3f ms  sum %.3f  both at once %.3f ms per pair of calls" % (
    os.environ.get("MCG_SCAN_WARPS", "27"), 1e3 * ts / K, 1e3 * tc / K, 1e3 * (ts + tc) / K, 1e3 * tb / K), flush=True)
