import os, sys, time, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
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
    for _ in range(2):
        ctx.featurise_score(blob, pin_off, pin_cov, cen_prof, cen_cov, argmin=h_arg, wmin=h_w)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); ctx.featurise_score(blob, pin_off, pin_cov, cen_prof, cen_cov, argmin=h_arg, wmin=h_w); ts.append(time.perf_counter() - t0)
    st = ctx.ingest_stats()
    print("%-28s e2e %.2f ms (min %.2f)  host %.0f MB" % (sys.argv[1], 1e3 * np.median(ts), 1e3 * min(ts), st["host_packed_bases"] / 1e6), flush=True)
else:
    for name, env in [("no waves", {"MCG_NO_WAVES": "1"}), ("waves /4", {"MCG_WAVE_DIV": "4"}), ("waves /8", {"MCG_WAVE_DIV": "8"}),
                      ("waves /12", {"MCG_WAVE_DIV": "12"}), ("waves /24", {"MCG_WAVE_DIV": "24"}), ("waves /12, 15 thr", {"MCG_WAVE_DIV": "12", "MCG_HOST_PACK_THREADS": "15"}),
                      ("waves /12, 14 thr", {"MCG_WAVE_DIV": "12", "MCG_HOST_PACK_THREADS": "14"})]:
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, name], env=e)
