#!/usr/bin/env python3
"""Headline benchmark: contigs/s for featurise + score (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): synthetic metaSPAdes-style contig set, 100 000 contigs
of 1-50 kb (log-uniform, ~1.28 Gbp), 10 coverage samples, 200 marker-seeded bins, per GPU.
One step = one pass of the hot path over that set: tetramer scan of the 2-bit store ->
normalise -> coverage terms -> rule-C scoring of every contig against every bin centroid
(label_prop_utils.assign_long) -> argmin.  At N > 1 every rank owns its own 100k-contig
shard (weak scaling), rank 0's centroids are broadcast and the assignments all-gathered
with NCCL inside the step.

`value`   inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`     the same work through mcg_featurise_score from pinned HOST buffers (ASCII bases,
          coverages, centroids in; assignments out), copies inside the timed region.
`roofline`/`kernels`  per-kernel device times from CUDA events on the launching stream.
`cpu_baseline`        the CPU oracle (C port of the reference arithmetic) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_CONTIGS = int(os.environ.get("MCG_BENCH_CONTIGS", 100000))
N_SAMPLES = int(os.environ.get("MCG_BENCH_SAMPLES", 10))
N_BINS = int(os.environ.get("MCG_BENCH_BINS", 200))
# length model: "spades" 1-50 kb log-uniform (configs[1], the default), "megahit" lognormal
# ~3 kb (configs[2]), "flye" 10 kb-5 Mb (configs[3]); the non-default ones are for checks only
LENGTH_KIND = os.environ.get("MCG_BENCH_KIND", "spades")
WORKLOAD = "synthetic %s-style %dk contigs, %d samples, %d bins per GPU" % (
    {"spades": "metaSPAdes (1-50 kb)", "megahit": "MEGAHIT (>= 1 kb, ~3 kb mean)",
     "flye": "Flye (10 kb-5 Mb)"}[LENGTH_KIND], N_CONTIGS // 1000, N_SAMPLES, N_BINS)
METRIC = "contigs/sec featurise+score"


def workload_tables(seed, n_contigs, n_samples, n_bins, shard=0):
    """Everything of the synthetic data set except the bases.  The
    community (genome composition and abundance tables) depends on ``seed`` only; ``shard`` > 0
    draws another set of contigs from the same community (the rows of another rank)."""
    from metacoag_b200 import synth
    rng = np.random.default_rng(seed)
    dinuc, abundance = synth.genome_tables(rng, n_bins, n_samples)
    if shard:
        rng = np.random.default_rng([seed, shard])
    trans = dinuc.reshape(n_bins, 4, 4)
    trans = (trans / trans.sum(axis=2, keepdims=True)).astype(np.float32)
    lengths = synth.sample_lengths(rng, n_contigs, LENGTH_KIND)
    genome_of = rng.integers(0, n_bins, n_contigs).astype(np.int32)
    offsets = np.zeros(n_contigs + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    noise = np.clip(rng.normal(1.0, 0.1, (n_contigs, n_samples)), 0.0, None)
    raw = abundance[genome_of] * noise
    raw[rng.random((n_contigs, n_samples)) < 0.01] = 0.0
    coverage = np.round(raw, 4)            # the "%.4f" of the abundance file
    coverage[coverage < 1e-4] = 1e-4       # feature_utils.py:152-153
    # marker-seeded bins: up to 20 contigs >= 3 kb of every genome
    members, seg = [], [0]
    order = np.argsort(genome_of, kind="stable")
    starts = np.searchsorted(genome_of[order], np.arange(n_bins + 1))
    for g in range(n_bins):
        ids = order[starts[g]:starts[g + 1]]
        ids = ids[lengths[ids] >= 3000][:20]
        if ids.size == 0:
            ids = order[starts[g]:starts[g] + 1] if starts[g + 1] > starts[g] else np.array([g])
        members.extend(int(i) for i in ids)
        seg.append(len(members))
    return dict(trans=trans, lengths=lengths, genome_of=genome_of, offsets=offsets,
                coverage=np.ascontiguousarray(coverage), members=np.array(members, np.int64),
                seg=np.array(seg, np.int64))


_REAL_STDOUT = None


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, lo, hi, nthreads):
    """The reference arithmetic for contigs lo:hi on the host cores (C port, pthreads)."""
    sub_off = offsets[lo:hi + 1] - offsets[lo]
    sub_blob = blob[int(offsets[lo]):int(offsets[hi])]
    _, freqs = oracle.count_kmers_batch(sub_blob, sub_off, want_counts=False, nthreads=nthreads)
    cov = coverage[lo:hi].copy()
    cov[cov < 1e-4] = 1e-4
    arg, w, _ = oracle.score_centroids(freqs, cov, cen_prof, cen_cov, nthreads=nthreads)
    return arg, w


def cpu_baseline(blob, offsets, coverage, cen_prof, cen_cov, budget_s):
    from oracle import oracle
    cores = oracle.max_threads()
    n = offsets.shape[0] - 1
    probe = min(n, 256)
    t0 = time.perf_counter()
    oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, 0, probe, 0)
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    sample = int(min(n, max(probe, rate * budget_s)))
    t0 = time.perf_counter()
    oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, 0, sample, 0)
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "contigs/s", "cores": cores, "kind": "port",
            "sample": "first %d contigs (%.1f Mbp) x %d bins, %d samples, %.1f s" % (
                sample, (offsets[sample] - offsets[0]) / 1e6, cen_prof.shape[0],
                coverage.shape[1], dt)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU arithmetic (oracle port; the reference itself is
    pure Python and cannot travel) on the same workload, bounded sample per step."""
    if rank != 0:
        return
    from metacoag_b200 import _lib
    from oracle import oracle
    tabs = workload_tables(1001, N_CONTIGS, N_SAMPLES, N_BINS)
    # the bases come from the same device generator as the CUDA arm (setup, not timed)
    ctx = _lib.Context(0)
    blob = ctx.synth_bases(tabs["offsets"], tabs["genome_of"], tabs["trans"], 1001)
    ctx.close()
    cores = oracle.max_threads()
    offsets, coverage = tabs["offsets"], tabs["coverage"]
    # bin centroids from the seed members, on the CPU as well (feature_utils.py:217-235)
    mem = tabs["members"]
    mem_off = np.zeros(mem.shape[0] + 1, dtype=np.int64)
    np.cumsum(tabs["lengths"][mem], out=mem_off[1:])
    mem_blob = np.concatenate([blob[offsets[i]:offsets[i + 1]] for i in mem])
    _, mem_prof = oracle.count_kmers_batch(mem_blob, mem_off, want_counts=False, nthreads=0)
    local_ids = np.arange(mem.shape[0], dtype=np.int64)
    cen_prof = oracle.segment_mean(mem_prof, local_ids, tabs["seg"])
    cen_cov = oracle.segment_mean(coverage[mem], local_ids, tabs["seg"])
    probe = 256
    t0 = time.perf_counter()
    oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, 0, probe, 0)
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    per_step_s = 120.0 / max(1, args.steps + args.warmup)
    sample = int(min(N_CONTIGS, max(probe, rate * per_step_s)))
    for _ in range(args.warmup):
        oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, 0, sample, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, 0, sample, 0)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    desc = "first %d contigs (%.1f Mbp) x %d bins per step, %d host threads" % (
        sample, offsets[sample] / 1e6, N_BINS, cores)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "contigs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": desc},
        "cpu_baseline": {"value": value, "unit": "contigs/s", "cores": cores, "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": "contigs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def main():
    # libraries (NCCL's version banner, for one) write to fd 1; the contract is ONE JSON line
    # on stdout, so everything else goes to stderr and the line is written to the saved fd
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from metacoag_b200 import _lib, parallel

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- setup (untimed) ---------------------------------------------------------------
    # every rank draws its own contigs from the SAME community (the bins rank 0 broadcasts are
    # the bins of all shards); until round 1h ranks > 0 drew their own community, so their contigs
    # matched no bin and their scorer ran 0.2 ms longer than rank 0's
    seed = 1001 + rank
    tabs = workload_tables(1001, N_CONTIGS, N_SAMPLES, N_BINS, shard=rank)
    offsets, coverage = tabs["offsets"], tabs["coverage"]
    n, n_bases = N_CONTIGS, int(offsets[-1])
    ctx = _lib.Context(local)
    # a non-default torch stream: the library, the NCCL collectives and the timing events
    # all run on it (handle 0 would mean "the context's own stream" to mcg_set_stream)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    blob = _lib.pinned_empty(n_bases)
    ctx.synth_bases(offsets, tabs["genome_of"], tabs["trans"], seed, out=blob)
    pin_cov = _lib.pinned_empty(coverage.nbytes).view(np.float64).reshape(coverage.shape)
    pin_cov[:] = coverage
    pin_off = _lib.pinned_empty(offsets.nbytes).view(np.int64)
    pin_off[:] = offsets

    ctx.load_contigs(blob, pin_off)
    ctx.load_coverage(pin_cov)
    ctx.scan()
    cen_prof, cen_cov = ctx.bin_profiles(tabs["members"], tabs["seg"])
    d_cen_prof = torch.from_numpy(cen_prof).to(dev)
    d_cen_cov = torch.from_numpy(cen_cov).to(dev)
    d_argmin = torch.empty(n, dtype=torch.int32, device=dev)
    d_wmin = torch.empty(n, dtype=torch.float64, device=dev)
    fp64_peak = ctx.fp64_peak()

    # N > 1: the two collectives of a step run on their own stream, and everything they touch is
    # double-buffered (centroid tables, assignment buffers, indexed by step parity).  The
    # centroid broadcast of step k overlaps the scorer of step k - 1 and the scan of step k; the
    # all-gather of step k overlaps the scan and the scorer of step k + 1.  A buffer is reused two
    # steps later, behind the event of its last reader.  Every step still does all of its work;
    # only the kernels stay on the critical path.
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    if world > 1:
        # the persistent scan fills every SM; a few SMs left free let the NCCL kernels start
        # under it instead of behind it
        ctx.set_reserved_sms(int(os.environ.get("MCG_BENCH_RESERVED_SMS", 0)))
    cen_bufs = [(d_cen_prof, d_cen_cov), (d_cen_prof.clone(), d_cen_cov.clone())]
    out_bufs = [(d_argmin, d_wmin), (torch.empty_like(d_argmin), torch.empty_like(d_wmin))]
    ev_bcast = [torch.cuda.Event(), torch.cuda.Event()]
    ev_scored = [torch.cuda.Event(), torch.cuda.Event()]
    ev_gathered = [torch.cuda.Event(), torch.cuda.Event()]
    gathered = [None]
    step_no = [0]

    exp = os.environ.get("MCG_BENCH_EXP", "")     # timing experiments only (not a bench line)

    # Peer-memory exchange (NVLink / NVSwitch stores + a signal-pad barrier, all on the compute
    # stream) when the GPUs can map each other's memory; the NCCL collectives otherwise.
    peer = None
    if world > 1 and os.environ.get("MCG_BENCH_EXCHANGE", "peer") == "peer":
        try:
            peer = parallel.PeerExchange(n, N_BINS, N_SAMPLES, dev)
        except Exception as err:   # no peer access on this box
            print("peer exchange unavailable (%s): NCCL collectives" % err, file=sys.stderr)
            peer = None

    def issue_broadcast(k):
        """Centroid tables of step k into buffer k & 1, on the collective stream."""
        if "nocoll" in exp:
            return
        q = k & 1
        with torch.cuda.stream(comm):
            comm.wait_event(ev_scored[q])       # the scorer of step k - 2 read these tables
            if peer is not None:
                if rank == 0:
                    peer.push_centroids(ctx, q, d_cen_prof, d_cen_cov, stream=comm)
                peer.barrier(channel=0)
            else:
                parallel.broadcast_centroids(cen_bufs[q][0], cen_bufs[q][1], src=0)
            ev_bcast[q].record(comm)

    def step():
        if world == 1:
            ctx.step_resident(d_argmin.data_ptr(), d_wmin.data_ptr())
            return d_argmin, d_wmin
        k = step_no[0]
        step_no[0] += 1
        par = k & 1
        cp, cc = peer.centroids(par) if peer is not None else cen_bufs[par]
        oa, ow = out_bufs[par]
        if k == 0:
            issue_broadcast(0)
        ctx.step_featurise()
        # the tables of the NEXT step are sent now: the exchange is queued in front of this step's
        # gather and runs under the scorer.  (Why overwriting parity (k + 1) & 1 is safe: the
        # gather of step k - 1, in front of it on the collective stream, ends with a rendezvous
        # every rank reaches after its scorer of step k - 1, the last reader of those tables.)
        issue_broadcast(k + 1)
        stream.wait_event(ev_bcast[par])
        stream.wait_event(ev_gathered[par])     # the exchange two steps back read these buffers
        if "nosetcen" not in exp:
            ctx.set_centroids_dev(cp.data_ptr(), cc.data_ptr(), N_BINS)
        ctx.step_score(oa.data_ptr(), ow.data_ptr())
        ev_scored[par].record(stream)
        if "nocoll" in exp:
            return oa, ow
        with torch.cuda.stream(comm):
            comm.wait_event(ev_scored[par])
            if peer is not None:
                peer.push_assignments(ctx, par, oa, ow, stream=comm)
                peer.barrier(channel=1)
                gathered[0] = peer.assignments(par)
            else:
                gathered[0] = parallel.gather_assignments(oa, ow, counts=[n] * world)
            ev_gathered[par].record(comm)
        return gathered[0]

    def drain():
        """The timed region ends when the last gather has landed."""
        if comm is not None:
            stream.wait_stream(comm)

    for _ in range(args.warmup):
        step()
    barrier()
    if peer is not None:
        # the gathered arrays the peer stores built == what the NCCL all-gather builds
        got_a, got_w = step()
        drain()
        torch.cuda.synchronize()
        last = out_bufs[(step_no[0] - 1) & 1]
        want_a, want_w = parallel.gather_assignments(last[0], last[1], counts=[n] * world)
        torch.cuda.synchronize()
        assert torch.equal(got_a, want_a) and torch.equal(got_w, want_w)
        barrier()

    # ---- per-kernel device times (CUDA events on the launching stream) ------------------
    ctx.profile(True)
    kernel_ms = {}
    reps = 5
    launches_per_step = 0
    for _ in range(reps):
        step()
        t = ctx.timings()
        launches_per_step = t.pop("launches")
        for k, v in t.items():
            kernel_ms[k] = kernel_ms.get(k, 0.0) + v / reps
    # The product path prunes the rule-C argmin (branch and bound on -log10 prob_comp,
    # include/mcg_b200.h mcg_set_prune): same argmin and weights, far fewer pairs scored.  The
    # FP64 roofline entries are computed from a pass with EVERY pair scored, so that their flops
    # are flops that ran; `value` and kernel_ms are the product path.
    ctx.set_prune(False)
    step()
    kernel_ms_all = {}
    for _ in range(reps):
        step()
        t = ctx.timings()
        t.pop("launches")
        for k, v in t.items():
            kernel_ms_all[k] = kernel_ms_all.get(k, 0.0) + v / reps
    res_all = step()
    drain()
    torch.cuda.synchronize()
    all_pairs_arg = res_all[0].clone()
    all_pairs_w = res_all[1].clone()
    ctx.set_prune(True)
    res = step()
    drain()
    torch.cuda.synchronize()
    assert torch.equal(all_pairs_arg, res[0]) and torch.equal(all_pairs_w, res[1])
    prune_stats = ctx.prune_stats()
    ctx.profile(False)
    barrier()

    # ---- timed region ------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    t_host = time.perf_counter()
    for _ in range(args.steps):
        step()
    host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / args.steps
    drain()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    value = n * world / (ms_per_step * 1e-3)

    # ---- end to end through the plug-in call, host buffers --------------------------------------
    h_arg = _lib.pinned_empty(4 * n).view(np.int32)
    h_w = _lib.pinned_empty(8 * n).view(np.float64)
    # the ingest splits the ASCII blob between the copy engine and host pack threads; the
    # ranks of one box share its cores
    ncpu = len(os.sched_getaffinity(0))
    pack_threads = int(os.environ.get("MCG_HOST_PACK_THREADS", max(0, ncpu // world - 1)))

    def time_e2e(threads, steps):
        ctx.set_host_pack_threads(threads)
        ctx.featurise_score(blob, pin_off, pin_cov, cen_prof, cen_cov, argmin=h_arg, wmin=h_w)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.featurise_score(blob, pin_off, pin_cov, cen_prof, cen_cov, argmin=h_arg, wmin=h_w)
            if world > 1:
                parallel.gather_assignments(torch.from_numpy(h_arg).to(dev, non_blocking=True),
                                            torch.from_numpy(h_w).to(dev, non_blocking=True),
                                            counts=[n] * world)
                torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / steps
        t_e2e = torch.tensor([sec], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        return float(t_e2e.item()), ctx.ingest_stats()

    # the copy engine alone (all bases cross the link as ASCII), for comparison
    e2e_dma_s, ing_dma = time_e2e(0, max(2, args.e2e_steps // 2))
    e2e_s, ingest = time_e2e(pack_threads, args.e2e_steps)
    ctx.profile(True)
    ctx.featurise_score(blob, pin_off, pin_cov, cen_prof, cen_cov, argmin=h_arg, wmin=h_w)
    e2e_ms = ctx.timings()
    e2e_ms.pop("launches")
    ctx.profile(False)
    side = offsets.nbytes + 16 * n + coverage.nbytes + cen_prof.nbytes + cen_cov.nbytes
    h2d = ingest["h2d_bytes"] + side
    d2h = 12 * n + 16

    # device results of the last resident step agree with the plug-in call
    ctx.step_resident(d_argmin.data_ptr(), d_wmin.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(d_argmin.cpu().numpy(), h_arg)

    # ---- roofline -----------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650"
    lens = tabs["lengths"]
    n_seg = int(np.sum((np.maximum(lens - 3, 0) + 255) // 256))
    scan_bytes = float(np.sum((lens + 3) // 4) + n * 136 * 4 + 4 * n_seg)
    pairs = float(n) * N_BINS
    score_flops = pairs * (431 + 14 * N_SAMPLES)
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
    # capture of this workload (profiles/traffic.json; null when this workload was not captured)
    try:
        captured = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        captured = {}
    wl_key = "%s-%d-%d-%d" % (LENGTH_KIND, N_CONTIGS, N_SAMPLES, N_BINS)

    def roof(kernel, bound, amount, ms, peak, unit, source, **extra):
        scale = 1e9 if unit == "GB/s" else 1e12
        traffic = captured.get(wl_key, {}).get(kernel.split()[0])
        r = {"kernel": kernel, "bound": bound, "achieved": amount / (ms * 1e-3) / scale,
             "peak": peak, "unit": unit, "peak_source": source, "traffic": traffic, "ms": ms}
        r["frac"] = r["achieved"] / peak
        r.update(extra)
        return r

    fp64_src = "DFMA microbenchmark on this device (mcg_fp64_peak); DMMA measures the same"
    scan_roof = roof("scan_kernel", "hbm", scan_bytes, kernel_ms["scan"], hbm_peak, "GB/s", hbm_src,
                     algorithmic_bytes=scan_bytes,
                     gbases_per_s=n_bases / (kernel_ms["scan"] * 1e-3) / 1e9,
                     note="bound by the L1TEX load/store data pipe, not HBM (DESIGN.md 4.1): two "
                          "shared-memory wavefronts per 32 windows is the floor of a "
                          "byte-counter histogram")
    # The scan is not HBM-bound: its counters live in shared memory and every window costs a
    # byte load and a byte store there.  What it fills is the L1TEX load/store data pipe, one
    # 128-byte wavefront per clock per SM (ncu: l1tex__data_pipe_lsu_wavefronts 88 % of peak);
    # the wavefront count per launch comes from the same ncu capture as `traffic`.
    wf = captured.get(wl_key, {}).get("scan_kernel_lsu_wavefronts")
    if wf:
        sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
        peak_wf = 148 * sm_clock
        scan_roof["binding"] = {"resource": "L1TEX LSU data pipe (shared + global wavefronts/s)",
                                "achieved": wf / (kernel_ms["scan"] * 1e-3), "peak": peak_wf,
                                "frac": wf / (kernel_ms["scan"] * 1e-3) / peak_wf,
                                "wavefronts_per_launch": wf,
                                "floor_wavefronts": 2.0 * n_bases / 32.0}
    kernels = [scan_roof]
    if kernel_ms.get("distance", 0.0) > 0.0:
        # SURVEY 8d splits F_pair = 431 + 14 S into 136*3 for the distance and the rest.  The
        # distance / probability timers hold ONE launch: with a chunked d^2 scratch (many rows x
        # many bins) that is the last chunk, whose pair count the library reports.
        chunk_pairs = float(prune_stats["pairs"]) if prune_stats.get("pairs") else pairs
        dist_flops = chunk_pairs * 136 * 3
        prob_flops = chunk_pairs * (431 - 136 * 3 + 14 * N_SAMPLES)
        dmma = chunk_pairs * 136 * 2 / (kernel_ms_all["distance"] * 1e-3) / 1e12
        kernels.append(roof("dist2_kernel (DMMA distance GEMM)", "fp64", dist_flops,
                            kernel_ms_all["distance"], fp64_peak, "TFLOP/s", fp64_src,
                            algorithmic_flops=dist_flops, pairs_per_launch=chunk_pairs,
                            dmma_tflops=dmma, dmma_frac=dmma / fp64_peak,
                            note="algorithmic flops count sub, mul, add per element (SURVEY 8d: "
                                 "3 x 136 per pair); the kernel computes |u|^2 + |v|^2 - 2 u.v, "
                                 "2 x 136 per pair on the FP64 tensor path, so `frac` can pass 1: "
                                 "dmma_frac is the executed fraction of the unit's peak"))
        kernels.append(roof("prob_kernel (Poisson / Gaussian terms, argmin)", "fp64", prob_flops,
                            kernel_ms_all["probability"], fp64_peak, "TFLOP/s", fp64_src,
                            algorithmic_flops=prob_flops,
                            pairs_per_launch=chunk_pairs,
                            gpairs_per_s=chunk_pairs / (kernel_ms_all["probability"] * 1e-3) / 1e9,
                            note="every pair scored (mcg_set_prune 0); the product path prunes: "
                                 "%.4f ms" % kernel_ms["probability"]))
    score_roof = roof("rule-C scoring (distance + probability kernels)", "fp64", score_flops,
                      kernel_ms_all["score"], fp64_peak, "TFLOP/s", fp64_src,
                      algorithmic_flops=score_flops,
                      gpairs_per_s=pairs / (kernel_ms_all["score"] * 1e-3) / 1e9,
                      note="every pair scored (mcg_set_prune 0); the product path prunes: "
                           "%.4f ms" % kernel_ms["score"])
    kernels.append(score_roof)
    dominant = max(kernels[:-1], key=lambda r: r["ms"])

    line = {
        "metric": METRIC, "value": value, "unit": "contigs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "contigs_per_gpu": n, "bases_per_gpu": n_bases,
                   "samples": N_SAMPLES, "bins": N_BINS,
                   "parallelism": "single GPU" if world == 1 else
                                  ("row-sharded x%d, " % world) +
                                  ("bin tables and assignments exchanged by peer-memory stores "
                                   "(NVLink) + signal barrier" if peer is not None
                                   else "NCCL broadcast + all-gather"),
                   "l2": "inputs larger than L2 (%.0f MB packed bases + %.0f MB profiles per step)"
                         % (n_bases / 4e6, n * 136 * 8 / 1e6)},
        "clocks": clocks,
        "e2e": {"value": n * world / e2e_s, "unit": "contigs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                "steps": args.e2e_steps, "device_ms": e2e_ms,
                "ingest": dict(ingest, note="ASCII blob split between the copy engine (+ pack "
                               "kernel) and host pack threads, meeting in the middle; contig "
                               "ranges are scanned and scored as they arrive, so the kernel "
                               "times sit inside device_ms.h2d"),
                "copy_engine_only": {"value": n * world / e2e_dma_s, "ms_per_step": e2e_dma_s * 1e3,
                                     "h2d_bytes_per_step": ing_dma["h2d_bytes"] + side}},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": dominant,
        "roofline_kernels": kernels,
        "kernel_ms": kernel_ms,
        "kernel_ms_every_pair_scored": kernel_ms_all,
        "host_enqueue_ms_per_step": host_enqueue_ms,
    }
    line["config"]["rule_c"] = ("branch and bound (bound_kernel + candidate kernels; exact: argmin "
                                "and weights checked bit-equal to the all-pairs pass in this run)")
    line["rule_c_candidates"] = dict(prune_stats, note="pairs scored after the bound / all pairs "
                                     "of the last rule-C chunk; used = the candidate list held "
                                     "them all (else prob_kernel scores every pair)")
    if N_SAMPLES > 30:
        # contigs the guarded rule-C path re-scored in the reference's operation order
        line["config"]["redo_contigs_per_step"] = int(ctx.last_redo_count())
    if rank == 0 and world == 1:
        line["cpu_baseline"] = cpu_baseline(blob, offsets, coverage, cen_prof, cen_cov,
                                            args.cpu_budget)
    if rank == 0:
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
