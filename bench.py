#!/usr/bin/env python3
"""Headline benchmark: contigs/s for featurise + score (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline workload (BASELINE.json configs[1]): synthetic metaSPAdes-style contig set, 100 000
contigs of 1-50 kb (log-uniform, ~1.28 Gbp), 10 coverage samples, 200 marker-seeded bins, per GPU.
One step = one pass of the hot path over that set: tetramer scan of the 2-bit store ->
normalise -> coverage terms -> rule-C scoring of every contig against every bin centroid
(label_prop_utils.assign_long) -> argmin.  At N > 1 every rank owns its own 100k-contig
shard (weak scaling), rank 0's centroids are sent to every rank and the assignments gathered
inside the step (peer-memory stores over NVLink, NCCL as the fallback).

`value`   inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`     the same work through the plug-in call from pinned HOST buffers (ASCII bases, coverages,
          centroids in; assignments out), copies inside the timed region.  N = 1: one context
          (mcg_featurise_score).  N > 1: ONE process -- rank 0 -- drives all N GPUs through the
          device pool (mcg_pool_featurise_score), which is how the single-process metacoag CLI
          uses a box; the other ranks wait on a host-side barrier.  `e2e.packed` is the same
          call with the contigs already in the 2-bit store a FASTA reader writes at parse time.
`configs` the other BASELINE.json configurations in the same run, each with its own device
          time, e2e, CPU baseline, roofline and an oracle check of seeded contigs:
            cfg3  1 M MEGAHIT-style contigs, 50 samples, 500 bins, STRONG-scaled over the N GPUs
            cfg5  5 M contigs, 100 samples, 1000 bins, strong-scaled, plus one round of
                  label-propagation candidates (bounded walks + rule B) over a 5 M-vertex graph
            cfg2-strains  the headline shape on a community of near-identical genome pairs:
                  the regime in which the rule-C branch and bound has the least to prune
`roofline`/`kernels`  per-kernel device times from CUDA events on the launching stream.
`cpu_baseline`        the CPU oracle (C port of the reference arithmetic) on a bounded sample,
                      and its results compared with the GPU's (`parity`).
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
KIND_TEXT = {"spades": "metaSPAdes (1-50 kb)", "megahit": "MEGAHIT (>= 1 kb, ~3 kb mean)",
             "flye": "Flye (10 kb-5 Mb)"}
WORKLOAD = "synthetic %s-style %dk contigs, %d samples, %d bins per GPU" % (
    KIND_TEXT[LENGTH_KIND], N_CONTIGS // 1000, N_SAMPLES, N_BINS)
METRIC = "contigs/sec featurise+score"
# the other BASELINE.json configurations, measured next to the headline (MCG_BENCH_CONFIGS=""
# turns them off, "cfg3" picks one; MCG_BENCH_SCALE shrinks them for a quick look)
CONFIGS = {
    "cfg3": dict(n=1000000, samples=50, bins=500, kind="megahit", seed=1003, steps=10,
                 e2e_steps=3, text="BASELINE configs[2]: synthetic MEGAHIT-style 1M contigs "
                 "(>= 1 kb), 50 samples, 500 bins, strong-scaled over the GPUs"),
    "cfg5": dict(n=5000000, samples=100, bins=1000, kind="megahit", seed=1005, steps=4,
                 e2e_steps=2, label_prop=True, text="BASELINE configs[4]: CAMI-scale synthetic "
                 "5M contigs, 100 samples, 1000 bins, strong-scaled over the GPUs, with one "
                 "round of label-propagation candidates"),
    "cfg2-strains": dict(n=100000, samples=10, bins=200, kind="spades", seed=1001, steps=20,
                         e2e_steps=0, strains=(0.15, 0.1), per_gpu=True,
                         text="configs[1] shape on a community of 100 near-identical genome "
                         "pairs (composition within d ~ 0.015, abundance within 10 %): the "
                         "hard case for the rule-C branch and bound"),
    "cfg2-strains-tight": dict(n=100000, samples=10, bins=200, kind="spades", seed=1001, steps=20,
                               e2e_steps=0, strains=(0.04, 0.02), per_gpu=True,
                               text="the same with pairs four times closer (composition within "
                               "d ~ 0.004, abundance within 2 %)"),
}


def community_tables(seed, n_bins, n_samples, strains=False):
    """Genome composition (first-order Markov transitions) and abundance tables of the
    synthetic community.  ``strains``: genome 2k+1 is a near copy of
    genome 2k (dinucleotide table perturbed by ~15 %, abundance within 10 %)."""
    from metacoag_b200 import synth
    rng = np.random.default_rng(seed)
    dinuc, abundance = synth.genome_tables(rng, n_bins, n_samples)
    if strains:
        eps, spread = strains if isinstance(strains, tuple) else (0.15, 0.1)
        r2 = np.random.default_rng([seed, 77])
        for g in range(1, n_bins, 2):
            dinuc[g] = dinuc[g - 1] * np.exp(eps * r2.normal(size=16))
            dinuc[g] /= dinuc[g].sum()
            abundance[g] = abundance[g - 1] * r2.uniform(1.0 - spread, 1.0 + spread, n_samples)
    trans = dinuc.reshape(n_bins, 4, 4)
    trans = (trans / trans.sum(axis=2, keepdims=True)).astype(np.float32)
    return rng, trans, abundance


def workload_tables(seed, n_contigs, n_samples, n_bins, shard=0, kind=None, strains=False,
                    coverage=True):
    """Everything of the synthetic data set except the bases.  The
    community (genome composition and abundance tables) depends on ``seed`` only; ``shard`` > 0
    draws another set of contigs from the same community (the rows of another rank).
    ``coverage=False`` leaves the [n, S] table to ``device_coverage`` (the 5M x 100 table of
    configs[4] takes numpy a quarter of a minute and 16 GB)."""
    from metacoag_b200 import synth
    rng, trans, abundance = community_tables(seed, n_bins, n_samples, strains)
    if shard:
        rng = np.random.default_rng([seed, shard])
    lengths = synth.sample_lengths(rng, n_contigs, kind or LENGTH_KIND)
    genome_of = rng.integers(0, n_bins, n_contigs).astype(np.int32)
    offsets = np.zeros(n_contigs + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    cov = None
    if coverage:
        noise = np.clip(rng.normal(1.0, 0.1, (n_contigs, n_samples)), 0.0, None)
        raw = abundance[genome_of] * noise
        raw[rng.random((n_contigs, n_samples)) < 0.01] = 0.0
        cov = np.round(raw, 4)            # the "%.4f" of the abundance file
        cov[cov < 1e-4] = 1e-4            # feature_utils.py:152-153
        cov = np.ascontiguousarray(cov)
    # marker-seeded bins: up to 20 contigs >= 3 kb of every genome
    order = np.argsort(genome_of, kind="stable")
    starts = np.searchsorted(genome_of[order], np.arange(n_bins + 1))
    members, seg = [], [0]
    for g in range(n_bins):
        ids = order[starts[g]:starts[g + 1]]
        ids = ids[lengths[ids] >= 3000][:20]
        if ids.size == 0:
            ids = order[starts[g]:starts[g] + 1] if starts[g + 1] > starts[g] else np.array([g])
        members.extend(int(i) for i in ids)
        seg.append(len(members))
    return dict(trans=trans, abundance=abundance, lengths=lengths, genome_of=genome_of,
                offsets=offsets, coverage=cov, members=np.array(members, np.int64),
                seg=np.array(seg, np.int64))


def device_coverage(tabs, n_samples, seed, shard, dev, out):
    """The coverage table of ``workload_tables`` drawn on the GPU (same recipe: abundance x
    N(1, 0.1^2) clipped at 0, 1 % zeros, four decimals, 1e-4 clamp) into the host array ``out``."""
    import torch
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) * 1000003 + int(shard))
    ab = torch.from_numpy(tabs["abundance"]).to(dev)
    g = torch.from_numpy(tabs["genome_of"].astype(np.int64)).to(dev)
    n = g.shape[0]
    view = torch.from_numpy(out)
    for lo in range(0, n, 1 << 20):
        hi = min(n, lo + (1 << 20))
        noise = torch.randn((hi - lo, n_samples), generator=gen, device=dev, dtype=torch.float64)
        raw = ab[g[lo:hi]] * torch.clamp(1.0 + 0.1 * noise, min=0.0)
        zero = torch.rand((hi - lo, n_samples), generator=gen, device=dev) < 0.01
        raw = torch.where(zero, torch.zeros_like(raw), raw)
        cov = torch.round(raw * 1e4) / 1e4
        cov = torch.clamp(cov, min=1e-4)
        view[lo:hi].copy_(cov)
    torch.cuda.synchronize()


_REAL_STDOUT = None


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def log(*a):
    print("[bench %.1fs]" % (time.perf_counter() - _T0), *a, file=sys.stderr, flush=True)


_T0 = time.perf_counter()


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


# ---- the CPU side: oracle port of the reference arithmetic (test infrastructure) ----------------
def oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, ids, nthreads):
    """The reference arithmetic for the contigs ``ids`` (any order) on the host cores (C port,
    pthreads): tetramer profiles, clamp, rule C against every centroid."""
    ids = np.asarray(ids, dtype=np.int64)
    lens = offsets[ids + 1] - offsets[ids]
    sub_off = np.zeros(ids.shape[0] + 1, dtype=np.int64)
    np.cumsum(lens, out=sub_off[1:])
    if ids.shape[0] and np.array_equal(ids, np.arange(ids[0], ids[0] + ids.shape[0])):
        sub_blob = blob[int(offsets[ids[0]]):int(offsets[ids[-1] + 1])]
    else:
        sub_blob = np.concatenate([blob[offsets[i]:offsets[i + 1]] for i in ids]) if ids.shape[0] \
            else np.zeros(0, np.uint8)
    _, freqs = oracle.count_kmers_batch(sub_blob, sub_off, want_counts=False, nthreads=nthreads)
    cov = coverage[ids].copy()
    cov[cov < 1e-4] = 1e-4
    arg, w, _ = oracle.score_centroids(freqs, cov, cen_prof, cen_cov, nthreads=nthreads)
    return arg, w


def compare(arg, w, want_arg, want_w):
    """`parity` entry: GPU results against the oracle's on the same contigs."""
    top = want_w == sys.float_info.max
    both = ~top & (w != sys.float_info.max)
    rel = np.abs(w[both] - want_w[both]) / np.abs(want_w[both]) if both.any() else np.zeros(1)
    return {"checked": int(want_arg.shape[0]), "argmin_equal": bool(np.array_equal(arg, want_arg)),
            "none_equal": bool(np.array_equal(w[top], want_w[top])),
            "max_rel_err": float(rel.max()), "tolerance": 1e-9,
            "ok": bool(np.array_equal(arg, want_arg) and np.array_equal(w[top], want_w[top])
                       and rel.max() <= 1e-9)}


def cpu_baseline(blob, offsets, coverage, cen_prof, cen_cov, budget_s, gpu_arg, gpu_w, seed=0):
    """(cpu_baseline, parity): the oracle on the first contigs of the set for ~budget_s seconds
    (the timed sample) plus 256 seeded contigs from anywhere in the set; every result is
    compared with what the GPU returned for the same contig."""
    from oracle import oracle
    cores = oracle.max_threads()
    n = offsets.shape[0] - 1
    probe = min(n, 256)
    t0 = time.perf_counter()
    oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, np.arange(probe), 0)
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    sample = int(min(n, max(probe, rate * budget_s)))
    t0 = time.perf_counter()
    arg, w = oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, np.arange(sample), 0)
    dt = time.perf_counter() - t0
    extra = np.sort(np.random.default_rng(seed).choice(n, min(n, 256), replace=False))
    arg2, w2 = oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, extra, 0)
    ids = np.concatenate([np.arange(sample), extra])
    parity = compare(gpu_arg[ids], gpu_w[ids], np.concatenate([arg, arg2]), np.concatenate([w, w2]))
    base = {"value": sample / dt, "unit": "contigs/s", "cores": cores, "kind": "port",
            "sample": "first %d contigs (%.1f Mbp) x %d bins, %d samples, %.1f s" % (
                sample, (offsets[sample] - offsets[0]) / 1e6, cen_prof.shape[0],
                coverage.shape[1], dt)}
    return base, parity


def _reference_package():
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "metacoag_utils")):
            return cand
    return None


def python_reference_worker(path):
    """Child process of python_reference_baseline: a fresh interpreter (the reference forks a
    multiprocessing.Pool; the parent holds CUDA contexts, NCCL threads and pinned memory)."""
    import concurrent.futures
    import importlib
    import tempfile
    pkg = _reference_package()
    sys.path.append(os.path.join(ROOT, "tests", "cli_shims"))    # Bio stand-in (site-packages first)
    sys.path.insert(0, pkg)
    fu = importlib.import_module("metacoag_utils.feature_utils")
    lpu = importlib.import_module("metacoag_utils.label_prop_utils")
    data = np.load(path)
    blob, offsets, coverage = data["blob"], data["offsets"], data["coverage"]
    cen_prof, cen_cov, n_score = data["cen_prof"], data["cen_cov"], int(data["n_score"])
    cores = os.cpu_count() or 1
    seqs = [blob[offsets[i]:offsets[i + 1]].tobytes().decode("latin-1")
            for i in range(offsets.shape[0] - 1)]
    lengths = {k: len(q) for k, q in enumerate(seqs)}
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        profiles = fu.get_tetramer_profiles(tmp + "/", seqs, "contigs.fa", lengths, 0, cores)
        scan_s = time.perf_counter() - t0
    covs = {k: [max(float(v), 1e-4) for v in coverage[k]] for k in range(n_score)}
    bin_prof = {b: cen_prof[b] for b in range(cen_prof.shape[0])}
    bin_cov = {b: cen_cov[b] for b in range(cen_cov.shape[0])}
    out = [None] * n_score

    def work(k):
        out[k] = lpu.assign_long(k, covs, profiles, bin_prof, bin_cov)
    t0 = time.perf_counter()
    with concurrent.futures.ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(work, range(n_score)))
    score_s = time.perf_counter() - t0
    emit({"scan_s": scan_s, "score_s": score_s, "cores": cores, "pkg": pkg,
          "arg": [-1 if r is None else int(r[1]) for r in out],
          "w": [repr(sys.float_info.max if r is None else float(r[2])) for r in out]})


def python_reference_baseline(blob, offsets, coverage, cen_prof, cen_cov, gpu_arg, gpu_w,
                              n_scan=1500, n_score=160, seed=0):
    """The reference's OWN Python code on this box's host cores, on seeded subsamples:
    feature_utils.get_tetramer_profiles with ``nthreads = cpu_count`` (its multiprocessing.Pool,
    feature_utils.py:73-119) and label_prop_utils.assign_long through the ThreadPoolExecutor of
    metacoag:959-999, as shipped.  The unmodified modules come from /root/reference (build
    container) or from baseline/_ref (the reference pip-installed once, unmodified, next to the
    repo: it travels to the GPU box); Bio is the stand-in of tests/cli_shims where BioPython is
    missing.  Results of the subsample are compared with the GPU's.  None when no copy exists."""
    import tempfile
    pkg = _reference_package()
    if pkg is None:
        return None
    n = offsets.shape[0] - 1
    rng = np.random.default_rng([seed, 31])
    ids = np.sort(rng.choice(n, min(n, n_scan), replace=False))
    n_score = min(n_score, ids.shape[0])
    lens = (offsets[ids + 1] - offsets[ids]).astype(np.int64)
    sub_off = np.zeros(ids.shape[0] + 1, dtype=np.int64)
    np.cumsum(lens, out=sub_off[1:])
    sub_blob = np.concatenate([blob[offsets[i]:offsets[i + 1]] for i in ids])
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "sample.npz")
        np.savez(path, blob=sub_blob, offsets=sub_off, coverage=coverage[ids[:n_score]],
                 cen_prof=cen_prof, cen_cov=cen_cov, n_score=n_score)
        env = {k: v for k, v in os.environ.items()
               if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "CUDA_VISIBLE_DEVICES")}
        env["CUDA_VISIBLE_DEVICES"] = ""
        try:
            proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--python-ref-worker",
                                   path], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                  text=True, timeout=600)
            res = json.loads(proc.stdout.strip().splitlines()[-1])
        except Exception as err:   # a broken copy is not a reason to lose the bench line
            return {"unavailable": "%s: %s" % (type(err).__name__, err)}
    scan_s, score_s, cores = res["scan_s"], res["score_s"], res["cores"]
    ref_arg = np.array(res["arg"], dtype=np.int32)
    ref_w = np.array([float(x) for x in res["w"]])
    bases = int(sub_off[-1])
    sub = ids[:n_score]
    mean_len = float(np.mean(np.diff(offsets)))
    per_contig = scan_s / max(1, bases) * mean_len + score_s / max(1, n_score)
    return {"value": 1.0 / per_contig, "unit": "contigs/s", "kind": "reference",
            "extrapolated": True, "cores": cores,
            "source": "metacoag_utils from %s" % ("/root/reference" if pkg == "/root/reference"
                                                  else "baseline/_ref (pip install of the reference)"),
            "scan": {"contigs": int(ids.shape[0]), "mbp": bases / 1e6, "seconds": scan_s,
                     "mbp_per_s": bases / 1e6 / scan_s,
                     "how": "feature_utils.get_tetramer_profiles, multiprocessing.Pool(%d)" % cores},
            "score": {"contigs": int(n_score), "bins": int(cen_prof.shape[0]),
                      "samples": int(coverage.shape[1]), "seconds": score_s,
                      "us_per_pair": score_s * 1e6 / max(1, n_score * cen_prof.shape[0]),
                      "how": "label_prop_utils.assign_long via ThreadPoolExecutor(%d), as metacoag:959-999" % cores},
            "sample": "seeded subsample; per-contig cost = scan seconds per base x mean contig "
                      "length + scoring seconds per contig, linear extrapolation",
            "parity_vs_gpu": compare(gpu_arg[sub], gpu_w[sub], ref_arg, ref_w)}


def config_dict(world, peer):
    """`config` of the headline line; the reference arm prints the same dict (the driver
    compares them key by key)."""
    return {"workload": WORKLOAD, "contigs_per_gpu": N_CONTIGS, "samples": N_SAMPLES,
            "bins": N_BINS}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU arithmetic (oracle port; the reference itself is
    pure Python and cannot travel) on the same workload, a bounded sample per step, all host
    threads.  This process loads oracle/ only: the tables are numpy, the bases come from the
    oracle's restatement of the library's generator (same bytes, tests/test_resident.py)."""
    if rank != 0:
        return
    from oracle import oracle
    tabs = workload_tables(1001, N_CONTIGS, N_SAMPLES, N_BINS)
    cores = oracle.max_threads()
    offsets, coverage = tabs["offsets"], tabs["coverage"]
    # bin centroids from the seed members, on the CPU as well (feature_utils.py:217-235); the
    # members can be anywhere in the set, so their bases are drawn too
    mem = tabs["members"]
    probe = 256
    need = int(max(probe, mem.max() + 1))
    blob = oracle.synth_bases(offsets, tabs["genome_of"], tabs["trans"], 1001, n_contigs=need)
    mem_off = np.zeros(mem.shape[0] + 1, dtype=np.int64)
    np.cumsum(tabs["lengths"][mem], out=mem_off[1:])
    mem_blob = np.concatenate([blob[offsets[i]:offsets[i + 1]] for i in mem])
    _, mem_prof = oracle.count_kmers_batch(mem_blob, mem_off, want_counts=False, nthreads=0)
    local_ids = np.arange(mem.shape[0], dtype=np.int64)
    cen_prof = oracle.segment_mean(mem_prof, local_ids, tabs["seg"])
    cen_cov = oracle.segment_mean(coverage[mem], local_ids, tabs["seg"])
    t0 = time.perf_counter()
    oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, np.arange(probe), 0)
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    per_step_s = 120.0 / max(1, args.steps + args.warmup)
    sample = int(min(N_CONTIGS, max(probe, rate * per_step_s)))
    if sample > need:
        blob = oracle.synth_bases(offsets, tabs["genome_of"], tabs["trans"], 1001, n_contigs=sample)
    ids = np.arange(sample)
    for _ in range(args.warmup):
        oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, ids, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_pass(oracle, blob, offsets, coverage, cen_prof, cen_cov, ids, 0)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    desc = "first %d contigs (%.1f Mbp) x %d bins per step, %d host threads" % (
        sample, offsets[sample] / 1e6, N_BINS, cores)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "contigs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus, None), "reference_sample": desc,
        "cpu_baseline": {"value": value, "unit": "contigs/s", "cores": cores, "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": "contigs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ---- roofline helpers --------------------------------------------------------------------------------
def load_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        peaks = {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 (B200_PROFILING.md)"
    return hbm, src


def captured_traffic():
    """dram bytes per launch from the committed `ncu --set full` capture (profiles/traffic.json),
    valid only for the kernels whose source has not changed since the capture (the file names the
    commit); null otherwise."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        return {}


def scan_algorithmic_bytes(lengths):
    """ceil(L / 4) packed input + 136 x 4 counts out + 4 B per 256-window segment
    (owner table); the f64 profile write belongs to the normalise kernel."""
    n_seg = int(np.sum((np.maximum(lengths - 3, 0) + 255) // 256))
    return float(np.sum((lengths + 3) // 4) + lengths.shape[0] * 136 * 4 + 4 * n_seg)


def roof(kernel, bound, amount, ms, peak, unit, source, traffic=None, **extra):
    scale = 1e9 if unit == "GB/s" else 1e12
    r = {"kernel": kernel, "bound": bound, "achieved": amount / (ms * 1e-3) / scale,
         "peak": peak, "unit": unit, "peak_source": source, "traffic": traffic, "ms": ms}
    r["frac"] = r["achieved"] / peak
    r.update(extra)
    return r


class Runtime:
    """torch / torch.distributed plumbing of one rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.host_pg = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            # a host-side barrier: while rank 0 drives every GPU through the pool the other
            # ranks must not sit in an NCCL kernel on theirs
            self.host_pg = dist.new_group(backend="gloo")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.host_pg)

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


# ---- whole data set on the host of the rank that drives the pool ----------------------------------
class HostSet:
    """All shards of a workload in pinned host memory (ASCII blob, offsets, coverage) plus the
    2-bit store a FASTA reader would have written at parse time."""

    def __init__(self, pool, shards, n_samples, seed):
        """shards: list of (tabs, shard index); shard r is generated on device r of the pool."""
        from metacoag_b200 import _lib
        ctxs = pool.contexts()
        sizes = [int(t["offsets"][-1]) for t, _ in shards]
        rows = [t["offsets"].shape[0] - 1 for t, _ in shards]
        self.n = int(sum(rows))
        self.n_bases = int(sum(sizes))
        self.blob = _lib.pinned_empty(self.n_bases)
        self.offsets = _lib.pinned_empty(8 * (self.n + 1)).view(np.int64)
        self.coverage = _lib.pinned_empty(8 * self.n * n_samples).view(np.float64).reshape(
            self.n, n_samples)
        self.genome_of = np.concatenate([t["genome_of"] for t, _ in shards])
        self.lengths = np.concatenate([t["lengths"] for t, _ in shards])
        self.row0 = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
        base0 = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        self.offsets[0] = 0
        for r, (t, _) in enumerate(shards):
            self.offsets[self.row0[r] + 1:self.row0[r + 1] + 1] = t["offsets"][1:] + base0[r]

        def draw(r):
            t, shard = shards[r]
            c = ctxs[r % len(ctxs)]
            c.synth_bases(t["offsets"], t["genome_of"], t["trans"], seed + shard,
                          out=self.blob[base0[r]:base0[r + 1]])
        threads = [threading.Thread(target=draw, args=(r,)) for r in range(len(shards))]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        import torch
        for r, (t, shard) in enumerate(shards):
            dst = self.coverage[self.row0[r]:self.row0[r + 1]]
            if t["coverage"] is not None:
                dst[:] = t["coverage"]
            else:
                device_coverage(t, n_samples, seed, shard, torch.device("cuda", ctxs[r % len(ctxs)].device),
                                dst)
        t0 = time.perf_counter()
        self.packed = _lib.pack_2bit_host(self.blob, self.offsets, threads=0, pinned=True)
        self.pack_s = time.perf_counter() - t0


def time_pool_call(rt, fn, steps):
    """Wall clock of `steps` calls of fn (after one untimed call)."""
    fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) / steps


def pool_e2e(pool, hs, cen_prof, cen_cov, steps):
    """e2e through the device pool from the host set: ASCII in (the reference's data type) and
    2-bit in (packed at parse time).  Returns (entry, argmin, wmin)."""
    from metacoag_b200 import _lib
    n = hs.n
    h_arg = _lib.pinned_empty(4 * n).view(np.int32)
    h_w = _lib.pinned_empty(8 * n).view(np.float64)
    side = hs.offsets.nbytes + hs.coverage.nbytes + (cen_prof.nbytes + cen_cov.nbytes) * pool.size
    ctxs = pool.contexts()

    def ascii_call():
        pool.featurise_score(hs.blob, hs.offsets, hs.coverage, cen_prof, cen_cov, argmin=h_arg,
                             wmin=h_w)

    def packed_call():
        pool.featurise_score(hs.packed, hs.offsets, hs.coverage, cen_prof, cen_cov,
                             encoding=_lib.ENC_2BIT, argmin=h_arg, wmin=h_w)

    sec = time_pool_call(None, ascii_call, steps)
    ing = [c.ingest_stats() for c in ctxs]
    arg, w = h_arg.copy(), h_w.copy()
    sec_p = time_pool_call(None, packed_call, steps)
    same = bool(np.array_equal(arg, h_arg) and np.array_equal(w, h_w))
    h2d = sum(i["h2d_bytes"] for i in ing) + side
    entry = {"value": n / sec, "unit": "contigs/s", "ms_per_step": sec * 1e3, "steps": steps,
             "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(12 * n),
             "api": "mcg_pool_featurise_score, one process, %d device(s), ASCII bases in" % pool.size,
             "host_pack_threads": [i["host_threads"] for i in ing],
             "host_packed_fraction": float(sum(i["host_packed_bases"] for i in ing) / max(1, hs.n_bases)),
             "packed": {"value": n / sec_p, "unit": "contigs/s", "ms_per_step": sec_p * 1e3,
                        "h2d_bytes_per_step": int(hs.packed.nbytes + side),
                        "d2h_bytes_per_step": int(12 * n), "same_results_as_ascii": same,
                        "note": "contigs already in the 2-bit store (written once by the FASTA "
                                "reader at parse time, %.0f ms for this set with all host "
                                "threads); a quarter of the bytes on the link" % (hs.pack_s * 1e3)}}
    return entry, arg, w


# ---- one round of label-propagation candidates at scale (configs[4]) ----------------------------
def label_prop_round(pool, hs, n_bins, seed, depth=10):
    """final_label_prop's batch work (label_prop_utils.py:419-446) for one labelling snapshot, at
    the size of the set: 60 % of the contigs carry the label of their genome, every other contig
    of at least 1 kb is a start node; bounded walks to labelled contigs over the CSR assembly
    graph (mcg_pool_bfs: graph and labels resident on every device, starts split), rule B for
    every (start, reached bin) item against the bin's 20 seed members (mcg_pool_score_member_items),
    best (depth, score) per start on the host, accepted labels patched on every device
    (mcg_pool_labels_set).  Features are the staged tables of the drop-in (replicated)."""
    from metacoag_b200 import _lib, synth
    from oracle import oracle
    n = hs.n
    t0 = time.perf_counter()
    indptr, indices = synth.graph_csr_large(seed, hs.genome_of)
    graph_s = time.perf_counter() - t0
    rng = np.random.default_rng([seed, 9])
    labelled = rng.random(n) < 0.6
    labels = np.where(labelled, hs.genome_of, -1).astype(np.int32)
    starts = np.flatnonzero(~labelled & (hs.lengths >= 1000)).astype(np.int64)
    # seed members: the first <= 20 labelled contigs >= 3 kb of every bin
    cand = np.flatnonzero(labelled & (hs.lengths >= 3000))
    order = cand[np.argsort(hs.genome_of[cand], kind="stable")]
    first = np.searchsorted(hs.genome_of[order], np.arange(n_bins + 1))
    seg = np.zeros(n_bins + 1, dtype=np.int64)
    members = []
    for b in range(n_bins):
        ids = order[first[b]:min(first[b + 1], first[b] + 20)]
        members.append(ids)
        seg[b + 1] = seg[b] + ids.shape[0]
    members = np.concatenate(members).astype(np.int64)

    # staging (untimed, once per stage): features replicated, graph + labels resident
    t0 = time.perf_counter()
    _, freqs = pool.tetramer_scan(hs.packed, hs.offsets, encoding=_lib.ENC_2BIT, want_counts=False)
    scan_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    pool.set_features(prof=freqs, cov=hs.coverage)
    pool.graph_load(indptr, indices)
    stage_s = time.perf_counter() - t0

    import torch
    dev = torch.device("cuda", pool.contexts()[0].device)
    d_starts = torch.from_numpy(starts).to(dev)

    def one_round():
        t = [time.perf_counter()]
        pool.labels_load(labels)
        t.append(time.perf_counter())
        hit_start, hit_count, hits = pool.bfs_resident(starts, depth)
        t.append(time.perf_counter())
        # group the hits into (start, bin) items: sorting / unique on the GPU (torch, plumbing),
        # 10^7 keys.  Hits of a start are contiguous; the starts' ranges come in completion order.
        d_hits = torch.from_numpy(hits).to(dev)
        d_count = torch.from_numpy(hit_count).to(dev).long()
        d_first = torch.from_numpy(hit_start).to(dev)
        owner = torch.repeat_interleave(torch.arange(starts.shape[0], device=dev), d_count)
        rank_in = torch.arange(owner.shape[0], device=dev) - torch.repeat_interleave(
            torch.cumsum(d_count, 0) - d_count, d_count)
        h = d_hits[d_first[owner] + rank_in].long()
        key = owner * n_bins + h[:, 1]
        uniq, inv = torch.unique(key, return_inverse=True)
        item_query = d_starts[uniq // n_bins].cpu().numpy()
        item_bin = (uniq % n_bins).to(torch.int32).cpu().numpy()
        t.append(time.perf_counter())
        score = pool.score_member_items(item_query, item_bin, members, seg, _lib.RULE_B)
        t.append(time.perf_counter())
        # best candidate per start: smallest (depth, score) (DataWrap, label_prop_utils.py:16-21):
        # two stable sorts, then the first hit of every start
        s_hit = torch.from_numpy(score).to(dev)[inv]
        o1 = torch.argsort(s_hit, stable=True)
        o2 = o1[torch.argsort(h[o1, 2], stable=True)]
        o3 = o2[torch.argsort(owner[o2], stable=True)]
        so = owner[o3]
        lead = torch.ones_like(so, dtype=torch.bool)
        lead[1:] = so[1:] != so[:-1]
        best = o3[lead]
        ok = s_hit[best] != sys.float_info.max          # final_label_prop's test (:468)
        acc_ids = d_starts[owner[best][ok]].cpu().numpy()
        acc_bins = h[best][ok][:, 1].to(torch.int32).cpu().numpy()
        t.append(time.perf_counter())
        pool.labels_set(acc_ids, acc_bins)
        t.append(time.perf_counter())
        return t, (hit_start, hit_count, hits, item_query, item_bin, score, acc_ids)

    one_round()
    t, (hit_start, hit_count, hits, item_query, item_bin, score, acc_ids) = one_round()
    ms = [(b - a) * 1e3 for a, b in zip(t[:-1], t[1:])]

    # parity + CPU side on seeded starts: the oracle's walk (pinned to the reference's
    # run_bfs_long) and its rule B
    pick = np.sort(np.random.default_rng([seed, 10]).choice(starts.shape[0], min(256, starts.shape[0]),
                                                            replace=False))
    t0 = time.perf_counter()
    want_walks = oracle.bfs_long_batch(indptr, indices, labels, starts[pick], depth)
    cpu_walk_us = (time.perf_counter() - t0) * 1e6 / max(1, pick.shape[0])
    walks_equal = all(
        [tuple(r) for r in hits[hit_start[i]:hit_start[i] + hit_count[i]].tolist()] == want
        for i, want in zip(pick, want_walks))
    sel = np.flatnonzero(np.isin(item_query, starts[pick]))[:512]
    rel, n_top, cpu_item_s = 0.0, 0, 0.0
    for j in sel:
        q, b = int(item_query[j]), int(item_bin[j])
        ids = np.concatenate([[q], members[seg[b]:seg[b + 1]]])
        sub = oracle_rows(oracle, hs, ids)          # profiles: not part of the timed scoring
        t0 = time.perf_counter()
        want = oracle.score_members(sub[0], sub[1], np.array([0]), np.arange(1, ids.shape[0]),
                                    np.array([0, ids.shape[0] - 1]), _lib.RULE_B, nthreads=1)[0, 0]
        cpu_item_s += time.perf_counter() - t0
        if want == sys.float_info.max or score[j] == sys.float_info.max:
            n_top += int(want != score[j])
        else:
            rel = max(rel, abs(score[j] - want) / abs(want))
    cpu_item_us = cpu_item_s * 1e6 / max(1, sel.shape[0])
    return {
        "vertices": int(n), "edges": int(indices.shape[0] // 2), "labelled": int(labelled.sum()),
        "starts": int(starts.shape[0]), "depth": depth, "hits": int(hits.shape[0]),
        "items": int(item_query.shape[0]), "seed_members": int(members.shape[0]),
        "accepted": int(acc_ids.shape[0]),
        "ms": {"labels_upload": ms[0], "walks": ms[1], "host_group_items": ms[2], "rule_b": ms[3],
               "host_best_per_start": ms[4], "labels_patch": ms[5], "round": sum(ms)},
        "starts_per_s": starts.shape[0] / (sum(ms) * 1e-3),
        "staging_untimed_s": {"synthetic_graph": graph_s, "scan_to_host": scan_s,
                              "features_and_graph_to_every_device": stage_s},
        "parity": {"walks_checked": int(pick.shape[0]), "walks_equal": bool(walks_equal),
                   "items_checked": int(sel.shape[0]), "max_rel_err": float(rel),
                   "max_weight_mismatches": int(n_top), "tolerance": 1e-9,
                   "ok": bool(walks_equal and rel <= 1e-9 and n_top == 0)},
        "cpu_port": {"walk_us_per_start": cpu_walk_us, "rule_b_us_per_item": cpu_item_us,
                     "cores": 1, "note": "C port of label_prop_utils.py:24-100, one thread, on "
                                         "the checked starts / items (walk scratch shared over "
                                         "the batch; profiles given)"},
        "api": "mcg_pool_bfs + mcg_pool_score_member_items + mcg_pool_labels_set, %d device(s)" % pool.size,
    }


def oracle_rows(oracle, hs, ids):
    """(profiles, clamped coverages) of the contigs ``ids`` by the oracle."""
    lens = hs.offsets[ids + 1] - hs.offsets[ids]
    off = np.zeros(ids.shape[0] + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    blob = np.concatenate([hs.blob[hs.offsets[i]:hs.offsets[i + 1]] for i in ids])
    _, prof = oracle.count_kmers_batch(blob, off, want_counts=False, nthreads=1)
    cov = hs.coverage[ids].copy()
    cov[cov < 1e-4] = 1e-4
    return prof, cov


# ---- the other BASELINE configurations -------------------------------------------------------------
def run_config(rt, name, cfg, pool, args, hbm_peak, hbm_src, fp64_peak):
    from metacoag_b200 import _lib, parallel
    torch, dist = rt.torch, rt.dist
    world, rank, dev = rt.world, rt.rank, rt.dev
    scale = float(os.environ.get("MCG_BENCH_SCALE", 1.0))
    per_gpu = bool(cfg.get("per_gpu"))
    n_total = int(cfg["n"] * scale) // 8 * 8
    n_rank = n_total if per_gpu else n_total // world
    S, B, seed = cfg["samples"], cfg["bins"], cfg["seed"]
    big = n_rank * S > 2e7
    log(name, "rank", rank, "tables")
    tabs = workload_tables(seed, n_rank, S, B, shard=rank, kind=cfg["kind"],
                           strains=cfg.get("strains", False), coverage=not big)
    offsets = tabs["offsets"]
    n_bases = int(offsets[-1])
    ctx = _lib.Context(rt.local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    blob = _lib.pinned_empty(n_bases)
    ctx.synth_bases(offsets, tabs["genome_of"], tabs["trans"], seed + rank, out=blob)
    cov = _lib.pinned_empty(8 * n_rank * S).view(np.float64).reshape(n_rank, S)
    if big:
        device_coverage(tabs, S, seed, rank, dev, cov)
    else:
        cov[:] = tabs["coverage"]
    log(name, "rank", rank, "load", n_rank, "contigs", n_bases, "bases")
    ctx.load_contigs(blob, offsets)
    ctx.load_coverage(cov)
    ctx.scan()
    # rank 0's marker-seeded bins are the bins of every rank
    fused = torch.zeros((B, 136 + S), dtype=torch.float64, device=dev)
    if rank == 0:
        cen_prof, cen_cov = ctx.bin_profiles(tabs["members"], tabs["seg"])
        fused.copy_(torch.from_numpy(np.concatenate([cen_prof, cen_cov], axis=1)))
    d_arg = torch.empty(n_rank, dtype=torch.int32, device=dev)
    d_w = torch.empty(n_rank, dtype=torch.float64, device=dev)
    d_cp = torch.empty((B, 136), dtype=torch.float64, device=dev)
    d_cc = torch.empty((B, S), dtype=torch.float64, device=dev)

    def step():
        ctx.step_featurise()
        if world > 1:
            dist.broadcast(fused, src=0)
        d_cp.copy_(fused[:, :136])
        d_cc.copy_(fused[:, 136:])
        ctx.set_centroids_dev(d_cp.data_ptr(), d_cc.data_ptr(), B)
        ctx.step_score(d_arg.data_ptr(), d_w.data_ptr())
        if world > 1 and not per_gpu:
            return parallel.gather_assignments(d_arg, d_w, counts=[n_rank] * world)
        return d_arg, d_w

    for _ in range(3):
        step()
    rt.barrier()
    # per-kernel times: the product path, then every pair scored (the flops of the FP64 roofline)
    ctx.profile(True)
    step()
    kms = ctx.timings()
    launches = kms.pop("launches")
    prune = ctx.prune_stats()
    redo = int(ctx.last_redo_count()) if S > 30 else None
    res = step()
    torch.cuda.synchronize()
    pruned = (res[0].clone(), res[1].clone())
    ctx.set_prune(False)
    step()
    step()
    kms_all = ctx.timings()
    kms_all.pop("launches")
    res = step()
    torch.cuda.synchronize()
    bound_exact = bool(torch.equal(pruned[0], res[0]) and torch.equal(pruned[1], res[1]))
    ctx.set_prune(True)
    ctx.profile(False)
    steps = cfg["steps"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rt.barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    rt.barrier()
    ms_step = rt.max_over_ranks(e0.elapsed_time(e1)) / steps
    n_all = n_rank * world
    entry = {
        "name": name, "workload": cfg["text"], "contigs": n_all, "contigs_per_gpu": n_rank,
        "bases_per_gpu": n_bases, "samples": S, "bins": B, "n_gpus": world,
        "scaling": "weak" if per_gpu else "strong",
        "ms_per_step": ms_step, "value": n_all / (ms_step * 1e-3), "unit": "contigs/s",
        "steps": steps, "gpu_launches_per_step": launches,
        "exchange": "none (single GPU)" if world == 1 else
                    "NCCL broadcast of the bin tables + all-gather of (bin, weight) on the compute stream",
        "kernel_ms": kms, "kernel_ms_every_pair_scored": kms_all,
        "rule_c_candidates": dict(
            prune, per_contig=(prune["candidates"] / max(1, prune["pairs"] // B)
                               if prune.get("pairs") else None),
            bound_is_exact=bound_exact,
            note="last rule-C chunk: pairs listed by bound_kernel / all pairs; bound_is_exact = "
                 "argmin and weights of the pruned pass bit-equal to the all-pairs pass"),
    }
    if redo is not None:
        entry["redo_contigs_per_step"] = redo
    # roofline of the dominant kernels
    scan_bytes = scan_algorithmic_bytes(tabs["lengths"])
    pairs = float(prune["pairs"]) if prune.get("pairs") else float(n_rank) * B
    entry["roofline"] = roof("scan_kernel", "hbm", scan_bytes, kms["scan"], hbm_peak, "GB/s", hbm_src,
                             algorithmic_bytes=scan_bytes,
                             gbases_per_s=n_bases / (kms["scan"] * 1e-3) / 1e9)
    if kms_all.get("probability", 0.0) > 0.0:
        flops = pairs * (431 - 136 * 3 + 14 * S)
        entry["roofline_scoring"] = roof(
            "prob_kernel (every pair scored, last chunk)", "fp64", flops, kms_all["probability"],
            fp64_peak, "TFLOP/s", "DFMA microbenchmark on this device (mcg_fp64_peak)",
            algorithmic_flops=flops, pairs_per_launch=pairs,
            product_path_ms=kms["probability"],
            note="SURVEY 8d flops of the Poisson / Gaussian part (23 + 14 S per pair); the product "
                 "path (branch and bound) takes product_path_ms for the same chunk")
    ctx.close()
    del blob, cov
    rt.host_barrier()

    # ---- through the pool, one process (rank 0) --------------------------------------------------
    e2e_steps = cfg["e2e_steps"]
    if rank == 0 and e2e_steps > 0:
        log(name, "host set for the pool")
        shards = [(tabs, 0)] if world == 1 else [
            (workload_tables(seed, n_rank, S, B, shard=r, kind=cfg["kind"],
                             strains=cfg.get("strains", False), coverage=not big), r)
            for r in range(world)]
        hs = HostSet(pool, shards, S, seed)
        log(name, "pool e2e")
        e2e, arg, w = pool_e2e(pool, hs, cen_prof, cen_cov, e2e_steps)
        entry["e2e"] = e2e
        log(name, "cpu baseline + parity")
        base, parity = cpu_baseline(hs.blob, hs.offsets, hs.coverage, cen_prof, cen_cov,
                                    float(os.environ.get("MCG_BENCH_CFG_CPU_S", 6.0)), arg, w,
                                    seed=seed)
        entry["cpu_baseline"] = base
        entry["parity"] = parity
        entry["e2e_vs_cpu_port"] = e2e["value"] / base["value"]
        if not os.environ.get("MCG_BENCH_NO_PYTHON_REF"):
            # ~3 s of the reference's Python per configuration (B x S decides the cost per contig)
            n_score = int(max(16, min(160, 3.0 / (B * (8e-6 + 0.8e-6 * S)))))
            entry["cpu_baseline_python"] = python_reference_baseline(
                hs.blob, hs.offsets, hs.coverage, cen_prof, cen_cov, arg, w, n_scan=1500,
                n_score=n_score, seed=seed)
        if cfg.get("label_prop"):
            log(name, "label-propagation round")
            entry["label_prop_round"] = label_prop_round(pool, hs, B, seed)
        del hs
    rt.host_barrier()
    if cfg.get("label_prop") and world > 1:
        # the all-gather of the updated bin_of_contig (int32 per contig) the per-rank layout needs
        lab = torch.zeros(n_rank, dtype=torch.int32, device=dev)
        out = torch.empty(n_rank * world, dtype=torch.int32, device=dev)
        for _ in range(3):
            dist.all_gather_into_tensor(out, lab)
        e0.record(stream)
        for _ in range(10):
            dist.all_gather_into_tensor(out, lab)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = rt.max_over_ranks(e0.elapsed_time(e1)) / 10
        if rank == 0 and "label_prop_round" in entry:
            entry["label_prop_round"]["nccl_allgather_bin_of_contig_ms"] = ms
    return entry


def main():
    # libraries (NCCL's version banner, for one) write to fd 1; the contract is ONE JSON line
    # on stdout, so everything else goes to stderr and the line is written to the saved fd
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if len(sys.argv) == 3 and sys.argv[1] == "--python-ref-worker":
        python_reference_worker(sys.argv[2])
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return

    rt = Runtime()
    torch, dist = rt.torch, rt.dist
    rank, world, local, dev = rt.rank, rt.world, rt.local, rt.dev
    from metacoag_b200 import _lib, parallel

    barrier = rt.barrier

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

    # N > 1: the two exchange steps of a step run on their own stream, and everything they touch is
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

    # Peer-memory exchange (NVLink / NVSwitch stores + a signal-pad barrier, all on the exchange
    # stream) when the GPUs can map each other's memory; the NCCL collectives otherwise.
    peer = None
    if world > 1 and os.environ.get("MCG_BENCH_EXCHANGE", "peer") == "peer":
        try:
            peer = parallel.PeerExchange(n, N_BINS, N_SAMPLES, dev)
        except Exception as err:   # no peer access on this box
            print("peer exchange unavailable (%s): NCCL collectives" % err, file=sys.stderr)
            peer = None

    def issue_broadcast(k):
        """Centroid tables of step k into buffer k & 1, on the exchange stream."""
        if "nocoll" in exp:
            return
        q = k & 1
        with torch.cuda.stream(comm):
            comm.wait_event(ev_scored[q])       # the scorer of step k - 2 read these tables
            if peer is not None:
                if rank == 0:
                    peer.push_centroids(ctx, q, d_cen_prof, d_cen_cov, stream=comm)
                peer.barrier(channel=0, stream=comm)
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
        # gather of step k - 1, in front of it on the exchange stream, ends with a rendezvous
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
                peer.barrier(channel=1, stream=comm)
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
    ms_per_step = rt.max_over_ranks(ms_total) / args.steps
    value = n * world / (ms_per_step * 1e-3)

    # ---- end to end through the plug-in call, host buffers --------------------------------------
    h_arg = _lib.pinned_empty(4 * n).view(np.int32)
    h_w = _lib.pinned_empty(8 * n).view(np.float64)
    pool = None
    if world == 1:
        ncpu = len(os.sched_getaffinity(0))
        pack_threads = int(os.environ.get("MCG_HOST_PACK_THREADS", max(0, ncpu - 1)))

        def time_e2e(threads, steps, **kw):
            ctx.set_host_pack_threads(threads)
            src = kw.pop("src", blob)
            ctx.featurise_score(src, pin_off, pin_cov, cen_prof, cen_cov, argmin=h_arg, wmin=h_w, **kw)
            t0 = time.perf_counter()
            for _ in range(steps):
                ctx.featurise_score(src, pin_off, pin_cov, cen_prof, cen_cov, argmin=h_arg,
                                    wmin=h_w, **kw)
            return (time.perf_counter() - t0) / steps, ctx.ingest_stats()

        # the copy engine alone (all bases cross the link as ASCII), for comparison
        e2e_dma_s, ing_dma = time_e2e(0, max(2, args.e2e_steps // 2))
        e2e_s, ingest = time_e2e(pack_threads, args.e2e_steps)
        ctx.profile(True)
        ctx.featurise_score(blob, pin_off, pin_cov, cen_prof, cen_cov, argmin=h_arg, wmin=h_w)
        e2e_ms = ctx.timings()
        e2e_ms.pop("launches")
        ctx.profile(False)
        e2e_arg, e2e_w = h_arg.copy(), h_w.copy()
        t0 = time.perf_counter()
        packed = _lib.pack_2bit_host(blob, pin_off, threads=0, pinned=True)
        pack_ms = (time.perf_counter() - t0) * 1e3
        e2e_p_s, _ = time_e2e(pack_threads, args.e2e_steps, src=packed, encoding=_lib.ENC_2BIT)
        packed_same = bool(np.array_equal(e2e_arg, h_arg) and np.array_equal(e2e_w, h_w))
        side = offsets.nbytes + 16 * n + coverage.nbytes + cen_prof.nbytes + cen_cov.nbytes
        e2e = {"value": n / e2e_s, "unit": "contigs/s", "h2d_bytes_per_step": ingest["h2d_bytes"] + side,
               "d2h_bytes_per_step": 12 * n + 16, "ms_per_step": e2e_s * 1e3,
               "steps": args.e2e_steps, "device_ms": e2e_ms,
               "api": "mcg_featurise_score, ASCII bases in",
               "ingest": dict(ingest, note="ASCII blob split between the copy engine (+ pack "
                              "kernel) and host pack threads, meeting in the middle; contig "
                              "ranges are scanned and scored as they arrive, so the kernel "
                              "times sit inside device_ms.h2d"),
               "copy_engine_only": {"value": n / e2e_dma_s, "ms_per_step": e2e_dma_s * 1e3,
                                    "h2d_bytes_per_step": ing_dma["h2d_bytes"] + side},
               "packed": {"value": n / e2e_p_s, "unit": "contigs/s", "ms_per_step": e2e_p_s * 1e3,
                          "h2d_bytes_per_step": int(packed.nbytes + side),
                          "d2h_bytes_per_step": 12 * n + 16, "same_results_as_ascii": packed_same,
                          "note": "contigs already in the 2-bit store (written once by the FASTA "
                                  "reader at parse time, %.0f ms for this set with all host "
                                  "threads); a quarter of the bytes on the link" % pack_ms}}
        del packed
        # device results of the last resident step agree with the plug-in call
        ctx.step_resident(d_argmin.data_ptr(), d_wmin.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_argmin.cpu().numpy(), e2e_arg)
        hs_blob, hs_off, hs_cov = blob, offsets, coverage
    else:
        # one process drives the box: rank 0, through the pool, on the N shards of the timed run
        rt.host_barrier()
        e2e = None
        if rank == 0:
            pool = _lib.Pool(list(range(world)))
            shards = [(tabs, 0)] + [(workload_tables(1001, N_CONTIGS, N_SAMPLES, N_BINS, shard=r), r)
                                    for r in range(1, world)]
            hs = HostSet(pool, shards, N_SAMPLES, 1001)
            e2e, e2e_arg, e2e_w = pool_e2e(pool, hs, cen_prof, cen_cov, args.e2e_steps)
            # the rows of shard 0 are the rows this rank scored resident
            assert np.array_equal(d_argmin.cpu().numpy(), e2e_arg[:n])
            hs_blob, hs_off, hs_cov = hs.blob, hs.offsets, hs.coverage
        rt.host_barrier()

    # ---- roofline -----------------------------------------------------------------------------------
    hbm_peak, hbm_src = load_peaks()
    lens = tabs["lengths"]
    scan_bytes = scan_algorithmic_bytes(lens)
    pairs = float(n) * N_BINS
    score_flops = pairs * (431 + 14 * N_SAMPLES)
    captured = captured_traffic()
    wl_key = "%s-%d-%d-%d" % (LENGTH_KIND, N_CONTIGS, N_SAMPLES, N_BINS)
    cap = captured.get(wl_key, {})

    fp64_src = "DFMA microbenchmark on this device (mcg_fp64_peak); DMMA measures the same"
    scan_roof = roof("scan_kernel", "hbm", scan_bytes, kernel_ms["scan"], hbm_peak, "GB/s", hbm_src,
                     traffic=cap.get("scan_kernel"),
                     algorithmic_bytes=scan_bytes,
                     gbases_per_s=n_bases / (kernel_ms["scan"] * 1e-3) / 1e9,
                     traffic_source=("%s (kernel source unchanged since commit %s)" % (
                         cap.get("source"), cap.get("commit")) if cap.get("source") else None),
                     note="bound by the L1TEX load/store data pipe, not HBM (DESIGN.md 4.1): two "
                          "shared-memory wavefronts per 32 windows is the floor of a "
                          "byte-counter histogram")
    # The scan is not HBM-bound: its counters live in shared memory and every window costs a
    # byte load and a byte store there.  What it fills is the L1TEX load/store data pipe, one
    # 128-byte wavefront per clock per SM (ncu: l1tex__data_pipe_lsu_wavefronts 88 % of peak);
    # the wavefront count per launch comes from the same ncu capture as `traffic`.
    wf = cap.get("scan_kernel_lsu_wavefronts")
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
        # The distance / probability timers hold ONE launch: with a chunked d^2 scratch (many rows x
        # many bins) that is the last chunk, whose pair count the library reports.
        chunk_pairs = float(prune_stats["pairs"]) if prune_stats.get("pairs") else pairs
        # the kernel computes |u|^2 + |v|^2 - 2 u.v: 2 x 136 flops per pair on the FP64 tensor
        # path are what it EXECUTES (`frac`); SURVEY 8d's difference form counts 3 x 136
        dmma_flops = chunk_pairs * 136 * 2
        kernels.append(roof("dist2_kernel (DMMA distance GEMM)", "fp64", dmma_flops,
                            kernel_ms_all["distance"], fp64_peak, "TFLOP/s", fp64_src,
                            traffic=cap.get("dist2_kernel"), executed_flops=dmma_flops,
                            pairs_per_launch=chunk_pairs,
                            survey_flops=chunk_pairs * 136 * 3,
                            survey_frac=chunk_pairs * 136 * 3 / (kernel_ms_all["distance"] * 1e-3) / 1e12 / fp64_peak,
                            note="frac = executed DMMA flops (2 x 136 per pair) / peak; survey_frac "
                                 "counts SURVEY 8d's sub, mul, add per element (3 x 136) and can pass 1"))
        prob_flops = chunk_pairs * (431 - 136 * 3 + 14 * N_SAMPLES)
        kernels.append(roof("prob_kernel (Poisson / Gaussian terms, argmin)", "fp64", prob_flops,
                            kernel_ms_all["probability"], fp64_peak, "TFLOP/s", fp64_src,
                            traffic=cap.get("prob_kernel"), algorithmic_flops=prob_flops,
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
        "config": config_dict(world, peer),
        "run": {"bases_per_gpu": n_bases,
                "parallelism": "single GPU" if world == 1 else
                               ("row-sharded x%d, " % world) +
                               ("bin tables and assignments exchanged by peer-memory stores "
                                "(NVLink) + signal barrier" if peer is not None
                                else "NCCL broadcast + all-gather"),
                "l2": "inputs larger than L2 (%.0f MB packed bases + %.0f MB profiles per step)"
                      % (n_bases / 4e6, n * 136 * 8 / 1e6),
                "rule_c": "branch and bound (bound_kernel + candidate kernels; exact: argmin "
                          "and weights checked bit-equal to the all-pairs pass in this run)"},
        "clocks": clocks,
        "gpu_launches": launches_per_step * args.steps,
        "roofline": dominant,
        "roofline_kernels": kernels,
        "kernel_ms": kernel_ms,
        "kernel_ms_every_pair_scored": kernel_ms_all,
        "host_enqueue_ms_per_step": host_enqueue_ms,
    }
    line["rule_c_candidates"] = dict(prune_stats, note="pairs scored after the bound / all pairs "
                                     "of the last rule-C chunk; used = the candidate list held "
                                     "them all (else prob_kernel scores every pair)")
    if N_SAMPLES > 30:
        # contigs the guarded rule-C path re-scored in the reference's operation order
        line["run"]["redo_contigs_per_step"] = int(ctx.last_redo_count())
    if rank == 0:
        line["e2e"] = e2e
        base, parity = cpu_baseline(hs_blob, hs_off, hs_cov, cen_prof, cen_cov, args.cpu_budget,
                                    e2e_arg, e2e_w)
        line["cpu_baseline"] = base
        line["parity"] = parity
        if not os.environ.get("MCG_BENCH_NO_PYTHON_REF"):
            line["cpu_baseline_python"] = python_reference_baseline(
                hs_blob, hs_off, hs_cov, cen_prof, cen_cov, e2e_arg, e2e_w)
        line["e2e"]["weak_scaling_note"] = (
            "e2e at N GPUs is one process (rank 0) driving all N through mcg_pool_*; the host "
            "reads N x %.2f GB of ASCII per step, so it is bound by host DRAM, not by the GPUs; "
            "e2e.packed moves a quarter of that" % (n_bases / 1e9))
    ctx.close()
    del blob
    rt.host_barrier()

    # ---- the other BASELINE configurations ----------------------------------------------------------
    wanted = os.environ.get("MCG_BENCH_CONFIGS", ",".join(CONFIGS))
    names = [c for c in wanted.split(",") if c in CONFIGS]
    configs = []
    if names:
        if rank == 0 and pool is None:
            pool = _lib.Pool(list(range(world)))
        for name in names:
            try:
                entry = run_config(rt, name, CONFIGS[name], pool, args, hbm_peak, hbm_src, fp64_peak)
            except Exception as err:   # a config that fails must not take the headline with it
                import traceback
                traceback.print_exc()
                entry = {"name": name, "error": "%s: %s" % (type(err).__name__, err)}
                if world > 1:
                    raise
            configs.append(entry)
    if rank == 0:
        line["configs"] = configs
        emit(line)
    if pool is not None:
        pool.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
