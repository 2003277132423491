"""ctypes face of oracle/libmcg_oracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under metacoag_b200/ does.
Parity status: pinned against tests/golden (generated from the live reference).
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmcg_oracle.so")
MAX_WEIGHT = sys.float_info.max
NPROF = 136

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "mcg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.orc_kmer_table.restype = ctypes.c_int
        L.orc_kmer_table.argtypes = [_u8p]
        L.orc_count_kmers_batch.restype = None
        L.orc_count_kmers_batch.argtypes = [_u8p, _i64p, ctypes.c_int64, ctypes.c_void_p, _f64p,
                                            ctypes.c_int]
        L.orc_segment_mean.restype = None
        L.orc_segment_mean.argtypes = [_f64p, _i64p, _i64p, ctypes.c_int64, ctypes.c_int, _f64p]
        for name in ("orc_normpdf",):
            getattr(L, name).restype = ctypes.c_double
            getattr(L, name).argtypes = [ctypes.c_double] * 3
        L.orc_comp_prob.restype = ctypes.c_double
        L.orc_comp_prob.argtypes = [ctypes.c_double]
        L.orc_lgamma.restype = ctypes.c_double
        L.orc_lgamma.argtypes = [ctypes.c_double]
        L.orc_euclidean.restype = ctypes.c_double
        L.orc_euclidean.argtypes = [_f64p, _f64p, ctypes.c_int]
        L.orc_cov_prob.restype = ctypes.c_double
        L.orc_cov_prob.argtypes = [_f64p, _f64p, ctypes.c_int]
        L.orc_pair_lp.restype = ctypes.c_int
        L.orc_pair_lp.argtypes = [_f64p, _f64p, _f64p, _f64p, ctypes.c_int,
                                  ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                  ctypes.POINTER(ctypes.c_double)]
        L.orc_score_members.restype = None
        L.orc_score_members.argtypes = [_f64p, _f64p, ctypes.c_int, _i64p, ctypes.c_int64, _i64p,
                                        _i64p, ctypes.c_int64, ctypes.c_int, _f64p, ctypes.c_int]
        L.orc_score_centroids.restype = None
        L.orc_score_centroids.argtypes = [_f64p, _f64p, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_int64, _f64p, _f64p, ctypes.c_int64,
                                          ctypes.c_void_p, _i32p, _f64p, ctypes.c_int]
        L.orc_cov_terms.restype = None
        L.orc_cov_terms.argtypes = [_f64p, ctypes.c_int64, _f64p, _f64p]
        L.orc_max_threads.restype = ctypes.c_int
        L.orc_synth_bases.restype = None
        L.orc_synth_bases.argtypes = [_i64p, ctypes.c_int64, _i32p,
                                      np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS"),
                                      ctypes.c_uint64, ctypes.c_double, ctypes.c_double, _u8p,
                                      ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        L.orc_bfs_long.restype = ctypes.c_int64
        L.orc_bfs_long.argtypes = [_i64p, _i32p, ctypes.c_int64, _i32p, ctypes.c_int64,
                                   ctypes.c_int, _i32p, ctypes.c_int64]
        L.orc_bfs_long_batch.restype = ctypes.c_int
        L.orc_bfs_long_batch.argtypes = [_i64p, _i32p, ctypes.c_int64, _i32p, _i64p, ctypes.c_int64,
                                         ctypes.c_int, _i32p, ctypes.c_int64, _i64p]
        _lib = L
    return _lib


def kmer_table():
    t = np.zeros(256, dtype=np.uint8)
    n = lib().orc_kmer_table(t)
    return t, n


def count_kmers_batch(blob, offsets, want_counts=True, nthreads=1):
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    if blob.shape[0] == 0:
        blob = np.zeros(1, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offsets.shape[0] - 1
    freqs = np.zeros((n, NPROF))
    counts = np.zeros((n, NPROF)) if want_counts else None
    lib().orc_count_kmers_batch(blob, offsets, n, counts.ctypes.data if want_counts else None,
                                freqs, nthreads)
    return counts, freqs


def segment_mean(rows, members, seg_offsets):
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    nseg = len(seg_offsets) - 1
    out = np.zeros((nseg, rows.shape[1]))
    lib().orc_segment_mean(rows, np.ascontiguousarray(members, dtype=np.int64),
                           np.ascontiguousarray(seg_offsets, dtype=np.int64), nseg,
                           rows.shape[1], out)
    return out


def normpdf(x, mean, sd):
    return lib().orc_normpdf(float(x), float(mean), float(sd))


def comp_prob(d):
    return lib().orc_comp_prob(float(d))


def euclidean(u, v):
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().orc_euclidean(u, v, u.shape[0])


def cov_prob(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return lib().orc_cov_prob(a, b, a.shape[0])


def pair_lp(pq, pt, cq, ct):
    """Returns (valid, lp, prob_comp, prob_cov) for one (query, target) pair."""
    lp, pc, pv = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    cq = np.ascontiguousarray(cq, dtype=np.float64)
    ok = lib().orc_pair_lp(np.ascontiguousarray(pq, dtype=np.float64),
                           np.ascontiguousarray(pt, dtype=np.float64), cq,
                           np.ascontiguousarray(ct, dtype=np.float64), cq.shape[0],
                           ctypes.byref(lp), ctypes.byref(pc), ctypes.byref(pv))
    return bool(ok), lp.value, pc.value, pv.value


def score_members(prof, cov, queries, members, seg_offsets, rule, nthreads=1):
    """rule: 0 = A (divide by all members), 1 = B (divide by valid members)."""
    prof = np.ascontiguousarray(prof, dtype=np.float64)
    cov = np.ascontiguousarray(cov, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.int64)
    nseg = len(seg_offsets) - 1
    out = np.zeros((queries.shape[0], nseg))
    lib().orc_score_members(prof, cov, cov.shape[1], queries, queries.shape[0],
                            np.ascontiguousarray(members, dtype=np.int64),
                            np.ascontiguousarray(seg_offsets, dtype=np.int64), nseg, rule, out,
                            nthreads)
    return out


def score_centroids(prof, cov, cen_prof, cen_cov, queries=None, want_weights=False, nthreads=1):
    """Rule C.  Returns (argmin int32 [-1 = None], min weight, weights or None)."""
    prof = np.ascontiguousarray(prof, dtype=np.float64)
    cov = np.ascontiguousarray(cov, dtype=np.float64)
    cen_prof = np.ascontiguousarray(cen_prof, dtype=np.float64)
    cen_cov = np.ascontiguousarray(cen_cov, dtype=np.float64)
    if queries is not None:
        queries = np.ascontiguousarray(queries, dtype=np.int64)
        nq, qptr = queries.shape[0], queries.ctypes.data
    else:
        nq, qptr = prof.shape[0], None
    B = cen_prof.shape[0]
    weights = np.zeros((nq, B)) if want_weights else None
    best = np.zeros(nq, dtype=np.int32)
    best_w = np.zeros(nq)
    lib().orc_score_centroids(prof, cov, cov.shape[1], qptr, nq, cen_prof, cen_cov, B,
                              weights.ctypes.data if want_weights else None, best, best_w,
                              nthreads)
    return best, best_w, weights


def cov_terms(cov):
    cov = np.ascontiguousarray(cov, dtype=np.float64)
    logc = np.zeros_like(cov)
    lgam = np.zeros_like(cov)
    lib().orc_cov_terms(cov.reshape(-1), cov.size, logc.reshape(-1), lgam.reshape(-1))
    return logc, lgam


def max_threads():
    return lib().orc_max_threads()


def strip_ascii(raw):
    """str.strip() on bytes: what count_kmers does before scanning (feature_utils.py:63)."""
    return raw.decode("ascii").strip().encode("ascii")


def synth_bases(offsets, genome_of, trans, seed, n_frac=1e-4, lower_frac=0.01, n_contigs=None,
                nthreads=0):
    """The synthetic contig generator of the CUDA library (mcg_synth_bases), restated: the
    bases of the first ``n_contigs`` contigs (all by default), byte for byte."""
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    genome_of = np.ascontiguousarray(genome_of, dtype=np.int32)
    trans = np.ascontiguousarray(trans, dtype=np.float32)
    n = offsets.shape[0] - 1
    end = int(offsets[n if n_contigs is None else n_contigs])
    out = np.zeros(max(end, 1), dtype=np.uint8)
    lib().orc_synth_bases(offsets, n, genome_of, trans, int(seed), float(n_frac), float(lower_frac),
                          out, 0, end, nthreads)
    return out[:end]


def bfs_long(indptr, indices, labels, start, depth_limit, max_hits=4096):
    """label_prop_utils.py:24-100 traversal: [(labelled node, bin, depth)] in pop order."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    while True:
        hits = np.zeros((max_hits, 3), dtype=np.int32)
        n = lib().orc_bfs_long(indptr, indices, indptr.shape[0] - 1, labels, int(start),
                               int(depth_limit), hits.reshape(-1), max_hits)
        if n < 0:
            raise MemoryError("orc_bfs_long")
        if n <= max_hits:
            return [tuple(r) for r in hits[:n].tolist()]
        max_hits = int(n)


def bfs_long_batch(indptr, indices, labels, starts, depth_limit, max_hits=64):
    """The walks of ``bfs_long`` for many start nodes on one thread with shared scratch: a list of
    hit lists.  Walks with more than ``max_hits`` hits are repeated one by one."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    n = starts.shape[0]
    hits = np.zeros((n, max_hits, 3), dtype=np.int32)
    count = np.zeros(n, dtype=np.int64)
    rc = lib().orc_bfs_long_batch(indptr, indices, indptr.shape[0] - 1, labels, starts, n,
                                  int(depth_limit), hits.reshape(-1), max_hits, count)
    if rc != 0:
        raise MemoryError("orc_bfs_long_batch")
    out = []
    for i in range(n):
        if count[i] <= max_hits:
            out.append([tuple(r) for r in hits[i, :count[i]].tolist()])
        else:
            out.append(bfs_long(indptr, indices, labels, int(starts[i]), depth_limit, int(count[i])))
    return out
