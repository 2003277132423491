"""ctypes binding of libmcg_b200.so (include/mcg_b200.h).

This is the only compute path of the package.  If the shared library is missing, or there
is no CUDA device, everything here raises: there is no CPU fallback and nothing in this
package imports ``oracle/``.
"""
import ctypes
import os
import sys
import threading

import numpy as np

NPROF = 136
MAX_WEIGHT = sys.float_info.max
ENC_ASCII, ENC_2BIT = 0, 1
RULE_A, RULE_B = 0, 1
N_TIMERS = 10
TIMER_NAMES = ("pack", "scan", "normalise", "cov_terms", "score", "h2d", "d2h", "other",
               "distance", "probability")

OK, ERR_CUDA, ERR_ARG, ERR_STATE, ERR_DOMAIN, ERR_NOMEM = 0, -1, -2, -3, -4, -5

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcg_b200.so")

_c = ctypes
_vp, _i64, _int, _dbl = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_double

# name -> (restype, argtypes); the list tests/test_abi.py checks against the header
SIGNATURES = {
    "mcg_device_count": (_int, []),
    "mcg_create": (_vp, [_int]),
    "mcg_destroy": (None, [_vp]),
    "mcg_last_error": (_c.c_char_p, [_vp]),
    "mcg_set_stream": (_int, [_vp, _vp]),
    "mcg_synchronize": (_int, [_vp]),
    "mcg_host_alloc": (_vp, [_c.c_size_t]),
    "mcg_host_free": (None, [_vp]),
    "mcg_kmer_table": (_int, [_vp]),
    "mcg_tetramer_scan": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp]),
    "mcg_load_contigs": (_int, [_vp, _vp, _vp, _i64, _int]),
    "mcg_scan": (_int, [_vp]),
    "mcg_get_profiles": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "mcg_set_profiles": (_int, [_vp, _vp, _i64]),
    "mcg_load_coverage": (_int, [_vp, _vp, _i64, _int]),
    "mcg_get_coverage_terms": (_int, [_vp, _vp, _vp, _vp]),
    "mcg_bin_profiles": (_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "mcg_normpdf": (_int, [_vp, _vp, _dbl, _dbl, _vp, _i64]),
    "mcg_tetramer_distance": (_int, [_vp, _vp, _vp, _i64, _int, _vp]),
    "mcg_comp_probability": (_int, [_vp, _vp, _vp, _i64]),
    "mcg_cov_probability": (_int, [_vp, _vp, _vp, _i64, _int, _vp]),
    "mcg_pair_scores": (_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "mcg_score_centroids": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp]),
    "mcg_score_members": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _int, _vp]),
    "mcg_bfs_candidates": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _int, _int, _vp, _vp,
                                  _vp, _vp]),
    "mcg_path_lengths": (_int, [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "mcg_connected_components": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "mcg_featurise_score": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _int, _vp, _vp, _i64, _vp,
                                   _vp]),
    "mcg_set_centroids": (_int, [_vp, _vp, _vp, _i64]),
    "mcg_set_centroids_dev": (_int, [_vp, _vp, _vp, _i64]),
    "mcg_step_resident": (_int, [_vp, _vp, _vp]),
    "mcg_step_featurise": (_int, [_vp]),
    "mcg_step_score": (_int, [_vp, _vp, _vp]),
    "mcg_get_assignments": (_int, [_vp, _vp, _vp]),
    "mcg_profile": (_int, [_vp, _int]),
    "mcg_timings": (_int, [_vp, _vp, _vp]),
    "mcg_synth_bases": (_int, [_vp, _vp, _i64, _vp, _vp, _int, _c.c_uint64, _dbl, _dbl, _vp]),
    "mcg_fp64_peak": (_int, [_vp, _vp]),
    "mcg_last_redo_count": (_int, [_vp, _vp]),
    "mcg_set_host_pack_threads": (_int, [_vp, _int]),
    "mcg_set_prune": (_int, [_vp, _int]),
    "mcg_set_reserved_sms": (_int, [_vp, _int]),
    "mcg_push_to_peers": (_int, [_vp, _vp, _c.c_size_t, _vp, _int, _vp]),
    "mcg_prune_stats": (_int, [_vp, _vp]),
    "mcg_ingest_stats": (_int, [_vp, _vp]),
    "mcg_pack_2bit_host": (_int, [_vp, _vp, _i64, _vp, _int]),
    "mcg_fasta_index": (_int, [_c.c_char_p, _vp, _vp, _vp]),
    "mcg_fasta_load": (_int, [_c.c_char_p, _vp, _vp, _vp, _vp, _i64, _i64, _i64]),
    "mcg_write_bin_fastas": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _c.c_int32]),
    "mcg_pack_2bit_host_mt": (_int, [_vp, _vp, _i64, _vp, _int]),
    "mcg_set_tc": (_int, [_vp, _int]),
    "mcg_tc_distances": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "mcg_score_member_items": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _int, _vp]),
    "mcg_graph_load": (_int, [_vp, _vp, _vp, _i64]),
    "mcg_labels_load": (_int, [_vp, _vp, _i64]),
    "mcg_labels_set": (_int, [_vp, _vp, _vp, _i64]),
    "mcg_bfs_resident": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _vp, _i64, _vp]),
    "mcg_pool_create": (_vp, [_vp, _int]),
    "mcg_pool_destroy": (None, [_vp]),
    "mcg_pool_size": (_int, [_vp]),
    "mcg_pool_ctx": (_vp, [_vp, _int]),
    "mcg_pool_last_error": (_c.c_char_p, [_vp]),
    "mcg_pool_bounds": (_int, [_vp, _vp]),
    "mcg_pool_tetramer_scan": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp]),
    "mcg_pool_featurise_score": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _int, _vp, _vp, _i64,
                                        _vp, _vp]),
    "mcg_pool_set_features": (_int, [_vp, _vp, _i64, _vp, _i64, _int]),
    "mcg_pool_bin_profiles": (_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "mcg_pool_score_centroids": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp]),
    "mcg_pool_score_members": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _int, _vp]),
    "mcg_pool_score_member_items": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _int, _vp]),
    "mcg_pool_graph_load": (_int, [_vp, _vp, _vp, _i64]),
    "mcg_pool_labels_load": (_int, [_vp, _vp, _i64]),
    "mcg_pool_labels_set": (_int, [_vp, _vp, _vp, _i64]),
    "mcg_pool_bfs": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _vp, _i64, _vp]),
}


class McgError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libmcg_b200: %s (code %d)" % (message, code))
        self.code = code


_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen the shared library and declare every prototype.  Raises if it is missing."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    "%s not found: build it with `python -m metacoag_b200.build` "
                    "(there is no CPU fallback)" % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (restype, argtypes) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def _arr(a, dtype, shape=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None:
        a = a.reshape(shape)
    return a


def packed_nbytes(offsets):
    """Bytes of the MCG_ENC_2BIT store of contigs with these offsets (16 per 64 bases, every
    contig padded to a multiple of 64)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    return int(np.sum(16 * ((np.diff(offsets) + 63) // 64)))


def pack_2bit_host(bases, offsets, level=-1, threads=None, pinned=False):
    """ASCII contigs -> MCG_ENC_2BIT bytes with the library's host packer (no device).
    ``threads`` (None: one thread, SIMD variant ``level``; 0: all hardware threads) selects the
    parallel packer a FASTA reader runs at parse time."""
    bases = _arr(bases, np.uint8)
    offsets = _arr(offsets, np.int64)
    n = offsets.shape[0] - 1
    nbytes = packed_nbytes(offsets)
    if pinned and nbytes:
        out = pinned_empty(nbytes)
        out[:] = 0
    else:
        out = np.zeros(nbytes, dtype=np.uint8)
    lib = load_library()
    if threads is None:
        rc = lib.mcg_pack_2bit_host(_ptr(bases), _ptr(offsets), n, _ptr(out), int(level))
    else:
        rc = lib.mcg_pack_2bit_host_mt(_ptr(bases), _ptr(offsets), n, _ptr(out), int(threads))
    if rc != OK:
        raise McgError(rc, "mcg_pack_2bit_host failed")
    return out


def device_count():
    return int(load_library().mcg_device_count())


def read_fasta(path, pinned=False):
    """FASTA file -> (names list[str], bases uint8[n_bases], offsets int64[n+1]) with the
    library's host reader (BioPython FASTA semantics, include/mcg_b200.h).  ``pinned`` puts
    the bases into page-locked memory (needs a CUDA device) so that mcg_load_contigs can hand
    them to the copy engine directly."""
    lib = load_library()
    raw = os.fsencode(path)
    n, nb, nn = _c.c_int64(0), _c.c_int64(0), _c.c_int64(0)
    rc = lib.mcg_fasta_index(raw, _c.byref(n), _c.byref(nb), _c.byref(nn))
    if rc != 0:
        raise FileNotFoundError(path)
    n, nb, nn = n.value, nb.value, nn.value
    bases = pinned_empty(nb) if (pinned and nb) else np.empty(nb, dtype=np.uint8)
    offsets = np.zeros(n + 1, dtype=np.int64)
    names = np.empty(nn, dtype=np.uint8)
    name_offsets = np.zeros(n + 1, dtype=np.int64)
    rc = lib.mcg_fasta_load(raw, _ptr(bases), _ptr(offsets), _ptr(names), _ptr(name_offsets),
                            n, nb, nn)
    if rc != 0:
        raise OSError("mcg_fasta_load(%s) failed (%d)" % (path, rc))
    text = names.tobytes().decode("latin-1")
    no = name_offsets.tolist()
    return [text[no[i]:no[i + 1]] for i in range(n)], bases, offsets


def write_bin_fastas(names, bases, offsets, file_of_record, paths):
    """Records of the contig blob -> one FASTA file per entry of ``paths``
    (``>{name}\\n{sequence}\\n`` in record order, metacoag:1236-1251); ``file_of_record[i]`` is
    the index of record i's file or -1.  Host-only."""
    lib = load_library()
    n = len(names)
    encoded = [s.encode("latin-1") for s in names]
    name_offsets = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum([len(b) for b in encoded], out=name_offsets[1:])
    name_blob = np.frombuffer(b"".join(encoded), dtype=np.uint8) if name_offsets[-1] else np.zeros(1, np.uint8)
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    if bases.shape[0] == 0:
        bases = np.zeros(1, np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    file_of_record = np.ascontiguousarray(file_of_record, dtype=np.int32)
    if offsets.shape[0] != n + 1 or file_of_record.shape[0] != n:
        raise ValueError("names, offsets and file_of_record disagree about the record count")
    raw = [os.fsencode(p) for p in paths]
    c_paths = (_c.c_char_p * max(1, len(raw)))(*raw)
    rc = lib.mcg_write_bin_fastas(_ptr(bases), _ptr(offsets), _ptr(name_blob), _ptr(name_offsets),
                                  n, _ptr(file_of_record), _c.cast(c_paths, _c.c_void_p),
                                  len(raw))
    if rc == -1000000000:
        raise ValueError("file_of_record holds an index outside paths")
    if rc != 0:
        raise OSError("could not write %s" % paths[-rc - 1])


class Context:
    """One ``mcg_ctx``: a CUDA device, its resident contig/bin tables and a stream."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.handle = self.lib.mcg_create(int(device))
        if not self.handle:
            msg = self.lib.mcg_last_error(None)
            raise McgError(ERR_CUDA, (msg or b"mcg_create failed").decode())
        self.device = int(device)
        self.n_contigs = 0
        self.n_samples = 0
        self.n_bins = 0

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mcg_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    _RENAMED = {"bfs": "mcg_bfs_resident"}

    def _fn(self, name):
        return getattr(self.lib, self._RENAMED.get(name, "mcg_" + name))

    def _last_error(self):
        return (self.lib.mcg_last_error(self.handle) or b"").decode()

    def _check(self, rc):
        if rc == OK:
            return
        msg = self._last_error()
        if rc == ERR_DOMAIN:
            # what the reference raises at matching_utils.py:39 when both Gaussians underflow
            raise ZeroDivisionError("float division by zero")
        raise McgError(rc, msg)

    # -- plumbing -------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._check(self.lib.mcg_set_stream(self.handle, cuda_stream))

    def synchronize(self):
        self._check(self.lib.mcg_synchronize(self.handle))

    def profile(self, on=True):
        self._check(self.lib.mcg_profile(self.handle, 1 if on else 0))

    def timings(self):
        ms = np.zeros(N_TIMERS, dtype=np.float32)
        launches = _c.c_int64(0)
        self._check(self.lib.mcg_timings(self.handle, ms.ctypes.data, _c.byref(launches)))
        out = dict(zip(TIMER_NAMES, (float(x) for x in ms)))
        out["launches"] = int(launches.value)
        return out

    def set_prune(self, enabled):
        """Rule-C branch and bound on / off (include/mcg_b200.h); results do not change."""
        self._check(self.lib.mcg_set_prune(self.handle, int(bool(enabled))))

    def set_tc(self, enabled):
        """Tensor-core distance bound of the rule-C branch and bound on / off; results do not
        change."""
        self._check(self.lib.mcg_set_tc(self.handle, int(bool(enabled))))

    def tc_distances(self, queries=None, nq=None):
        """(d2a float32 [nq, n_bins], margin float32 [nq]) of the tensor-core distance bound
        against the resident centroids."""
        if queries is not None:
            queries = _arr(queries, np.int64)
            nq = queries.shape[0]
        d2a = np.zeros((nq, self.n_bins), dtype=np.float32)
        margin = np.zeros(nq, dtype=np.float32)
        self._check(self.lib.mcg_tc_distances(self.handle, _ptr(queries), nq, _ptr(d2a), _ptr(margin)))
        return d2a, margin

    def set_reserved_sms(self, sms):
        """SMs the persistent scan kernel leaves free for collectives on another stream."""
        self._check(self.lib.mcg_set_reserved_sms(self.handle, int(sms)))

    def push_to_peers(self, src_ptr, nbytes, peer_ptrs, stream=None):
        """``nbytes`` from device address ``src_ptr`` to every address of ``peer_ptrs`` (peer
        memory mapped into this process), one kernel on ``stream`` (a cudaStream_t handle;
        default: the context's stream)."""
        arr = (_c.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        self._check(self.lib.mcg_push_to_peers(self.handle, int(src_ptr), int(nbytes), arr,
                                               len(peer_ptrs), stream))

    def prune_stats(self):
        """The candidate list of the last pruned rule-C chunk (include/mcg_b200.h)."""
        st = (_c.c_int64 * 4)()
        self._check(self.lib.mcg_prune_stats(self.handle, st))
        return {"candidates": st[0], "capacity": st[1], "pairs": st[2], "used": bool(st[3])}

    def set_host_pack_threads(self, threads):
        """Host threads that pack ASCII chunks during a load (0 = copy engine only)."""
        self._check(self.lib.mcg_set_host_pack_threads(self.handle, int(threads)))

    def ingest_stats(self):
        """Last ASCII load: h2d bytes of the bases, host-packed / kernel-packed bases, threads."""
        st = np.zeros(4, dtype=np.int64)
        self._check(self.lib.mcg_ingest_stats(self.handle, _ptr(st)))
        return {"h2d_bytes": int(st[0]), "host_packed_bases": int(st[1]),
                "gpu_packed_bases": int(st[2]), "host_threads": int(st[3])}

    def last_redo_count(self):
        """Contigs the last S > 30 rule-C call re-scored in the reference's order."""
        n = _c.c_int64(0)
        self._check(self.lib.mcg_last_redo_count(self.handle, _c.byref(n)))
        return n.value

    def fp64_peak(self):
        t = _c.c_double(0.0)
        self._check(self.lib.mcg_fp64_peak(self.handle, _c.byref(t)))
        return t.value

    # -- featurisation ----------------------------------------------------------------
    def tetramer_scan(self, bases, offsets, encoding=ENC_ASCII, want_counts=True, want_freqs=True):
        bases = _arr(bases, np.uint8)
        offsets = _arr(offsets, np.int64)
        n = offsets.shape[0] - 1
        counts = np.zeros((n, NPROF), dtype=np.uint32) if want_counts else None
        freqs = np.zeros((n, NPROF), dtype=np.float64) if want_freqs else None
        self._check(self._fn("tetramer_scan")(self.handle, _ptr(bases), _ptr(offsets), n,
                                               encoding, _ptr(counts), _ptr(freqs)))
        self.n_contigs = n
        return counts, freqs

    def load_contigs(self, bases, offsets, encoding=ENC_ASCII):
        bases = _arr(bases, np.uint8)
        offsets = _arr(offsets, np.int64)
        n = offsets.shape[0] - 1
        self._check(self.lib.mcg_load_contigs(self.handle, _ptr(bases), _ptr(offsets), n,
                                              encoding))
        self.n_contigs = n

    def scan(self):
        self._check(self.lib.mcg_scan(self.handle))

    def get_profiles(self, first=0, n=None, want_counts=False):
        n = self.n_contigs - first if n is None else n
        counts = np.zeros((n, NPROF), dtype=np.uint32) if want_counts else None
        freqs = np.zeros((n, NPROF), dtype=np.float64)
        self._check(self.lib.mcg_get_profiles(self.handle, first, n, _ptr(counts), _ptr(freqs)))
        return (counts, freqs) if want_counts else freqs

    def set_profiles(self, prof):
        prof = _arr(prof, np.float64)
        if prof.ndim != 2 or prof.shape[1] != NPROF:
            raise ValueError("profiles must be [n,136]")
        self._check(self.lib.mcg_set_profiles(self.handle, _ptr(prof), prof.shape[0]))
        self.n_contigs = prof.shape[0]

    def load_coverage(self, cov):
        cov = _arr(cov, np.float64)
        if cov.ndim != 2:
            raise ValueError("coverage must be [n,S]")
        self._check(self.lib.mcg_load_coverage(self.handle, _ptr(cov), cov.shape[0],
                                               cov.shape[1]))
        self.n_samples = cov.shape[1]
        self._cov_rows = cov.shape[0]

    def coverage_terms(self):
        shape = (self._cov_rows, self.n_samples)
        cov, logc, lgam = (np.zeros(shape) for _ in range(3))
        self._check(self.lib.mcg_get_coverage_terms(self.handle, _ptr(cov), _ptr(logc),
                                                    _ptr(lgam)))
        return cov, logc, lgam

    def bin_profiles(self, members, seg_offsets):
        members = _arr(members, np.int64)
        seg_offsets = _arr(seg_offsets, np.int64)
        nseg = seg_offsets.shape[0] - 1
        cen_prof = np.zeros((nseg, NPROF))
        cen_cov = np.zeros((nseg, self.n_samples))
        self._check(self._fn("bin_profiles")(self.handle, _ptr(members), _ptr(seg_offsets), nseg,
                                              _ptr(cen_prof), _ptr(cen_cov)))
        self.n_bins = nseg
        return cen_prof, cen_cov

    # -- elementwise primitives ----------------------------------------------------------
    def normpdf(self, x, mean, sd):
        x = _arr(x, np.float64).reshape(-1)
        out = np.zeros_like(x)
        self._check(self.lib.mcg_normpdf(self.handle, _ptr(x), float(mean), float(sd), _ptr(out),
                                         x.shape[0]))
        return out

    def tetramer_distance(self, u, v):
        u = _arr(u, np.float64)
        v = _arr(v, np.float64)
        if u.ndim == 1:
            u, v = u[None, :], v[None, :]
        out = np.zeros(u.shape[0])
        self._check(self.lib.mcg_tetramer_distance(self.handle, _ptr(u), _ptr(v), u.shape[0],
                                                   u.shape[1], _ptr(out)))
        return out

    def comp_probability(self, dist):
        dist = _arr(dist, np.float64).reshape(-1)
        out = np.zeros_like(dist)
        self._check(self.lib.mcg_comp_probability(self.handle, _ptr(dist), _ptr(out),
                                                  dist.shape[0]))
        return out

    def cov_probability(self, cov1, cov2):
        cov1 = _arr(cov1, np.float64)
        cov2 = _arr(cov2, np.float64)
        if cov1.ndim == 1:
            cov1, cov2 = cov1[None, :], cov2[None, :]
        out = np.zeros(cov1.shape[0])
        self._check(self.lib.mcg_cov_probability(self.handle, _ptr(cov1), _ptr(cov2),
                                                 cov1.shape[0], cov1.shape[1], _ptr(out)))
        return out

    # -- scoring ------------------------------------------------------------------------------
    def pair_scores(self, queries, targets, want=("dist", "pc", "pv", "lp")):
        queries = _arr(queries, np.int64)
        targets = _arr(targets, np.int64)
        shape = (queries.shape[0], targets.shape[0])
        bufs = {k: (np.zeros(shape) if k in want else None) for k in ("dist", "pc", "pv", "lp")}
        self._check(self.lib.mcg_pair_scores(self.handle, _ptr(queries), shape[0], _ptr(targets),
                                             shape[1], _ptr(bufs["dist"]), _ptr(bufs["pc"]),
                                             _ptr(bufs["pv"]), _ptr(bufs["lp"])))
        return bufs

    def score_centroids(self, queries=None, nq=None, cen_prof=None, cen_cov=None,
                        want_weights=False):
        """Rule C.  Returns (argmin int32 with -1 = None, wmin, weights or None)."""
        if queries is not None:
            queries = _arr(queries, np.int64)
            nq = queries.shape[0]
        elif nq is None:
            nq = self.n_contigs
        if cen_prof is not None:
            cen_prof = _arr(cen_prof, np.float64)
            cen_cov = _arr(cen_cov, np.float64)
            n_bins = cen_prof.shape[0]
            if cen_prof.shape != (n_bins, NPROF) or cen_cov.shape != (n_bins, self.n_samples):
                raise ValueError("centroid shapes must be [B,136] and [B,S]")
            self.n_bins = n_bins
        n_bins = self.n_bins
        weights = np.zeros((nq, n_bins)) if want_weights else None
        argmin = np.zeros(nq, dtype=np.int32)
        wmin = np.zeros(nq)
        self._check(self._fn("score_centroids")(self.handle, _ptr(queries), nq, _ptr(cen_prof),
                                                 _ptr(cen_cov), n_bins, _ptr(weights),
                                                 _ptr(argmin), _ptr(wmin)))
        return argmin, wmin, weights

    def score_members(self, queries, members, seg_offsets, rule):
        queries = _arr(queries, np.int64)
        members = _arr(members, np.int64)
        seg_offsets = _arr(seg_offsets, np.int64)
        nseg = seg_offsets.shape[0] - 1
        weights = np.zeros((queries.shape[0], nseg))
        self._check(self._fn("score_members")(self.handle, _ptr(queries), queries.shape[0],
                                               _ptr(members), _ptr(seg_offsets), nseg, int(rule),
                                               _ptr(weights)))
        return weights

    def bfs_candidates(self, indptr, indices, labels, starts, depth_limit, max_hits=64):
        indptr = _arr(indptr, np.int64)
        indices = _arr(indices, np.int32)
        labels = _arr(labels, np.int32)
        starts = _arr(starts, np.int64)
        n_vertices = indptr.shape[0] - 1
        ns = starts.shape[0]
        while True:
            node = np.zeros((ns, max_hits), dtype=np.int32)
            bin_ = np.zeros((ns, max_hits), dtype=np.int32)
            depth = np.zeros((ns, max_hits), dtype=np.int32)
            n_hits = np.zeros(ns, dtype=np.int32)
            self._check(self.lib.mcg_bfs_candidates(
                self.handle, _ptr(indptr), _ptr(indices), n_vertices, _ptr(labels), _ptr(starts),
                ns, int(depth_limit), int(max_hits), _ptr(node), _ptr(bin_), _ptr(depth),
                _ptr(n_hits)))
            if ns == 0 or int(n_hits.max()) <= max_hits:
                return node, bin_, depth, n_hits
            max_hits = int(n_hits.max())

    def path_lengths(self, indptr, indices, sources, targets):
        """[len(sources), len(targets)] int32: vertices on a shortest path, 0 = unreachable."""
        indptr = _arr(indptr, np.int64)
        indices = _arr(indices, np.int32)
        sources = _arr(sources, np.int64)
        targets = _arr(targets, np.int64)
        out = np.zeros((sources.shape[0], targets.shape[0]), dtype=np.int32)
        self._check(self.lib.mcg_path_lengths(
            self.handle, _ptr(indptr), _ptr(indices), indptr.shape[0] - 1, _ptr(sources),
            sources.shape[0], _ptr(targets), targets.shape[0], _ptr(out)))
        return out

    def connected_components(self, indptr, indices):
        """int32[n]: smallest vertex id of every vertex's component."""
        indptr = _arr(indptr, np.int64)
        indices = _arr(indices, np.int32)
        labels = np.zeros(indptr.shape[0] - 1, dtype=np.int32)
        self._check(self.lib.mcg_connected_components(self.handle, _ptr(indptr), _ptr(indices),
                                                      indptr.shape[0] - 1, _ptr(labels)))
        return labels

    def score_member_items(self, item_query, item_bin, members, seg_offsets, rule):
        """Rules A / B for listed (contig, bin index) items: float64[n_items]."""
        item_query = _arr(item_query, np.int64)
        item_bin = _arr(item_bin, np.int32)
        members = _arr(members, np.int64)
        seg_offsets = _arr(seg_offsets, np.int64)
        weights = np.zeros(item_query.shape[0])
        self._check(self._fn("score_member_items")(
            self.handle, _ptr(item_query), _ptr(item_bin), item_query.shape[0], _ptr(members),
            _ptr(seg_offsets), seg_offsets.shape[0] - 1, int(rule), _ptr(weights)))
        return weights

    # -- resident assembly graph (label propagation) ---------------------------------------------
    def graph_load(self, indptr, indices):
        indptr = _arr(indptr, np.int64)
        indices = _arr(indices, np.int32)
        self._check(self._fn("graph_load")(self.handle, _ptr(indptr), _ptr(indices),
                                           indptr.shape[0] - 1))
        self.n_vertices = indptr.shape[0] - 1

    def labels_load(self, labels):
        labels = _arr(labels, np.int32)
        self._check(self._fn("labels_load")(self.handle, _ptr(labels), labels.shape[0]))

    def labels_set(self, ids, bins):
        ids = _arr(ids, np.int64).reshape(-1)
        bins = _arr(bins, np.int32).reshape(-1)
        self._check(self._fn("labels_set")(self.handle, _ptr(ids), _ptr(bins), ids.shape[0]))

    def bfs_resident(self, starts, depth_limit):
        """Walks from ``starts`` over the resident graph: (hit_start int64[n], hit_count
        int32[n], hits int32[total, 3] = (labelled node, bin, depth)), hits of a start
        contiguous and in the reference's pop order."""
        starts = _arr(starts, np.int64).reshape(-1)
        ns = starts.shape[0]
        hit_start = np.zeros(ns, dtype=np.int64)
        hit_count = np.zeros(ns, dtype=np.int32)
        cap = 4 * ns + 256
        total = _c.c_int64(0)
        while True:
            hits = np.zeros((cap, 3), dtype=np.int32)
            self._check(self._fn("bfs")(
                self.handle, _ptr(starts), ns, int(depth_limit), _ptr(hit_start), _ptr(hit_count),
                _ptr(hits), cap, _c.byref(total)))
            if total.value <= cap:
                return hit_start, hit_count, hits[:total.value]
            cap = int(total.value)

    # -- fused / resident ---------------------------------------------------------------------
    def featurise_score(self, bases, offsets, cov, cen_prof, cen_cov, encoding=ENC_ASCII,
                        argmin=None, wmin=None):
        offsets = _arr(offsets, np.int64)
        n = offsets.shape[0] - 1
        cov = _arr(cov, np.float64)
        cen_prof = _arr(cen_prof, np.float64)
        cen_cov = _arr(cen_cov, np.float64)
        if argmin is None:
            argmin = np.zeros(n, dtype=np.int32)
        if wmin is None:
            wmin = np.zeros(n)
        self._check(self._fn("featurise_score")(
            self.handle, _ptr(bases), _ptr(offsets), n, encoding, _ptr(cov), cov.shape[1],
            _ptr(cen_prof), _ptr(cen_cov), cen_prof.shape[0], _ptr(argmin), _ptr(wmin)))
        self.n_contigs, self.n_samples, self.n_bins = n, cov.shape[1], cen_prof.shape[0]
        self._cov_rows = n
        return argmin, wmin

    def set_centroids(self, cen_prof, cen_cov):
        cen_prof = _arr(cen_prof, np.float64)
        cen_cov = _arr(cen_cov, np.float64)
        self._check(self.lib.mcg_set_centroids(self.handle, _ptr(cen_prof), _ptr(cen_cov),
                                               cen_prof.shape[0]))
        self.n_bins = cen_prof.shape[0]

    def set_centroids_dev(self, d_cen_prof, d_cen_cov, n_bins):
        """Centroids from device pointers (ints), stream-ordered, no host sync."""
        self._check(self.lib.mcg_set_centroids_dev(self.handle, d_cen_prof, d_cen_cov,
                                                   int(n_bins)))
        self.n_bins = int(n_bins)

    def step_resident(self, d_argmin=None, d_wmin=None):
        self._check(self.lib.mcg_step_resident(self.handle, d_argmin, d_wmin))

    def step_featurise(self):
        self._check(self.lib.mcg_step_featurise(self.handle))

    def step_score(self, d_argmin=None, d_wmin=None):
        self._check(self.lib.mcg_step_score(self.handle, d_argmin, d_wmin))

    def get_assignments(self):
        argmin = np.zeros(self.n_contigs, dtype=np.int32)
        wmin = np.zeros(self.n_contigs)
        self._check(self.lib.mcg_get_assignments(self.handle, _ptr(argmin), _ptr(wmin)))
        return argmin, wmin

    def synth_bases(self, offsets, genome_of, trans, seed, n_frac=1e-4, lower_frac=0.01,
                    out=None):
        offsets = _arr(offsets, np.int64)
        genome_of = _arr(genome_of, np.int32)
        trans = _arr(trans, np.float32)
        n = offsets.shape[0] - 1
        if out is None:
            out = np.zeros(int(offsets[-1]), dtype=np.uint8)
        self._check(self.lib.mcg_synth_bases(self.handle, _ptr(offsets), n, _ptr(genome_of),
                                             _ptr(trans), trans.shape[0], int(seed),
                                             float(n_frac), float(lower_frac), _ptr(out)))
        return out


def kmer_table():
    """feature_utils.py:38-57 compute_kmer_inds(4) as a uint8[256] table."""
    table = np.zeros(256, dtype=np.uint8)
    rc = load_library().mcg_kmer_table(table.ctypes.data)
    if rc != OK:
        raise McgError(rc, "mcg_kmer_table")
    return table


def pinned_empty(n_bytes):
    """A uint8 numpy array backed by page-locked memory (freed with the array)."""
    lib = load_library()
    p = lib.mcg_host_alloc(int(max(1, n_bytes)))
    if not p:
        raise MemoryError("mcg_host_alloc(%d) failed" % n_bytes)
    buf = (_c.c_uint8 * int(max(1, n_bytes))).from_address(p)
    arr = np.frombuffer(buf, dtype=np.uint8, count=int(n_bytes))
    _PINNED[p] = buf
    import weakref
    weakref.finalize(arr, _free_pinned, p)
    return arr


_PINNED = {}


def _free_pinned(p):
    _PINNED.pop(p, None)
    try:
        load_library().mcg_host_free(p)
    except Exception:
        pass


_default = {}
_default_lock = threading.Lock()


def default_context(device=None):
    """Process-wide context for the drop-in modules (one per device)."""
    if device is None:
        device = int(os.environ.get("MCG_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _default_lock:
        ctx = _default.get(device)
        if ctx is None:
            ctx = Context(device)
            _default[device] = ctx
        return ctx


class Pool(Context):
    """One ``mcg_pool``: every GPU of the box (or the first ``MCG_DEVICES``) behind one caller.
    The pooled calls have the signatures of the per-context ones (include/mcg_b200.h); every
    other method runs on the first device's context."""

    _POOLED = {"tetramer_scan", "featurise_score", "bin_profiles", "score_centroids",
               "score_members", "score_member_items", "graph_load", "labels_load", "labels_set",
               "bfs"}

    def __init__(self, devices=None):
        self.lib = load_library()
        if devices is None:
            self.pool = self.lib.mcg_pool_create(None, 0)
        else:
            arr = (_c.c_int * len(devices))(*[int(d) for d in devices])
            self.pool = self.lib.mcg_pool_create(arr, len(devices))
        if not self.pool:
            msg = self.lib.mcg_last_error(None)
            raise McgError(ERR_CUDA, (msg or b"mcg_pool_create failed").decode())
        self.size = int(self.lib.mcg_pool_size(self.pool))
        self.first = self.lib.mcg_pool_ctx(self.pool, 0)
        self.handle = self.first
        self.device = 0
        self.n_contigs = self.n_samples = self.n_bins = 0

    def close(self):
        if getattr(self, "pool", None):
            self.lib.mcg_pool_destroy(self.pool)
            self.pool = None
            self.handle = None

    def contexts(self):
        """Borrowed per-device handles (valid while the pool lives)."""
        out = []
        for i in range(self.size):
            c = Context.__new__(Context)
            c.lib, c.handle, c.device = self.lib, self.lib.mcg_pool_ctx(self.pool, i), i
            c.n_contigs = c.n_samples = c.n_bins = 0
            c.close = lambda: None
            out.append(c)
        return out

    def _fn(self, name):
        if name in self._POOLED:
            fn = getattr(self.lib, "mcg_pool_" + name)
            pool = self.pool

            def call(_handle, *args):
                self._pooled_call = True
                return fn(pool, *args)
            return call
        return Context._fn(self, name)

    def _last_error(self):
        if getattr(self, "_pooled_call", False):
            self._pooled_call = False
            return (self.lib.mcg_pool_last_error(self.pool) or b"").decode()
        return Context._last_error(self)

    def bounds(self):
        out = np.zeros(self.size + 1, dtype=np.int64)
        self.lib.mcg_pool_bounds(self.pool, _ptr(out))
        return out

    def set_features(self, prof=None, cov=None):
        """Feature tables of every contig, replicated on every device."""
        if prof is not None:
            prof = _arr(prof, np.float64)
            if prof.ndim != 2 or prof.shape[1] != NPROF:
                raise ValueError("profiles must be [n,136]")
            self.n_contigs = prof.shape[0]
        if cov is not None:
            cov = _arr(cov, np.float64)
            if cov.ndim != 2:
                raise ValueError("coverage must be [n,S]")
            self.n_samples = cov.shape[1]
            self._cov_rows = cov.shape[0]
        self._pooled_call = True
        self._check(self.lib.mcg_pool_set_features(
            self.pool, _ptr(prof), 0 if prof is None else prof.shape[0], _ptr(cov),
            0 if cov is None else cov.shape[0], 0 if cov is None else cov.shape[1]))

    def set_profiles(self, prof):
        self.set_features(prof=prof)

    def load_coverage(self, cov):
        self.set_features(cov=cov)


_pool = None


def default_pool():
    """Process-wide device pool of the drop-in modules.  Under torchrun (LOCAL_RANK set) every
    process owns its one device; a single process takes every device (MCG_DEVICES limits)."""
    global _pool
    with _default_lock:
        if _pool is None:
            if "LOCAL_RANK" in os.environ and "MCG_DEVICES" not in os.environ:
                _pool = Pool([int(os.environ.get("MCG_DEVICE", os.environ["LOCAL_RANK"]))])
            elif "MCG_DEVICE" in os.environ:
                _pool = Pool([int(os.environ["MCG_DEVICE"])])
            else:
                _pool = Pool()
        return _pool
