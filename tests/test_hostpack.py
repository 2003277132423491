"""The host half of the contig ingest (metacoag_b200/csrc/mcg_hostpack.cpp): every SIMD
variant against a numpy restatement of the MCG_ENC_2BIT layout (include/mcg_b200.h), and the
hybrid upload against the copy-engine-only upload on the GPU."""
import numpy as np
import pytest

from metacoag_b200 import _lib


def numpy_pack(blob, offsets):
    """feature_utils.py:21,31-35: A0 C1 G2 T3, anything else 0; 16 bases per little-endian
    32-bit word, every contig zero-padded to 64 bases."""
    lut = np.zeros(256, dtype=np.uint64)
    lut[ord("C")], lut[ord("G")], lut[ord("T")] = 1, 2, 3
    out = []
    for c in range(len(offsets) - 1):
        codes = lut[blob[offsets[c]:offsets[c + 1]]]
        pad = np.zeros(64 * ((codes.shape[0] + 63) // 64), dtype=np.uint64)
        pad[:codes.shape[0]] = codes
        words = (pad.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint64))).sum(axis=1)
        out.append(words.astype(np.uint32))
    return np.concatenate(out).view(np.uint8) if out else np.zeros(0, dtype=np.uint8)


def random_contigs(seed, lens):
    rng = np.random.default_rng(seed)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    p = np.array([.2, .2, .2, .2, .03, .03, .03, .03, .02, .02, .02, .02])
    blob = rng.choice(np.frombuffer(b"ACGTacgtNnX-", dtype=np.uint8), int(offsets[-1]), p=p / p.sum())
    return blob, offsets


@pytest.mark.parametrize("level", [-1, 0, 1, 2, 3])
def test_host_pack_variants(level):
    rng = np.random.default_rng(5)
    lens = np.concatenate([np.arange(0, 200), rng.integers(1, 5000, 300), [0, 64, 128, 1 << 16]])
    blob, offsets = random_contigs(6, lens)
    assert np.array_equal(_lib.pack_2bit_host(blob, offsets, level), numpy_pack(blob, offsets))


def test_host_pack_empty():
    assert _lib.pack_2bit_host(np.zeros(0, np.uint8), np.zeros(1, np.int64)).shape == (0,)


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 3, 16])
def test_hybrid_upload_matches_copy_engine_upload(threads):
    """>= 16 MB of ASCII: host threads pack from the back while the copy engine takes the
    front; the tetramer counts must not depend on who packed what."""
    rng = np.random.default_rng(threads)
    lens = np.concatenate([rng.integers(1, 60000, 1200), [3_000_000, 5, 0, 63, 64, 65]])
    blob, offsets = random_contigs(threads, lens)
    assert blob.shape[0] > (16 << 20)
    ctx = _lib.Context(0)
    try:
        ctx.set_host_pack_threads(0)
        ctx.load_contigs(blob, offsets)
        assert ctx.ingest_stats()["host_packed_bases"] == 0
        ctx.scan()
        want = ctx.get_profiles(want_counts=True)[0]
        ctx.set_host_pack_threads(threads)
        for _ in range(3):   # the split is timing-dependent: try it more than once
            ctx.load_contigs(blob, offsets)
            st = ctx.ingest_stats()
            assert st["host_threads"] == threads
            assert st["host_packed_bases"] + st["gpu_packed_bases"] == blob.shape[0]
            ctx.scan()
            got = ctx.get_profiles(want_counts=True)[0]
            assert np.array_equal(got, want)
        assert np.array_equal(want.sum(axis=1), np.maximum(lens - 3, 0))
    finally:
        ctx.close()
