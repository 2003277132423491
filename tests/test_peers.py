"""mcg_push_to_peers (the multi-GPU exchange as stores into peer memory) on
one GPU: every "peer" is a buffer of the same device, which exercises the kernel and the
argument checks; bench.py checks the real thing against the NCCL all-gather on every N > 1 run."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from metacoag_b200 import _lib  # noqa: E402

pytestmark = pytest.mark.gpu


def test_push_to_peers_copies_the_block_everywhere():
    import torch
    dev = torch.device("cuda", 0)
    ctx = _lib.Context(0)
    try:
        rng = np.random.default_rng(3)
        for words, shift in ((4096, 0), (4099, 0), (1, 0), (4096, 4), (12, 8), (1 << 20, 0)):
            src_all = torch.from_numpy(rng.integers(0, 2 ** 31, words + 8, dtype=np.int64)
                                       .astype(np.int32)).to(dev)
            src = src_all[shift // 4:shift // 4 + words]
            peers = [torch.zeros(words + 8, dtype=torch.int32, device=dev) for _ in range(5)]
            torch.cuda.synchronize()
            # destinations at different alignments: 16-byte and 4-byte paths
            offs = [0, 0, 1, 2, 4] if shift == 0 else [1, 1, 1, 1, 1]
            ctx.push_to_peers(src.data_ptr(), 4 * words,
                              [p.data_ptr() + 4 * o for p, o in zip(peers, offs)])
            ctx.synchronize()
            for p, o in zip(peers, offs):
                assert torch.equal(p[o:o + words], src)
                assert int(p[:o].abs().sum()) == 0 and int(p[o + words:].abs().sum()) == 0
        with pytest.raises(_lib.McgError):
            ctx.push_to_peers(src.data_ptr(), 6, [peers[0].data_ptr()])          # not whole words
        with pytest.raises(_lib.McgError):
            ctx.push_to_peers(src.data_ptr(), 8, [peers[0].data_ptr()] * 17)     # too many peers
        with pytest.raises(_lib.McgError):
            ctx.push_to_peers(src.data_ptr(), 8, [peers[0].data_ptr(), 0])       # a missing buffer
    finally:
        ctx.close()
