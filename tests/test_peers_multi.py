"""The peer-memory exchange on REAL peers: one process per GPU of the box (skipped below two
devices), checked against the NCCL all-gather / broadcast it replaces (tests/peer_worker.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from metacoag_b200 import _lib  # noqa: E402

pytestmark = pytest.mark.gpu


def test_peer_exchange_matches_nccl_on_every_gpu_of_the_box():
    n = _lib.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
         "--master-addr", "127.0.0.1", "--master-port", "29541",
         os.path.join(ROOT, "tests", "peer_worker.py")],
        env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    out = proc.stdout
    assert proc.returncode == 0, out[-4000:]
    if "PEER-UNAVAILABLE" in out:
        pytest.skip("no peer access between the GPUs of this box")
    assert "PEER-OK" in out, out[-4000:]
