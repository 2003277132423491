"""What the shipped library is made of, read from its SASS (`cuobjdump -sass`, no GPU needed):
the distance bound really runs on the 5th-generation tensor cores with TMA operands and tensor
memory, the float64 paths on DMMA / DFMA, and nothing was compiled for another architecture.
The full per-kernel histogram is `tools/sass_histogram.py` -> profiles/r02_sass_histogram.txt."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from metacoag_b200 import build as mcg_build  # noqa: E402


@pytest.fixture(scope="module")
def sass():
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    lib = mcg_build.build_library()
    out = subprocess.run([tool, "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), [])
        elif cur is not None:
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                cur.append(m.group(1))
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    return kernels, archs


def ops(kernels, name):
    hits = [v for k, v in kernels.items() if name in k]
    assert hits, "no kernel named *%s*" % name
    return [op for v in hits for op in v]


def test_only_sm_100a(sass):
    assert sass[1] == {"sm_100a"}


def test_distance_bound_is_tcgen05_tma_tmem(sass):
    tc = ops(sass[0], "tc_dist_kernel")
    assert any(op.startswith("UTCHMMA") for op in tc)          # tcgen05.mma kind::f16
    assert any(op.startswith("UTMALDG") for op in tc)          # cp.async.bulk.tensor (TMA loads)
    assert any(op.startswith("LDTM") for op in tc)             # tcgen05.ld from tensor memory
    assert any(op.startswith("UTCBAR") for op in tc)           # tcgen05.commit -> mbarrier
    assert any(op.startswith("SYNCS") for op in tc)            # mbarrier try_wait / arrive
    assert not any(op.startswith(("HMMA", "IMMA")) for op in tc)   # no legacy mma.sync in it


def test_float64_paths(sass):
    assert any(op.startswith("DMMA") for op in ops(sass[0], "dist2_kernel"))
    assert any(op.startswith("DMMA") for op in ops(sass[0], "row_norms_kernel"))
    assert any(op.startswith("DFMA") for op in ops(sass[0], "cand_score_kernel"))
    scan = ops(sass[0], "scan_kernel")
    assert sum(op.startswith("LDS") for op in scan) > 50 and not any(op.startswith("ATOMS") for op in scan)
