"""Bin x bin merge scoring (inline in the reference's CLI,
metacoag:1090-1147) against tests/golden/merge.json, which tests/golden/make_merge.py
produced with the reference's own matching_utils functions: same edges in the same order,
same removals, for four thresholds."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from metacoag_b200 import _backend, matching_utils  # noqa: E402


def load(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "merge.json")))
    bins = {int(b): v for b, v in g["bins"].items()}
    markers = {int(b): v for b, v in g["bin_markers"].items()}
    prof = {int(b): np.array(v) for b, v in g["seed_prof"].items()}
    cov = {int(b): np.array(v) for b, v in g["seed_cov"].items()}
    return g, bins, markers, prof, cov


def check(golden_dir):
    g, bins, markers, prof, cov = load(golden_dir)
    n_edges = 0
    for case in g["cases"]:
        edges, to_rem = matching_utils.bin_merge_candidates(
            bins, markers, prof, cov, case["w_intra"], g["n_mg"], g["bin_mg_threshold"])
        assert [list(e) for e in edges] == case["edges"]
        assert to_rem == case["bins_to_rem"]
        n_edges += len(edges)
    assert n_edges > 0


def test_merge_host_logic(golden_dir):
    from oracle_backend import OracleBackend
    old = _backend.set_backend(OracleBackend())
    try:
        check(golden_dir)
    finally:
        _backend.set_backend(old)


def test_merge_rejects_zero_coverage(golden_dir):
    from oracle_backend import OracleBackend
    g, bins, markers, prof, cov = load(golden_dir)
    cov[0] = np.array(cov[0])
    cov[0][1] = 0.0
    old = _backend.set_backend(OracleBackend())
    try:
        with pytest.raises(ValueError):     # math.log(0) in get_cov_probability
            matching_utils.bin_merge_candidates(bins, markers, prof, cov, 5.0, 108, 0.33333)
    finally:
        _backend.set_backend(old)


@pytest.mark.gpu
def test_merge_gpu(golden_dir):
    _backend.set_backend(None)
    check(golden_dir)
