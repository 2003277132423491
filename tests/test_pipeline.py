"""The drop-in modules against the reference, stage by stage, on the golden pipeline.

tests/golden/pipeline.{npz,json} hold a synthetic assembly and the state the REFERENCE's own
functions produced after every binning stage (tests/golden/make_pipeline.py).  The same
driver runs metacoag_b200's modules:
  - test_host_logic (no GPU): with the oracle-backed back end -- checks the host-side
    control flow (matching, heap loops, marker rules, dict staging) bit for bit;
  - test_gpu_pipeline (GPU): with the real back end -- identical bins end to end.
"""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

import pipeline_driver  # noqa: E402
from _refload import ShimGraph  # noqa: E402  (a plain-Python graph; nothing is imported from the reference)

from metacoag_b200 import _backend, feature_utils, label_prop_utils, matching_utils  # noqa: E402


def load_case(golden_dir):
    data = np.load(os.path.join(golden_dir, "pipeline.npz"))
    meta = json.load(open(os.path.join(golden_dir, "pipeline.json")))
    n = data["lengths"].shape[0]
    blob, offsets = data["blob"], data["offsets"]
    sequences = [blob[offsets[i]:offsets[i + 1]].tobytes().decode("ascii") for i in range(n)]
    contig_lengths = {i: int(data["lengths"][i]) for i in range(n)}
    min_length = meta["min_length"]
    coverages = {i: [float(x) for x in data["coverage"][i]] for i in range(n)
                 if contig_lengths[i] >= min_length}
    markers = {int(k): v for k, v in meta["markers"].items()}
    graph = ShimGraph()
    graph.add_vertices(n)
    graph.add_edges([(int(a), int(b)) for a, b in data["edges"]])
    graph.simplify()
    return sequences, coverages, contig_lengths, markers, graph, meta


def run(golden_dir, tmp_path):
    sequences, coverages, contig_lengths, markers, graph, meta = load_case(golden_dir)
    states = pipeline_driver.run_pipeline(feature_utils, matching_utils, label_prop_utils,
                                          sequences, coverages, contig_lengths, markers, graph,
                                          meta["n_samples"], str(tmp_path) + "/",
                                          min_length=meta["min_length"])
    return states, meta["states"]


def compare(states, want, exact_floats):
    for stage in ("profile_ids",):
        assert states[stage] == want[stage]
    for stage in ("initial", "match_contigs", "further_match_contigs", "label_prop",
                  "assign_to_bins", "final_label_prop"):
        got = {str(b): m for b, m in states[stage].items()}
        assert got == want[stage], stage
    assert {str(c): b for c, b in states["bin_of_contig"].items()} == want["bin_of_contig"]
    # centroids: bit-exact (feature_utils.get_bin_profiles is adds and one division)
    for b, (prof, cov) in states["centroids"].items():
        wp, wc = want["centroids"][str(b)]
        assert prof == wp and cov == wc, b
    # assign_long: same contig, same bin, weight within 1e-9 (bit-exact with the oracle)
    assert len(states["assign_long"]) == len(want["assign_long"])
    for (c, b, w), (wc_, wb, ww) in zip(states["assign_long"], want["assign_long"]):
        assert (c, b) == (wc_, wb)
        if exact_floats:
            assert w == ww
        else:
            assert abs(w - ww) <= 1e-9 * abs(ww)


def test_host_logic(golden_dir, tmp_path):
    from oracle_backend import OracleBackend
    old = _backend.set_backend(OracleBackend())
    try:
        states, want = run(golden_dir, tmp_path)
    finally:
        _backend.set_backend(old)
    compare(states, want, exact_floats=True)
    grew = sum(len(m) for m in states["final_label_prop"].values()) - \
        sum(len(m) for m in states["assign_to_bins"].values())
    assert grew > 0 and len(states["match_contigs"]) > len(states["initial"])


def test_pickle_cache_roundtrip(golden_dir, tmp_path):
    """get_tetramer_profiles writes the reference's cache file (same name, same dict) and a
    second call reads it back instead of scanning (feature_utils.py:79-111)."""
    import pickle
    from oracle_backend import OracleBackend
    old = _backend.set_backend(OracleBackend())
    try:
        seqs = ["ACGTACGTTTGACCA" * 80, "ACG", "GGGCCCAAATTT" * 100]
        lengths = {0: len(seqs[0]), 1: 3, 2: len(seqs[2])}
        out = str(tmp_path) + "/pre"
        first = feature_utils.get_tetramer_profiles(out, seqs, "x/y/contigs.fa", lengths, 1000, 4)
        cache = out + "contigs.fa.normalized_contig_tetramers.pickle"
        assert os.path.isfile(cache)
        assert sorted(first) == [0, 2]
        stored = pickle.load(open(cache, "rb"))
        assert sorted(stored) == [0, 1, 2] and stored[0].shape == (136,)
        _backend.set_backend(None.__class__())   # any scan would now fail
        again = feature_utils.get_tetramer_profiles(out, seqs, "x/y/contigs.fa", lengths, 1000, 4)
        assert all(np.array_equal(first[k], again[k]) for k in first)
    finally:
        _backend.set_backend(old)


@pytest.mark.gpu
def test_gpu_pipeline(golden_dir, tmp_path):
    _backend.set_backend(None)      # the real back end
    states, want = run(golden_dir, tmp_path)
    compare(states, want, exact_floats=False)


@pytest.mark.gpu
def test_gpu_scalar_primitives_match_reference_values():
    """The one-value-per-call API the CLI's bin-merge loop uses (metacoag:1108-1125)."""
    _backend.set_backend(None)
    assert abs(matching_utils.get_comp_probability(0.0) - 0.9790291495280203) < 1e-12
    assert matching_utils.get_comp_probability(0.21) == 0.0
    with pytest.raises(ZeroDivisionError):
        matching_utils.get_comp_probability(1.40)
    with pytest.raises(ValueError):
        matching_utils.get_cov_probability([0.0, 1.0], [1.0, 1.0])
    got = matching_utils.normpdf(0.01, matching_utils.MU_INTRA, matching_utils.SIGMA_INTRA)
    want = np.exp(-(0.01 ** 2) / (2 * matching_utils.SIGMA_INTRA ** 2)) / (
        matching_utils.SIGMA_INTRA * (2 * np.pi) ** 0.5)
    assert abs(got - want) <= 1e-12 * want
    assert matching_utils.get_tetramer_distance([0.0, 3.0], [4.0, 0.0]) == 5.0
