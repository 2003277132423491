"""Host ingest: the library's FASTA reader and the drop-in
get_cov_len / get_cov_len_megahit against tests/golden/ingest/expected.json, which
tests/golden/make_ingest.py produced with the reference's own functions
(feature_utils.py:122-214)."""
import json
import os

import numpy as np
import pytest

from metacoag_b200 import _lib, feature_utils


@pytest.fixture(scope="module")
def ingest(golden_dir):
    d = os.path.join(golden_dir, "ingest")
    return d, json.load(open(os.path.join(d, "expected.json")))


def _check(got, want):
    sequences, coverages, contig_lengths, n_samples = got
    assert list(sequences) == want["sequences"]
    assert len(sequences) == len(want["sequences"])
    assert {str(k): list(v) for k, v in coverages.items()} == want["coverages"]
    assert {str(k): v for k, v in contig_lengths.items()} == want["contig_lengths"]
    assert n_samples == want["n_samples"]


@pytest.mark.parametrize("min_length", [1000, 60])
def test_get_cov_len_matches_reference(ingest, min_length, monkeypatch):
    d, g = ingest
    if min_length == 60:
        monkeypatch.setenv("MCG_FASTA_RANGE_BYTES", "2000")     # ~30 parallel ranges
    got = feature_utils.get_cov_len(os.path.join(d, "contigs.fasta"), g["contig_names_rev"],
                                    min_length, os.path.join(d, "abundance.tsv"))
    _check(got, g["cases"]["plain_%d" % min_length])
    assert got[1][10 ** 9] == [0] and got[2][10 ** 9] == 0      # the reference's defaultdicts


@pytest.mark.parametrize("min_length", [1000, 60])
def test_get_cov_len_megahit_matches_reference(ingest, min_length):
    d, g = ingest
    got = feature_utils.get_cov_len_megahit(
        os.path.join(d, "contigs.fasta"), g["contig_names_rev_megahit"],
        g["graph_to_contig_map_rev"], min_length, os.path.join(d, "abundance.tsv"))
    _check(got, g["cases"]["megahit_%d" % min_length])


def test_unknown_contig_name_raises_keyerror(ingest):
    d, g = ingest
    names = dict(g["contig_names_rev"])
    names.pop(next(iter(names)))
    with pytest.raises(KeyError):
        feature_utils.get_cov_len(os.path.join(d, "contigs.fasta"), names, 1000,
                                  os.path.join(d, "abundance.tsv"))


def test_no_long_contig_exits(ingest):
    d, g = ingest
    with pytest.raises(SystemExit):
        feature_utils.get_cov_len(os.path.join(d, "contigs.fasta"), g["contig_names_rev"],
                                  10 ** 7, os.path.join(d, "abundance.tsv"))


@pytest.mark.parametrize("range_bytes", [None, "7", "64"])
def test_native_reader_matches_python_restatement(tmp_path, monkeypatch, range_bytes):
    """mcg_fasta_load against the pure-Python SimpleFastaParser restatement on hostile text;
    with MCG_FASTA_RANGE_BYTES the file is cut into many record-aligned ranges that are
    parsed in parallel (the default only does that above 8 MB)."""
    if range_bytes:
        monkeypatch.setenv("MCG_FASTA_RANGE_BYTES", range_bytes)
    rng = np.random.default_rng(3)
    pieces = [">", ">>", "> ", ">a b", "ACGT", "acgtn", " ", "  ", "\t", "\n", "\r\n", "\r",
              "\n\n", "NNNN", ">x\ty z", "A C G T", "GATTACA" * 30, "\x0b", ";c"]
    for trial in range(300):
        text = "".join(rng.choice(pieces, rng.integers(0, 40)))
        path = tmp_path / ("t%d.fa" % trial)
        path.write_bytes(text.encode("latin-1"))
        want = list(feature_utils._fasta_records(str(path)))
        names, blob, offsets = _lib.read_fasta(str(path))
        got = [(names[i], blob[offsets[i]:offsets[i + 1]].tobytes().decode("latin-1"))
               for i in range(len(names))]
        assert got == want, repr(text)


def test_missing_file():
    with pytest.raises(FileNotFoundError):
        _lib.read_fasta("/nonexistent/contigs.fasta")


def test_contig_blob_behaves_like_the_list():
    blob = np.frombuffer(b"ACGTTTGGA", dtype=np.uint8)
    cb = feature_utils.ContigBlob(blob, [0, 4, 4, 9])
    assert len(cb) == 3 and list(cb) == ["ACGT", "", "TTGGA"]
    assert cb[0] == "ACGT" and cb[-1] == "TTGGA" and cb[1:] == ["", "TTGGA"]
    with pytest.raises(IndexError):
        cb[3]
    assert not cb.needs_strip()
    assert feature_utils.ContigBlob(np.frombuffer(b"\tACGT", dtype=np.uint8), [0, 5]).needs_strip()
    b2, o2 = feature_utils._ascii_blob(cb)
    assert b2 is blob and np.array_equal(o2, [0, 4, 4, 9])
