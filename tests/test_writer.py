"""Result writers against the outputs of the reference CLI's own writer
block (metacoag:1180-1257), executed unchanged by tests/golden/make_writer.py on the hostile
ingest FASTA.  Host-only code: runs without a GPU."""
import json
import os

import numpy as np
import pytest

from metacoag_b200 import _lib, output_utils

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "writer.json")))
FASTA = os.path.join(HERE, "golden", "ingest", "contigs.fasta")


def _int_keys(d):
    return None if d is None else {int(k): v for k, v in d.items()}


@pytest.mark.parametrize("assembler", ["spades", "megahit"])
def test_writers_match_reference_block(tmp_path, assembler):
    gold = GOLD[assembler]
    spec = gold["inputs"]
    bins = {int(b): v for b, v in spec["bins"].items()}
    out = str(tmp_path)
    bins_path, lq_path = output_utils.make_output_dirs(out, spec["prefix"])
    final_bins, lowq_bins, count = output_utils.write_contig_to_bin(
        out, spec["prefix"], spec["delimiter"], [tuple(c) for c in spec["bin_cliques"]], bins,
        spec["bins_to_rem"], spec["bin_clique_size"], spec["min_bin_size"],
        _int_keys(spec["contig_names"]), assembler=assembler,
        contig_descriptions=spec["contig_descriptions"],
        graph_to_contig_map=spec["graph_to_contig_map"])
    assert count == gold["final_bin_count"]
    assert final_bins == _int_keys(gold["final_bins"])
    assert lowq_bins == _int_keys(gold["lowq_bins"])
    output_utils.write_bin_fastas(FASTA, final_bins, lowq_bins, bins_path, lq_path,
                                  spec["prefix"], spec["contig_names_rev"], assembler=assembler,
                                  graph_to_contig_map_rev=spec["graph_to_contig_map_rev"])
    found = {}
    for root, _, names in os.walk(out):
        for name in names:
            path = os.path.join(root, name)
            found[os.path.relpath(path, out)] = open(path, newline="").read()
    assert sorted(found) == sorted(gold["files"])
    for name, text in gold["files"].items():
        assert found[name] == text, name


def test_writer_edge_cases(tmp_path):
    names = ["a", "", "c c", "d"]
    seqs = [b"ACGT", b"", b"nnNN", b"T" * 100000]
    offsets = np.zeros(5, dtype=np.int64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    blob = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    paths = [str(tmp_path / ("f%d.fa" % i)) for i in range(3)]
    open(paths[2], "w").write("stale contents that must be truncated")
    _lib.write_bin_fastas(names, blob, offsets, [1, 1, -1, 0], paths)
    assert open(paths[0]).read() == ">d\n" + "T" * 100000 + "\n"
    assert open(paths[1]).read() == ">a\nACGT\n>\n\n"
    assert open(paths[2]).read() == ""
    with pytest.raises(ValueError):
        _lib.write_bin_fastas(names, blob, offsets, [3, 0, 0, 0], paths)
    with pytest.raises(OSError):
        _lib.write_bin_fastas(names, blob, offsets, [0, 0, 0, 0], [str(tmp_path / "no" / "dir.fa")])
    _lib.write_bin_fastas([], np.zeros(0, np.uint8), np.zeros(1, np.int64), [], [])
