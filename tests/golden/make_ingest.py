#!/usr/bin/env python3
"""Generate tests/golden/ingest/* from the LIVE reference (build container).

A small FASTA / abundance pair with the awkward cases (CRLF and bare-CR line ends, blank
lines, text before the first record, lower case and N, spaces and trailing blanks inside
sequence lines, a header with a description, an empty record, no final newline; abundance
values below the 1e-4 clamp, exponents, a contig listed twice, contigs below min_length) is
pushed through the reference's get_cov_len and get_cov_len_megahit (feature_utils.py:122-214,
Bio.SeqIO replaced by the stand-in of _refload.py) and the results are stored as JSON.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refload import load_reference  # noqa: E402

OUT = os.path.join(HERE, "ingest")


def main():
    fu, _, _ = load_reference()
    rng = np.random.default_rng(31)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)

    def seq(n):
        return alphabet[rng.integers(0, 4, n)].tobytes().decode()

    records = []
    for i in range(40):
        n = int(rng.choice([0, 3, 4, 59, 60, 61, 700, 1000, 1001, 2500, 6000]))
        records.append(("NODE_%d_length_%d_cov_%.3f" % (i + 1, n, rng.uniform(1, 50)), seq(n)))
    parts = ["; a comment line before the first record\n", "\n"]
    for i, (name, s) in enumerate(records):
        style = i % 8
        header = ">" + name + (" some description here" if style == 1 else "") + ("   " if style == 2 else "")
        eol = "\r\n" if style == 3 else ("\r" if style == 4 else "\n")
        parts.append(header + eol)
        width = [60, 80, 70, 60, 60, 1 << 20, 61, 60][style]
        body = s
        if style == 5:
            body = s.lower()
        if style == 6 and len(s) > 100:
            body = s[:50] + "NNNNnnnn" + s[58:]
            records[i] = (name, body)
        if style == 5:
            records[i] = (name, body)
        for k in range(0, len(body), width):
            line = body[k:k + width]
            if style == 7 and len(line) > 20:
                line = line[:10] + " " + line[10:] + " \t"     # inner space, trailing blanks
            parts.append(line + eol)
        if style == 0:
            parts.append(eol)                                     # blank line after the record
    text = "".join(parts)
    text = text.rstrip("\n")                                      # no final newline
    with open(os.path.join(OUT, "contigs.fasta"), "w", newline="") as fh:
        fh.write(text)

    names = [n for n, _ in records]
    contig_names_rev = {n: i for i, n in enumerate(names)}
    graph_names = {n: "k141_%d" % (7 * i + 3) for i, n in enumerate(names)}   # FASTA name -> graph name
    contig_names_rev_megahit = {g: i for i, g in enumerate(graph_names.values())}
    lines = []
    for i, n in enumerate(names):
        vals = rng.lognormal(2.0, 1.5, 3)
        cells = ["%.4f" % v for v in vals]
        if i % 5 == 0:
            cells[0] = "0.0000"
        if i % 7 == 0:
            cells[1] = "3e-05"
        if i % 9 == 0:
            cells[2] = "1.25E+2"
        lines.append(n + "\t" + "\t".join(cells) + "\n")
    lines.append(lines[17])                                       # a contig listed twice
    with open(os.path.join(OUT, "abundance.tsv"), "w") as fh:
        fh.writelines(lines)

    golden = {"contig_names_rev": contig_names_rev, "graph_to_contig_map_rev": graph_names,
              "contig_names_rev_megahit": contig_names_rev_megahit, "cases": {}}
    for min_length in (1000, 60):
        s, cov, lens, ns = fu.get_cov_len(os.path.join(OUT, "contigs.fasta"), contig_names_rev,
                                          min_length, os.path.join(OUT, "abundance.tsv"))
        golden["cases"]["plain_%d" % min_length] = {
            "sequences": list(s), "coverages": {str(k): v for k, v in cov.items()},
            "contig_lengths": {str(k): v for k, v in lens.items()}, "n_samples": ns}
        s, cov, lens, ns = fu.get_cov_len_megahit(
            os.path.join(OUT, "contigs.fasta"), contig_names_rev_megahit, graph_names, min_length,
            os.path.join(OUT, "abundance.tsv"))
        golden["cases"]["megahit_%d" % min_length] = {
            "sequences": list(s), "coverages": {str(k): v for k, v in cov.items()},
            "contig_lengths": {str(k): v for k, v in lens.items()}, "n_samples": ns}
    with open(os.path.join(OUT, "expected.json"), "w") as fh:
        json.dump(golden, fh)
    c = golden["cases"]["plain_1000"]
    print("ingest golden: %d records, %d with coverage, %d samples; total bases %d" % (
        len(c["sequences"]), len(c["coverages"]), c["n_samples"], sum(len(x) for x in c["sequences"])))


if __name__ == "__main__":
    main()
