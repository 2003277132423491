#!/usr/bin/env python3
"""Generate tests/golden/writer.json from the LIVE reference (build container).

The result writers live inline in the reference CLI's main() (metacoag:1180-1257), so they
cannot be imported.  This script reads that block out of /root/reference/metacoag at run time
(from ``final_bins = {}`` to the last ``close()``), dedents it and executes it unchanged with
hand-made inputs over tests/golden/ingest/contigs.fasta (the hostile FASTA of make_ingest.py),
for the spades and the megahit flavour, and stores every output file's text.  Nothing of the
reference's source is written into this repository.
"""
import csv
import json
import logging
import os
import sys
import tempfile
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from _refload import REFERENCE_ROOT, load_reference, _parse_fasta  # noqa: E402


def reference_block():
    lines = open(os.path.join(REFERENCE_ROOT, "metacoag")).read().split("\n")
    first = next(i for i, l in enumerate(lines) if l.strip() == "final_bins = {}")
    last = max(i for i, l in enumerate(lines) if l.strip() == "bin_files[c].close()")
    return textwrap.dedent("\n".join(lines[first:last + 1]))


def inputs(assembler):
    fasta = os.path.join(HERE, "ingest", "contigs.fasta")
    ids = [r.id for r in _parse_fasta(fasta, "fasta")]
    n = len(ids)
    if assembler == "megahit":
        graph_to_contig_map_rev = {name: "k141_%d" % (7 * i + 3) for i, name in enumerate(ids)}
        graph_to_contig_map = {g: c for c, g in graph_to_contig_map_rev.items()}
        contig_names = {i: graph_to_contig_map_rev[name] for i, name in enumerate(ids)}
        contig_descriptions = {name: name + " flag=1 multi=2.0 len=%d" % i for i, name in enumerate(ids)}
    else:
        graph_to_contig_map_rev = graph_to_contig_map = contig_descriptions = None
        contig_names = {i: name for i, name in enumerate(ids)}
    contig_names_rev = {v: k for k, v in contig_names.items()}
    # six bins over the records (some records in no bin); bin 4 is marked for removal, bin 5 is
    # too small, bins 0 and 1 form one clique
    bins = {b: [i for i in range(n) if i % 7 == b] for b in range(6)}
    bins[2] = list(reversed(bins[2]))
    bin_cliques = [(0, 1), (2,), (3,), (4,), (5,)]
    bins_to_rem = [4]
    bin_clique_size = {"0_1": 500000, "2": 300000, "3": 200000, "4": 900000, "5": 10}
    return dict(contigs_file=fasta, bins=bins, bin_cliques=bin_cliques, bins_to_rem=bins_to_rem,
                bin_clique_size=bin_clique_size, min_bin_size=200000, assembler=assembler,
                contig_names=contig_names, contig_names_rev=contig_names_rev,
                graph_to_contig_map=graph_to_contig_map,
                graph_to_contig_map_rev=graph_to_contig_map_rev,
                contig_descriptions=contig_descriptions, prefix="run1_", delimiter="\t")


def main():
    load_reference()
    import Bio.SeqIO as SeqIO
    block = reference_block()
    out = {}
    for assembler in ("spades", "megahit"):
        with tempfile.TemporaryDirectory() as tmp:
            ns = inputs(assembler)
            ns.update(output_path=tmp, output_bins_path=f"{tmp}/run1_bins",
                      lq_output_bins_path=f"{tmp}/run1_low_quality_bins", csv=csv, SeqIO=SeqIO,
                      tqdm=lambda it, **kw: it, logger=logging.getLogger("golden"),
                      final_bin_count=0)
            os.makedirs(ns["output_bins_path"])
            os.makedirs(ns["lq_output_bins_path"])
            exec(compile(block, "metacoag[writer block]", "exec"), ns)
            files = {}
            for root, _, names in os.walk(tmp):
                for name in names:
                    path = os.path.join(root, name)
                    files[os.path.relpath(path, tmp)] = open(path, newline="").read()
            spec = inputs(assembler)
            spec.pop("contigs_file")
            out[assembler] = {"inputs": spec, "files": files, "final_bin_count": ns["final_bin_count"],
                              "final_bins": {str(k): v for k, v in ns["final_bins"].items()},
                              "lowq_bins": {str(k): v for k, v in ns["lowq_bins"].items()}}
    with open(os.path.join(HERE, "writer.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    print({k: sorted(v["files"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
