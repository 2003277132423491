"""TEST INFRASTRUCTURE: a synthetic assembly written out as the files the ``metacoag`` CLI reads
(SURVEY.md section 8d): contigs FASTA (80-column lines, lower case and N runs included), GFA
assembly graph, abundance TSV (``%.4f``, some exact zeros -> the 1e-4 clamp), and a
``{contigs}.hmmout`` in HMMER's domtblout layout so that the CLI never looks for FragGeneScan /
hmmsearch (metacoag:568).  Flavours follow the CLI's ``--assembler`` (metacoag:337-413):

  custom   ``S`` / ``L`` lines name the contigs directly (graph_utils.py:344-378)
  spades   ``contigs.paths`` (one single-segment path per contig, forward and reverse entry) +
           ``L`` lines between the segments (graph_utils.py:12-112)
  megahit  FASTA ids ``k141_<i>`` with ``NODE_<i>_...`` graph names in the GFA
           (graph_utils.py:283-341; exercises get_cov_len_megahit)

Both tests/golden/make_cli.py (reference run, build container) and tests/test_cli.py (drop-in
run) call ``write_dataset`` with the same arguments; the golden file records a SHA-256 of the
written inputs so that a drifting generator fails loudly instead of comparing different data.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from metacoag_b200 import synth  # noqa: E402

FLAVOURS = {
    # flavour: (seed, genomes, contigs, samples)
    "custom": (11, 5, 220, 3),
    "spades": (12, 6, 260, 2),
    "megahit": (13, 4, 200, 4),
}


def _names(flavour, asm):
    n = asm.n_contigs
    if flavour == "custom":
        return ["ctg%04d" % i for i in range(n)], None
    if flavour == "spades":
        return ["NODE_%d_length_%d_cov_%.4f" % (i + 1, asm.lengths[i], 10.0 + i % 7)
                for i in range(n)], None
    # megahit: FASTA ids k141_<i>; the graph names the same contigs NODE_<i>_length_.._ID_..
    fasta = ["k141_%d" % i for i in range(n)]
    graph = ["NODE_%d_length_%d_cov_1.0000_ID_%d" % (i + 1, asm.lengths[i], 2 * i + 1)
             for i in range(n)]
    return fasta, graph


def write_dataset(dirpath, flavour):
    seed, n_genomes, n_contigs, n_samples = FLAVOURS[flavour]
    rng = np.random.default_rng(seed)
    lengths = np.exp(rng.uniform(np.log(500), np.log(60000), n_contigs)).astype(np.int64)
    asm = synth.make_assembly(seed=seed, n_genomes=n_genomes, n_contigs=n_contigs,
                              n_samples=n_samples, lengths=lengths, n_frac=2e-4, lower_frac=0.02)
    names, graph_names = _names(flavour, asm)
    os.makedirs(dirpath, exist_ok=True)
    files = {}

    contigs = os.path.join(dirpath, "final.contigs.fa" if flavour == "megahit" else "contigs.fasta")
    with open(contigs, "w") as fh:
        for i, name in enumerate(names):
            seq = asm.sequence(i)
            extra = " flag=1 multi=2.0 len=%d" % len(seq) if flavour == "megahit" else ""
            fh.write(">%s%s\n" % (name, extra))
            for k in range(0, len(seq), 80):
                fh.write(seq[k:k + 80] + "\n")
    files["contigs"] = contigs

    abundance = os.path.join(dirpath, "abundance.tsv")
    with open(abundance, "w") as fh:
        for i, name in enumerate(names):
            fh.write(name + "\t" + "\t".join(asm.coverage_text[i]) + "\n")
    files["abundance"] = abundance

    graph = os.path.join(dirpath, "graph.gfa")
    with open(graph, "w") as fh:
        if flavour == "custom":
            for name in names:
                fh.write("S\t%s\t*\n" % name)
            for a, b in asm.edges:
                fh.write("L\t%s\t+\t%s\t+\t0M\n" % (names[a], names[b]))
        elif flavour == "spades":
            for i in range(n_contigs):
                fh.write("S\t%d\t*\n" % (i + 1))
            for a, b in asm.edges:
                fh.write("L\t%d\t+\t%d\t+\t0M\n" % (a + 1, b + 1))
        else:
            for i, gname in enumerate(graph_names):
                fh.write("S\t%s\t%s\n" % (gname, asm.sequence(i)))
            for a, b in asm.edges:
                fh.write("L\t%s\t+\t%s\t+\t0M\n" % (graph_names[a], graph_names[b]))
    files["graph"] = graph

    if flavour == "spades":
        paths = os.path.join(dirpath, "contigs.paths")
        with open(paths, "w") as fh:
            for i, name in enumerate(names):
                fh.write("%s\n%d+\n%s'\n%d-\n" % (name, i + 1, name, i + 1))
        files["paths"] = paths

    hmmout = contigs + ".hmmout"
    with open(hmmout, "w") as fh:
        fh.write("# synthetic domtblout (tests/cli_data.py)\n")
        for c in sorted(asm.contig_markers):
            for k, gene in enumerate(asm.contig_markers[c]):
                target = "%s_%d_%d_+" % (names[c], 10 + 3 * k, 400 + 3 * k)
                # a short hit on every seventh gene: dropped by mg_threshold (hmm to - hmm from)
                hmm_to = 250 if k % 7 else 90
                fh.write("%s - 300 %s PF%05d 260 1e-30 100.0 0.0 1 1 1e-30 1e-30 100.0 0.0 "
                         "1 %d 1 300 1 300 0.95 synthetic marker\n" % (target, gene, k, hmm_to))
    files["hmmout"] = hmmout

    digest = hashlib.sha256()
    for key in sorted(files):
        with open(files[key], "rb") as fh:
            digest.update(fh.read())
    files["sha256"] = digest.hexdigest()
    files["n_contigs"] = n_contigs
    files["names"] = names
    return files


def cli_args(flavour, files, output, nthreads=4):
    args = ["--assembler", flavour, "--graph", files["graph"], "--contigs", files["contigs"],
            "--abundance", files["abundance"], "--output", output, "--nthreads", str(nthreads)]
    if flavour == "spades":
        args += ["--paths", files["paths"]]
    return args


def read_result(output):
    """(contig name -> bin name from contig_to_bin.tsv, sorted record counts of the per-bin FASTA
    files) -- what tests/test_metacoag.py:72-79 of the reference looks at, plus the assignment."""
    import csv
    assign = {}
    with open(os.path.join(output, "contig_to_bin.tsv")) as fh:
        for row in csv.reader(fh, delimiter=","):
            if row:
                assign[row[0]] = row[1]
    counts = []
    bins_dir = os.path.join(output, "bins")
    for f in sorted(os.listdir(bins_dir)) if os.path.isdir(bins_dir) else []:
        with open(os.path.join(bins_dir, f)) as fh:
            counts.append(sum(1 for line in fh if line.startswith(">")))
    return assign, sorted(counts)


def partition(assign):
    """Bin names depend on clique enumeration order; the partition does not."""
    groups = {}
    for contig, b in assign.items():
        groups.setdefault(b, []).append(contig)
    return sorted(sorted(g) for g in groups.values())
