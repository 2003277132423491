"""The result writers of the metacoag CLI (``metacoag:1180-1251``), which the reference keeps
inline in ``main()``: ``contig_to_bin.tsv`` and one FASTA file per final / low-quality bin.

The reference parses the contigs file a second time with ``Bio.SeqIO`` and writes record by
record through Python file objects.  Here the records are the ones ``feature_utils.get_cov_len``
already staged (``ContigBlob`` + names): the library gathers every bin's
records from the blob and writes the files in parallel (``mcg_write_bin_fastas``, host code).
"""
import csv
import os

import numpy as np

from . import _lib
from .feature_utils import ContigBlob, _read_contigs


def write_contig_to_bin(output_path, prefix, delimiter, bin_cliques, bins, bins_to_rem,
                        bin_clique_size, min_bin_size, contig_names, assembler="spades",
                        contig_descriptions=None, graph_to_contig_map=None):
    """``{output_path}/{prefix}contig_to_bin.tsv`` and the two contig -> bin-name maps
    (metacoag:1180-1229): a clique of merged bins is written when it is not a lone bin marked
    for removal and its size reaches ``min_bin_size``; everything else is low quality.
    Returns ``(final_bins, lowq_bins, final_bin_count)``."""
    final_bins, lowq_bins = {}, {}
    final_bin_count = 0
    with open(f"{output_path}/{prefix}contig_to_bin.tsv", mode="w") as out_file:
        writer = csv.writer(out_file, delimiter=delimiter, quotechar='"',
                            quoting=csv.QUOTE_MINIMAL)
        for bin_clique in bin_cliques:
            bin_name = "_".join(str(x) for x in list(bin_clique))
            lone_removed = len(bin_clique) == 1 and bin_clique[0] in bins_to_rem
            if not lone_removed and bin_clique_size[bin_name] >= min_bin_size:
                final_bin_count += 1
                for b in bin_clique:
                    for contig in bins[b]:
                        final_bins[contig] = bin_name
                        if assembler == "megahit":
                            name = contig_descriptions[graph_to_contig_map[contig_names[contig]]]
                        else:
                            name = contig_names[contig]
                        writer.writerow([name, "bin_" + bin_name])
            else:
                for b in bin_clique:
                    for contig in bins[b]:
                        lowq_bins[contig] = bin_name
    return final_bins, lowq_bins, final_bin_count


def write_bin_fastas(contigs_file, final_bins, lowq_bins, output_bins_path, lq_output_bins_path,
                     prefix, contig_names_rev, assembler="spades", graph_to_contig_map_rev=None,
                     sequences=None, names=None):
    """One FASTA per bin name (metacoag:1231-1257): ``{output_bins_path}/{prefix}bin_{name}.fasta``
    for the names in ``final_bins``, ``{lq_output_bins_path}/{prefix}bin_{name}_seqs.fasta`` for
    those in ``lowq_bins``; a record goes to its final bin if it has one, else to its
    low-quality bin, in the order of the contigs file, as ``>{id}\\n{sequence}\\n``.  A name
    present in both maps shares one file object in the reference (the low-quality path wins);
    that is kept.  ``sequences`` / ``names`` are the staged ``ContigBlob`` and record ids; when
    absent the file is read with the library's FASTA reader.  Returns the paths written."""
    if sequences is None or names is None or not isinstance(sequences, ContigBlob):
        names, sequences = _read_contigs(contigs_file)
    path_of = {}
    for bin_name in set(final_bins.values()):
        path_of[bin_name] = f"{output_bins_path}/{prefix}bin_{bin_name}.fasta"
    created_only = []
    for bin_name in set(lowq_bins.values()):
        if bin_name in path_of:
            created_only.append(path_of[bin_name])       # opened "w+" and never written
        path_of[bin_name] = f"{lq_output_bins_path}/{prefix}bin_{bin_name}_seqs.fasta"
    order = list(path_of)
    index_of = {name: i for i, name in enumerate(order)}
    file_of_record = np.full(len(names), -1, dtype=np.int32)
    for n, record_id in enumerate(names):
        if assembler == "megahit":
            contig_num = contig_names_rev[graph_to_contig_map_rev[record_id]]
        else:
            contig_num = contig_names_rev[record_id]
        if contig_num in final_bins:
            file_of_record[n] = index_of[final_bins[contig_num]]
        elif contig_num in lowq_bins:
            file_of_record[n] = index_of[lowq_bins[contig_num]]
    paths = [path_of[name] for name in order]
    for path in created_only:
        open(path, "w+").close()
    _lib.write_bin_fastas(names, sequences.blob, sequences.offsets, file_of_record, paths)
    return paths + created_only


def make_output_dirs(output_path, prefix):
    """``{output_path}/{prefix}bins`` and ``.../{prefix}low_quality_bins`` (metacoag:1166-1173)."""
    output_bins_path = f"{output_path}/{prefix}bins"
    lq_output_bins_path = f"{output_path}/{prefix}low_quality_bins"
    os.makedirs(output_bins_path, exist_ok=True)
    os.makedirs(lq_output_bins_path, exist_ok=True)
    return output_bins_path, lq_output_bins_path
