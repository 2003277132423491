"""Drop-in for ``metacoag_utils/feature_utils.py`` (MetaCoAG v1.1.5) on the B200 path.

Same public names, positional parameters and return types as the reference module; the
arithmetic runs in libmcg_b200.so (tetramer scan, normalisation, per-bin means), the file
parsing stays on the host.  Citations are ``feature_utils.py:line`` of the reference.
"""
import itertools
import logging
import os
import pickle
import sys
from collections import defaultdict

import numpy as np

from . import _backend, _lib

# same logger as the reference so that messages land in the same log (feature_utils.py:15)
logger = logging.getLogger("MetaCoaAG 1.1.5")

complements = {"A": "T", "C": "G", "G": "C", "T": "A"}
nt_bits = {"A": 0, "C": 1, "G": 2, "T": 3}
VERY_SMALL_VAL = 0.0001


def get_rc(seq):
    """Reverse complement; characters without a complement pass through (:25-28)."""
    return "".join(complements.get(ch, ch) for ch in reversed(seq))


def mer2bits(kmer):
    """MSB-first 2-bit code of a k-mer, anything but ACGT counting as A (:31-35)."""
    code = 0
    for ch in kmer:
        code = (code << 2) | nt_bits.get(ch, 0)
    return code


def compute_kmer_inds(k):
    """``({code: slot}, n_slots)``: a code and its reverse complement share a slot, slots are
    handed out in ascending code order (:38-57).  k = 4 comes from the library's table, the
    one the scan kernel folds with."""
    if k == 4:
        table = _lib.kmer_table()
        return {code: int(table[code]) for code in range(256)}, _lib.NPROF
    slots, n_slots = {}, 0
    for letters in itertools.product("ACGT", repeat=k):
        kmer = "".join(letters)
        code, rc = mer2bits(kmer), mer2bits(get_rc(kmer))
        if rc in slots:
            slots[code] = slots[rc]
        else:
            slots[code] = n_slots
            n_slots += 1
    return slots, n_slots


class ContigBlob:
    """The contig sequences as one byte blob + offsets that still reads like the reference's
    ``sequences`` list (``len``, indexing, slicing and iteration give ``str``): what
    get_cov_len / get_cov_len_megahit return here.  get_tetramer_profiles hands ``blob`` and
    ``offsets`` to the C ABI as they are instead of joining 10^5..10^6 Python strings
    (SURVEY.md 8f item 1).  The ASCII bytes stay: the writer of the per-bin FASTA files
    (metacoag:1236-1251) needs them."""

    def __init__(self, blob, offsets, packed=None):
        self.blob = blob
        self.offsets = np.asarray(offsets, dtype=np.int64)
        # the same contigs in the library's 2-bit layout, packed once at parse time: every later
        # upload (get_tetramer_profiles) carries a quarter of the bytes
        self.packed = packed

    def __len__(self):
        return self.offsets.shape[0] - 1

    def _one(self, i):
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("list index out of range")
        return self.blob[self.offsets[i]:self.offsets[i + 1]].tobytes().decode("latin-1")

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(j) for j in range(*i.indices(len(self)))]
        return self._one(int(i))

    def __iter__(self):
        for i in range(len(self)):
            yield self._one(i)

    def needs_strip(self):
        """count_kmers strips the sequence (:63); true when that would change any record
        (lines are right-stripped by the reader, so only a leading tab-like byte can)."""
        n = len(self)
        if n == 0 or self.blob.shape[0] == 0:
            return False
        lens = np.diff(self.offsets)
        first = self.blob[np.minimum(self.offsets[:-1], self.blob.shape[0] - 1)][lens > 0]
        last = self.blob[np.maximum(self.offsets[1:] - 1, 0)][lens > 0]
        edge = np.concatenate([first, last])
        return bool((((edge >= 9) & (edge <= 13)) | (edge == 32) | ((edge >= 28) & (edge <= 31))).any())


def _ascii_blob(sequences):
    """One contiguous uint8 blob + offsets from the reference's list of strings, with the
    ``str.strip()`` count_kmers applies (:63)."""
    if isinstance(sequences, ContigBlob) and not sequences.needs_strip():
        return sequences.blob, sequences.offsets
    parts = []
    for seq in sequences:
        seq = str(seq)
        if seq[:1].isspace() or seq[-1:].isspace():
            seq = seq.strip()
        parts.append(seq.encode("latin-1", "replace"))
    offsets = np.zeros(len(parts) + 1, dtype=np.int64)
    if parts:
        np.cumsum([len(p) for p in parts], out=offsets[1:])
    blob = np.frombuffer(b"".join(parts), dtype=np.uint8) if parts else np.zeros(0, np.uint8)
    return blob, offsets


def count_kmers(args):
    """``(counts float64[136], counts / max(1, sum(counts)))`` of one sequence (:60-70)."""
    seq, k, kmer_inds, kmer_count_len = args
    if k != 4 or kmer_count_len != _lib.NPROF:
        raise ValueError("the B200 scan kernel counts canonical tetramers (k = 4) only")
    blob, offsets = _ascii_blob([seq])
    counts, freqs = _backend.backend().tetramer_scan(blob, offsets)
    return counts[0].astype(np.float64), freqs[0]


def get_tetramer_profiles(output_path, sequences, contigs_file, contig_lengths, min_length,
                          nthreads):
    """Normalised tetramer profile of every contig of at least ``min_length`` (:73-119).

    The pickle cache keeps the reference's file name, including its missing path separator
    (``f"{output_path}{basename}..."``), and its format (``{i: float64[136]}`` over ALL
    records), so either implementation can pick up the other's cache.  ``nthreads`` is
    accepted for signature compatibility; one kernel launch scans every contig.
    """
    basename = contigs_file.split("/")[-1]
    cache = f"{output_path}{basename}.normalized_contig_tetramers.pickle"
    if os.path.isfile(cache):
        with open(cache, "rb") as handle:
            normalized = pickle.load(handle)
        dense = None
    else:
        blob, offsets = _ascii_blob(sequences)
        packed = getattr(sequences, "packed", None)
        if packed is not None and blob is not sequences.blob:
            packed = None          # the blob was rebuilt (strip): the parse-time packing is stale
        if len(sequences):
            _, freqs = _backend.backend().tetramer_scan(blob, offsets, packed=packed)
        else:
            freqs = np.zeros((0, _lib.NPROF))
        normalized = {i: row for i, row in enumerate(freqs)}
        with open(cache, "wb") as handle:
            pickle.dump(normalized, handle, protocol=pickle.HIGHEST_PROTOCOL)
        dense = freqs
    # the reference's dict of the long contigs; it remembers the dense table it is a view of,
    # so that staging it for the scorers is not a Python loop over contigs (_backend.stage)
    table = _backend.ProfileTable((i, normalized[i]) for i in range(len(normalized))
                                  if contig_lengths[i] >= min_length)
    table.dense = dense
    return table


def _fasta_records(path):
    """(name, sequence) per record, pure Python: BioPython's SimpleFastaParser restated (the
    reference reads FASTA with Bio.SeqIO, :131, :182) -- text before the first '>' skipped,
    name = first word of the header (``record.id``), lines right-stripped and joined, spaces
    removed.  The library's reader (mcg_fasta_load) is checked against this one."""
    name, chunks = None, []
    with open(path, encoding="latin-1") as handle:
        for line in handle:
            if line[:1] == ">":
                if name is not None:
                    yield name, "".join(chunks).replace(" ", "").replace("\r", "")
                words = line[1:].rstrip().split(None, 1)
                name, chunks = (words[0] if words else ""), []
            elif name is not None:
                chunks.append(line.rstrip())
    if name is not None:
        yield name, "".join(chunks).replace(" ", "").replace("\r", "")


def _read_contigs(contigs_file):
    """(names, sequences) of a FASTA file through the library's host reader: the sequences
    come back as a ContigBlob (pinned when a CUDA device is there, so that the upload needs
    no staging copy)."""
    on_device = _lib.device_count() > 0
    names, blob, offsets = _lib.read_fasta(contigs_file, pinned=on_device)
    packed = None
    if on_device and blob.shape[0]:
        packed = _lib.pack_2bit_host(blob, offsets, threads=0, pinned=True)
    return names, ContigBlob(blob, offsets, packed)


def _read_abundance(abundance_file, contig_num_of, contig_lengths, min_length, coverages):
    """Abundance TSV ``name<TAB>c1<TAB>...`` -> coverages of the long contigs, values below
    1e-4 raised to 1e-4; a contig listed twice accumulates its values (:142-158, :189-205).
    Exits like the reference when no contig is long enough.  Returns the number of samples."""
    with open(abundance_file, "r") as handle:
        for line in handle:
            fields = line.strip().split("\t")
            contig_num = contig_num_of(fields[0])
            if contig_lengths[contig_num] >= min_length:
                for text in fields[1:]:
                    value = float(text)
                    if value < VERY_SMALL_VAL:
                        value = VERY_SMALL_VAL
                    if contig_num not in coverages:
                        coverages[contig_num] = [value]
                    else:
                        coverages[contig_num].append(value)
    if len(coverages) == 0:
        logger.error(f"Could not find any contigs longer than {min_length}bp.")
        logger.info("Exiting MetaCoAG... Bye...!")
        sys.exit(1)
    return len(coverages[next(iter(coverages))])


def get_cov_len(contigs_file, contig_names_rev, min_length, abundance_file):
    """``(sequences, coverages, contig_lengths, n_samples)`` (:122-165).  ``coverages`` is a
    ``defaultdict(lambda: [0])`` and ``contig_lengths`` a ``defaultdict(int)`` like the
    reference's; an unknown contig name raises KeyError."""
    _backend.reset()      # a pipeline run starts here: nothing staged by an earlier run survives
    coverages = defaultdict(lambda: [0])
    contig_lengths = defaultdict(int)
    names, sequences = _read_contigs(contigs_file)
    for name, length in zip(names, np.diff(sequences.offsets).tolist()):
        contig_lengths[contig_names_rev[name]] = length
    n_samples = _read_abundance(abundance_file, lambda name: contig_names_rev[name],
                                contig_lengths, min_length, coverages)
    return sequences, coverages, contig_lengths, n_samples


def get_cov_len_megahit(contigs_file, contig_names_rev, graph_to_contig_map_rev, min_length,
                        abundance_file):
    """MEGAHIT flavour: FASTA names map to graph names first (:168-214)."""
    _backend.reset()
    coverages, contig_lengths = {}, {}
    names, sequences = _read_contigs(contigs_file)
    for name, length in zip(names, np.diff(sequences.offsets).tolist()):
        contig_lengths[contig_names_rev[graph_to_contig_map_rev[name]]] = length
    n_samples = _read_abundance(
        abundance_file, lambda name: contig_names_rev[graph_to_contig_map_rev[name]],
        contig_lengths, min_length, coverages)
    return sequences, coverages, contig_lengths, n_samples


def get_bin_profiles(bins, coverages, normalized_tetramer_profiles):
    """Per-bin mean tetramer and coverage profile (:217-235): ``({b: f64[136]}, {b: f64[S]})``."""
    order = list(bins)
    if not order:
        return {}, {}
    members, seg = _backend.csr_bins(bins, order)
    cen_prof, cen_cov = _backend.backend().bin_profiles(normalized_tetramer_profiles, coverages,
                                                        members, seg)
    return ({b: cen_prof[i] for i, b in enumerate(order)},
            {b: cen_cov[i] for i, b in enumerate(order)})
