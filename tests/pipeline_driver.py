"""The binning stages of the ``metacoag`` CLI that sit on the featurise + score path, as one
function parameterised by the three modules it calls.

tests/golden/make_pipeline.py runs it with the reference's own modules (imported from
/root/reference in the build container) and stores every intermediate state; the tests run
it with ``metacoag_b200``'s drop-in modules and compare.  The stage order, the glue between
the stages and the thresholds follow metacoag:640-1053.
"""
import math
import operator
import sys

MAX_WEIGHT = sys.float_info.max


def marker_tables(contig_markers):
    """marker_contigs {gene: [contigs]} and marker_contig_counts {gene: n} in first-seen
    order of the genes, as marker_gene_utils builds them from the .hmmout file."""
    marker_contigs, counts = {}, {}
    for contig in contig_markers:
        for gene in contig_markers[contig]:
            marker_contigs.setdefault(gene, []).append(contig)
    for gene in marker_contigs:
        counts[gene] = len(marker_contigs[gene])
    return marker_contigs, counts


def marker_gene_iterations(contig_markers, marker_contigs, marker_contig_counts):
    """smg_iteration (metacoag:654-695): genes grouped by their contig count, largest first;
    inside a group ordered by the total number of marker genes on their contigs."""
    smg_iteration, n = {}, 0
    for g_count in sorted(set(marker_contig_counts.values()), reverse=True):
        totals = {}
        for gene in marker_contig_counts:
            if marker_contig_counts[gene] == g_count:
                totals[gene] = sum(len(contig_markers[c]) for c in marker_contigs[gene])
        for gene, _ in sorted(totals.items(), key=operator.itemgetter(1), reverse=True):
            smg_iteration[n] = marker_contigs[gene]
            n += 1
    return smg_iteration


def snapshot(bins):
    return {int(b): [int(c) for c in members] for b, members in bins.items()}


def run_pipeline(fu, mu, lpu, sequences, coverages, contig_lengths, contig_markers, graph,
                 n_samples, workdir, min_length=1000, p_intra=0.1, p_inter=0.01, d_limit=20,
                 depth=10, nthreads=2):
    """Returns {stage name: state}.  ``coverages`` / ``contig_lengths`` / ``contig_markers``
    are the dicts get_cov_len and the marker-gene parser would hand to the CLI."""
    out = {}
    w_intra = -(math.log(p_intra, 10)) * (n_samples + 1)       # metacoag:540-541
    w_inter = -(math.log(p_inter, 10)) * (n_samples + 1)

    profiles = fu.get_tetramer_profiles(workdir, sequences, "contigs.fasta", contig_lengths,
                                        min_length, nthreads)
    out["profile_ids"] = sorted(int(c) for c in profiles)

    marker_contigs, marker_contig_counts = marker_tables(contig_markers)
    smg_iteration = marker_gene_iterations(contig_markers, marker_contigs, marker_contig_counts)

    bins, bin_of_contig, bin_markers, binned_with_markers = {}, {}, {}, []
    for i, contig in enumerate(smg_iteration[0]):                # metacoag:700-718
        binned_with_markers.append(contig)
        bins[i] = [contig]
        bin_of_contig[contig] = i
        bin_markers[i] = contig_markers[contig]
    n_bins = len(bins)
    out["initial"] = snapshot(bins)

    contig_names = {}
    bins, bin_of_contig, n_bins, bin_markers, binned_with_markers = mu.match_contigs(
        smg_iteration, bins, n_bins, bin_of_contig, binned_with_markers, bin_markers,
        contig_markers, contig_lengths, contig_names, profiles, coverages, graph, w_intra,
        w_inter, d_limit)
    out["match_contigs"] = snapshot(bins)

    unbinned = list(set(contig_markers.keys()) - set(binned_with_markers))   # metacoag:775-789
    by_length = sorted({c: contig_lengths[c] for c in unbinned}.items(),
                       key=operator.itemgetter(1), reverse=True)
    bins, bin_of_contig, n_bins, bin_markers, binned_with_markers = mu.further_match_contigs(
        by_length, min_length, bins, n_bins, bin_of_contig, binned_with_markers, bin_markers,
        contig_markers, profiles, coverages, w_intra)
    out["further_match_contigs"] = snapshot(bins)

    smg_bin_counts = [len(bins[b]) for b in bins]                 # metacoag:832-836
    cen_prof, cen_cov = fu.get_bin_profiles(bins, coverages, profiles)
    out["centroids"] = {int(b): ([float(x) for x in cen_prof[b]], [float(x) for x in cen_cov[b]])
                        for b in cen_prof}

    # the two label_prop calls of metacoag:884-927; non_isolated has the list-of-lists shape
    # graph_utils.get_non_isolated returns, which makes both calls no-ops
    non_isolated = [sorted(contig_lengths)]
    for d, w in ((1, w_intra), (depth, w_inter)):
        bins, bin_of_contig, bin_markers, binned_with_markers = lpu.label_prop(
            bin_of_contig, bins, contig_markers, bin_markers, binned_with_markers,
            smg_bin_counts, non_isolated, contig_lengths, min_length, graph, profiles, coverages,
            d, w)
    out["label_prop"] = snapshot(bins)

    all_contigs = sorted(contig_lengths)                          # metacoag:947-1022
    long_unbinned = [c for c in all_contigs
                     if c not in bin_of_contig and contig_lengths[c] >= min_length]
    assigned = [lpu.assign_long(c, coverages, profiles, cen_prof, cen_cov)
                for c in long_unbinned]
    put_to_bins = [x for x in assigned if x is not None]
    out["assign_long"] = [[int(c), int(b), float(w)] for c, b, w in put_to_bins]
    if put_to_bins:
        bins, bin_of_contig, bin_markers, binned_with_markers = lpu.assign_to_bins(
            put_to_bins, bins, bin_of_contig, bin_markers, binned_with_markers, contig_markers,
            contig_lengths)
    out["assign_to_bins"] = snapshot(bins)

    bins, bin_of_contig, bin_markers, binned_with_markers = lpu.final_label_prop(
        bin_of_contig, bins, contig_markers, bin_markers, binned_with_markers, smg_bin_counts,
        contig_lengths, min_length, graph, profiles, coverages, depth, MAX_WEIGHT)
    out["final_label_prop"] = snapshot(bins)
    out["bin_of_contig"] = {int(c): int(b) for c, b in bin_of_contig.items()}
    return out
