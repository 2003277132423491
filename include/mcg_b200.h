/*
 * mcg_b200.h -- C ABI of libmcg_b200.so, the B200 (sm_100a) featurise + score path
 * of MetaCoAG.
 *
 * The reference (MetaCoAG v1.1.5) is pure Python and defines NO FFI of its own
 * (SURVEY.md section 8b): its boundary is the module API of
 * metacoag_utils/{feature_utils,matching_utils,label_prop_utils}.py.  Each entry
 * point below names the reference function(s) (file:line under the reference
 * root) whose arithmetic it replaces; metacoag_b200/*.py keeps those Python
 * signatures and binds these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every host buffer is caller-owned,
 *     C-contiguous and alive for the duration of the call;
 *   - device memory is owned by the context (mcg_ctx) and freed by mcg_destroy;
 *   - every call returns 0 (MCG_OK) or a negative code; mcg_last_error() gives
 *     the message.  Nothing throws or exits across this boundary;
 *   - a context is bound to one CUDA device (mcg_pool: one context per device of the box
 *     behind one caller) and is internally mutex-guarded
 *     (label_prop_utils.assign_long is entered from several Python threads,
 *     metacoag:959-999, and ctypes drops the GIL);
 *   - a call makes the context's device current for its duration and puts the caller's
 *     current device back before it returns;
 *   - there is no CPU fallback: without a CUDA device mcg_create() fails.
 */
#ifndef MCG_B200_H
#define MCG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCG_NPROF 136 /* canonical tetramers, feature_utils.py:38-57 */

enum {
    MCG_OK = 0,
    MCG_ERR_CUDA = -1,   /* CUDA runtime failure (message has the CUDA error string) */
    MCG_ERR_ARG = -2,    /* bad argument */
    MCG_ERR_STATE = -3,  /* called before the data it needs was loaded */
    MCG_ERR_DOMAIN = -4, /* the reference would raise here: ZeroDivisionError in
                            get_comp_probability (matching_utils.py:36-39, both Gaussians
                            underflow at distance >~ 1.39) */
    MCG_ERR_NOMEM = -5
};

enum { MCG_ENC_ASCII = 0, MCG_ENC_2BIT = 1 };
enum { MCG_RULE_A = 0, MCG_RULE_B = 1 };

typedef struct mcg_ctx mcg_ctx;

/* ---- context ---------------------------------------------------------------- */
int mcg_device_count(void);
mcg_ctx *mcg_create(int device);       /* NULL on failure; mcg_last_error(NULL) says why */
void mcg_destroy(mcg_ctx *ctx);
const char *mcg_last_error(const mcg_ctx *ctx);
/* Run on a caller-provided CUDA stream (e.g. torch's current stream, so that the
 * NCCL collectives the host issues are stream-ordered after the kernels); NULL
 * restores the context's own stream. */
int mcg_set_stream(mcg_ctx *ctx, void *cuda_stream);
int mcg_synchronize(mcg_ctx *ctx);
/* Pinned host memory for the buffers handed to the calls below (optional). */
void *mcg_host_alloc(size_t bytes);
void mcg_host_free(void *p);

/* ---- featurisation ---------------------------------------------------------- */
/* feature_utils.py:38-57 compute_kmer_inds(4): code -> slot of the canonical
 * tetramer, code = b0<<6|b1<<4|b2<<2|b3 with A0 C1 G2 T3 (feature_utils.py:31-35).
 * Pure host arithmetic (a 256-entry table); the kernels use the same table. */
int mcg_kmer_table(uint8_t table[256]);

/* feature_utils.py:60-70 count_kmers / :73-119 get_tetramer_profiles.
 * bases: concatenated contigs; offsets[n+1] in bases.  MCG_ENC_ASCII: one byte
 * per base, A0 C1 G2 T3 and ANY other byte 0 (lower case and N included,
 * feature_utils.py:21).  MCG_ENC_2BIT: base i of contig c is bits [2j,2j+2) of
 * 32-bit word pk_word(c) + i/16, j = i%16, each contig starting on a 16-byte
 * boundary: pk_word(c) = sum_{k<c} 4*ceil(len_k/64)  (offsets still in bases).
 * counts [n,136] uint32 and/or freqs [n,136] float64 (either may be NULL);
 * freqs = counts / max(1, len-3), IEEE division, bit-exact with the reference.
 * Stateless: nothing stays resident. */
int mcg_tetramer_scan(mcg_ctx *ctx, const uint8_t *bases, const int64_t *offsets, int64_t n,
                      int encoding, uint32_t *counts, double *freqs);

/* Resident form: upload + pack the contigs once, scan on the device, keep the
 * normalised profiles in the context as the contig table the scorers read. */
int mcg_load_contigs(mcg_ctx *ctx, const uint8_t *bases, const int64_t *offsets, int64_t n,
                     int encoding);
int mcg_scan(mcg_ctx *ctx);
/* ASCII loads of 16 MB and more are split between the copy engine (ASCII chunks, packed to
 * 2 bits by a kernel) and `threads` host threads that pack chunks themselves so that only
 * a quarter of their bytes cross the link; the two sides work towards each other, so the
 * split follows their measured speeds, and the packed store is bit-identical either way.
 * threads = 0 turns the host side off; -1 (default) takes MCG_HOST_PACK_THREADS from the
 * environment, else the hardware threads minus one (the caller feeds the copy engine).  (Replaces the per-sequence pickling
 * to a worker pool of feature_utils.py:97-104.) */
int mcg_set_host_pack_threads(mcg_ctx *ctx, int threads);
/* Rule C at scale (mcg_score_centroids / mcg_step_* / mcg_featurise_score without the weight
 * matrix) is a branch and bound.  log prob_cov <= 0 and is at most a number U of the contig
 * alone (matching_utils.py:42-66: a Poisson term is largest when the two coverages agree), and
 * -log prob_comp grows with the tetramer distance (matching_utils.py:36-39).  So once a contig
 * has the weight W of one bin (its nearest centroid), bins beyond the distance where
 * -log10 prob_comp passes W + U / ln 10 have a weight strictly above W: they are listed out, the
 * few that remain are scored exactly as before, and argmin / minimum weight
 * (label_prop_utils.py:340-345, first minimum) are the same bits with and without.  On by
 * default; 0 scores every pair (measurement, tests). */
int mcg_set_prune(mcg_ctx *ctx, int enabled);
/* The distances the branch and bound rules bins out with come from the 5th-generation tensor
 * cores (tcgen05.mma on an fp16 hi/lo split of the centred profiles, fp32 accumulators in tensor
 * memory, operands by TMA; csrc/mcg_tc.cu) together with a rigorous error margin; every pair
 * that is not ruled out gets its float64 distance (matching_utils.py:28-29) again, so results
 * keep their bits.  On by default; 0 = the float64 distance matrix of round 1 (DMMA) for all
 * pairs (measurement, tests). */
int mcg_set_tc(mcg_ctx *ctx, int enabled);
/* Parity / diagnostic form of that bound: d2a [nq, n_bins] fp32 against the resident centroids
 * (mcg_set_centroids / mcg_bin_profiles) and margin [nq]: |d2a - d^2| <= margin[q] for every bin,
 * d^2 the float64 squared tetramer distance.  queries NULL = rows 0..nq-1. */
int mcg_tc_distances(mcg_ctx *ctx, const int64_t *queries, int64_t nq, float *d2a, float *margin);
/* The scan kernel is persistent, one CTA per SM, and its shared memory leaves no room for
 * another kernel's CTA: a collective launched on another stream (the NCCL broadcast of the bin
 * tables, the all-gather of the assignments) would start only when the scan ends.  This keeps
 * `sms` SMs out of the scan's grid so that it starts at once (multi-GPU callers: 2-4). */
int mcg_set_reserved_sms(mcg_ctx *ctx, int sms);
/* The exchange steps of the path (bin tables from rank 0 to every rank, per-contig
 * (bin, weight) from every rank to every rank) as stores into peer memory: the caller maps one
 * buffer per rank into every process (symmetric memory: CUDA VMM / IPC; metacoag_b200/parallel.py
 * uses torch's allocator for it) and passes, per rank, the address where this rank's block goes.
 * One kernel on `stream` writes `bytes` (a multiple of 4) from d_src to all `world`
 * destinations over NVLink / NVSwitch; 16-byte stores when all addresses and the size allow.
 * Arrival is the caller's business (a barrier over the same symmetric allocation). */
#define MCG_MAX_PEERS 16
int mcg_push_to_peers(mcg_ctx *ctx, const void *d_src, size_t bytes, void *const *peer_dst,
                      int world, void *stream /* cudaStream_t; NULL = the context's stream */);
/* The last pruned rule-C chunk (synchronises the stream): [0] pairs listed as candidates,
 * [1] capacity of the list, [2] all pairs of the chunk, [3] 1 when the list was used, 0 when it
 * overflowed and every pair was scored (or pruning is off). */
int mcg_prune_stats(mcg_ctx *ctx, int64_t stats[4]);
/* The host-side packer on its own: ASCII contigs -> the MCG_ENC_2BIT layout described above
 * (packed has 16 * sum_c ceil(len_c / 64) bytes).  Pure host code, no device needed.
 * level -1 = best SIMD variant of this CPU, 0 scalar, 1 AVX2, 2 AVX-512BW, 3 AVX2 with a
 * masked 512-bit tail; a level the CPU lacks falls back to the next lower one. */
int mcg_pack_2bit_host(const uint8_t *bases, const int64_t *offsets, int64_t n, uint8_t *packed,
                       int level);
/* The same with `threads` host threads (<= 0: all hardware threads), contig ranges of equal
 * bases each: the pass a FASTA reader runs once at parse time (feature_utils.py:131-138) so that
 * every later upload of the contigs carries 2 bits per base instead of 8. */
int mcg_pack_2bit_host_mt(const uint8_t *bases, const int64_t *offsets, int64_t n, uint8_t *packed,
                          int threads);
/* feature_utils.py:131-138, 182-186: the contigs are read with Bio.SeqIO.parse(path,
 * "fasta"), keeping record.id and str(record.seq).  These two host-only calls produce the same
 * names and sequences (BioPython's SimpleFastaParser rules: text before the first '>' skipped,
 * id = first word of the header, lines right-stripped, ' ' removed, "\n" / "\r\n" / "\r" line
 * ends) as one byte blob + offsets, ready for mcg_load_contigs, instead of a list of Python
 * strings (the file is cut at record starts and the ranges are parsed in parallel).
 * mcg_fasta_index sizes the outputs; mcg_fasta_load fills caller-owned buffers
 * (bases[n_bases], offsets[n_records + 1], names[name_bytes], name_offsets[n_records + 1]).
 * Return 0, -1 if the file cannot be read, -2 if it changed between the two calls. */
int mcg_fasta_index(const char *path, int64_t *n_records, int64_t *n_bases, int64_t *name_bytes);
int mcg_fasta_load(const char *path, uint8_t *bases, int64_t *offsets, uint8_t *names,
                   int64_t *name_offsets, int64_t n_records, int64_t n_bases, int64_t name_bytes);
/* metacoag:1210-1251 (the per-bin FASTA writer, inline in the CLI): the reference parses the
 * contigs file again and writes ">{record.id}\n{str(record.seq)}\n" for every binned record,
 * in file order, to {prefix}bins/{prefix}bin_{name}.fasta or
 * {prefix}low_quality_bins/{prefix}bin_{name}_seqs.fasta.  Here the records come from the
 * blob mcg_fasta_load filled: file_of_record[i] is the index into paths of record i's output
 * (-1 = the record is in no bin).  Outputs are truncated ("w+") and written in parallel, one
 * worker per file; an output without records is created empty.  Host-only.
 * Returns 0, -(1 + i) when paths[i] could not be written, -1000000000 for a bad file index. */
int mcg_write_bin_fastas(const uint8_t *bases, const int64_t *offsets, const uint8_t *names,
                         const int64_t *name_offsets, int64_t n_records,
                         const int32_t *file_of_record, const char *const *paths,
                         int32_t n_files);
/* Last ASCII load: stats[0] bytes copied host->device for the bases, [1] bases packed by
 * host threads, [2] bases packed by the kernel, [3] host threads used. */
int mcg_ingest_stats(mcg_ctx *ctx, int64_t stats[4]);
int mcg_get_profiles(mcg_ctx *ctx, int64_t first, int64_t n, uint32_t *counts, double *freqs);
/* Contig table from host profiles (the dict the Python API carries around). */
int mcg_set_profiles(mcg_ctx *ctx, const double *prof /* [n,136] */, int64_t n);

/* feature_utils.py:150-153 (clamp < 1e-4 -> 1e-4) plus the per-(contig, sample)
 * terms get_cov_probability recomputes for every pair (matching_utils.py:48-54):
 * log(c) and lgamma(c + 1) -- the latter as CPython's math.lgamma computes it
 * (Lanczos, g = 6.0246800407767296), not libm's.  cov [n,S]. */
int mcg_load_coverage(mcg_ctx *ctx, const double *cov, int64_t n, int n_samples);
int mcg_get_coverage_terms(mcg_ctx *ctx, double *cov, double *logc, double *lgam);

/* feature_utils.py:217-235 get_bin_profiles: per bin, the mean of the member rows
 * (rows added in member order, one division).  members: contig ids, CSR by
 * seg_offsets[nseg+1].  Outputs [nseg,136] and [nseg,S] (either may be NULL). */
int mcg_bin_profiles(mcg_ctx *ctx, const int64_t *members, const int64_t *seg_offsets,
                     int64_t nseg, double *cen_prof, double *cen_cov);

/* ---- scoring ---------------------------------------------------------------- */
/* Elementwise forms of matching_utils.py:21-66 (normpdf, get_tetramer_distance,
 * get_comp_probability, get_cov_probability) over n independent inputs. */
int mcg_normpdf(mcg_ctx *ctx, const double *x, double mean, double sd, double *out, int64_t n);
int mcg_tetramer_distance(mcg_ctx *ctx, const double *u, const double *v, int64_t n, int width,
                          double *out);
int mcg_comp_probability(mcg_ctx *ctx, const double *dist, double *out, int64_t n);
int mcg_cov_probability(mcg_ctx *ctx, const double *cov1, const double *cov2, int64_t n,
                        int n_samples, double *out);

/* All-pairs primitives between resident contigs: queries[nq] x targets[nt] ->
 * dist / prob_comp / prob_cov / lp, each [nq,nt] or NULL; lp = -(log10 pc + log10 pv)
 * when pc*pv > 0 else MAX_WEIGHT (DBL_MAX).  Diagnostic / parity form of the inner
 * loop of matching_utils.py:123-147. */
int mcg_pair_scores(mcg_ctx *ctx, const int64_t *queries, int64_t nq, const int64_t *targets,
                    int64_t nt, double *dist, double *pc, double *pv, double *lp);

/* Rule C -- label_prop_utils.py:308-345 assign_long for a batch of contigs
 * (metacoag:947-999 calls it once per long unbinned contig) and the bin x bin merge
 * scoring of metacoag:1104-1140 when the queries are themselves centroids.
 * queries: contig ids or NULL for 0..nq-1.  cen_prof [B,136], cen_cov [B,S]
 * (NULL, NULL = the centroids left resident by mcg_bin_profiles).
 * weights [nq,B] or NULL; argmin[nq] = first index of the minimum, -1 where the
 * reference returns None (minimum is MAX_WEIGHT); wmin[nq]. */
int mcg_score_centroids(mcg_ctx *ctx, const int64_t *queries, int64_t nq, const double *cen_prof,
                        const double *cen_cov, int64_t n_bins, double *weights, int32_t *argmin,
                        double *wmin);

/* Rules A and B -- the (query x bin x member) loops of matching_utils.py:117-155,
 * 368-398 (A: sum / n_members) and label_prop_utils.py:54-86 (B: sum / n_valid).
 * members/seg_offsets: CSR of the member contigs to score per bin.  weights [nq,nseg]. */
int mcg_score_members(mcg_ctx *ctx, const int64_t *queries, int64_t nq, const int64_t *members,
                      const int64_t *seg_offsets, int64_t nseg, int rule, double *weights);

/* ---- label-propagation frontier ---------------------------------------------- */
/* label_prop_utils.py:24-100 run_bfs_long for a batch of start nodes over the
 * CSR assembly graph (neighbours in ascending id order, what igraph gives after
 * simplify()): bounded FIFO BFS that stops at labelled nodes and reproduces the
 * reference's depth-overwrite behaviour (:93-98).  labels[v] = bin or -1.
 * For start node i, hits are written at hit_offsets[i] .. : (labelled node, bin,
 * depth); the score of (start, bin) is rule B and comes from mcg_score_members.
 * max_hits_per_node bounds the per-node output; n_hits[i] is the true count
 * (a count above the bound means the caller must retry with a larger bound). */
int mcg_bfs_candidates(mcg_ctx *ctx, const int64_t *indptr, const int32_t *indices,
                       int64_t n_vertices, const int32_t *labels, const int64_t *starts,
                       int64_t n_starts, int depth_limit, int max_hits_per_node,
                       int32_t *hit_node, int32_t *hit_bin, int32_t *hit_depth,
                       int32_t *n_hits);

/* Rules A / B for LISTED (query, bin) items instead of the dense [queries, bins] matrix:
 * label_prop_utils.py:54-86 (run_bfs_long) scores a candidate only against the bins its walk
 * reached, a handful of the B bins.  item_query[i] = contig id, item_bin[i] = index into
 * seg_offsets; weights[i] as mcg_score_members would give it (distance in the reference's
 * difference form, matching_utils.py:28-29). */
int mcg_score_member_items(mcg_ctx *ctx, const int64_t *item_query, const int32_t *item_bin,
                           int64_t n_items, const int64_t *members, const int64_t *seg_offsets,
                           int64_t nseg, int rule, double *weights);

/* The label-propagation loops (label_prop_utils.py:230-303, 448-511) call run_bfs_long again for
 * the contigs around EVERY accepted contig.  These calls keep the CSR assembly graph and the
 * labels in the context, so that such a call moves its start nodes and its hits, not the graph:
 *   mcg_graph_load   once per graph (labels start as all -1);
 *   mcg_labels_load  all labels (labels[v] = bin of contig v or -1), once per stage;
 *   mcg_labels_set   labels[ids[i]] = bins[i] -- one entry after every acceptance
 *                    (label_prop_utils.py:270-271, 472-473);
 *   mcg_bfs_resident the walks of mcg_bfs_candidates from `starts`, hits as flat lists: start i
 *                    owns hits[hit_start[i] .. + hit_count[i]) as (labelled node, bin, depth)
 *                    int32 triples in the reference's pop order.  *total = triples of all
 *                    starts; when it exceeds hit_cap nothing else is valid and the caller
 *                    comes back with a larger buffer. */
int mcg_graph_load(mcg_ctx *ctx, const int64_t *indptr, const int32_t *indices, int64_t n_vertices);
int mcg_labels_load(mcg_ctx *ctx, const int32_t *labels, int64_t n_vertices);
int mcg_labels_set(mcg_ctx *ctx, const int64_t *ids, const int32_t *bins, int64_t n);
int mcg_bfs_resident(mcg_ctx *ctx, const int64_t *starts, int64_t n_starts, int depth_limit,
                     int64_t *hit_start, int32_t *hit_count, int32_t *hits /* [hit_cap][3] */,
                     int64_t hit_cap, int64_t *total);

/* matching_utils.py:191-197 and :285-292 (graph gate of match_contigs): for every
 * (source, target) pair, len(assembly_graph.get_shortest_paths(source, to=target)[0]) -- the
 * number of vertices on a shortest path over the CSR assembly graph, 1 when source ==
 * target, 0 when the target cannot be reached.  out is [n_sources, n_targets].  All sources
 * advance together (bit-parallel multi-source BFS, 32 per pass) instead of one igraph BFS
 * per pair. */
int mcg_path_lengths(mcg_ctx *ctx, const int64_t *indptr, const int32_t *indices,
                     int64_t n_vertices, const int64_t *sources, int64_t n_sources,
                     const int64_t *targets, int64_t n_targets, int32_t *out);

/* graph_utils.py:407-443 get_connected_components / get_non_isolated: the reference grows
 * the component of every binned contig by repeated neighbour sweeps (quadratic in the
 * component size, per start vertex).  labels[v] = smallest vertex id of the component of v
 * over the undirected CSR graph; the host groups vertices by label. */
int mcg_connected_components(mcg_ctx *ctx, const int64_t *indptr, const int32_t *indices,
                             int64_t n_vertices, int32_t *labels);

/* ---- fused plug-in call and the resident step ---------------------------------- */
/* Featurise + score in one call from HOST buffers: upload, pack, scan, normalise,
 * coverage terms, rule-C scoring against the given centroids, download
 * (argmin, wmin).  Leaves profiles and coverage terms resident. */
int mcg_featurise_score(mcg_ctx *ctx, const uint8_t *bases, const int64_t *offsets, int64_t n,
                        int encoding, const double *cov, int n_samples, const double *cen_prof,
                        const double *cen_cov, int64_t n_bins, int32_t *argmin, double *wmin);

/* Same work on data already resident (mcg_load_contigs + coverage uploaded with
 * mcg_load_coverage, centroids with mcg_set_centroids): scan -> normalise -> coverage
 * terms -> rule C; results stay on the device (d_argmin int32[n], d_wmin f64[n] are
 * DEVICE pointers, or NULL to use context-owned buffers). */
int mcg_set_centroids(mcg_ctx *ctx, const double *cen_prof, const double *cen_cov,
                      int64_t n_bins);
/* Same, from DEVICE buffers (e.g. the tensors an NCCL broadcast just filled);
 * stream-ordered on the context's stream, no host synchronisation. */
int mcg_set_centroids_dev(mcg_ctx *ctx, const void *d_cen_prof, const void *d_cen_cov,
                          int64_t n_bins);
int mcg_step_resident(mcg_ctx *ctx, void *d_argmin, void *d_wmin);
/* The two halves of mcg_step_resident -- scan + normalise + coverage terms, then rule C against
 * the resident centroids -- for callers that put something between them, e.g. an NCCL
 * broadcast of the centroids that runs on another stream while the scan is in flight
 * (mcg_set_centroids_dev after it, before mcg_step_score). */
int mcg_step_featurise(mcg_ctx *ctx);
int mcg_step_score(mcg_ctx *ctx, void *d_argmin, void *d_wmin);
int mcg_get_assignments(mcg_ctx *ctx, int32_t *argmin, double *wmin);

/* Per-kernel device times of the last mcg_step_resident / mcg_featurise_score,
 * measured with CUDA events on the launching stream when profiling is on.
 * Slots: 0 pack, 1 scan, 2 normalise (+ row norms), 3 coverage terms, 4 score (all scoring
 * kernels), 5 h2d (+ overlapped packing), 6 d2h, 7 other, 8 distance GEMM, 9 probability
 * kernel (8 and 9 are the two halves of slot 4 on the large rule-C path). */
#define MCG_N_TIMERS 10
int mcg_profile(mcg_ctx *ctx, int on);
int mcg_timings(mcg_ctx *ctx, float *ms /* [MCG_N_TIMERS] */, int64_t *launches);

/* Rule C with more than 30 samples (label_prop_utils.py:308-345 on the large batches):
 * the kernels score in the log domain and re-score, in the reference's operation order,
 * only the contigs whose answer could depend on a Poisson product that leaves the normal
 * double range (matching_utils.py:62-66 underflow).  This returns how many contigs the
 * last such call re-scored (0 when the path was not taken); it synchronises the stream. */
int mcg_last_redo_count(mcg_ctx *ctx, int64_t *count);

/* ---- device pool: the GPUs of one box behind one caller --------------- */
/* The reference is ONE process (metacoag:191; its only fan-out is the thread pool around
 * assign_long, metacoag:959-999), so a drop-in cannot ask for one process per GPU.  A pool owns
 * one context and one host thread per device and splits every call itself:
 *   - from raw bases (mcg_pool_tetramer_scan, mcg_pool_featurise_score): contiguous contig
 *     ranges balanced by total bases, every device uploads / scans / scores its own rows and
 *     writes its slice of the caller's arrays (mcg_pool_bounds: the ranges of the last call);
 *   - staged features (mcg_pool_set_features + the scorers): tables replicated on every device,
 *     QUERIES split evenly;
 *   - label propagation: graph and labels replicated, start nodes split, hits concatenated in
 *     start order.
 * Arguments, results and errors are those of the per-context calls of the same name; results do
 * not depend on the number of devices (bit for bit).  devices = NULL: all devices, or the first
 * MCG_DEVICES of them.  One pool call runs at a time (internally serialised). */
typedef struct mcg_pool mcg_pool;
mcg_pool *mcg_pool_create(const int *devices, int n);   /* NULL on failure */
void mcg_pool_destroy(mcg_pool *pool);
int mcg_pool_size(const mcg_pool *pool);
mcg_ctx *mcg_pool_ctx(mcg_pool *pool, int i);           /* context of the i-th device */
const char *mcg_pool_last_error(const mcg_pool *pool);
int mcg_pool_bounds(mcg_pool *pool, int64_t *bounds /* [size + 1] */);
int mcg_pool_tetramer_scan(mcg_pool *pool, const uint8_t *bases, const int64_t *offsets, int64_t n,
                           int encoding, uint32_t *counts, double *freqs);
int mcg_pool_featurise_score(mcg_pool *pool, const uint8_t *bases, const int64_t *offsets,
                             int64_t n, int encoding, const double *cov, int n_samples,
                             const double *cen_prof, const double *cen_cov, int64_t n_bins,
                             int32_t *argmin, double *wmin);
int mcg_pool_set_features(mcg_pool *pool, const double *prof /* [n_prof,136] or NULL */,
                          int64_t n_prof, const double *cov /* [n_cov,S] or NULL */, int64_t n_cov,
                          int n_samples);
int mcg_pool_bin_profiles(mcg_pool *pool, const int64_t *members, const int64_t *seg_offsets,
                          int64_t nseg, double *cen_prof, double *cen_cov);
int mcg_pool_score_centroids(mcg_pool *pool, const int64_t *queries, int64_t nq,
                             const double *cen_prof, const double *cen_cov, int64_t n_bins,
                             double *weights, int32_t *argmin, double *wmin);
int mcg_pool_score_members(mcg_pool *pool, const int64_t *queries, int64_t nq,
                           const int64_t *members, const int64_t *seg_offsets, int64_t nseg,
                           int rule, double *weights);
int mcg_pool_score_member_items(mcg_pool *pool, const int64_t *item_query, const int32_t *item_bin,
                                int64_t n_items, const int64_t *members, const int64_t *seg_offsets,
                                int64_t nseg, int rule, double *weights);
int mcg_pool_graph_load(mcg_pool *pool, const int64_t *indptr, const int32_t *indices,
                        int64_t n_vertices);
int mcg_pool_labels_load(mcg_pool *pool, const int32_t *labels, int64_t n_vertices);
int mcg_pool_labels_set(mcg_pool *pool, const int64_t *ids, const int32_t *bins, int64_t n);
int mcg_pool_bfs(mcg_pool *pool, const int64_t *starts, int64_t n_starts, int depth_limit,
                 int64_t *hit_start, int32_t *hit_count, int32_t *hits, int64_t hit_cap,
                 int64_t *total);

/* ---- synthetic workloads (bench / tests) ------------------ */
/* Fill a device-generated ASCII contig set: per-genome first-order Markov bases,
 * n_frac of positions overwritten with 'N', lower_frac of contigs lower-cased.
 * bases (host, may be pinned) receives sum(len) bytes. */
int mcg_synth_bases(mcg_ctx *ctx, const int64_t *offsets, int64_t n, const int32_t *genome_of,
                    const float *trans /* [n_genomes,4,4] row-stochastic */, int n_genomes,
                    uint64_t seed, double n_frac, double lower_frac, uint8_t *bases);

/* FP64 FMA throughput of this device (dependent-free DFMA loop), the roofline
 * denominator of the scoring kernel.  TFLOP/s. */
int mcg_fp64_peak(mcg_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* MCG_B200_H */
