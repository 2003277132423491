// Internal declarations shared by the translation units of libmcg_b200.so.
// Public ABI: include/mcg_b200.h.
#pragma once

#include <cuda_runtime.h>

#include <cfloat>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "mcg_b200.h"

#define MCG_MAX_WEIGHT DBL_MAX   // matching_utils.py:15 MAX_WEIGHT = sys.float_info.max

// A grow-only device allocation.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

// Per-contig descriptor read by the scan kernel (one 16-byte load).
struct __align__(16) ContigMeta {
    int64_t pk_word;    // first 32-bit word of the contig in the packed store
    int32_t seg_first;  // first scan segment (256 windows each) of the contig
    int32_t nwin;       // tetramer windows = max(0, len - 3)
};

enum { T_PACK = 0, T_SCAN, T_NORM, T_COV, T_SCORE, T_H2D, T_D2H, T_OTHER, T_DIST, T_PROB };

struct mcg_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // H2D of contig chunks, overlapped with packing
    cudaStream_t copy_stream2 = nullptr;  // H2D of host-packed chunks
    cudaEvent_t copy2_done = nullptr;
    cudaStream_t cov_stream = nullptr;    // coverage rows of a wave: H2D + coverage terms
    cudaEvent_t cov_ev = nullptr;
    void *pinned_pack = nullptr;          // host-packed contigs, staged for the copy engine
    size_t pinned_pack_cap = 0;
    int reserved_sms = 0;                 // SMs the persistent scan leaves free (mcg_set_reserved_sms)
    int prune = 1;                        // rule-C branch and bound (mcg_set_prune)
    int host_pack_threads = -1;           // -1: MCG_HOST_PACK_THREADS or the hardware threads - 1
    int64_t ingest_stats[4] = {};         // last ASCII load: h2d bytes, host-packed bases,
                                          // GPU-packed bases, host threads
    cudaEvent_t copy_ev[8] = {};
    cudaEvent_t copy_start = nullptr;
    std::mutex mu;
    std::string err;

    // ---- contig store (2 bits per base) and scan work list -----------------
    int64_t n_contigs = 0;     // contigs in the packed store
    int64_t n_words = 0;       // packed 32-bit words (incl. per-contig padding)
    int64_t n_seg = 0;         // scan segments
    int64_t n_bases = 0;
    DevBuf packed;             // uint32 [n_words + 4]
    DevBuf meta;               // ContigMeta [n_contigs]
    DevBuf seg_contig;         // int32 [n_seg]
    DevBuf ascii;              // staging for MCG_ENC_ASCII uploads
    DevBuf offsets;            // int64 [n_contigs + 1] (bases)
    DevBuf tile_counter;       // unsigned long long [1]

    // ---- contig feature table ------------------------------------------------
    int64_t n_prof = 0;
    DevBuf counts;             // uint32 [n_prof,136]
    DevBuf prof;               // f64 [n_prof,136]
    DevBuf prof_norm;          // f64 [n_prof] squared row norms
    int64_t n_cov = 0;
    int n_samples = 0;
    DevBuf cov, logc, lgam;    // f64 [n_cov,S] each

    // ---- bin tables -------------------------------------------------------------
    int64_t n_bins = 0;
    int cen_samples = 0;
    DevBuf cen_prof;           // f64 [B,136]
    DevBuf cen_norm;           // f64 [B]
    DevBuf cen_cov, cen_logc, cen_lgam;   // f64 [B,S]

    // ---- assembly graph + labels, resident for a label-propagation stage ------------
    int64_t g_vertices = 0, g_edges = 0;
    DevBuf g_indptr, g_indices;   // int64 [V + 1], int32 [E]
    DevBuf g_labels;              // int32 [V] bin of the contig or -1
    DevBuf g_work, g_block, g_starts, g_patch;   // walk scratch (grow-only)
    void *g_pinned = nullptr;     // pinned mirror of the head of g_block
    size_t g_pinned_cap = 0;

    // ---- tensor-core bound (mcg_tc.cu): fp16 operands and their norms ----------------
    DevBuf tc_a, tc_b;            // __half [rows,448]
    DevBuf tc_aaux, tc_baux;      // float: |x|^2, |x| per row (+ max |x_v| behind the bins)
    DevBuf tc_d2a, tc_mrow;       // float [rows, ldf] approximate d^2 of a rule-C chunk, margin [rows]
    int tc = 1;                   // rule-C branch and bound takes its distances from tcgen05 (mcg_set_tc)

    // ---- results / scratch ---------------------------------------------------------
    DevBuf argmin, wmin;       // int32 [n], f64 [n]
    DevBuf status;             // int32 [4] device-side flags (0: domain error)
    DevBuf scratch[6];
    bool force_split = false;  // rule C always on the two-kernel path (waves: results must not depend on batch size)
    bool force_split_calls = false;   // the same across calls: set by the device pool around a sharded call
    bool cand_valid = false;   // scratch[3] starts with the candidate count of the last pruned rule-C chunk
    long long cand_cap = 0, cand_pairs = 0;
    bool redo_valid = false;   // scratch[4] starts with the redo count of the last guarded rule-C call
    void *pinned = nullptr;    // small pinned staging block
    size_t pinned_cap = 0;

    // ---- timing -----------------------------------------------------------------------
    bool profiling = false;
    cudaEvent_t ev[2 * MCG_N_TIMERS] = {};
    bool ev_used[MCG_N_TIMERS] = {};
    float ms[MCG_N_TIMERS] = {};
    int64_t launches = 0;      // kernels launched since mcg_profile(ctx, 1)
};

int mcg_fail(mcg_ctx *ctx, int code, const char *fmt, ...);
void mcg_set_global_error(const char *msg);

#define MCG_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return mcg_fail((ctx), MCG_ERR_CUDA, "%s failed: %s (%s:%d)", #call,              \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                     \
    } while (0)

#define MCG_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != MCG_OK) return rc__; \
    } while (0)

#define MCG_KERNEL_CHECK(ctx)                                                                 \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess)                                                               \
            return mcg_fail((ctx), MCG_ERR_CUDA, "kernel launch failed: %s (%s:%d)",          \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                     \
        (ctx)->launches++;                                                                    \
    } while (0)

int mcg_reserve(mcg_ctx *ctx, DevBuf &buf, size_t bytes);
int mcg_pinned_reserve(mcg_ctx *ctx, size_t bytes);
void mcg_timer_begin(mcg_ctx *ctx, int slot);
void mcg_timer_end(mcg_ctx *ctx, int slot);
int mcg_check_status(mcg_ctx *ctx);   // reads the device status flags, maps to MCG_ERR_DOMAIN

struct TimerScope {
    mcg_ctx *ctx;
    int slot;
    TimerScope(mcg_ctx *c, int s) : ctx(c), slot(s) { mcg_timer_begin(c, s); }
    ~TimerScope() { mcg_timer_end(ctx, slot); }
};

// ---- launchers implemented in the kernel translation units (all asynchronous on
// ctx->stream; pointers are device pointers) -------------------------------------------

// mcg_scan.cu
int mcg_k_pack_ascii(mcg_ctx *ctx, const uint8_t *d_ascii, const int64_t *d_offsets,
                     const ContigMeta *d_meta, int64_t n_contigs, int64_t word0, int64_t n_words,
                     uint32_t *d_packed);
int mcg_k_fill_seg_contig(mcg_ctx *ctx, const ContigMeta *d_meta, int64_t n_contigs,
                          int64_t n_seg, int32_t *d_seg_contig);
int mcg_k_scan_range(mcg_ctx *ctx, const uint32_t *d_packed, const ContigMeta *d_meta,
                     const int32_t *d_seg_contig, int64_t seg0, int64_t seg1, uint32_t *d_counts);
int mcg_k_scan(mcg_ctx *ctx, const uint32_t *d_packed, const ContigMeta *d_meta,
               const int32_t *d_seg_contig, int64_t n_seg, int64_t n_contigs,
               uint32_t *d_counts);
int mcg_k_normalise(mcg_ctx *ctx, const uint32_t *d_counts, const ContigMeta *d_meta,
                    int64_t n_contigs, double *d_prof);
int mcg_k_synth(mcg_ctx *ctx, const int64_t *d_offsets, int64_t n, const int32_t *d_genome_of,
                const float *d_trans, uint64_t seed, double n_frac, double lower_frac,
                uint8_t *d_bases, int64_t n_bases);

// mcg_score.cu
struct ScoreSide {
    const double *prof;   // [rows,136]
    const double *cov;    // [rows,S]
    const double *logc;   // [rows,S]
    const double *lgam;   // [rows,S]
    const double *norm;   // [rows] squared L2 norm of each profile row
    const int64_t *index; // gather index or nullptr (identity)
    int64_t first;        // offset into index (or first row when index is nullptr)
    int64_t n;            // rows addressed on this side
};
struct ScoreOut {
    // rule C outputs
    double *weights = nullptr;  // [na,nb] rule-C weights
    int32_t *argmin = nullptr;  // [na]
    double *wmin = nullptr;     // [na]
    // raw outputs
    double *dist = nullptr, *pc = nullptr, *pv = nullptr, *lp = nullptr;  // [na,nb]
};
int mcg_k_cov_terms(mcg_ctx *ctx, double *d_cov, double *d_logc, double *d_lgam, int64_t count);
int mcg_k_row_norms(mcg_ctx *ctx, const double *d_prof, int64_t rows, double *d_norm);
int mcg_k_normalise_norms(mcg_ctx *ctx, const uint32_t *d_counts, const ContigMeta *d_meta,
                          int64_t rows, double *d_prof, double *d_norm);
int mcg_k_pairs(mcg_ctx *ctx, const ScoreSide &a, const ScoreSide &b, int n_samples,
                const ScoreOut &out, int32_t *d_status);
int mcg_k_aggregate_members(mcg_ctx *ctx, const double *d_lp, int64_t nq, int64_t nm,
                            const int64_t *d_seg_offsets, int64_t nseg, int rule,
                            double *d_weights);
int mcg_k_member_items(mcg_ctx *ctx, const ScoreSide &table, const int64_t *d_item_query,
                       const int32_t *d_item_bin, int64_t n_items, const int64_t *d_members,
                       const int64_t *d_seg_offsets, int n_samples, int rule, double *d_weights,
                       int32_t *d_status);
int mcg_k_segment_mean(mcg_ctx *ctx, const double *d_rows, int width, const int64_t *d_members,
                       const int64_t *d_seg_offsets, int64_t nseg, double *d_out);
int mcg_k_normpdf(mcg_ctx *ctx, const double *d_x, double mean, double sd, double *d_out,
                  int64_t n);
int mcg_k_distance(mcg_ctx *ctx, const double *d_u, const double *d_v, int64_t n, int width,
                   double *d_out);
int mcg_k_comp_prob(mcg_ctx *ctx, const double *d_dist, double *d_out, int64_t n,
                    int32_t *d_status);
int mcg_k_cov_prob(mcg_ctx *ctx, const double *d_c1, const double *d_c2, int64_t n, int S,
                   double *d_out);
int mcg_k_fp64_peak(mcg_ctx *ctx, double *tflops);
int mcg_k_prune_stats(mcg_ctx *ctx, int64_t stats[4]);

// mcg_tc.cu
// halves of mcg_bfs_resident for the device pool (mcg_api.cu)
int mcg_bfs_run(mcg_ctx *ctx, const int64_t *starts, int64_t n_starts, int depth_limit,
                int64_t hit_cap, int64_t *total);
int mcg_bfs_fetch(mcg_ctx *ctx, int64_t n_starts, int64_t *hit_start, int32_t *hit_count,
                  int32_t *hits, int64_t total);
long long mcg_tc_ldf(long long bins);
int mcg_k_tc_prep_bins(mcg_ctx *ctx, const ScoreSide &b);
int mcg_k_tc_dist(mcg_ctx *ctx, const ScoreSide &a, long long bins, float *d_d2a, long long ldf,
                  float *d_mrow, unsigned long long *d_tilemin, long long nbt64);

// mcg_frontier.cu
int mcg_k_bfs(mcg_ctx *ctx, const int64_t *d_indptr, const int32_t *d_indices, int64_t n_vertices,
              const int32_t *d_labels, const int64_t *d_starts, int64_t n_starts, int depth_limit,
              int max_hits, int32_t *d_hit_node, int32_t *d_hit_bin, int32_t *d_hit_depth,
              int32_t *d_n_hits, int32_t *d_work, int64_t work_stride);
int mcg_k_labels_set(mcg_ctx *ctx, int32_t *d_labels, const int64_t *d_ids, const int32_t *d_bins,
                     int64_t n);
int64_t mcg_walk_chunk(void);
size_t mcg_walk_work_bytes(int64_t n_starts);
int mcg_k_walks(mcg_ctx *ctx, const int64_t *d_indptr, const int32_t *d_indices,
                const int32_t *d_labels, const int64_t *d_starts, int64_t n_starts,
                int depth_limit, int32_t *d_work, char *d_block, int64_t cap);
int mcg_k_walks_redo(mcg_ctx *ctx, const int64_t *d_indptr, const int32_t *d_indices,
                     const int32_t *d_labels, const int64_t *d_starts, int64_t n_starts,
                     int n_redo, int depth_limit, int32_t *d_work, int64_t stride, char *d_block,
                     int64_t cap);
int mcg_k_components(mcg_ctx *ctx, const int64_t *d_indptr, const int32_t *d_indices,
                     int64_t n, int32_t *d_work, int *d_flag, int32_t *d_labels);
int mcg_path_words(int64_t ns);   // mask words per vertex; work space = (4 W + 3) n int32
int mcg_k_path_lengths(mcg_ctx *ctx, const int64_t *d_indptr, const int32_t *d_indices,
                       int64_t n_vertices, const int32_t *d_tslot, const int64_t *d_sources,
                       int64_t ns, int64_t nt, int32_t *d_work, int *d_counts, int32_t *d_out);
