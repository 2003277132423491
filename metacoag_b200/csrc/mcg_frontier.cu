// Label-propagation frontier over the CSR assembly graph.
//
//   reference: label_prop_utils.py:24-100 run_bfs_long.  For one start node it runs a FIFO
//   BFS bounded by `threhold`, does not expand labelled nodes (other than the start), and
//   reports every labelled node it pops together with depth[node].  Two details of the
//   Python are reproduced on purpose because they change the reported depths:
//     - a node is appended to the queue every time an unvisited neighbour relation is
//       seen (membership of the queue is not checked), so it can be popped several times;
//     - depth[neighbour] is overwritten on every such sighting, also when the new depth
//       exceeds the bound (:93-98).
//   The score attached to a hit (rule B) is static per (start, bin) and comes from
//   mcg_score_members; this kernel produces the (labelled node, bin, depth) triples in
//   the reference's pop order.
//
// One thread walks one start node; thousands of starts run side by side.  Per start the
// queue and an open-addressing (vertex -> visited, depth) table live in a slice of a
// global work buffer, so the footprint does not depend on the graph size.
#include "mcg_internal.cuh"

namespace {

struct Slot {
    int32_t key;   // vertex + 1, 0 = empty
    int32_t val;   // depth << 1 | visited
};

__device__ __forceinline__ int find_or_insert(Slot *table, int cap_mask, int vertex, bool *full) {
    uint32_t h = static_cast<uint32_t>(vertex) * 2654435761u;
    int pos = static_cast<int>((h >> 7) & cap_mask);
    for (int probes = 0; probes <= cap_mask; ++probes) {
        const int key = table[pos].key;
        if (key == vertex + 1) return pos;
        if (key == 0) {
            table[pos].key = vertex + 1;
            table[pos].val = 0;
            return pos;
        }
        pos = (pos + 1) & cap_mask;
    }
    *full = true;
    return 0;
}

__global__ void bfs_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                           const int32_t *__restrict__ labels, const int64_t *__restrict__ starts,
                           long long n_starts, int depth_limit, int max_hits,
                           int32_t *__restrict__ hit_node, int32_t *__restrict__ hit_bin,
                           int32_t *__restrict__ hit_depth, int32_t *__restrict__ n_hits,
                           int32_t *__restrict__ work, long long work_stride, int queue_cap,
                           int table_cap) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n_starts) return;
    int32_t *queue = work + i * work_stride;
    Slot *table = reinterpret_cast<Slot *>(queue + queue_cap);
    const int mask = table_cap - 1;
    for (int t = 0; t < table_cap; ++t) table[t].key = 0;

    const int start = static_cast<int>(starts[i]);
    bool overflow = false;
    int head = 0, tail = 0, n_visited = 0, hits = 0;
    queue[tail++] = start;
    table[find_or_insert(table, mask, start, &overflow)].val = 0;

    while (head < tail && !overflow) {
        const int v = queue[head++];
        const int sv = find_or_insert(table, mask, v, &overflow);
        if (overflow) break;
        int val = table[sv].val;
        if (!(val & 1)) { val |= 1; table[sv].val = val; ++n_visited; }
        const int dv = val >> 1;
        const int lab = labels[v];
        if (lab >= 0 && n_visited > 1) {
            if (hits < max_hits) {
                const long long o = i * max_hits + hits;
                hit_node[o] = v;
                hit_bin[o] = lab;
                hit_depth[o] = dv;
            }
            ++hits;
        } else {
            for (int64_t e = indptr[v]; e < indptr[v + 1]; ++e) {
                const int nb = indices[e];
                const int sn = find_or_insert(table, mask, nb, &overflow);
                if (overflow) break;
                const int nval = table[sn].val;
                if (nval & 1) continue;                     // already visited
                table[sn].val = (dv + 1) << 1;              // depth overwrite, visited = 0
                if (dv + 1 > depth_limit) continue;
                if (tail >= queue_cap) { overflow = true; break; }
                queue[tail++] = nb;
            }
        }
    }
    n_hits[i] = overflow ? -2 : hits;
}

}  // namespace

int mcg_k_bfs(mcg_ctx *ctx, const int64_t *d_indptr, const int32_t *d_indices, int64_t n_vertices,
              const int32_t *d_labels, const int64_t *d_starts, int64_t n_starts, int depth_limit,
              int max_hits, int32_t *d_hit_node, int32_t *d_hit_bin, int32_t *d_hit_depth,
              int32_t *d_n_hits, int32_t *d_work, int64_t work_stride) {
    (void)n_vertices;
    if (n_starts == 0) return MCG_OK;
    // half of the slice is the queue, the rest a power-of-two table of 8-byte slots
    const int queue_cap = static_cast<int>(work_stride / 2);
    int table_cap = 1;
    while (2LL * table_cap * 2 <= work_stride - queue_cap) table_cap *= 2;
    const int threads = 64;
    const int64_t blocks = (n_starts + threads - 1) / threads;
    bfs_kernel<<<static_cast<unsigned>(blocks), threads, 0, ctx->stream>>>(
        d_indptr, d_indices, d_labels, d_starts, n_starts, depth_limit, max_hits, d_hit_node,
        d_hit_bin, d_hit_depth, d_n_hits, d_work, work_stride, queue_cap, table_cap);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

// ---- the same walk over a RESIDENT graph, hits as flat lists ------------------------------------
//
// label_prop_utils.py:448-511 re-runs run_bfs_long for the few contigs around every accepted
// contig.  With the CSR graph and the labels resident in the context (mcg_graph_load,
// mcg_labels_load / mcg_labels_set) such a call moves only its start nodes and its hits, and a
// stage with millions of start nodes needs no per-start output slots: every walk buffers its
// hits, reserves a range of one flat list with a single atomicAdd when it ends, and copies
// them there (same pop order).  The visited / depth table of a walk sits in SHARED memory
// (128 slots per walk, slot-major so that the threads of a warp use different banks for the
// same slot): bounded walks touch tens of vertices, and the table was what the first version
// paid global-memory latency for.  A walk that does not fit (table 3/4 full, queue or hit
// buffer full) is listed and repeated by the same code over a table in global memory.
namespace {

constexpr int kWalkThreads = 64;
constexpr int kWalkSlots = 128;                 // shared table slots per walk
constexpr int kWalkQueue = 384;                 // queue entries per walk
constexpr int kWalkHits = 64;                   // buffered hits per walk
constexpr int kWalkSlice = kWalkQueue + 2 * kWalkHits;   // int32 of global work space per walk

struct WalkOut {
    long long *total;        // [1] entries of the flat list reserved so far (may pass cap)
    int *n_redo;             // [1]
    int32_t *redo;           // walks that did not fit (index into starts)
    long long *start;        // [n_starts] first entry of the walk's hits
    int32_t *count;          // [n_starts] hits of the walk; -2 while it waits in `redo`
    int32_t *hits;           // [cap][3] (labelled node, bin, depth)
    long long cap;
};

struct SharedTable {
    int2 *base;              // this thread's column: slot s at base[s * kWalkThreads]
    int used;
    __device__ __forceinline__ int capacity() const { return kWalkSlots; }
    __device__ __forceinline__ int limit() const { return kWalkSlots * 3 / 4; }
    __device__ __forceinline__ int2 &at(int pos) { return base[pos * kWalkThreads]; }
};

struct GlobalTable {
    int2 *base;
    int cap;                 // power of two
    int used;
    __device__ __forceinline__ int capacity() const { return cap; }
    __device__ __forceinline__ int limit() const { return cap / 4 * 3; }
    __device__ __forceinline__ int2 &at(int pos) { return base[pos]; }
};

// slot of `vertex` (x = vertex + 1, y = depth << 1 | visited), inserted if absent; -1 = full
template <class Table> __device__ __forceinline__ int table_slot(Table &tb, int vertex) {
    const int mask = tb.capacity() - 1;
    int pos = static_cast<int>(((static_cast<uint32_t>(vertex) * 2654435761u) >> 7) & mask);
    for (int probes = 0; probes <= mask; ++probes) {
        const int key = tb.at(pos).x;
        if (key == vertex + 1) return pos;
        if (key == 0) {
            if (tb.used >= tb.limit()) return -1;
            tb.at(pos) = make_int2(vertex + 1, 0);
            ++tb.used;
            return pos;
        }
        pos = (pos + 1) & mask;
    }
    return -1;
}

// One bounded walk (label_prop_utils.py:37-53, 92-100).  Hits go to tmp as (node, depth) pairs.
// Returns the number of hits, or -1 when a table / queue / hit buffer was too small.
template <class Table>
__device__ int walk_one(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                        const int32_t *__restrict__ labels, int start, int depth_limit,
                        int32_t *queue, int queue_cap, int32_t *tmp, int tmp_cap, Table &tb) {
    int head = 0, tail = 0, n_visited = 0, hits = 0;
    queue[tail++] = start;
    if (table_slot(tb, start) < 0) return -1;
    while (head < tail) {
        const int v = queue[head++];
        const int sv = table_slot(tb, v);
        if (sv < 0) return -1;
        int val = tb.at(sv).y;
        if (!(val & 1)) { val |= 1; tb.at(sv).y = val; ++n_visited; }
        const int dv = val >> 1;
        if (labels[v] >= 0 && n_visited > 1) {
            if (hits >= tmp_cap) return -1;
            tmp[2 * hits] = v;
            tmp[2 * hits + 1] = dv;
            ++hits;
            continue;
        }
        for (int64_t e = indptr[v]; e < indptr[v + 1]; ++e) {
            const int nb = indices[e];
            const int sn = table_slot(tb, nb);
            if (sn < 0) return -1;
            if (tb.at(sn).y & 1) continue;                 // already visited
            tb.at(sn).y = (dv + 1) << 1;                   // depth overwrite, visited = 0
            if (dv + 1 > depth_limit) continue;
            if (tail >= queue_cap) return -1;
            queue[tail++] = nb;
        }
    }
    return hits;
}

__device__ __forceinline__ void walk_emit(const WalkOut &out, long long i, int hits,
                                          const int32_t *tmp, const int32_t *__restrict__ labels) {
    const long long base = hits ? atomicAdd(reinterpret_cast<unsigned long long *>(out.total),
                                            static_cast<unsigned long long>(hits)) : 0;
    out.start[i] = base;
    out.count[i] = hits;
    if (base + hits > out.cap) return;   // the caller sees total > cap and comes back
    for (int h = 0; h < hits; ++h) {
        int32_t *o = out.hits + 3 * (base + h);
        o[0] = tmp[2 * h];
        o[1] = labels[tmp[2 * h]];
        o[2] = tmp[2 * h + 1];
    }
}

__global__ void __launch_bounds__(kWalkThreads)
walk_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
            const int32_t *__restrict__ labels, const int64_t *__restrict__ starts,
            long long first, long long n_starts, int depth_limit, int32_t *__restrict__ work,
            WalkOut out) {
    extern __shared__ int2 walk_tab[];
    const int t = threadIdx.x;
    for (int s = 0; s < kWalkSlots; ++s) walk_tab[s * kWalkThreads + t].x = 0;
    const long long local = blockIdx.x * static_cast<long long>(kWalkThreads) + t;
    const long long i = first + local;
    if (i >= n_starts) return;
    int32_t *queue = work + local * kWalkSlice;
    int32_t *tmp = queue + kWalkQueue;
    SharedTable tb = {walk_tab + t, 0};
    const int hits = walk_one(indptr, indices, labels, static_cast<int>(starts[i]), depth_limit,
                              queue, kWalkQueue, tmp, kWalkHits, tb);
    if (hits < 0) {
        out.count[i] = -2;
        out.redo[atomicAdd(out.n_redo, 1)] = static_cast<int32_t>(i);
        return;
    }
    walk_emit(out, i, hits, tmp, labels);
}

// The listed walks again, table / queue / hit buffer in a slice of global memory (stride a power
// of two >= 1024 int32): [queue stride/2][hits 2 x stride/16][table 2 x stride/8 slots].
__global__ void __launch_bounds__(64)
walk_redo_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                 const int32_t *__restrict__ labels, const int64_t *__restrict__ starts,
                 int n_redo, int depth_limit, int32_t *__restrict__ work, long long stride,
                 WalkOut out, int *__restrict__ n_failed) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_redo) return;
    const long long i = out.redo[j];
    if (out.count[i] != -2) return;      // done by an earlier pass
    const int q = static_cast<int>(stride / 2), h = static_cast<int>(stride / 16),
              c = static_cast<int>(stride / 8);
    int32_t *queue = work + j * stride;
    int32_t *tmp = queue + q;
    GlobalTable tb = {reinterpret_cast<int2 *>(tmp + 2 * h), c, 0};
    for (int s = 0; s < c; ++s) tb.base[s].x = 0;
    const int hits = walk_one(indptr, indices, labels, static_cast<int>(starts[i]), depth_limit,
                              queue, q, tmp, h, tb);
    if (hits < 0) { atomicAdd(n_failed, 1); return; }
    walk_emit(out, i, hits, tmp, labels);
}

__global__ void labels_set_kernel(int32_t *__restrict__ labels, const int64_t *__restrict__ ids,
                                  const int32_t *__restrict__ bins, long long n) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i < n) labels[ids[i]] = bins[i];
}

}  // namespace

int mcg_k_labels_set(mcg_ctx *ctx, int32_t *d_labels, const int64_t *d_ids, const int32_t *d_bins,
                     int64_t n) {
    if (n == 0) return MCG_OK;
    labels_set_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(
        d_labels, d_ids, d_bins, n);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int64_t mcg_walk_chunk(void) { return 131072; }
size_t mcg_walk_work_bytes(int64_t n_starts) {
    const int64_t chunk = n_starts < mcg_walk_chunk() ? n_starts : mcg_walk_chunk();
    const int64_t padded = (chunk + kWalkThreads - 1) / kWalkThreads * kWalkThreads;
    return sizeof(int32_t) * static_cast<size_t>(padded) * kWalkSlice;
}

// block: [total i64][n_redo i32][n_failed i32][start i64 x n][count i32 x n (padded to 8 bytes)]
//        [redo i32 x n (padded)][hits i32 x 3 cap]
int mcg_k_walks(mcg_ctx *ctx, const int64_t *d_indptr, const int32_t *d_indices,
                const int32_t *d_labels, const int64_t *d_starts, int64_t n_starts,
                int depth_limit, int32_t *d_work, char *d_block, int64_t cap) {
    static bool attr[64] = {};
    const size_t smem = sizeof(int2) * kWalkSlots * kWalkThreads;
    if (!(ctx->device < 64 && attr[ctx->device])) {
        MCG_CUDA(ctx, cudaFuncSetAttribute(walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
        if (ctx->device < 64) attr[ctx->device] = true;
    }
    const int64_t n8 = (n_starts + 1) / 2 * 2;
    WalkOut out;
    out.total = reinterpret_cast<long long *>(d_block);
    out.n_redo = reinterpret_cast<int *>(d_block + 8);
    out.start = reinterpret_cast<long long *>(d_block + 16);
    out.count = reinterpret_cast<int32_t *>(out.start + n_starts);
    out.redo = out.count + n8;
    out.hits = out.redo + n8;
    out.cap = cap;
    MCG_CUDA(ctx, cudaMemsetAsync(d_block, 0, 16, ctx->stream));
    const int64_t chunk = mcg_walk_chunk();
    for (int64_t first = 0; first < n_starts; first += chunk) {
        const int64_t m = n_starts - first < chunk ? n_starts - first : chunk;
        walk_kernel<<<static_cast<unsigned>((m + kWalkThreads - 1) / kWalkThreads), kWalkThreads,
                      smem, ctx->stream>>>(d_indptr, d_indices, d_labels, d_starts, first,
                                           n_starts, depth_limit, d_work, out);
        MCG_KERNEL_CHECK(ctx);
    }
    return MCG_OK;
}

// The walks walk_kernel listed, with `stride` int32 of global work space each (n_redo known to
// the host).  *d_block + 12 counts the walks that still did not fit.
int mcg_k_walks_redo(mcg_ctx *ctx, const int64_t *d_indptr, const int32_t *d_indices,
                     const int32_t *d_labels, const int64_t *d_starts, int64_t n_starts,
                     int n_redo, int depth_limit, int32_t *d_work, int64_t stride, char *d_block,
                     int64_t cap) {
    const int64_t n8 = (n_starts + 1) / 2 * 2;
    WalkOut out;
    out.total = reinterpret_cast<long long *>(d_block);
    out.n_redo = reinterpret_cast<int *>(d_block + 8);
    out.start = reinterpret_cast<long long *>(d_block + 16);
    out.count = reinterpret_cast<int32_t *>(out.start + n_starts);
    out.redo = out.count + n8;
    out.hits = out.redo + n8;
    out.cap = cap;
    MCG_CUDA(ctx, cudaMemsetAsync(d_block + 12, 0, 4, ctx->stream));
    walk_redo_kernel<<<static_cast<unsigned>((n_redo + 63) / 64), 64, 0, ctx->stream>>>(
        d_indptr, d_indices, d_labels, d_starts, n_redo, depth_limit, d_work, stride, out,
        reinterpret_cast<int *>(d_block + 12));
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

// ---- shortest-path lengths for the graph gate of match_contigs ----------------------------------
//
//   reference: matching_utils.py:191-197 and :285-292 call, for a matched contig v and every
//   member w of the bin, assembly_graph.get_shortest_paths(v, to=w) and add len(path): the
//   number of VERTICES on a shortest path (1 when v == w), 0 when w cannot be reached.  igraph
//   runs one BFS per (v, w) pair.
//
// Here all sources of a launch advance together, one bit each (bit-parallel multi-source BFS,
// W 32-bit words per vertex): a vertex carries the mask of sources that have reached it, a
// level expands only the vertices whose mask grew (compact frontier lists, so a level costs
// its frontier, not the graph), and the first time bit s arrives at a target vertex the pair
// (s, target) gets level + 1 vertices.  One cooperative launch walks all levels with two
// grid-wide barriers per level, whatever the number of sources.
#include <cooperative_groups.h>

namespace {
namespace cg = cooperative_groups;

struct PathArgs {
    const int64_t *indptr;
    const int32_t *indices;
    uint32_t *visited, *front_a, *front_b, *next;   // [n, W] source masks, zero on entry
    int32_t *mark;                                  // [n] level at which the vertex was listed
    int32_t *list_a, *list_b;                       // [n] frontier vertex lists
    int *counts;                                    // [2] list lengths, zero on entry
    const int32_t *tslot;                           // [n] column of the vertex in `out`, or -1
    const int64_t *sources;                         // [ns] (ns <= 32 W)
    int ns, W;
    long long nt;
    int32_t *out;                                   // [ns, nt], zero on entry
};

__global__ void __launch_bounds__(256) path_lengths_kernel(PathArgs a) {
    cg::grid_group grid = cg::this_grid();
    const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const int W = a.W;

    if (tid < a.ns) {
        const long long s = a.sources[tid];
        const uint32_t bit = 1u << (tid & 31);
        const int w = static_cast<int>(tid >> 5);
        atomicOr(&a.visited[s * W + w], bit);
        atomicOr(&a.front_a[s * W + w], bit);
        // (the same vertex may be given as several sources: it is listed once)
        if (atomicExch(&a.mark[s], -1) != -1) a.list_a[atomicAdd(&a.counts[0], 1)] = static_cast<int>(s);
        const int col = a.tslot[s];
        if (col >= 0) a.out[tid * a.nt + col] = 1;
    }
    grid.sync();

    uint32_t *front_cur = a.front_a, *front_new = a.front_b;
    int32_t *list_cur = a.list_a, *list_new = a.list_b;
    int cur = 0;
    for (int level = 1;; ++level) {
        const int n_cur = a.counts[cur];
        if (n_cur == 0) break;
        // expand: offer this level's new bits to the neighbours that do not have them yet
        for (long long i = tid; i < n_cur; i += stride) {
            const long long v = list_cur[i];
            for (int64_t e = a.indptr[v]; e < a.indptr[v + 1]; ++e) {
                const long long u = a.indices[e];
                bool any = false;
                for (int w = 0; w < W; ++w) {
                    const uint32_t nw = front_cur[v * W + w] & ~a.visited[u * W + w];
                    if (nw) {
                        atomicOr(&a.next[u * W + w], nw);
                        any = true;
                    }
                }
                if (any && atomicExch(&a.mark[u], level) != level)
                    list_new[atomicAdd(&a.counts[cur ^ 1], 1)] = static_cast<int>(u);
            }
        }
        grid.sync();
        // every thread has read counts[cur] (n_cur) before that barrier: it can be cleared now
        // and serves as the output counter of the next level's expansion
        if (tid == 0) a.counts[cur] = 0;
        // commit: the offered bits are all new (visited did not change during the expansion)
        const int n_new = a.counts[cur ^ 1];
        for (long long i = tid; i < n_new; i += stride) {
            const long long u = list_new[i];
            const int col = a.tslot[u];
            for (int w = 0; w < W; ++w) {
                const uint32_t nx = a.next[u * W + w];
                a.next[u * W + w] = 0;
                a.visited[u * W + w] |= nx;
                front_new[u * W + w] = nx;
                if (col >= 0)
                    for (uint32_t m = nx; m; m &= m - 1)
                        a.out[static_cast<long long>(32 * w + __ffs(m) - 1) * a.nt + col] = level + 1;
            }
        }
        for (long long i = tid; i < n_cur; i += stride) {
            const long long v = list_cur[i];
            for (int w = 0; w < W; ++w) front_cur[v * W + w] = 0;
        }
        grid.sync();
        cur ^= 1;
        uint32_t *tf = front_cur; front_cur = front_new; front_new = tf;
        int32_t *tl = list_cur; list_cur = list_new; list_new = tl;
    }
}

}  // namespace

// d_work: path_work_words(W) * n_vertices int32 for the W this call picks (see
// mcg_path_words); d_out [ns, nt] int32.  More sources than 32 W go in several launches.
// Measured (100 k vertices, 200 sources): W = 7 in one launch took 508 ms against 46 ms for seven
// launches of W = 1 -- a vertex re-enters the frontier once per arrival level of any source,
// and with W words every entry walks all W of them; the barriers were never the cost.
int mcg_path_words(int64_t ns) {
    (void)ns;
    return 1;
}

int mcg_k_path_lengths(mcg_ctx *ctx, const int64_t *d_indptr, const int32_t *d_indices,
                       int64_t n_vertices, const int32_t *d_tslot, const int64_t *d_sources,
                       int64_t ns, int64_t nt, int32_t *d_work, int *d_counts, int32_t *d_out) {
    if (ns == 0 || nt == 0) return MCG_OK;
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        int supported = 0;
        MCG_CUDA(ctx, cudaDeviceGetAttribute(&supported, cudaDevAttrCooperativeLaunch, ctx->device));
        if (!supported) return mcg_fail(ctx, MCG_ERR_CUDA, "cooperative launch not supported");
        MCG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm,
                                                                    path_lengths_kernel, 256, 0));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        // frontiers are thousands of vertices, not millions: more CTAs only make the grid
        // barrier (two per level) slower
        if (blocks_per_sm > 1) blocks_per_sm = 1;
    }
    MCG_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(int32_t) * ns * nt, ctx->stream));
    const int W = mcg_path_words(ns);
    const size_t n = static_cast<size_t>(n_vertices);
    for (int64_t s0 = 0; s0 < ns; s0 += 32LL * W) {
        MCG_CUDA(ctx, cudaMemsetAsync(d_work, 0, sizeof(int32_t) * (4 * W + 1) * n, ctx->stream));
        MCG_CUDA(ctx, cudaMemsetAsync(d_counts, 0, 2 * sizeof(int), ctx->stream));
        PathArgs a;
        a.indptr = d_indptr;
        a.indices = d_indices;
        a.visited = reinterpret_cast<uint32_t *>(d_work);
        a.front_a = a.visited + n * W;
        a.front_b = a.front_a + n * W;
        a.next = a.front_b + n * W;
        a.mark = d_work + 4 * n * W;
        a.list_a = a.mark + n;
        a.list_b = a.list_a + n;
        a.counts = d_counts;
        a.tslot = d_tslot;
        a.sources = d_sources + s0;
        a.ns = static_cast<int>(ns - s0 < 32LL * W ? ns - s0 : 32LL * W);
        a.W = W;
        a.nt = nt;
        a.out = d_out + s0 * nt;
        void *params[] = {&a};
        MCG_CUDA(ctx, cudaLaunchCooperativeKernel(reinterpret_cast<void *>(path_lengths_kernel),
                                                  dim3(ctx->sm_count * blocks_per_sm), dim3(256),
                                                  params, 0, ctx->stream));
        ctx->launches++;
    }
    return MCG_OK;
}

// ---- connected components (graph_utils.get_non_isolated) -------------------------------------------
//
//   reference: graph_utils.py:407-443 grows, for every binned contig i, the component of i by
//   repeated neighbour sweeps over a Python list (`neighbour not in component` is a linear scan,
//   so a component of c vertices costs O(c^2) per start vertex) and hands it to a process pool.
//
// Here one labelling of the whole CSR graph (FastSV: stochastic + aggressive hooking onto the
// grandparent, then shortcutting, all with atomicMin on a double-buffered parent array)
// gives every vertex the smallest vertex id of its component in O(log n) rounds; the host
// groups vertices by label.
namespace {

__global__ void cc_hook_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                               long long n, const int32_t *__restrict__ f, int32_t *__restrict__ fn) {
    const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (v >= n) return;
    const int32_t fv = f[v];
    const int32_t gv = f[fv];
    atomicMin(&fn[v], gv);                            // shortcutting
    for (int64_t e = indptr[v]; e < indptr[v + 1]; ++e) {
        const int32_t u = indices[e];
        const int32_t gu = f[f[u]];
        atomicMin(&fn[fv], gu);                       // stochastic hooking
        atomicMin(&fn[v], gu);                        // aggressive hooking
    }
}

__global__ void cc_diff_kernel(const int32_t *__restrict__ f, const int32_t *__restrict__ fn,
                               long long n, int *__restrict__ changed) {
    const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (v < n && f[v] != fn[v]) *changed = 1;
}

__global__ void cc_init_kernel(int32_t *__restrict__ f, int32_t *__restrict__ fn, long long n) {
    const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (v < n) { f[v] = static_cast<int32_t>(v); fn[v] = static_cast<int32_t>(v); }
}

}  // namespace

// d_work: 2 * n int32; on return d_labels[v] = smallest vertex id of v's component.
int mcg_k_components(mcg_ctx *ctx, const int64_t *d_indptr, const int32_t *d_indices,
                     int64_t n, int32_t *d_work, int *d_flag, int32_t *d_labels) {
    if (n == 0) return MCG_OK;
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((n + threads - 1) / threads);
    int32_t *f = d_work, *fn = d_work + n;
    cc_init_kernel<<<blocks, threads, 0, ctx->stream>>>(f, fn, n);
    MCG_KERNEL_CHECK(ctx);
    for (int round = 0; round < 64; ++round) {
        cc_hook_kernel<<<blocks, threads, 0, ctx->stream>>>(d_indptr, d_indices, n, f, fn);
        MCG_KERNEL_CHECK(ctx);
        MCG_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
        cc_diff_kernel<<<blocks, threads, 0, ctx->stream>>>(f, fn, n, d_flag);
        MCG_KERNEL_CHECK(ctx);
        int changed = 0;
        MCG_CUDA(ctx, cudaMemcpyAsync(&changed, d_flag, sizeof(int), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        MCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        MCG_CUDA(ctx, cudaMemcpyAsync(f, fn, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice,
                                      ctx->stream));
        if (!changed) {
            MCG_CUDA(ctx, cudaMemcpyAsync(d_labels, fn, sizeof(int32_t) * n,
                                          cudaMemcpyDeviceToDevice, ctx->stream));
            return MCG_OK;
        }
    }
    return mcg_fail(ctx, MCG_ERR_STATE, "connected components did not converge in 64 rounds");
}
