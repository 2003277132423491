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
