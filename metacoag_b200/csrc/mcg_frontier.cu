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
