// mcg_pool: the GPUs of one box behind ONE caller (include/mcg_b200.h, "device pool").
//
// The reference is one Python process (metacoag:191) whose only fan-out is a thread pool around
// label_prop_utils.assign_long (metacoag:959-999).  A drop-in therefore cannot ask its caller to
// run one process per GPU: the pool owns one mcg_ctx and one host worker thread per device and
// splits every call itself --
//   * featurisation from raw bases: contiguous contig ranges balanced by TOTAL BASES (the scan is
//     per base; contig lengths span 1 kb .. 5 Mb), each device uploads / packs / scans / scores
//     its own rows and writes its slice of the caller's output arrays;
//   * scoring of staged features (rules A / B / C from the dict-shaped API): the feature tables
//     are small against 180 GB (5.4 + 12 GB at 5 M contigs x 100 samples), so they are
//     replicated and the QUERIES are split;
//   * label propagation: graph and labels replicated, start nodes split, hit lists concatenated
//     in start order.
// No data-path collective exists: the bin tables come from the caller's host memory to every
// device (234 KB - 3.5 MB), results go straight to the caller's arrays.  Results do not depend on
// the number of devices: the kernels a call would take on one device are forced on every slice.
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <thread>
#include <vector>

#include "mcg_internal.cuh"

struct mcg_pool {
    std::vector<mcg_ctx *> ctx;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::function<int(int)> job;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
    std::vector<int> rc;
    std::mutex call_mu;              // one pool call at a time
    std::string err;
    std::vector<int64_t> bounds;     // contig ranges of the last sharded call
};

namespace {

int pool_fail(mcg_pool *pool, int code, const char *msg) {
    pool->err = msg;
    mcg_set_global_error(msg);
    return code;
}

void worker_main(mcg_pool *pool, int i) {
    cudaSetDevice(pool->ctx[i]->device);
    uint64_t seen = 0;
    for (;;) {
        std::function<int(int)> job;
        {
            std::unique_lock<std::mutex> lock(pool->mu);
            pool->cv_work.wait(lock, [&] { return pool->stop || pool->generation != seen; });
            if (pool->stop) return;
            seen = pool->generation;
            job = pool->job;
        }
        const int rc = job(i);
        {
            std::lock_guard<std::mutex> lock(pool->mu);
            pool->rc[i] = rc;
            if (--pool->pending == 0) pool->cv_done.notify_all();
        }
    }
}

// fn(i) on the worker of every device; the first failure wins.
int pool_run(mcg_pool *pool, const std::function<int(int)> &fn) {
    const int n = static_cast<int>(pool->ctx.size());
    int rc = MCG_OK, bad = -1;
    if (n == 1) {
        rc = fn(0);
        bad = 0;
    } else {
        {
            std::lock_guard<std::mutex> lock(pool->mu);
            pool->job = fn;
            pool->pending = n;
            ++pool->generation;
        }
        pool->cv_work.notify_all();
        {
            std::unique_lock<std::mutex> lock(pool->mu);
            pool->cv_done.wait(lock, [&] { return pool->pending == 0; });
        }
        for (int i = 0; i < n && rc == MCG_OK; ++i)
            if (pool->rc[i] != MCG_OK) { rc = pool->rc[i]; bad = i; }
    }
    if (rc != MCG_OK && bad >= 0) {
        pool->err = "device " + std::to_string(pool->ctx[bad]->device) + ": " + pool->ctx[bad]->err;
        mcg_set_global_error(pool->err.c_str());
    }
    return rc;
}

// contiguous ranges with (nearly) equal total bases
void partition_by_bases(const int64_t *offsets, int64_t n, int parts, std::vector<int64_t> *bounds) {
    bounds->assign(static_cast<size_t>(parts) + 1, 0);
    (*bounds)[parts] = n;
    const int64_t total = n > 0 ? offsets[n] - offsets[0] : 0;
    for (int r = 1; r < parts; ++r) {
        const int64_t target = offsets[0] + total / parts * r;
        (*bounds)[r] = std::lower_bound(offsets, offsets + n + 1, target) - offsets;
        if ((*bounds)[r] > n) (*bounds)[r] = n;
        if ((*bounds)[r] < (*bounds)[r - 1]) (*bounds)[r] = (*bounds)[r - 1];
    }
}

void split_even(int64_t n, int parts, std::vector<int64_t> *bounds) {
    bounds->assign(static_cast<size_t>(parts) + 1, 0);
    for (int r = 0; r <= parts; ++r) (*bounds)[r] = n / parts * r + std::min<int64_t>(r, n % parts);
}

// 32-bit words of contigs [c0, c1) in the MCG_ENC_2BIT layout
int64_t packed_words(const int64_t *offsets, int64_t c0, int64_t c1) {
    int64_t words = 0;
    for (int64_t c = c0; c < c1; ++c) words += 4 * ((offsets[c + 1] - offsets[c] + 63) / 64);
    return words;
}

struct CallGuard {
    std::unique_lock<std::mutex> lock;
    explicit CallGuard(mcg_pool *p) : lock(p->call_mu) {}
};

}  // namespace

extern "C" mcg_pool *mcg_pool_create(const int *devices, int n) {
    std::vector<int> devs;
    if (devices && n > 0) {
        devs.assign(devices, devices + n);
    } else {
        const int count = mcg_device_count();
        int want = count;
        if (const char *env = getenv("MCG_DEVICES"))   // "4" = the first four devices
            if (atoi(env) > 0) want = std::min(count, atoi(env));
        for (int d = 0; d < want; ++d) devs.push_back(d);
    }
    if (devs.empty()) {
        mcg_set_global_error("no CUDA device: libmcg_b200 has no CPU path");
        return nullptr;
    }
    if (devs.size() > MCG_MAX_PEERS) devs.resize(MCG_MAX_PEERS);
    mcg_pool *pool = new mcg_pool();
    for (int d : devs) {
        mcg_ctx *c = mcg_create(d);
        if (!c) {
            for (mcg_ctx *x : pool->ctx) mcg_destroy(x);
            delete pool;
            return nullptr;
        }
        pool->ctx.push_back(c);
    }
    const int nd = static_cast<int>(pool->ctx.size());
    pool->rc.assign(nd, 0);
    // the devices of the pool share the host cores for the ASCII -> 2 bit side of an upload
    const int hw = static_cast<int>(std::thread::hardware_concurrency());
    const char *env = getenv("MCG_HOST_PACK_THREADS");
    if (nd > 1 && !(env && *env))
        for (mcg_ctx *c : pool->ctx) c->host_pack_threads = std::max(0, (hw - nd) / nd);
    if (nd > 1)
        for (int i = 0; i < nd; ++i) pool->workers.emplace_back(worker_main, pool, i);
    return pool;
}

extern "C" void mcg_pool_destroy(mcg_pool *pool) {
    if (!pool) return;
    {
        std::lock_guard<std::mutex> lock(pool->mu);
        pool->stop = true;
    }
    pool->cv_work.notify_all();
    for (std::thread &t : pool->workers) t.join();
    for (mcg_ctx *c : pool->ctx) mcg_destroy(c);
    delete pool;
}

extern "C" int mcg_pool_size(const mcg_pool *pool) {
    return pool ? static_cast<int>(pool->ctx.size()) : 0;
}

extern "C" mcg_ctx *mcg_pool_ctx(mcg_pool *pool, int i) {
    if (!pool || i < 0 || i >= static_cast<int>(pool->ctx.size())) return nullptr;
    return pool->ctx[i];
}

extern "C" const char *mcg_pool_last_error(const mcg_pool *pool) {
    return pool ? pool->err.c_str() : mcg_last_error(nullptr);
}

extern "C" int mcg_pool_bounds(mcg_pool *pool, int64_t *bounds) {
    if (!pool || !bounds) return MCG_ERR_ARG;
    CallGuard guard(pool);
    const size_t n = pool->ctx.size() + 1;
    for (size_t i = 0; i < n; ++i) bounds[i] = i < pool->bounds.size() ? pool->bounds[i] : 0;
    return MCG_OK;
}

// ---- featurisation from raw bases: rows sharded by bases ----------------------------------------
namespace {

struct Shard {
    const uint8_t *bases;
    std::vector<int64_t> offsets;   // rebased to 0
    int64_t c0, c1;
};

// the slice of (bases, offsets) device i owns; 2-bit slices start at their packed word
int make_shards(mcg_pool *pool, const uint8_t *bases, const int64_t *offsets, int64_t n,
                int encoding, std::vector<Shard> *shards) {
    const int nd = static_cast<int>(pool->ctx.size());
    partition_by_bases(offsets, n, nd, &pool->bounds);
    shards->resize(nd);
    std::vector<int64_t> words(nd, 0);
    int rc = pool_run(pool, [&](int i) -> int {
        Shard &s = (*shards)[i];
        s.c0 = pool->bounds[i];
        s.c1 = pool->bounds[i + 1];
        s.offsets.resize(static_cast<size_t>(s.c1 - s.c0) + 1);
        const int64_t base = offsets[s.c0];
        for (int64_t c = s.c0; c <= s.c1; ++c) s.offsets[c - s.c0] = offsets[c] - base;
        if (encoding == MCG_ENC_2BIT) words[i] = packed_words(offsets, s.c0, s.c1);
        return MCG_OK;
    });
    if (rc != MCG_OK) return rc;
    int64_t w = 0;
    for (int i = 0; i < nd; ++i) {
        Shard &s = (*shards)[i];
        s.bases = encoding == MCG_ENC_2BIT ? bases + 4 * w : bases + offsets[s.c0];
        w += words[i];
    }
    return MCG_OK;
}

}  // namespace

extern "C" int mcg_pool_tetramer_scan(mcg_pool *pool, const uint8_t *bases, const int64_t *offsets,
                                      int64_t n, int encoding, uint32_t *counts, double *freqs) {
    if (!pool) return MCG_ERR_ARG;
    CallGuard guard(pool);
    if (n < 0 || (n > 0 && !offsets)) return pool_fail(pool, MCG_ERR_ARG, "bad contig arguments");
    if (n == 0) return MCG_OK;
    if (offsets[0] != 0) return pool_fail(pool, MCG_ERR_ARG, "offsets[0] must be 0");
    std::vector<Shard> shards;
    MCG_TRY(make_shards(pool, bases, offsets, n, encoding, &shards));
    return pool_run(pool, [&](int i) -> int {
        const Shard &s = shards[i];
        if (s.c1 == s.c0) return MCG_OK;
        return mcg_tetramer_scan(pool->ctx[i], s.bases, s.offsets.data(), s.c1 - s.c0, encoding,
                                 counts ? counts + s.c0 * MCG_NPROF : nullptr,
                                 freqs ? freqs + s.c0 * MCG_NPROF : nullptr);
    });
}

extern "C" int mcg_pool_featurise_score(mcg_pool *pool, const uint8_t *bases, const int64_t *offsets,
                                        int64_t n, int encoding, const double *cov, int n_samples,
                                        const double *cen_prof, const double *cen_cov,
                                        int64_t n_bins, int32_t *argmin, double *wmin) {
    if (!pool) return MCG_ERR_ARG;
    CallGuard guard(pool);
    if (n <= 0 || !offsets || !cov || n_samples <= 0)
        return pool_fail(pool, MCG_ERR_ARG, "bad arguments");
    if (offsets[0] != 0) return pool_fail(pool, MCG_ERR_ARG, "offsets[0] must be 0");
    std::vector<Shard> shards;
    MCG_TRY(make_shards(pool, bases, offsets, n, encoding, &shards));
    return pool_run(pool, [&](int i) -> int {
        const Shard &s = shards[i];
        if (s.c1 == s.c0) return MCG_OK;
        mcg_ctx *c = pool->ctx[i];
        // the kernels the whole set takes on one device, whatever the size of the slice
        c->force_split_calls = n >= 2048;
        const int rc = mcg_featurise_score(c, s.bases, s.offsets.data(), s.c1 - s.c0, encoding,
                                           cov + s.c0 * n_samples, n_samples, cen_prof, cen_cov,
                                           n_bins, argmin ? argmin + s.c0 : nullptr,
                                           wmin ? wmin + s.c0 : nullptr);
        c->force_split_calls = false;
        return rc;
    });
}

// ---- staged features: tables replicated, queries split ----------------------------------------------
extern "C" int mcg_pool_set_features(mcg_pool *pool, const double *prof, int64_t n_prof,
                                     const double *cov, int64_t n_cov, int n_samples) {
    if (!pool) return MCG_ERR_ARG;
    CallGuard guard(pool);
    return pool_run(pool, [&](int i) -> int {
        mcg_ctx *c = pool->ctx[i];
        if (prof) MCG_TRY(mcg_set_profiles(c, prof, n_prof));
        if (cov) MCG_TRY(mcg_load_coverage(c, cov, n_cov, n_samples));
        return MCG_OK;
    });
}

extern "C" int mcg_pool_bin_profiles(mcg_pool *pool, const int64_t *members,
                                     const int64_t *seg_offsets, int64_t nseg, double *cen_prof,
                                     double *cen_cov) {
    if (!pool) return MCG_ERR_ARG;
    CallGuard guard(pool);
    // B x (136 + S) numbers: every device computes (and keeps) all of them, device 0 reports
    return pool_run(pool, [&](int i) -> int {
        return mcg_bin_profiles(pool->ctx[i], members, seg_offsets, nseg, i == 0 ? cen_prof : nullptr,
                                i == 0 ? cen_cov : nullptr);
    });
}

extern "C" int mcg_pool_score_centroids(mcg_pool *pool, const int64_t *queries, int64_t nq,
                                        const double *cen_prof, const double *cen_cov,
                                        int64_t n_bins, double *weights, int32_t *argmin,
                                        double *wmin) {
    if (!pool) return MCG_ERR_ARG;
    CallGuard guard(pool);
    if (nq < 0) return pool_fail(pool, MCG_ERR_ARG, "bad arguments");
    const int nd = static_cast<int>(pool->ctx.size());
    std::vector<int64_t> cut;
    split_even(nq, nd, &cut);
    return pool_run(pool, [&](int i) -> int {
        mcg_ctx *c = pool->ctx[i];
        const int64_t q0 = cut[i], q1 = cut[i + 1];
        std::vector<int64_t> iota;
        const int64_t *q = queries ? queries + q0 : nullptr;
        if (!queries && q0 > 0) {
            iota.resize(static_cast<size_t>(q1 - q0));
            for (int64_t k = q0; k < q1; ++k) iota[k - q0] = k;
            q = iota.data();
        }
        c->force_split_calls = nq >= 2048;
        // every device takes the centroids even when it has no queries (later calls with
        // cen_prof = NULL read the resident tables)
        const int rc = mcg_score_centroids(c, q, q1 - q0, cen_prof, cen_cov, n_bins,
                                           weights ? weights + q0 * n_bins : nullptr,
                                           argmin ? argmin + q0 : nullptr, wmin ? wmin + q0 : nullptr);
        c->force_split_calls = false;
        return rc;
    });
}

extern "C" int mcg_pool_score_members(mcg_pool *pool, const int64_t *queries, int64_t nq,
                                      const int64_t *members, const int64_t *seg_offsets,
                                      int64_t nseg, int rule, double *weights) {
    if (!pool) return MCG_ERR_ARG;
    CallGuard guard(pool);
    if (nq < 0 || !weights) return pool_fail(pool, MCG_ERR_ARG, "bad arguments");
    const int nd = static_cast<int>(pool->ctx.size());
    std::vector<int64_t> cut;
    split_even(nq, nd, &cut);
    return pool_run(pool, [&](int i) -> int {
        const int64_t q0 = cut[i], q1 = cut[i + 1];
        if (q1 == q0) return MCG_OK;
        return mcg_score_members(pool->ctx[i], queries + q0, q1 - q0, members, seg_offsets, nseg,
                                 rule, weights + q0 * nseg);
    });
}

extern "C" int mcg_pool_score_member_items(mcg_pool *pool, const int64_t *item_query,
                                           const int32_t *item_bin, int64_t n_items,
                                           const int64_t *members, const int64_t *seg_offsets,
                                           int64_t nseg, int rule, double *weights) {
    if (!pool) return MCG_ERR_ARG;
    CallGuard guard(pool);
    if (n_items < 0) return pool_fail(pool, MCG_ERR_ARG, "bad arguments");
    const int nd = static_cast<int>(pool->ctx.size());
    // a handful of items (the contigs around one accepted contig) stay on one device
    std::vector<int64_t> cut;
    split_even(n_items, n_items < 4096 ? 1 : nd, &cut);
    cut.resize(static_cast<size_t>(nd) + 1, n_items);
    return pool_run(pool, [&](int i) -> int {
        const int64_t a = cut[i], b = cut[i + 1];
        if (b == a) return MCG_OK;
        return mcg_score_member_items(pool->ctx[i], item_query + a, item_bin + a, b - a, members,
                                      seg_offsets, nseg, rule, weights + a);
    });
}

// ---- label propagation: graph and labels replicated, start nodes split ---------------------------------
extern "C" int mcg_pool_graph_load(mcg_pool *pool, const int64_t *indptr, const int32_t *indices,
                                   int64_t n_vertices) {
    if (!pool) return MCG_ERR_ARG;
    CallGuard guard(pool);
    return pool_run(pool, [&](int i) -> int {
        return mcg_graph_load(pool->ctx[i], indptr, indices, n_vertices);
    });
}

extern "C" int mcg_pool_labels_load(mcg_pool *pool, const int32_t *labels, int64_t n_vertices) {
    if (!pool) return MCG_ERR_ARG;
    CallGuard guard(pool);
    return pool_run(pool, [&](int i) -> int {
        return mcg_labels_load(pool->ctx[i], labels, n_vertices);
    });
}

extern "C" int mcg_pool_labels_set(mcg_pool *pool, const int64_t *ids, const int32_t *bins,
                                   int64_t n) {
    if (!pool) return MCG_ERR_ARG;
    CallGuard guard(pool);
    // a patch of a few entries is a kernel launch per device: cheaper from this thread than
    // through the workers
    if (n <= 16) {
        for (mcg_ctx *c : pool->ctx) {
            const int rc = mcg_labels_set(c, ids, bins, n);
            if (rc != MCG_OK) {
                pool->err = c->err;
                return rc;
            }
        }
        return MCG_OK;
    }
    return pool_run(pool, [&](int i) -> int { return mcg_labels_set(pool->ctx[i], ids, bins, n); });
}

extern "C" int mcg_pool_bfs(mcg_pool *pool, const int64_t *starts, int64_t n_starts,
                            int depth_limit, int64_t *hit_start, int32_t *hit_count, int32_t *hits,
                            int64_t hit_cap, int64_t *total) {
    if (!pool) return MCG_ERR_ARG;
    CallGuard guard(pool);
    if (n_starts < 0 || !total) return pool_fail(pool, MCG_ERR_ARG, "bad arguments");
    const int nd = static_cast<int>(pool->ctx.size());
    if (nd == 1 || n_starts < 4096)
        return pool_run(pool, [&](int i) -> int {
            if (i != 0) return MCG_OK;
            return mcg_bfs_resident(pool->ctx[0], starts, n_starts, depth_limit, hit_start, hit_count,
                                    hits, hit_cap, total);
        });
    std::vector<int64_t> cut;
    split_even(n_starts, nd, &cut);
    // every device walks its share and keeps the hits; the totals lay the shares end to end in
    // start order, and every device copies its list straight to its place in the caller's
    std::vector<int64_t> part_total(nd, 0);
    const int64_t share = hit_cap / nd + 1;
    int rc = pool_run(pool, [&](int i) -> int {
        const int64_t a = cut[i], b = cut[i + 1];
        if (b == a) return MCG_OK;
        int64_t cap = share;
        for (;;) {
            MCG_TRY(mcg_bfs_run(pool->ctx[i], starts + a, b - a, depth_limit, cap, &part_total[i]));
            if (part_total[i] <= cap) return MCG_OK;
            cap = part_total[i];
        }
    });
    if (rc != MCG_OK) return rc;
    std::vector<int64_t> base(nd + 1, 0);
    for (int i = 0; i < nd; ++i) base[i + 1] = base[i] + part_total[i];
    *total = base[nd];
    if (base[nd] > hit_cap) return MCG_OK;
    rc = pool_run(pool, [&](int i) -> int {
        const int64_t a = cut[i], b = cut[i + 1];
        if (b == a) return MCG_OK;
        MCG_TRY(mcg_bfs_fetch(pool->ctx[i], b - a, hit_start + a, hit_count + a, hits + 3 * base[i],
                              part_total[i]));
        if (base[i])
            for (int64_t k = a; k < b; ++k) hit_start[k] += base[i];
        return MCG_OK;
    });
    if (rc != MCG_OK) return rc;
    return MCG_OK;
}
