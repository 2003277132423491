// C ABI of libmcg_b200.so (include/mcg_b200.h): context, staging and the host-facing entry
// points.  The kernels live in mcg_scan.cu, mcg_score.cu and mcg_frontier.cu.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>

#include "mcg_internal.cuh"

// ---- errors ---------------------------------------------------------------------------
static std::mutex g_err_mu;
static std::string g_err;

void mcg_set_global_error(const char *msg) {
    std::lock_guard<std::mutex> lock(g_err_mu);
    g_err = msg;
}

int mcg_fail(mcg_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    mcg_set_global_error(buf);
    return code;
}

// ---- small helpers ----------------------------------------------------------------------
int mcg_reserve(mcg_ctx *ctx, DevBuf &buf, size_t bytes) {
    if (bytes <= buf.cap && buf.p) return MCG_OK;
    if (buf.p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(buf.p);
        buf.p = nullptr;
        buf.cap = 0;
    }
    size_t want = bytes < 256 ? 256 : bytes;
    cudaError_t e = cudaMalloc(&buf.p, want);
    if (e != cudaSuccess) {
        buf.p = nullptr;
        return mcg_fail(ctx, MCG_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want,
                        cudaGetErrorString(e));
    }
    buf.cap = want;
    return MCG_OK;
}

int mcg_pinned_reserve(mcg_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pinned_cap) return MCG_OK;
    if (ctx->pinned) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_cap = 0;
    }
    cudaError_t e = cudaMallocHost(&ctx->pinned, bytes);
    if (e != cudaSuccess)
        return mcg_fail(ctx, MCG_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes,
                        cudaGetErrorString(e));
    ctx->pinned_cap = bytes;
    return MCG_OK;
}

void mcg_timer_begin(mcg_ctx *ctx, int slot) {
    if (!ctx->profiling) return;
    cudaEventRecord(ctx->ev[2 * slot], ctx->stream);
}

void mcg_timer_end(mcg_ctx *ctx, int slot) {
    if (!ctx->profiling) return;
    cudaEventRecord(ctx->ev[2 * slot + 1], ctx->stream);
    ctx->ev_used[slot] = true;
}

int mcg_check_status(mcg_ctx *ctx) {
    int32_t flags[4] = {0, 0, 0, 0};
    MCG_CUDA(ctx, cudaMemcpyAsync(flags, ctx->status.p, sizeof flags, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    MCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flags[0]) cudaMemsetAsync(ctx->status.p, 0, sizeof flags, ctx->stream);
    if (flags[0] & 2)
        return mcg_fail(ctx, MCG_ERR_ARG, "coverage values must be finite (NaN or inf found)");
    if (flags[0] & 1)
        return mcg_fail(ctx, MCG_ERR_DOMAIN,
                        "float division by zero (get_comp_probability: both Gaussians underflow)");
    return MCG_OK;
}

static int h2d(mcg_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return MCG_OK;
    MCG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return MCG_OK;
}
static int d2h(mcg_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return MCG_OK;
    MCG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return MCG_OK;
}
static int sync(mcg_ctx *ctx) {
    MCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCG_OK;
}

// Serialises the calls of a context and makes its device current for the call; the caller's
// current device is put back on return (a pool talks to several devices from one thread, and a
// host program -- torch, for one -- keeps its own notion of the current device).
struct Guard {
    mcg_ctx *ctx;
    std::unique_lock<std::mutex> lock;
    int prev = -1;
    explicit Guard(mcg_ctx *c) : ctx(c), lock(c->mu) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != c->device) cudaSetDevice(c->device);
    }
    ~Guard() {
        if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    }
};

#define MCG_ENTER(ctx)                                   \
    if (!(ctx)) return MCG_ERR_ARG;                      \
    Guard guard__(ctx)

// ---- context ------------------------------------------------------------------------------
extern "C" int mcg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" mcg_ctx *mcg_create(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        mcg_set_global_error(
            "no CUDA device: libmcg_b200 has no CPU path (cudaGetDeviceCount failed or 0)");
        cudaGetLastError();
        return nullptr;
    }
    if (device < 0 || device >= n) {
        mcg_set_global_error("device index out of range");
        return nullptr;
    }
    int prev_device = -1;
    if (cudaGetDevice(&prev_device) != cudaSuccess) prev_device = -1;
    if (cudaSetDevice(device) != cudaSuccess) {
        mcg_set_global_error("cudaSetDevice failed");
        return nullptr;
    }
    struct Restore {
        int prev, cur;
        ~Restore() { if (prev >= 0 && prev != cur) cudaSetDevice(prev); }
    } restore{prev_device, device};
    mcg_ctx *ctx = new mcg_ctx();
    ctx->device = device;
    if (const char *env = getenv("MCG_PRUNE")) ctx->prune = atoi(env) != 0;
    if (const char *env = getenv("MCG_TC")) ctx->tc = atoi(env) != 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        mcg_set_global_error("cudaStreamCreate failed");
        delete ctx;
        return nullptr;
    }
    ctx->stream = ctx->own_stream;
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->copy2_done, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->cov_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->cov_ev, cudaEventDisableTiming);
    for (int i = 0; i < 8; ++i) cudaEventCreateWithFlags(&ctx->copy_ev[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->copy_start, cudaEventDisableTiming);
    for (int i = 0; i < 2 * MCG_N_TIMERS; ++i) cudaEventCreate(&ctx->ev[i]);
    if (mcg_reserve(ctx, ctx->status, 4 * sizeof(int32_t)) != MCG_OK) {
        delete ctx;
        return nullptr;
    }
    cudaMemsetAsync(ctx->status.p, 0, 4 * sizeof(int32_t), ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return ctx;
}

extern "C" void mcg_destroy(mcg_ctx *ctx) {
    if (!ctx) return;
    int prev_device = -1;
    if (cudaGetDevice(&prev_device) != cudaSuccess) prev_device = -1;
    struct Restore {
        int prev, cur;
        ~Restore() { if (prev >= 0 && prev != cur) cudaSetDevice(prev); }
    } restore{prev_device, ctx->device};
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->packed, &ctx->meta, &ctx->seg_contig, &ctx->ascii, &ctx->offsets,
                      &ctx->tile_counter, &ctx->counts, &ctx->prof, &ctx->cov, &ctx->logc,
                      &ctx->lgam, &ctx->cen_prof, &ctx->cen_cov, &ctx->cen_logc, &ctx->cen_lgam,
                      &ctx->argmin, &ctx->wmin, &ctx->status, &ctx->prof_norm, &ctx->cen_norm,
                      &ctx->g_indptr, &ctx->g_indices, &ctx->g_labels, &ctx->g_work, &ctx->g_block,
                      &ctx->g_starts, &ctx->g_patch, &ctx->tc_a, &ctx->tc_b, &ctx->tc_aaux,
                      &ctx->tc_baux, &ctx->tc_d2a, &ctx->tc_mrow};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (DevBuf &b : ctx->scratch)
        if (b.p) cudaFree(b.p);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->g_pinned) cudaFreeHost(ctx->g_pinned);
    if (ctx->pinned_pack) cudaFreeHost(ctx->pinned_pack);
    if (ctx->copy2_done) cudaEventDestroy(ctx->copy2_done);
    if (ctx->cov_ev) cudaEventDestroy(ctx->cov_ev);
    if (ctx->cov_stream) cudaStreamDestroy(ctx->cov_stream);
    if (ctx->copy_stream2) cudaStreamDestroy(ctx->copy_stream2);
    for (int i = 0; i < 2 * MCG_N_TIMERS; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 8; ++i)
        if (ctx->copy_ev[i]) cudaEventDestroy(ctx->copy_ev[i]);
    if (ctx->copy_start) cudaEventDestroy(ctx->copy_start);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" const char *mcg_last_error(const mcg_ctx *ctx) {
    if (ctx) return ctx->err.c_str();
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lock(g_err_mu);
    copy = g_err;
    return copy.c_str();
}

extern "C" int mcg_set_stream(mcg_ctx *ctx, void *cuda_stream) {
    MCG_ENTER(ctx);
    MCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return MCG_OK;
}

extern "C" int mcg_synchronize(mcg_ctx *ctx) {
    MCG_ENTER(ctx);
    return sync(ctx);
}

extern "C" void *mcg_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void mcg_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ---- canonical tetramer table (feature_utils.py:38-57) ----------------------------------------
extern "C" int mcg_kmer_table(uint8_t table[256]) {
    // A code and its reverse complement share a slot; slots are handed out in ascending
    // code order (the reference walks the sorted 4-mers, which is ascending code order).
    int slot_of[256];
    for (int i = 0; i < 256; ++i) slot_of[i] = -1;
    int next = 0;
    for (int code = 0; code < 256; ++code) {
        int rc = 0, c = code;
        for (int i = 0; i < 4; ++i) { rc = (rc << 2) | (3 - (c & 3)); c >>= 2; }
        if (slot_of[rc] >= 0) slot_of[code] = slot_of[rc];
        else slot_of[code] = next++;
        table[code] = static_cast<uint8_t>(slot_of[code]);
    }
    return next == MCG_NPROF ? MCG_OK : MCG_ERR_STATE;
}

// ---- contig store ---------------------------------------------------------------------------------
// Host-side layout pass: per-contig packed word offset, first segment and window count.
static int build_layout(mcg_ctx *ctx, const int64_t *offsets, int64_t n,
                        std::vector<ContigMeta> &meta, int64_t *n_words, int64_t *n_seg) {
    meta.resize(static_cast<size_t>(n));
    int64_t words = 0, segs = 0;
    for (int64_t c = 0; c < n; ++c) {
        const int64_t len = offsets[c + 1] - offsets[c];
        if (len < 0) return mcg_fail(ctx, MCG_ERR_ARG, "offsets must be non-decreasing");
        if (len > 0x7fffff00LL) return mcg_fail(ctx, MCG_ERR_ARG, "contig longer than 2^31 bases");
        const int64_t nwin = len >= 4 ? len - 3 : 0;
        if (segs > 0x7fffffffLL) return mcg_fail(ctx, MCG_ERR_ARG, "too many scan segments");
        meta[c].pk_word = words;
        meta[c].seg_first = static_cast<int32_t>(segs);
        meta[c].nwin = static_cast<int32_t>(nwin);
        words += 4 * ((len + 63) / 64);
        segs += (nwin + 255) / 256;
    }
    if (segs > 0x7fffffffLL) return mcg_fail(ctx, MCG_ERR_ARG, "too many scan segments");
    *n_words = words;
    *n_seg = segs;
    return MCG_OK;
}

// MCG_TRACE=1: host wall-clock marks of the ingest phases on stderr (development aid)
static double trace_now() {
    return std::chrono::duration<double, std::milli>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
}
static bool trace_on() {
    static const bool on = getenv("MCG_TRACE") != nullptr;
    return on;
}
#define MCG_MARK(label) do { if (trace_on()) fprintf(stderr, "[mcg] %-22s %.3f ms\n", label, trace_now()); } while (0)

// ---- ASCII ingest: copy engine + GPU pack kernel from the front, host threads + 4x smaller
// copies from the back ------------------------------------------------------------------------
extern "C" void mcg_hostpack_range(const uint8_t *bases, const int64_t *offsets, int64_t c0,
                                   int64_t c1, const void *first_word, int64_t word_stride_bytes,
                                   uint8_t *packed, int level);

static int host_pack_threads(mcg_ctx *ctx, int64_t n_bases) {
    if (n_bases < (16 << 20)) return 0;   // small loads: the plain upload is already short
    int t = ctx->host_pack_threads;
    if (t < 0) {
        const char *env = getenv("MCG_HOST_PACK_THREADS");
        if (env && *env) t = atoi(env);
        // one core stays with the calling thread: it feeds the copy engine, and when it is
        // starved by its own workers the split degenerates (measured: 12.3 ms with 15 of 16
        // threads, 14.5-15 ms with 16)
        else {
            // the ranks of one box (torchrun exports LOCAL_WORLD_SIZE) share its cores
            int local_world = 1;
            const char *lw = getenv("LOCAL_WORLD_SIZE");
            if (lw && atoi(lw) > 0) local_world = atoi(lw);
            t = static_cast<int>(std::thread::hardware_concurrency()) / local_world - 1;
        }
    }
    return t < 0 ? 0 : (t > 256 ? 256 : t);
}

// The contig blob is cut into chunks of whole contigs (~1 MB of bases).  The calling thread
// feeds the copy engine with runs of ASCII chunks taken from the front (~8 MB per copy, at
// most two in flight, each followed by its pack kernel on the compute stream); `threads` host
// threads take single chunks from the back, pack them into pinned memory (mcg_hostpack.cpp)
// and the calling thread copies the packed result, adjacent chunks in one copy.  Both sides
// stop when they meet, so the split follows the measured speeds of the link and of the host,
// and the small host chunks keep the tail (workers finishing while the link idles) short.
// Output is the same packed store either way.
//
// `wave` (optional): called on the compute stream's side with a contig range [c0, c1) whose
// packed words are complete in stream order -- the caller scans and scores it while the rest
// of the blob is still on its way.  Ranges come from both ends and cover every contig once.
typedef std::function<int(int64_t, int64_t)> WaveFn;

static int upload_ascii(mcg_ctx *ctx, const uint8_t *bases, const int64_t *offsets, int64_t n,
                        const std::vector<ContigMeta> &meta, int64_t n_words,
                        const WaveFn *wave = nullptr, int64_t wave_bases = 0) {
    if (n == 0) return MCG_OK;
    const int64_t n_bases = offsets[n];
    const int threads = host_pack_threads(ctx, n_bases);
    struct Chunk { int64_t c0, c1; };
    std::vector<Chunk> chunks;
    const int64_t chunk_bytes = threads > 0 ? (1 << 20) : (32 << 20);
    const int64_t copy_bytes = threads > 0 ? (8 << 20) : (32 << 20);   // ASCII bytes per copy
    const int64_t packed_copy_words = (1 << 20) / 4;   // packed copies are batched to >= 1 MB
    for (int64_t c0 = 0; c0 < n;) {
        int64_t c1 = c0;
        while (c1 < n && offsets[c1] - offsets[c0] < chunk_bytes) ++c1;
        chunks.push_back({c0, c1});
        c0 = c1;
    }
    const int64_t n_chunks = static_cast<int64_t>(chunks.size());
    if (n_chunks >= 0x7fffffffLL) return mcg_fail(ctx, MCG_ERR_ARG, "too many contig chunks");
    uint8_t *stage = nullptr;
    if (threads > 0) {
        const size_t need = sizeof(uint32_t) * static_cast<size_t>(n_words);
        if (need > ctx->pinned_pack_cap) {
            if (ctx->pinned_pack) {
                cudaStreamSynchronize(ctx->copy_stream2);
                cudaFreeHost(ctx->pinned_pack);
                ctx->pinned_pack = nullptr;
                ctx->pinned_pack_cap = 0;
            }
            if (cudaMallocHost(&ctx->pinned_pack, need) != cudaSuccess)
                return mcg_fail(ctx, MCG_ERR_NOMEM, "cudaMallocHost(%zu) failed", need);
            ctx->pinned_pack_cap = need;
        }
        stage = static_cast<uint8_t *>(ctx->pinned_pack);
    }
    MCG_CUDA(ctx, cudaEventRecord(ctx->copy_start, ctx->stream));
    MCG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_start, 0));
    MCG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream2, ctx->copy_start, 0));

    // Unclaimed chunks are [front, end): one 64-bit word, claimed by compare-and-swap from
    // either side, so that the polling caller never holds a lock the workers need.
    std::atomic<uint64_t> range(static_cast<uint64_t>(n_chunks));   // front << 32 | end
    std::vector<std::atomic<int64_t>> ready(static_cast<size_t>(n_chunks));   // packed chunks, in
    for (auto &r : ready) r.store(-1, std::memory_order_relaxed);             // completion order
    std::atomic<int64_t> ready_tail(0);
    std::atomic<int> workers_left(threads);
    auto claim_back = [&]() -> int64_t {
        uint64_t v = range.load(std::memory_order_relaxed);
        for (;;) {
            const uint64_t f = v >> 32, e = v & 0xffffffffu;
            if (f >= e) return -1;
            if (range.compare_exchange_weak(v, (f << 32) | (e - 1), std::memory_order_acq_rel))
                return static_cast<int64_t>(e - 1);
        }
    };
    auto claim_front_run = [&](Chunk *run) -> bool {
        uint64_t v = range.load(std::memory_order_relaxed);
        for (;;) {
            const uint64_t f = v >> 32, e = v & 0xffffffffu;
            if (f >= e) return false;
            uint64_t k = f;
            run->c0 = chunks[f].c0;
            do run->c1 = chunks[k++].c1;
            while (k < e && offsets[run->c1] - offsets[run->c0] < copy_bytes);
            if (range.compare_exchange_weak(v, (k << 32) | e, std::memory_order_acq_rel)) return true;
        }
    };
    int rc = MCG_OK;
    int64_t h2d_bytes = 0, host_bases = 0, gpu_bases = 0;
    // contigs [front_done, front_c) are packed in stream order but not yet handed to `wave`;
    // contigs [back_c, back_done) are copied on copy_stream2 but not yet handed to `wave`
    int64_t front_c = 0, front_done = 0, back_c = n, back_done = n;
    // The link is one resource: at most kWindow copies (ASCII runs and packed runs together)
    // are in flight, and packed runs go first -- they carry four times the bases per byte.
    // Without the shared window the ASCII side kept the copy engine busy on its own and the
    // packed copies piled up behind it (4.6 ms of drain after the host had finished).
    constexpr int kWindow = 3;
    int inflight[kWindow] = {-1, -1, -1};   // copy_ev slots of the copies on the wire
    int slot = 0;
    auto free_slot = [&]() -> int * {
        for (int &f : inflight)
            if (f < 0) return &f;
        return nullptr;
    };
    auto word_range = [&](const Chunk &ch, int64_t *w0, int64_t *w1) {
        *w0 = meta[ch.c0].pk_word;
        *w1 = ch.c1 < n ? meta[ch.c1].pk_word : n_words;
    };
    // the next ASCII run from the front into a free slot, its pack kernel behind it
    auto feed_ascii = [&]() -> bool {
        int *f = free_slot();
        if (!f) return false;
        Chunk ch;
        if (!claim_front_run(&ch)) return false;
        const size_t bytes = static_cast<size_t>(offsets[ch.c1] - offsets[ch.c0]);
        cudaError_t e = cudaSuccess;
        if (bytes)
            e = cudaMemcpyAsync(ctx->ascii.as<uint8_t>() + offsets[ch.c0], bases + offsets[ch.c0],
                                bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->copy_ev[slot], ctx->copy_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[slot], 0);
        if (e != cudaSuccess) {
            rc = mcg_fail(ctx, MCG_ERR_CUDA, "contig upload failed: %s", cudaGetErrorString(e));
            return false;
        }
        int64_t w0, w1;
        word_range(ch, &w0, &w1);
        rc = mcg_k_pack_ascii(ctx, ctx->ascii.as<uint8_t>(), ctx->offsets.as<int64_t>(),
                              ctx->meta.as<ContigMeta>(), n, w0, w1 - w0, ctx->packed.as<uint32_t>());
        if (rc != MCG_OK) return false;
        *f = slot;
        slot = (slot + 1) & 7;
        h2d_bytes += static_cast<int64_t>(bytes);
        gpu_bases += static_cast<int64_t>(bytes);
        front_c = ch.c1;
        if (wave && offsets[front_c] - offsets[front_done] >= wave_bases) {
            rc = (*wave)(front_done, front_c);
            front_done = front_c;
        }
        return rc == MCG_OK;
    };
    // the copy engine starts on the front before the pack threads are even created
    MCG_MARK("upload starts");
    feed_ascii();
    feed_ascii();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads && rc == MCG_OK; ++t)
        pool.emplace_back([&]() {
            for (;;) {
                const int64_t k = claim_back();
                if (k < 0) break;
                mcg_hostpack_range(bases, offsets, chunks[k].c0, chunks[k].c1, &meta[0].pk_word,
                                   sizeof(ContigMeta), stage, -1);
                ready[ready_tail.fetch_add(1, std::memory_order_relaxed)].store(
                    k, std::memory_order_release);
            }
            workers_left.fetch_sub(1, std::memory_order_release);
        });
    if (static_cast<int>(pool.size()) < threads) workers_left.store(0);

    // host-packed chunks complete roughly from the back; they are copied as runs that end at
    // `up_hi`, the highest chunk not yet copied
    std::vector<char> packed_done(static_cast<size_t>(n_chunks), 0);
    int64_t ready_head = 0, up_hi = n_chunks - 1;
    for (bool done = false; !done && rc == MCG_OK;) {
        bool progressed = false;
        bool link_idle = true;
        for (int &f : inflight) {   // retire the copies that have left
            if (f >= 0 && cudaEventQuery(ctx->copy_ev[f]) != cudaErrorNotReady) { f = -1; progressed = true; }
            if (f >= 0) link_idle = false;
        }
        // everything claimed and every worker gone: the remaining packed chunks are final
        const bool finishing = workers_left.load(std::memory_order_acquire) == 0 &&
                               (range.load(std::memory_order_acquire) >> 32) >=
                                   (range.load(std::memory_order_acquire) & 0xffffffffu);
        while (ready_head < n_chunks) {
            const int64_t k = ready[ready_head].load(std::memory_order_acquire);
            if (k < 0) break;
            packed_done[k] = 1;
            ++ready_head;
        }
        // packed runs first
        while (rc == MCG_OK && up_hi >= 0 && packed_done[up_hi]) {
            int *f = free_slot();
            if (!f) break;
            int64_t lo = up_hi;
            while (lo > 0 && packed_done[lo - 1]) --lo;
            const Chunk run = {chunks[lo].c0, chunks[up_hi].c1};
            int64_t w0, w1;
            word_range(run, &w0, &w1);
            // small runs wait for their neighbours unless the link has nothing else to do
            if (!finishing && !link_idle && w1 - w0 < packed_copy_words) break;
            const size_t bytes = sizeof(uint32_t) * static_cast<size_t>(w1 - w0);
            cudaError_t e = cudaSuccess;
            if (bytes)
                e = cudaMemcpyAsync(ctx->packed.as<uint32_t>() + w0, stage + 4 * w0, bytes,
                                    cudaMemcpyHostToDevice, ctx->copy_stream2);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->copy_ev[slot], ctx->copy_stream2);
            if (e != cudaSuccess) {
                rc = mcg_fail(ctx, MCG_ERR_CUDA, "packed upload failed: %s", cudaGetErrorString(e));
                break;
            }
            const int ev = slot;
            *f = slot;
            slot = (slot + 1) & 7;
            link_idle = false;
            h2d_bytes += static_cast<int64_t>(bytes);
            host_bases += offsets[run.c1] - offsets[run.c0];
            up_hi = lo - 1;
            progressed = true;
            back_c = run.c0;
            if (wave && offsets[back_done] - offsets[back_c] >= wave_bases) {
                e = cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[ev], 0);
                if (e != cudaSuccess)
                    rc = mcg_fail(ctx, MCG_ERR_CUDA, "wave sync failed: %s", cudaGetErrorString(e));
                else
                    rc = (*wave)(back_c, back_done);
                back_done = back_c;
            }
        }
        // then ASCII into whatever is left of the window
        while (rc == MCG_OK && feed_ascii()) progressed = true;
        // done: nothing left to claim, workers gone, and every packed chunk handed over (the
        // chunks below up_hi that are not packed_done went to the copy engine as ASCII)
        done = finishing && ready_head == ready_tail.load(std::memory_order_acquire) &&
               (up_hi < 0 || !packed_done[up_hi]);
        if (!progressed && !done) std::this_thread::yield();
    }
    MCG_MARK("all chunks queued");
    if (rc != MCG_OK) range.store(0);   // let the workers run out of chunks
    for (std::thread &t : pool) t.join();
    MCG_MARK("workers joined");
    if (rc != MCG_OK) return rc;
    MCG_CUDA(ctx, cudaEventRecord(ctx->copy2_done, ctx->copy_stream2));
    MCG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy2_done, 0));
    if (wave) {   // what is left at both ends (the two sides have met: front_c == back_c)
        if (front_c != back_c) return mcg_fail(ctx, MCG_ERR_STATE, "upload sides did not meet");
        MCG_TRY((*wave)(front_done, front_c));
        MCG_TRY((*wave)(back_c, back_done));
    }
    ctx->ingest_stats[0] = h2d_bytes;
    ctx->ingest_stats[1] = host_bases;
    ctx->ingest_stats[2] = gpu_bases;
    ctx->ingest_stats[3] = threads;
    return MCG_OK;
}

// `wave_factory` (optional, ASCII only): given the host layout, returns the function the upload
// calls for every contig range that has arrived (see upload_ascii).
typedef std::function<WaveFn(const std::vector<ContigMeta> &, int64_t)> WaveFactory;

static int load_contigs_locked(mcg_ctx *ctx, const uint8_t *bases, const int64_t *offsets,
                               int64_t n, int encoding, const WaveFactory *wave_factory = nullptr) {
    if (n < 0 || (!offsets && n > 0)) return mcg_fail(ctx, MCG_ERR_ARG, "bad contig arguments");
    if (encoding != MCG_ENC_ASCII && encoding != MCG_ENC_2BIT)
        return mcg_fail(ctx, MCG_ERR_ARG, "unknown encoding %d", encoding);
    std::vector<ContigMeta> meta;
    int64_t n_words = 0, n_seg = 0;
    static const int64_t zero_off[1] = {0};
    if (n == 0) offsets = zero_off;
    if (offsets[0] != 0) return mcg_fail(ctx, MCG_ERR_ARG, "offsets[0] must be 0");
    MCG_TRY(build_layout(ctx, offsets, n, meta, &n_words, &n_seg));
    MCG_MARK("layout built");
    const int64_t n_bases = offsets[n] - offsets[0];
    if (n_bases > 0 && !bases) return mcg_fail(ctx, MCG_ERR_ARG, "bases is NULL");

    MCG_TRY(mcg_reserve(ctx, ctx->packed, sizeof(uint32_t) * (n_words + 32)));
    MCG_TRY(mcg_reserve(ctx, ctx->meta, sizeof(ContigMeta) * std::max<int64_t>(n, 1)));
    MCG_TRY(mcg_reserve(ctx, ctx->seg_contig, sizeof(int32_t) * std::max<int64_t>(n_seg, 1)));
    MCG_TRY(mcg_pinned_reserve(ctx, sizeof(ContigMeta) * std::max<int64_t>(n, 1)));
    memcpy(ctx->pinned, meta.data(), sizeof(ContigMeta) * n);
    {
        TimerScope t(ctx, T_H2D);
        MCG_TRY(h2d(ctx, ctx->meta.p, ctx->pinned, sizeof(ContigMeta) * n));
        // the words after the last contig are read as halo / by idle lanes
        MCG_CUDA(ctx, cudaMemsetAsync(ctx->packed.as<uint32_t>() + n_words, 0,
                                      sizeof(uint32_t) * 32, ctx->stream));
        MCG_TRY(mcg_k_fill_seg_contig(ctx, ctx->meta.as<ContigMeta>(), n, n_seg,
                                      ctx->seg_contig.as<int32_t>()));
        if (encoding == MCG_ENC_ASCII) {
            MCG_TRY(mcg_reserve(ctx, ctx->ascii, static_cast<size_t>(n_bases) + 256));
            MCG_TRY(mcg_reserve(ctx, ctx->offsets, sizeof(int64_t) * (n + 1)));
            MCG_TRY(h2d(ctx, ctx->offsets.p, offsets, sizeof(int64_t) * (n + 1)));
            if (wave_factory) {
                ctx->n_contigs = n;
                ctx->n_words = n_words;
                ctx->n_seg = n_seg;
                ctx->n_bases = n_bases;
                const WaveFn wave = (*wave_factory)(meta, n_seg);
                // ~12 waves, none below 32 MB of bases
                const char *div_env = getenv("MCG_WAVE_DIV");
                const int64_t div = div_env && atoi(div_env) > 0 ? atoi(div_env) : 12;
                const int64_t wave_bases = std::max<int64_t>(32 << 20, n_bases / div);
                MCG_TRY(upload_ascii(ctx, bases, offsets, n, meta, n_words, &wave, wave_bases));
            } else {
                MCG_TRY(upload_ascii(ctx, bases, offsets, n, meta, n_words));
            }
        } else if (wave_factory && n_words > 0) {
            // 2-bit input (packed at parse time): the copy is cut at contig boundaries into
            // ~16 MB runs on the copy stream, and every run is scanned and scored behind its
            // own event while the next one is on the wire
            ctx->n_contigs = n;
            ctx->n_words = n_words;
            ctx->n_seg = n_seg;
            ctx->n_bases = n_bases;
            const WaveFn wave = (*wave_factory)(meta, n_seg);
            const int64_t run_words = std::max<int64_t>((4 << 20), n_words / 12);
            MCG_CUDA(ctx, cudaEventRecord(ctx->copy_start, ctx->stream));
            MCG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_start, 0));
            const uint32_t *src = reinterpret_cast<const uint32_t *>(bases);
            int slot = 0;
            for (int64_t c0 = 0; c0 < n;) {
                const int64_t w0 = meta[c0].pk_word;
                // first contig whose words start at or beyond w0 + run_words
                int64_t lo = c0 + 1, hi = n;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) / 2;
                    if (meta[mid].pk_word - w0 < run_words) lo = mid + 1; else hi = mid;
                }
                const int64_t c1 = lo;
                const int64_t w1 = c1 < n ? meta[c1].pk_word : n_words;
                if (w1 > w0)
                    MCG_CUDA(ctx, cudaMemcpyAsync(ctx->packed.as<uint32_t>() + w0, src + w0,
                                                  sizeof(uint32_t) * (w1 - w0), cudaMemcpyHostToDevice,
                                                  ctx->copy_stream));
                MCG_CUDA(ctx, cudaEventRecord(ctx->copy_ev[slot], ctx->copy_stream));
                MCG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[slot], 0));
                slot = (slot + 1) & 7;
                MCG_TRY(wave(c0, c1));
                c0 = c1;
            }
            ctx->ingest_stats[0] = static_cast<int64_t>(sizeof(uint32_t)) * n_words;
            ctx->ingest_stats[1] = ctx->ingest_stats[2] = ctx->ingest_stats[3] = 0;
        } else {
            MCG_TRY(h2d(ctx, ctx->packed.p, bases, sizeof(uint32_t) * n_words));
            ctx->ingest_stats[0] = static_cast<int64_t>(sizeof(uint32_t)) * n_words;
            ctx->ingest_stats[1] = ctx->ingest_stats[2] = ctx->ingest_stats[3] = 0;
        }
    }
    // the pinned meta block must not be overwritten before the copy has left
    MCG_MARK("before final sync");
    MCG_TRY(sync(ctx));
    MCG_MARK("after final sync");
    ctx->n_contigs = n;
    ctx->n_words = n_words;
    ctx->n_seg = n_seg;
    ctx->n_bases = n_bases;
    return MCG_OK;
}

static int scan_locked(mcg_ctx *ctx) {
    const int64_t n = ctx->n_contigs;
    MCG_TRY(mcg_reserve(ctx, ctx->counts, sizeof(uint32_t) * MCG_NPROF * std::max<int64_t>(n, 1)));
    MCG_TRY(mcg_reserve(ctx, ctx->prof, sizeof(double) * MCG_NPROF * std::max<int64_t>(n, 1)));
    {
        TimerScope t(ctx, T_SCAN);
        MCG_TRY(mcg_k_scan(ctx, ctx->packed.as<uint32_t>(), ctx->meta.as<ContigMeta>(),
                           ctx->seg_contig.as<int32_t>(), ctx->n_seg, n,
                           ctx->counts.as<uint32_t>()));
    }
    MCG_TRY(mcg_reserve(ctx, ctx->prof_norm, sizeof(double) * std::max<int64_t>(n, 1)));
    {
        TimerScope t(ctx, T_NORM);
        MCG_TRY(mcg_k_normalise_norms(ctx, ctx->counts.as<uint32_t>(), ctx->meta.as<ContigMeta>(),
                                      n, ctx->prof.as<double>(), ctx->prof_norm.as<double>()));
    }
    ctx->n_prof = n;
    return MCG_OK;
}

extern "C" int mcg_load_contigs(mcg_ctx *ctx, const uint8_t *bases, const int64_t *offsets,
                                int64_t n, int encoding) {
    MCG_ENTER(ctx);
    return load_contigs_locked(ctx, bases, offsets, n, encoding);
}

extern "C" int mcg_scan(mcg_ctx *ctx) {
    MCG_ENTER(ctx);
    return scan_locked(ctx);
}

static int get_profiles_locked(mcg_ctx *ctx, int64_t first, int64_t n, uint32_t *counts,
                               double *freqs) {
    if (first < 0 || n < 0 || first + n > ctx->n_prof)
        return mcg_fail(ctx, MCG_ERR_ARG, "profile range [%lld,%lld) outside 0..%lld",
                        (long long)first, (long long)(first + n), (long long)ctx->n_prof);
    TimerScope t(ctx, T_D2H);
    if (counts) {
        if (!ctx->counts.p || ctx->n_contigs != ctx->n_prof)
            return mcg_fail(ctx, MCG_ERR_STATE, "no scan counts resident");
        MCG_TRY(d2h(ctx, counts, ctx->counts.as<uint32_t>() + first * MCG_NPROF,
                    sizeof(uint32_t) * MCG_NPROF * n));
    }
    if (freqs)
        MCG_TRY(d2h(ctx, freqs, ctx->prof.as<double>() + first * MCG_NPROF,
                    sizeof(double) * MCG_NPROF * n));
    return sync(ctx);
}

extern "C" int mcg_get_profiles(mcg_ctx *ctx, int64_t first, int64_t n, uint32_t *counts,
                                double *freqs) {
    MCG_ENTER(ctx);
    return get_profiles_locked(ctx, first, n, counts, freqs);
}

extern "C" int mcg_tetramer_scan(mcg_ctx *ctx, const uint8_t *bases, const int64_t *offsets,
                                 int64_t n, int encoding, uint32_t *counts, double *freqs) {
    MCG_ENTER(ctx);
    MCG_TRY(load_contigs_locked(ctx, bases, offsets, n, encoding));
    MCG_TRY(scan_locked(ctx));
    return get_profiles_locked(ctx, 0, n, counts, freqs);
}

extern "C" int mcg_set_profiles(mcg_ctx *ctx, const double *prof, int64_t n) {
    MCG_ENTER(ctx);
    if (n < 0 || (n > 0 && !prof)) return mcg_fail(ctx, MCG_ERR_ARG, "bad profile arguments");
    MCG_TRY(mcg_reserve(ctx, ctx->prof, sizeof(double) * MCG_NPROF * std::max<int64_t>(n, 1)));
    MCG_TRY(h2d(ctx, ctx->prof.p, prof, sizeof(double) * MCG_NPROF * n));
    MCG_TRY(mcg_reserve(ctx, ctx->prof_norm, sizeof(double) * std::max<int64_t>(n, 1)));
    MCG_TRY(mcg_k_row_norms(ctx, ctx->prof.as<double>(), n, ctx->prof_norm.as<double>()));
    MCG_TRY(sync(ctx));
    ctx->n_prof = n;
    return MCG_OK;
}

// ---- coverage -----------------------------------------------------------------------------------------
// room for [n, S] coverages and their terms; `upload` = copy them now, on the compute stream
static int load_coverage_locked(mcg_ctx *ctx, const double *cov, int64_t n, int S,
                                bool upload = true) {
    if (n < 0 || S <= 0 || (n > 0 && !cov)) return mcg_fail(ctx, MCG_ERR_ARG, "bad coverage arguments");
    const size_t bytes = sizeof(double) * static_cast<size_t>(std::max<int64_t>(n, 1)) * S;
    MCG_TRY(mcg_reserve(ctx, ctx->cov, bytes));
    MCG_TRY(mcg_reserve(ctx, ctx->logc, bytes));
    MCG_TRY(mcg_reserve(ctx, ctx->lgam, bytes));
    if (upload) {
        TimerScope t(ctx, T_H2D);
        MCG_TRY(h2d(ctx, ctx->cov.p, cov, sizeof(double) * n * S));
    }
    ctx->n_cov = n;
    ctx->n_samples = S;
    return MCG_OK;
}

// Coverage rows [c0, c1) of a wave: copy and coverage terms on the side stream, so that they run
// under the kernels of the waves in front; the compute stream waits for them where the wave's
// scoring starts.  (The scans need no coverage: with the table sent up-front the first scan
// waited for S x 8 bytes per contig -- 4 GB at 5 M contigs x 100 samples -- on the same link.)
static int wave_coverage(mcg_ctx *ctx, const double *cov, int64_t c0, int64_t c1) {
    const int S = ctx->n_samples;
    const size_t off = static_cast<size_t>(c0) * S;
    const int64_t count = (c1 - c0) * S;
    MCG_CUDA(ctx, cudaMemcpyAsync(ctx->cov.as<double>() + off, cov + off, sizeof(double) * count,
                                  cudaMemcpyHostToDevice, ctx->cov_stream));
    cudaStream_t keep = ctx->stream;
    ctx->stream = ctx->cov_stream;
    const int rc = mcg_k_cov_terms(ctx, ctx->cov.as<double>() + off, ctx->logc.as<double>() + off,
                                   ctx->lgam.as<double>() + off, count);
    ctx->stream = keep;
    MCG_TRY(rc);
    MCG_CUDA(ctx, cudaEventRecord(ctx->cov_ev, ctx->cov_stream));
    MCG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->cov_ev, 0));
    return MCG_OK;
}

static int cov_terms_locked(mcg_ctx *ctx) {
    TimerScope t(ctx, T_COV);
    return mcg_k_cov_terms(ctx, ctx->cov.as<double>(), ctx->logc.as<double>(),
                           ctx->lgam.as<double>(), ctx->n_cov * ctx->n_samples);
}

extern "C" int mcg_load_coverage(mcg_ctx *ctx, const double *cov, int64_t n, int n_samples) {
    MCG_ENTER(ctx);
    MCG_TRY(load_coverage_locked(ctx, cov, n, n_samples));
    MCG_TRY(cov_terms_locked(ctx));
    return mcg_check_status(ctx);
}

extern "C" int mcg_get_coverage_terms(mcg_ctx *ctx, double *cov, double *logc, double *lgam) {
    MCG_ENTER(ctx);
    const size_t bytes = sizeof(double) * ctx->n_cov * ctx->n_samples;
    if (cov) MCG_TRY(d2h(ctx, cov, ctx->cov.p, bytes));
    if (logc) MCG_TRY(d2h(ctx, logc, ctx->logc.p, bytes));
    if (lgam) MCG_TRY(d2h(ctx, lgam, ctx->lgam.p, bytes));
    return sync(ctx);
}

// ---- bin tables ------------------------------------------------------------------------------------------
static int need_features(mcg_ctx *ctx) {
    if (ctx->n_prof == 0 || !ctx->prof.p)
        return mcg_fail(ctx, MCG_ERR_STATE, "no profiles resident (mcg_scan / mcg_set_profiles)");
    if (ctx->n_cov == 0 || !ctx->cov.p)
        return mcg_fail(ctx, MCG_ERR_STATE, "no coverage resident (mcg_load_coverage)");
    return MCG_OK;
}

static int check_ids(mcg_ctx *ctx, const int64_t *ids, int64_t n, int64_t limit, const char *what) {
    for (int64_t i = 0; i < n; ++i)
        if (ids[i] < 0 || ids[i] >= limit)
            return mcg_fail(ctx, MCG_ERR_ARG, "%s[%lld] = %lld outside 0..%lld", what, (long long)i,
                            (long long)ids[i], (long long)limit);
    return MCG_OK;
}

static int set_centroids_locked(mcg_ctx *ctx, const double *cen_prof, const double *cen_cov,
                                int64_t B, int S) {
    if (B <= 0 || !cen_prof || !cen_cov) return mcg_fail(ctx, MCG_ERR_ARG, "bad centroid arguments");
    MCG_TRY(mcg_reserve(ctx, ctx->cen_prof, sizeof(double) * MCG_NPROF * B));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_cov, sizeof(double) * S * B));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_logc, sizeof(double) * S * B));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_lgam, sizeof(double) * S * B));
    {
        TimerScope t(ctx, T_H2D);
        MCG_TRY(h2d(ctx, ctx->cen_prof.p, cen_prof, sizeof(double) * MCG_NPROF * B));
        MCG_TRY(h2d(ctx, ctx->cen_cov.p, cen_cov, sizeof(double) * S * B));
    }
    // centroid coverages are means of clamped values, so the clamp in the kernel is inert
    MCG_TRY(mcg_k_cov_terms(ctx, ctx->cen_cov.as<double>(), ctx->cen_logc.as<double>(),
                            ctx->cen_lgam.as<double>(), B * S));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_norm, sizeof(double) * B));
    MCG_TRY(mcg_k_row_norms(ctx, ctx->cen_prof.as<double>(), B, ctx->cen_norm.as<double>()));
    ctx->n_bins = B;
    ctx->cen_samples = S;
    return MCG_OK;
}

extern "C" int mcg_set_centroids(mcg_ctx *ctx, const double *cen_prof, const double *cen_cov,
                                 int64_t n_bins) {
    MCG_ENTER(ctx);
    if (ctx->n_samples <= 0) return mcg_fail(ctx, MCG_ERR_STATE, "load coverage before centroids");
    MCG_TRY(set_centroids_locked(ctx, cen_prof, cen_cov, n_bins, ctx->n_samples));
    return sync(ctx);
}

extern "C" int mcg_set_centroids_dev(mcg_ctx *ctx, const void *d_cen_prof, const void *d_cen_cov,
                                     int64_t B) {
    MCG_ENTER(ctx);
    const int S = ctx->n_samples;
    if (S <= 0) return mcg_fail(ctx, MCG_ERR_STATE, "load coverage before centroids");
    if (B <= 0 || !d_cen_prof || !d_cen_cov) return mcg_fail(ctx, MCG_ERR_ARG, "bad centroid arguments");
    MCG_TRY(mcg_reserve(ctx, ctx->cen_prof, sizeof(double) * MCG_NPROF * B));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_cov, sizeof(double) * S * B));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_logc, sizeof(double) * S * B));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_lgam, sizeof(double) * S * B));
    MCG_CUDA(ctx, cudaMemcpyAsync(ctx->cen_prof.p, d_cen_prof, sizeof(double) * MCG_NPROF * B,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    MCG_CUDA(ctx, cudaMemcpyAsync(ctx->cen_cov.p, d_cen_cov, sizeof(double) * S * B,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    MCG_TRY(mcg_k_cov_terms(ctx, ctx->cen_cov.as<double>(), ctx->cen_logc.as<double>(),
                            ctx->cen_lgam.as<double>(), B * S));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_norm, sizeof(double) * B));
    MCG_TRY(mcg_k_row_norms(ctx, ctx->cen_prof.as<double>(), B, ctx->cen_norm.as<double>()));
    ctx->n_bins = B;
    ctx->cen_samples = S;
    return MCG_OK;
}

extern "C" int mcg_bin_profiles(mcg_ctx *ctx, const int64_t *members, const int64_t *seg_offsets,
                                int64_t nseg, double *cen_prof, double *cen_cov) {
    MCG_ENTER(ctx);
    MCG_TRY(need_features(ctx));
    if (nseg <= 0 || !seg_offsets) return mcg_fail(ctx, MCG_ERR_ARG, "bad segment arguments");
    const int64_t nm = seg_offsets[nseg];
    if (seg_offsets[0] != 0) return mcg_fail(ctx, MCG_ERR_ARG, "seg_offsets[0] must be 0");
    for (int64_t b = 0; b < nseg; ++b)
        if (seg_offsets[b + 1] <= seg_offsets[b])
            return mcg_fail(ctx, MCG_ERR_ARG, "bin %lld has no members", (long long)b);
    MCG_TRY(check_ids(ctx, members, nm, std::min(ctx->n_prof, ctx->n_cov), "members"));
    const int S = ctx->n_samples;
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], sizeof(int64_t) * nm));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[1], sizeof(int64_t) * (nseg + 1)));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_prof, sizeof(double) * MCG_NPROF * nseg));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_cov, sizeof(double) * S * nseg));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_logc, sizeof(double) * S * nseg));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_lgam, sizeof(double) * S * nseg));
    MCG_TRY(h2d(ctx, ctx->scratch[0].p, members, sizeof(int64_t) * nm));
    MCG_TRY(h2d(ctx, ctx->scratch[1].p, seg_offsets, sizeof(int64_t) * (nseg + 1)));
    MCG_TRY(mcg_k_segment_mean(ctx, ctx->prof.as<double>(), MCG_NPROF,
                               ctx->scratch[0].as<int64_t>(), ctx->scratch[1].as<int64_t>(), nseg,
                               ctx->cen_prof.as<double>()));
    MCG_TRY(mcg_k_segment_mean(ctx, ctx->cov.as<double>(), S, ctx->scratch[0].as<int64_t>(),
                               ctx->scratch[1].as<int64_t>(), nseg, ctx->cen_cov.as<double>()));
    if (cen_prof) MCG_TRY(d2h(ctx, cen_prof, ctx->cen_prof.p, sizeof(double) * MCG_NPROF * nseg));
    if (cen_cov) MCG_TRY(d2h(ctx, cen_cov, ctx->cen_cov.p, sizeof(double) * S * nseg));
    MCG_TRY(mcg_k_cov_terms(ctx, ctx->cen_cov.as<double>(), ctx->cen_logc.as<double>(),
                            ctx->cen_lgam.as<double>(), nseg * S));
    MCG_TRY(mcg_reserve(ctx, ctx->cen_norm, sizeof(double) * nseg));
    MCG_TRY(mcg_k_row_norms(ctx, ctx->cen_prof.as<double>(), nseg, ctx->cen_norm.as<double>()));
    ctx->n_bins = nseg;
    ctx->cen_samples = S;
    return sync(ctx);
}

// ---- elementwise primitives ----------------------------------------------------------------------------------
extern "C" int mcg_normpdf(mcg_ctx *ctx, const double *x, double mean, double sd, double *out,
                           int64_t n) {
    MCG_ENTER(ctx);
    if (n < 0 || (n > 0 && (!x || !out))) return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], sizeof(double) * n));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[1], sizeof(double) * n));
    MCG_TRY(h2d(ctx, ctx->scratch[0].p, x, sizeof(double) * n));
    MCG_TRY(mcg_k_normpdf(ctx, ctx->scratch[0].as<double>(), mean, sd,
                          ctx->scratch[1].as<double>(), n));
    MCG_TRY(d2h(ctx, out, ctx->scratch[1].p, sizeof(double) * n));
    return sync(ctx);
}

extern "C" int mcg_tetramer_distance(mcg_ctx *ctx, const double *u, const double *v, int64_t n,
                                     int width, double *out) {
    MCG_ENTER(ctx);
    if (n < 0 || width <= 0 || (n > 0 && (!u || !v || !out)))
        return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    const size_t bytes = sizeof(double) * n * width;
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], bytes));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[1], bytes));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[2], sizeof(double) * n));
    MCG_TRY(h2d(ctx, ctx->scratch[0].p, u, bytes));
    MCG_TRY(h2d(ctx, ctx->scratch[1].p, v, bytes));
    MCG_TRY(mcg_k_distance(ctx, ctx->scratch[0].as<double>(), ctx->scratch[1].as<double>(), n,
                           width, ctx->scratch[2].as<double>()));
    MCG_TRY(d2h(ctx, out, ctx->scratch[2].p, sizeof(double) * n));
    return sync(ctx);
}

extern "C" int mcg_comp_probability(mcg_ctx *ctx, const double *dist, double *out, int64_t n) {
    MCG_ENTER(ctx);
    if (n < 0 || (n > 0 && (!dist || !out))) return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], sizeof(double) * n));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[1], sizeof(double) * n));
    MCG_TRY(h2d(ctx, ctx->scratch[0].p, dist, sizeof(double) * n));
    MCG_TRY(mcg_k_comp_prob(ctx, ctx->scratch[0].as<double>(), ctx->scratch[1].as<double>(), n,
                            ctx->status.as<int32_t>()));
    MCG_TRY(d2h(ctx, out, ctx->scratch[1].p, sizeof(double) * n));
    return mcg_check_status(ctx);
}

extern "C" int mcg_cov_probability(mcg_ctx *ctx, const double *cov1, const double *cov2, int64_t n,
                                   int n_samples, double *out) {
    MCG_ENTER(ctx);
    if (n < 0 || n_samples <= 0 || (n > 0 && (!cov1 || !cov2 || !out)))
        return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    const size_t bytes = sizeof(double) * n * n_samples;
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], bytes));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[1], bytes));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[2], sizeof(double) * n));
    MCG_TRY(h2d(ctx, ctx->scratch[0].p, cov1, bytes));
    MCG_TRY(h2d(ctx, ctx->scratch[1].p, cov2, bytes));
    MCG_TRY(mcg_k_cov_prob(ctx, ctx->scratch[0].as<double>(), ctx->scratch[1].as<double>(), n,
                           n_samples, ctx->scratch[2].as<double>()));
    MCG_TRY(d2h(ctx, out, ctx->scratch[2].p, sizeof(double) * n));
    return sync(ctx);
}

// ---- scoring -----------------------------------------------------------------------------------------------------
static ScoreSide contig_side(mcg_ctx *ctx, const int64_t *d_index, int64_t n) {
    ScoreSide s;
    s.prof = ctx->prof.as<double>();
    s.cov = ctx->cov.as<double>();
    s.logc = ctx->logc.as<double>();
    s.lgam = ctx->lgam.as<double>();
    s.norm = ctx->prof_norm.as<double>();
    s.index = d_index;
    s.first = 0;
    s.n = n;
    return s;
}

static ScoreSide centroid_side(mcg_ctx *ctx) {
    ScoreSide s;
    s.prof = ctx->cen_prof.as<double>();
    s.cov = ctx->cen_cov.as<double>();
    s.logc = ctx->cen_logc.as<double>();
    s.lgam = ctx->cen_lgam.as<double>();
    s.norm = ctx->cen_norm.as<double>();
    s.index = nullptr;
    s.first = 0;
    s.n = ctx->n_bins;
    return s;
}

extern "C" int mcg_pair_scores(mcg_ctx *ctx, const int64_t *queries, int64_t nq,
                               const int64_t *targets, int64_t nt, double *dist, double *pc,
                               double *pv, double *lp) {
    MCG_ENTER(ctx);
    MCG_TRY(need_features(ctx));
    if (nq <= 0 || nt <= 0 || !queries || !targets) return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    const int64_t limit = std::min(ctx->n_prof, ctx->n_cov);
    MCG_TRY(check_ids(ctx, queries, nq, limit, "queries"));
    MCG_TRY(check_ids(ctx, targets, nt, limit, "targets"));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], sizeof(int64_t) * nq));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[1], sizeof(int64_t) * nt));
    MCG_TRY(h2d(ctx, ctx->scratch[0].p, queries, sizeof(int64_t) * nq));
    MCG_TRY(h2d(ctx, ctx->scratch[1].p, targets, sizeof(int64_t) * nt));
    const size_t mat = sizeof(double) * nq * nt;
    double *host[4] = {dist, pc, pv, lp};
    double *dev[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < 4; ++i)
        if (host[i]) {
            MCG_TRY(mcg_reserve(ctx, ctx->scratch[2 + i], mat));
            dev[i] = ctx->scratch[2 + i].as<double>();
        }
    ScoreOut out;
    out.dist = dev[0]; out.pc = dev[1]; out.pv = dev[2]; out.lp = dev[3];
    {
        TimerScope t(ctx, T_SCORE);
        MCG_TRY(mcg_k_pairs(ctx, contig_side(ctx, ctx->scratch[0].as<int64_t>(), nq),
                            contig_side(ctx, ctx->scratch[1].as<int64_t>(), nt), ctx->n_samples,
                            out, ctx->status.as<int32_t>()));
    }
    for (int i = 0; i < 4; ++i)
        if (host[i]) MCG_TRY(d2h(ctx, host[i], dev[i], mat));
    // a NaN prob_comp is reported in pc; the caller decides whether it is an error
    MCG_TRY(sync(ctx));
    cudaMemsetAsync(ctx->status.p, 0, 4 * sizeof(int32_t), ctx->stream);
    return MCG_OK;
}

static int score_centroids_locked(mcg_ctx *ctx, const int64_t *d_queries, int64_t nq,
                                  double *d_weights, int32_t *d_argmin, double *d_wmin) {
    ScoreOut out;
    out.weights = d_weights;
    out.argmin = d_argmin;
    out.wmin = d_wmin;
    TimerScope t(ctx, T_SCORE);
    return mcg_k_pairs(ctx, contig_side(ctx, d_queries, nq), centroid_side(ctx), ctx->n_samples,
                       out, ctx->status.as<int32_t>());
}

extern "C" int mcg_score_centroids(mcg_ctx *ctx, const int64_t *queries, int64_t nq,
                                   const double *cen_prof, const double *cen_cov, int64_t n_bins,
                                   double *weights, int32_t *argmin, double *wmin) {
    MCG_ENTER(ctx);
    MCG_TRY(need_features(ctx));
    if (nq < 0) return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    const int64_t limit = std::min(ctx->n_prof, ctx->n_cov);
    if (cen_prof || cen_cov) {
        MCG_TRY(set_centroids_locked(ctx, cen_prof, cen_cov, n_bins, ctx->n_samples));
    } else if (ctx->n_bins == 0) {
        return mcg_fail(ctx, MCG_ERR_STATE, "no centroids resident");
    }
    if (nq == 0) return MCG_OK;
    const int64_t *d_queries = nullptr;
    if (queries) {
        MCG_TRY(check_ids(ctx, queries, nq, limit, "queries"));
        MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], sizeof(int64_t) * nq));
        MCG_TRY(h2d(ctx, ctx->scratch[0].p, queries, sizeof(int64_t) * nq));
        d_queries = ctx->scratch[0].as<int64_t>();
    } else if (nq > limit) {
        return mcg_fail(ctx, MCG_ERR_ARG, "nq exceeds the resident contig table");
    }
    const int64_t B = ctx->n_bins;
    MCG_TRY(mcg_reserve(ctx, ctx->argmin, sizeof(int32_t) * nq));
    MCG_TRY(mcg_reserve(ctx, ctx->wmin, sizeof(double) * nq));
    double *d_weights = nullptr;
    if (weights) {
        MCG_TRY(mcg_reserve(ctx, ctx->scratch[2], sizeof(double) * nq * B));
        d_weights = ctx->scratch[2].as<double>();
    }
    MCG_TRY(score_centroids_locked(ctx, d_queries, nq, d_weights, ctx->argmin.as<int32_t>(),
                                   ctx->wmin.as<double>()));
    {
        TimerScope t(ctx, T_D2H);
        if (weights) MCG_TRY(d2h(ctx, weights, d_weights, sizeof(double) * nq * B));
        if (argmin) MCG_TRY(d2h(ctx, argmin, ctx->argmin.p, sizeof(int32_t) * nq));
        if (wmin) MCG_TRY(d2h(ctx, wmin, ctx->wmin.p, sizeof(double) * nq));
    }
    return mcg_check_status(ctx);
}

extern "C" int mcg_score_members(mcg_ctx *ctx, const int64_t *queries, int64_t nq,
                                 const int64_t *members, const int64_t *seg_offsets, int64_t nseg,
                                 int rule, double *weights) {
    MCG_ENTER(ctx);
    MCG_TRY(need_features(ctx));
    if (nq < 0 || nseg < 0 || !seg_offsets || !weights || (rule != MCG_RULE_A && rule != MCG_RULE_B))
        return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    if (nq == 0 || nseg == 0) return MCG_OK;
    const int64_t nm = seg_offsets[nseg];
    const int64_t limit = std::min(ctx->n_prof, ctx->n_cov);
    MCG_TRY(check_ids(ctx, queries, nq, limit, "queries"));
    MCG_TRY(check_ids(ctx, members, nm, limit, "members"));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], sizeof(int64_t) * nq));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[1], sizeof(int64_t) * std::max<int64_t>(nm, 1)));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[2], sizeof(int64_t) * (nseg + 1)));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[3], sizeof(double) * nq * std::max<int64_t>(nm, 1)));
    ctx->redo_valid = false;
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[4], sizeof(double) * nq * nseg));
    MCG_TRY(h2d(ctx, ctx->scratch[0].p, queries, sizeof(int64_t) * nq));
    MCG_TRY(h2d(ctx, ctx->scratch[1].p, members, sizeof(int64_t) * nm));
    MCG_TRY(h2d(ctx, ctx->scratch[2].p, seg_offsets, sizeof(int64_t) * (nseg + 1)));
    {
        TimerScope t(ctx, T_SCORE);
        if (nm > 0) {
            ScoreOut out;
            out.lp = ctx->scratch[3].as<double>();
            MCG_TRY(mcg_k_pairs(ctx, contig_side(ctx, ctx->scratch[0].as<int64_t>(), nq),
                                contig_side(ctx, ctx->scratch[1].as<int64_t>(), nm),
                                ctx->n_samples, out, ctx->status.as<int32_t>()));
        }
        MCG_TRY(mcg_k_aggregate_members(ctx, ctx->scratch[3].as<double>(), nq, nm,
                                        ctx->scratch[2].as<int64_t>(), nseg, rule,
                                        ctx->scratch[4].as<double>()));
    }
    MCG_TRY(d2h(ctx, weights, ctx->scratch[4].p, sizeof(double) * nq * nseg));
    return mcg_check_status(ctx);
}

// ---- frontier ---------------------------------------------------------------------------------------------------------
extern "C" int mcg_bfs_candidates(mcg_ctx *ctx, const int64_t *indptr, const int32_t *indices,
                                  int64_t n_vertices, const int32_t *labels, const int64_t *starts,
                                  int64_t n_starts, int depth_limit, int max_hits_per_node,
                                  int32_t *hit_node, int32_t *hit_bin, int32_t *hit_depth,
                                  int32_t *n_hits) {
    MCG_ENTER(ctx);
    if (n_vertices <= 0 || n_starts < 0 || !indptr || !labels || max_hits_per_node <= 0)
        return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    if (n_starts == 0) return MCG_OK;
    MCG_TRY(check_ids(ctx, starts, n_starts, n_vertices, "starts"));
    const int64_t n_edges = indptr[n_vertices];
    DevBuf d_indptr, d_indices, d_labels, d_starts, d_hits, d_nhits, d_work;
    int rc = MCG_OK;
    int64_t work_stride = 1024;
    auto cleanup = [&]() {
        for (DevBuf *b : {&d_indptr, &d_indices, &d_labels, &d_starts, &d_hits, &d_nhits, &d_work})
            if (b->p) cudaFree(b->p);
    };
#define BFS_TRY(expr) do { rc = (expr); if (rc != MCG_OK) { cleanup(); return rc; } } while (0)
    BFS_TRY(mcg_reserve(ctx, d_indptr, sizeof(int64_t) * (n_vertices + 1)));
    BFS_TRY(mcg_reserve(ctx, d_indices, sizeof(int32_t) * std::max<int64_t>(n_edges, 1)));
    BFS_TRY(mcg_reserve(ctx, d_labels, sizeof(int32_t) * n_vertices));
    BFS_TRY(mcg_reserve(ctx, d_starts, sizeof(int64_t) * n_starts));
    BFS_TRY(mcg_reserve(ctx, d_hits, sizeof(int32_t) * 3 * n_starts * max_hits_per_node));
    BFS_TRY(mcg_reserve(ctx, d_nhits, sizeof(int32_t) * n_starts));
    BFS_TRY(h2d(ctx, d_indptr.p, indptr, sizeof(int64_t) * (n_vertices + 1)));
    BFS_TRY(h2d(ctx, d_indices.p, indices, sizeof(int32_t) * n_edges));
    BFS_TRY(h2d(ctx, d_labels.p, labels, sizeof(int32_t) * n_vertices));
    BFS_TRY(h2d(ctx, d_starts.p, starts, sizeof(int64_t) * n_starts));
    int32_t *hn = d_hits.as<int32_t>();
    int32_t *hb = hn + n_starts * max_hits_per_node;
    int32_t *hd = hb + n_starts * max_hits_per_node;
    std::vector<int32_t> counts(static_cast<size_t>(n_starts));
    for (;;) {
        // the per-start slice grows until no walk overflows its queue / table
        if (d_work.p) { cudaFree(d_work.p); d_work = DevBuf(); }
        BFS_TRY(mcg_reserve(ctx, d_work, sizeof(int32_t) * work_stride * n_starts));
        BFS_TRY(mcg_k_bfs(ctx, d_indptr.as<int64_t>(), d_indices.as<int32_t>(), n_vertices,
                          d_labels.as<int32_t>(), d_starts.as<int64_t>(), n_starts, depth_limit,
                          max_hits_per_node, hn, hb, hd, d_nhits.as<int32_t>(),
                          d_work.as<int32_t>(), work_stride));
        BFS_TRY(d2h(ctx, counts.data(), d_nhits.p, sizeof(int32_t) * n_starts));
        BFS_TRY(sync(ctx));
        bool overflow = false;
        for (int64_t i = 0; i < n_starts; ++i) overflow |= counts[i] == -2;
        if (!overflow) break;
        work_stride *= 4;
        if (sizeof(int32_t) * work_stride * n_starts > (size_t(8) << 30)) {
            cleanup();
            return mcg_fail(ctx, MCG_ERR_NOMEM, "BFS work buffer would exceed 8 GiB");
        }
    }
    const size_t hb_bytes = sizeof(int32_t) * n_starts * max_hits_per_node;
    if (hit_node) BFS_TRY(d2h(ctx, hit_node, hn, hb_bytes));
    if (hit_bin) BFS_TRY(d2h(ctx, hit_bin, hb, hb_bytes));
    if (hit_depth) BFS_TRY(d2h(ctx, hit_depth, hd, hb_bytes));
    BFS_TRY(sync(ctx));
    if (n_hits) memcpy(n_hits, counts.data(), sizeof(int32_t) * n_starts);
    cleanup();
#undef BFS_TRY
    return MCG_OK;
}

// ---- resident assembly graph: label propagation at scale ------------------------------------------
extern "C" int mcg_graph_load(mcg_ctx *ctx, const int64_t *indptr, const int32_t *indices,
                              int64_t n_vertices) {
    MCG_ENTER(ctx);
    if (n_vertices <= 0 || !indptr) return mcg_fail(ctx, MCG_ERR_ARG, "bad graph arguments");
    if (n_vertices > 0x7ffffff0LL) return mcg_fail(ctx, MCG_ERR_ARG, "too many vertices");
    const int64_t n_edges = indptr[n_vertices];
    if (indptr[0] != 0 || n_edges < 0 || (n_edges > 0 && !indices))
        return mcg_fail(ctx, MCG_ERR_ARG, "bad CSR arrays");
    for (int64_t v = 0; v < n_vertices; ++v)
        if (indptr[v + 1] < indptr[v]) return mcg_fail(ctx, MCG_ERR_ARG, "indptr must not decrease");
    for (int64_t e = 0; e < n_edges; ++e)
        if (indices[e] < 0 || indices[e] >= n_vertices)
            return mcg_fail(ctx, MCG_ERR_ARG, "edge endpoint %d out of range", indices[e]);
    MCG_TRY(mcg_reserve(ctx, ctx->g_indptr, sizeof(int64_t) * (n_vertices + 1)));
    MCG_TRY(mcg_reserve(ctx, ctx->g_indices, sizeof(int32_t) * std::max<int64_t>(n_edges, 1)));
    MCG_TRY(mcg_reserve(ctx, ctx->g_labels, sizeof(int32_t) * n_vertices));
    MCG_TRY(h2d(ctx, ctx->g_indptr.p, indptr, sizeof(int64_t) * (n_vertices + 1)));
    MCG_TRY(h2d(ctx, ctx->g_indices.p, indices, sizeof(int32_t) * n_edges));
    MCG_CUDA(ctx, cudaMemsetAsync(ctx->g_labels.p, 0xff, sizeof(int32_t) * n_vertices, ctx->stream));
    MCG_TRY(sync(ctx));
    ctx->g_vertices = n_vertices;
    ctx->g_edges = n_edges;
    return MCG_OK;
}

extern "C" int mcg_labels_load(mcg_ctx *ctx, const int32_t *labels, int64_t n_vertices) {
    MCG_ENTER(ctx);
    if (ctx->g_vertices == 0) return mcg_fail(ctx, MCG_ERR_STATE, "no graph resident (mcg_graph_load)");
    if (!labels || n_vertices != ctx->g_vertices)
        return mcg_fail(ctx, MCG_ERR_ARG, "labels must cover the %lld vertices of the resident graph",
                        (long long)ctx->g_vertices);
    MCG_TRY(h2d(ctx, ctx->g_labels.p, labels, sizeof(int32_t) * n_vertices));
    return sync(ctx);
}

struct LabelPatch {
    long long id[16];
    int bin[16];
    int n;
};
static __global__ void labels_patch_kernel(int32_t *labels, LabelPatch p) {
    if (threadIdx.x < p.n) labels[p.id[threadIdx.x]] = p.bin[threadIdx.x];
}

extern "C" int mcg_labels_set(mcg_ctx *ctx, const int64_t *ids, const int32_t *bins, int64_t n) {
    MCG_ENTER(ctx);
    if (ctx->g_vertices == 0) return mcg_fail(ctx, MCG_ERR_STATE, "no graph resident (mcg_graph_load)");
    if (n < 0 || (n > 0 && (!ids || !bins))) return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    MCG_TRY(check_ids(ctx, ids, n, ctx->g_vertices, "ids"));
    if (n == 0) return MCG_OK;
    if (n <= 16) {
        // the usual call (one accepted contig): the patch travels in the kernel arguments
        LabelPatch p;
        p.n = static_cast<int>(n);
        for (int i = 0; i < p.n; ++i) { p.id[i] = ids[i]; p.bin[i] = bins[i]; }
        labels_patch_kernel<<<1, 32, 0, ctx->stream>>>(ctx->g_labels.as<int32_t>(), p);
        MCG_KERNEL_CHECK(ctx);
        return MCG_OK;
    }
    MCG_TRY(mcg_reserve(ctx, ctx->g_patch, 12 * static_cast<size_t>(n)));
    int64_t *d_ids = ctx->g_patch.as<int64_t>();
    int32_t *d_bins = reinterpret_cast<int32_t *>(d_ids + n);
    MCG_TRY(h2d(ctx, d_ids, ids, sizeof(int64_t) * n));
    MCG_TRY(h2d(ctx, d_bins, bins, sizeof(int32_t) * n));
    MCG_TRY(mcg_k_labels_set(ctx, ctx->g_labels.as<int32_t>(), d_ids, d_bins, n));
    return sync(ctx);   // ids / bins are the caller's again
}

namespace {
struct WalkHead { long long total; int n_redo, n_failed; };

size_t walk_head_bytes(int64_t n_starts) {
    const int64_t n8 = (n_starts + 1) / 2 * 2;
    return 16 + 8 * static_cast<size_t>(n_starts) + 8 * static_cast<size_t>(n8);
}

// The walks that did not fit their shared-memory table again, over tables in global memory.
int walk_redo(mcg_ctx *ctx, int64_t n_starts, int depth_limit, int64_t hit_cap, WalkHead *head) {
    char *blk = ctx->g_block.as<char>();
    for (int64_t stride = 4096;; stride *= 4) {
        const size_t bytes = sizeof(int32_t) * static_cast<size_t>(stride) * head->n_redo;
        if (bytes > (size_t(8) << 30))
            return mcg_fail(ctx, MCG_ERR_NOMEM, "BFS work buffer would exceed 8 GiB");
        MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], bytes));
        MCG_TRY(mcg_k_walks_redo(ctx, ctx->g_indptr.as<int64_t>(), ctx->g_indices.as<int32_t>(),
                                 ctx->g_labels.as<int32_t>(), ctx->g_starts.as<int64_t>(),
                                 n_starts, head->n_redo, depth_limit,
                                 ctx->scratch[0].as<int32_t>(), stride, blk, hit_cap));
        WalkHead again;
        MCG_TRY(d2h(ctx, &again, blk, sizeof again));
        MCG_TRY(sync(ctx));
        head->total = again.total;
        if (again.n_failed == 0) return MCG_OK;
    }
}

int walk_check(mcg_ctx *ctx, const int64_t *starts, int64_t n_starts, const int64_t *hit_start,
               const int32_t *hit_count, const int32_t *hits, int64_t hit_cap, const int64_t *total) {
    if (ctx->g_vertices == 0) return mcg_fail(ctx, MCG_ERR_STATE, "no graph resident (mcg_graph_load)");
    if (n_starts < 0 || hit_cap < 0 || !total || (n_starts > 0 && (!starts || !hit_start || !hit_count)) ||
        (hit_cap > 0 && !hits))
        return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    if (n_starts > 0x7fffffffLL) return mcg_fail(ctx, MCG_ERR_ARG, "too many start nodes");
    return MCG_OK;
}

// starts up, walks (and the redo passes) done; ranges and hits stay in ctx->g_block
int walk_run(mcg_ctx *ctx, const int64_t *starts, int64_t n_starts, int depth_limit,
             int64_t hit_cap, WalkHead *head, bool read_head) {
    MCG_TRY(check_ids(ctx, starts, n_starts, ctx->g_vertices, "starts"));
    MCG_TRY(mcg_reserve(ctx, ctx->g_block, walk_head_bytes(n_starts) + 12 * static_cast<size_t>(hit_cap)));
    MCG_TRY(mcg_reserve(ctx, ctx->g_starts, sizeof(int64_t) * n_starts));
    MCG_TRY(mcg_reserve(ctx, ctx->g_work, mcg_walk_work_bytes(n_starts)));
    MCG_TRY(h2d(ctx, ctx->g_starts.p, starts, sizeof(int64_t) * n_starts));
    {
        TimerScope t(ctx, T_OTHER);
        MCG_TRY(mcg_k_walks(ctx, ctx->g_indptr.as<int64_t>(), ctx->g_indices.as<int32_t>(),
                            ctx->g_labels.as<int32_t>(), ctx->g_starts.as<int64_t>(), n_starts,
                            depth_limit, ctx->g_work.as<int32_t>(), ctx->g_block.as<char>(), hit_cap));
    }
    if (!read_head) return MCG_OK;
    MCG_TRY(d2h(ctx, head, ctx->g_block.p, sizeof *head));
    MCG_TRY(sync(ctx));
    if (head->n_redo > 0) MCG_TRY(walk_redo(ctx, n_starts, depth_limit, hit_cap, head));
    return MCG_OK;
}

int walk_fetch(mcg_ctx *ctx, int64_t n_starts, int64_t *hit_start, int32_t *hit_count,
               int32_t *hits, int64_t total) {
    TimerScope t(ctx, T_D2H);
    const char *blk = ctx->g_block.as<char>();
    MCG_TRY(d2h(ctx, hit_start, blk + 16, sizeof(int64_t) * n_starts));
    MCG_TRY(d2h(ctx, hit_count, blk + 16 + 8 * n_starts, sizeof(int32_t) * n_starts));
    MCG_TRY(d2h(ctx, hits, blk + walk_head_bytes(n_starts), 12 * static_cast<size_t>(total)));
    return sync(ctx);
}
}  // namespace

// The two halves of mcg_bfs_resident for the device pool: every device walks its share of the
// start nodes, the totals give every share its place in the caller's hit list, and every device
// copies straight there (no staging vectors on the host, no second memcpy).
int mcg_bfs_run(mcg_ctx *ctx, const int64_t *starts, int64_t n_starts, int depth_limit,
                int64_t hit_cap, int64_t *total) {
    MCG_ENTER(ctx);
    *total = 0;
    if (n_starts == 0) return MCG_OK;
    if (ctx->g_vertices == 0) return mcg_fail(ctx, MCG_ERR_STATE, "no graph resident (mcg_graph_load)");
    if (n_starts > 0x7fffffffLL) return mcg_fail(ctx, MCG_ERR_ARG, "too many start nodes");
    WalkHead head;
    MCG_TRY(walk_run(ctx, starts, n_starts, depth_limit, hit_cap, &head, true));
    *total = head.total;
    return MCG_OK;
}

int mcg_bfs_fetch(mcg_ctx *ctx, int64_t n_starts, int64_t *hit_start, int32_t *hit_count,
                  int32_t *hits, int64_t total) {
    MCG_ENTER(ctx);
    if (n_starts == 0) return MCG_OK;
    return walk_fetch(ctx, n_starts, hit_start, hit_count, hits, total);
}

extern "C" int mcg_bfs_resident(mcg_ctx *ctx, const int64_t *starts, int64_t n_starts,
                                int depth_limit, int64_t *hit_start, int32_t *hit_count,
                                int32_t *hits, int64_t hit_cap, int64_t *total) {
    MCG_ENTER(ctx);
    MCG_TRY(walk_check(ctx, starts, n_starts, hit_start, hit_count, hits, hit_cap, total));
    *total = 0;
    if (n_starts == 0) return MCG_OK;
    WalkHead head;
    // Small calls (the few contigs around one accepted contig) come back in ONE copy and one
    // synchronisation: header, ranges and the first hits speculatively, through pinned memory.
    const bool small = n_starts <= 4096;
    if (!small) {
        MCG_TRY(walk_run(ctx, starts, n_starts, depth_limit, hit_cap, &head, true));
        *total = head.total;
        if (head.total > hit_cap) return MCG_OK;   // the caller comes back with a larger buffer
        return walk_fetch(ctx, n_starts, hit_start, hit_count, hits, head.total);
    }
    MCG_TRY(walk_run(ctx, starts, n_starts, depth_limit, hit_cap, &head, false));
    const size_t head_bytes = walk_head_bytes(n_starts);
    const int64_t spec_hits = std::min<int64_t>(hit_cap, 4 * n_starts + 256);
    const size_t spec_bytes = head_bytes + 12 * static_cast<size_t>(spec_hits);
    if (spec_bytes > ctx->g_pinned_cap) {
        if (ctx->g_pinned) cudaFreeHost(ctx->g_pinned);
        ctx->g_pinned = nullptr;
        ctx->g_pinned_cap = 0;
        const size_t want = std::max<size_t>(spec_bytes, 1 << 16);
        if (cudaMallocHost(&ctx->g_pinned, want) != cudaSuccess)
            return mcg_fail(ctx, MCG_ERR_NOMEM, "cudaMallocHost(%zu) failed", want);
        ctx->g_pinned_cap = want;
    }
    MCG_TRY(d2h(ctx, ctx->g_pinned, ctx->g_block.p, spec_bytes));
    MCG_TRY(sync(ctx));
    memcpy(&head, ctx->g_pinned, sizeof head);
    bool redone = false;
    if (head.n_redo > 0) {
        redone = true;
        MCG_TRY(walk_redo(ctx, n_starts, depth_limit, hit_cap, &head));
    }
    *total = head.total;
    if (head.total > hit_cap) return MCG_OK;   // the caller comes back with a larger buffer
    if (!redone && head.total <= spec_hits) {
        const char *src = static_cast<const char *>(ctx->g_pinned);
        memcpy(hit_start, src + 16, sizeof(int64_t) * n_starts);
        memcpy(hit_count, src + 16 + 8 * n_starts, sizeof(int32_t) * n_starts);
        if (head.total) memcpy(hits, src + head_bytes, 12 * static_cast<size_t>(head.total));
        return MCG_OK;
    }
    return walk_fetch(ctx, n_starts, hit_start, hit_count, hits, head.total);
}

// Rules A / B for listed (query, bin) items: label_prop_utils.py:54-86 scores a candidate only
// against the bins its walk reached.
extern "C" int mcg_score_member_items(mcg_ctx *ctx, const int64_t *item_query,
                                      const int32_t *item_bin, int64_t n_items,
                                      const int64_t *members, const int64_t *seg_offsets,
                                      int64_t nseg, int rule, double *weights) {
    MCG_ENTER(ctx);
    MCG_TRY(need_features(ctx));
    if (n_items < 0 || nseg <= 0 || !seg_offsets || (rule != MCG_RULE_A && rule != MCG_RULE_B) ||
        (n_items > 0 && (!item_query || !item_bin || !weights)))
        return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    if (n_items == 0) return MCG_OK;
    const int64_t nm = seg_offsets[nseg];
    const int64_t limit = std::min(ctx->n_prof, ctx->n_cov);
    if (seg_offsets[0] != 0) return mcg_fail(ctx, MCG_ERR_ARG, "seg_offsets[0] must be 0");
    for (int64_t b = 0; b < nseg; ++b)
        if (seg_offsets[b + 1] < seg_offsets[b])
            return mcg_fail(ctx, MCG_ERR_ARG, "seg_offsets must not decrease");
    MCG_TRY(check_ids(ctx, item_query, n_items, limit, "item_query"));
    MCG_TRY(check_ids(ctx, members, nm, limit, "members"));
    for (int64_t i = 0; i < n_items; ++i)
        if (item_bin[i] < 0 || item_bin[i] >= nseg)
            return mcg_fail(ctx, MCG_ERR_ARG, "item_bin[%lld] = %d outside 0..%lld", (long long)i,
                            item_bin[i], (long long)nseg);
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], sizeof(int64_t) * n_items));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[1], sizeof(int64_t) * std::max<int64_t>(nm, 1)));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[2], sizeof(int64_t) * (nseg + 1)));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[3], sizeof(int32_t) * n_items));
    ctx->cand_valid = false;
    ctx->redo_valid = false;
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[4], sizeof(double) * n_items));
    MCG_TRY(h2d(ctx, ctx->scratch[0].p, item_query, sizeof(int64_t) * n_items));
    MCG_TRY(h2d(ctx, ctx->scratch[1].p, members, sizeof(int64_t) * nm));
    MCG_TRY(h2d(ctx, ctx->scratch[2].p, seg_offsets, sizeof(int64_t) * (nseg + 1)));
    MCG_TRY(h2d(ctx, ctx->scratch[3].p, item_bin, sizeof(int32_t) * n_items));
    {
        TimerScope t(ctx, T_SCORE);
        MCG_TRY(mcg_k_member_items(ctx, contig_side(ctx, nullptr, 0), ctx->scratch[0].as<int64_t>(),
                                   ctx->scratch[3].as<int32_t>(), n_items,
                                   ctx->scratch[1].as<int64_t>(), ctx->scratch[2].as<int64_t>(),
                                   ctx->n_samples, rule, ctx->scratch[4].as<double>(),
                                   ctx->status.as<int32_t>()));
    }
    MCG_TRY(d2h(ctx, weights, ctx->scratch[4].p, sizeof(double) * n_items));
    return mcg_check_status(ctx);
}

extern "C" int mcg_path_lengths(mcg_ctx *ctx, const int64_t *indptr, const int32_t *indices,
                                int64_t n_vertices, const int64_t *sources, int64_t n_sources,
                                const int64_t *targets, int64_t n_targets, int32_t *out) {
    MCG_ENTER(ctx);
    if (n_vertices <= 0 || n_sources < 0 || n_targets < 0 || !indptr || !out)
        return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    if (n_sources == 0 || n_targets == 0) return MCG_OK;
    if (n_vertices > 0x7fffffffLL) return mcg_fail(ctx, MCG_ERR_ARG, "too many vertices");
    MCG_TRY(check_ids(ctx, sources, n_sources, n_vertices, "sources"));
    MCG_TRY(check_ids(ctx, targets, n_targets, n_vertices, "targets"));
    const int64_t n_edges = indptr[n_vertices];
    if (n_edges > 0 && !indices) return mcg_fail(ctx, MCG_ERR_ARG, "indices is NULL");
    // column of every distinct target vertex; duplicates in `targets` share a column
    std::vector<int32_t> tslot(static_cast<size_t>(n_vertices), -1);
    std::vector<int64_t> col_of(static_cast<size_t>(n_targets));
    int64_t n_cols = 0;
    for (int64_t j = 0; j < n_targets; ++j) {
        int32_t &slot = tslot[static_cast<size_t>(targets[j])];
        if (slot < 0) slot = static_cast<int32_t>(n_cols++);
        col_of[j] = slot;
    }
    DevBuf d_indptr, d_indices, d_tslot, d_sources, d_work, d_counts, d_out;
    int rc = MCG_OK;
    auto cleanup = [&]() {
        for (DevBuf *b : {&d_indptr, &d_indices, &d_tslot, &d_sources, &d_work, &d_counts, &d_out})
            if (b->p) cudaFree(b->p);
    };
#define PL_TRY(expr) do { rc = (expr); if (rc != MCG_OK) { cleanup(); return rc; } } while (0)
    PL_TRY(mcg_reserve(ctx, d_indptr, sizeof(int64_t) * (n_vertices + 1)));
    PL_TRY(mcg_reserve(ctx, d_indices, sizeof(int32_t) * std::max<int64_t>(n_edges, 1)));
    PL_TRY(mcg_reserve(ctx, d_tslot, sizeof(int32_t) * n_vertices));
    PL_TRY(mcg_reserve(ctx, d_sources, sizeof(int64_t) * n_sources));
    PL_TRY(mcg_reserve(ctx, d_work, sizeof(int32_t) * (4 * mcg_path_words(n_sources) + 3) * n_vertices));
    PL_TRY(mcg_reserve(ctx, d_counts, 2 * sizeof(int)));
    PL_TRY(mcg_reserve(ctx, d_out, sizeof(int32_t) * n_sources * n_cols));
    PL_TRY(h2d(ctx, d_indptr.p, indptr, sizeof(int64_t) * (n_vertices + 1)));
    if (n_edges) PL_TRY(h2d(ctx, d_indices.p, indices, sizeof(int32_t) * n_edges));
    PL_TRY(h2d(ctx, d_tslot.p, tslot.data(), sizeof(int32_t) * n_vertices));
    PL_TRY(h2d(ctx, d_sources.p, sources, sizeof(int64_t) * n_sources));
    {
        TimerScope t(ctx, T_OTHER);
        PL_TRY(mcg_k_path_lengths(ctx, d_indptr.as<int64_t>(), d_indices.as<int32_t>(), n_vertices,
                                  d_tslot.as<int32_t>(), d_sources.as<int64_t>(), n_sources, n_cols,
                                  d_work.as<int32_t>(), d_counts.as<int>(), d_out.as<int32_t>()));
    }
    std::vector<int32_t> dense(static_cast<size_t>(n_sources * n_cols));
    PL_TRY(d2h(ctx, dense.data(), d_out.p, sizeof(int32_t) * n_sources * n_cols));
    PL_TRY(sync(ctx));
    for (int64_t i = 0; i < n_sources; ++i)
        for (int64_t j = 0; j < n_targets; ++j)
            out[i * n_targets + j] = dense[static_cast<size_t>(i * n_cols + col_of[j])];
    cleanup();
#undef PL_TRY
    return MCG_OK;
}

extern "C" int mcg_connected_components(mcg_ctx *ctx, const int64_t *indptr, const int32_t *indices,
                                        int64_t n_vertices, int32_t *labels) {
    MCG_ENTER(ctx);
    if (n_vertices < 0 || !indptr || !labels) return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    if (n_vertices == 0) return MCG_OK;
    if (n_vertices > 0x7fffffffLL) return mcg_fail(ctx, MCG_ERR_ARG, "too many vertices");
    const int64_t n_edges = indptr[n_vertices];
    if (n_edges > 0 && !indices) return mcg_fail(ctx, MCG_ERR_ARG, "indices is NULL");
    for (int64_t e = 0; e < n_edges; ++e)
        if (indices[e] < 0 || indices[e] >= n_vertices)
            return mcg_fail(ctx, MCG_ERR_ARG, "edge endpoint %d out of range", indices[e]);
    DevBuf d_indptr, d_indices, d_work, d_flag, d_labels;
    int rc = MCG_OK;
    auto cleanup = [&]() {
        for (DevBuf *b : {&d_indptr, &d_indices, &d_work, &d_flag, &d_labels})
            if (b->p) cudaFree(b->p);
    };
#define CC_TRY(expr) do { rc = (expr); if (rc != MCG_OK) { cleanup(); return rc; } } while (0)
    CC_TRY(mcg_reserve(ctx, d_indptr, sizeof(int64_t) * (n_vertices + 1)));
    CC_TRY(mcg_reserve(ctx, d_indices, sizeof(int32_t) * std::max<int64_t>(n_edges, 1)));
    CC_TRY(mcg_reserve(ctx, d_work, sizeof(int32_t) * 2 * n_vertices));
    CC_TRY(mcg_reserve(ctx, d_flag, sizeof(int)));
    CC_TRY(mcg_reserve(ctx, d_labels, sizeof(int32_t) * n_vertices));
    CC_TRY(h2d(ctx, d_indptr.p, indptr, sizeof(int64_t) * (n_vertices + 1)));
    if (n_edges) CC_TRY(h2d(ctx, d_indices.p, indices, sizeof(int32_t) * n_edges));
    {
        TimerScope t(ctx, T_OTHER);
        CC_TRY(mcg_k_components(ctx, d_indptr.as<int64_t>(), d_indices.as<int32_t>(), n_vertices,
                                d_work.as<int32_t>(), d_flag.as<int>(), d_labels.as<int32_t>()));
    }
    CC_TRY(d2h(ctx, labels, d_labels.p, sizeof(int32_t) * n_vertices));
    CC_TRY(sync(ctx));
    cleanup();
#undef CC_TRY
    return MCG_OK;
}

// ---- fused plug-in call and the resident step ---------------------------------------------------------------------------
static int step_locked(mcg_ctx *ctx, int32_t *d_argmin, double *d_wmin) {
    if (ctx->n_contigs == 0) return mcg_fail(ctx, MCG_ERR_STATE, "no contigs resident");
    if (ctx->n_cov != ctx->n_contigs)
        return mcg_fail(ctx, MCG_ERR_STATE, "coverage rows (%lld) != contigs (%lld)",
                        (long long)ctx->n_cov, (long long)ctx->n_contigs);
    if (ctx->n_bins == 0) return mcg_fail(ctx, MCG_ERR_STATE, "no centroids resident");
    MCG_TRY(scan_locked(ctx));
    MCG_TRY(cov_terms_locked(ctx));
    const int64_t n = ctx->n_contigs;
    if (!d_argmin) {
        MCG_TRY(mcg_reserve(ctx, ctx->argmin, sizeof(int32_t) * n));
        d_argmin = ctx->argmin.as<int32_t>();
    }
    if (!d_wmin) {
        MCG_TRY(mcg_reserve(ctx, ctx->wmin, sizeof(double) * n));
        d_wmin = ctx->wmin.as<double>();
    }
    return score_centroids_locked(ctx, nullptr, n, nullptr, d_argmin, d_wmin);
}

// The two halves of mcg_step_resident, for callers that put something between them (the
// multi-GPU bench receives the broadcast centroids while the scan is running).
extern "C" int mcg_step_featurise(mcg_ctx *ctx) {
    MCG_ENTER(ctx);
    if (ctx->n_contigs == 0) return mcg_fail(ctx, MCG_ERR_STATE, "no contigs resident");
    if (ctx->n_cov != ctx->n_contigs)
        return mcg_fail(ctx, MCG_ERR_STATE, "coverage rows (%lld) != contigs (%lld)",
                        (long long)ctx->n_cov, (long long)ctx->n_contigs);
    MCG_TRY(scan_locked(ctx));
    return cov_terms_locked(ctx);
}

extern "C" int mcg_step_score(mcg_ctx *ctx, void *d_argmin, void *d_wmin) {
    MCG_ENTER(ctx);
    MCG_TRY(need_features(ctx));
    if (ctx->n_bins == 0) return mcg_fail(ctx, MCG_ERR_STATE, "no centroids resident");
    const int64_t n = ctx->n_contigs;
    int32_t *a = static_cast<int32_t *>(d_argmin);
    double *w = static_cast<double *>(d_wmin);
    if (!a) {
        MCG_TRY(mcg_reserve(ctx, ctx->argmin, sizeof(int32_t) * n));
        a = ctx->argmin.as<int32_t>();
    }
    if (!w) {
        MCG_TRY(mcg_reserve(ctx, ctx->wmin, sizeof(double) * n));
        w = ctx->wmin.as<double>();
    }
    return score_centroids_locked(ctx, nullptr, n, nullptr, a, w);
}

extern "C" int mcg_step_resident(mcg_ctx *ctx, void *d_argmin, void *d_wmin) {
    MCG_ENTER(ctx);
    return step_locked(ctx, static_cast<int32_t *>(d_argmin), static_cast<double *>(d_wmin));
}

extern "C" int mcg_get_assignments(mcg_ctx *ctx, int32_t *argmin, double *wmin) {
    MCG_ENTER(ctx);
    const int64_t n = ctx->n_contigs;
    {
        TimerScope t(ctx, T_D2H);
        if (argmin) MCG_TRY(d2h(ctx, argmin, ctx->argmin.p, sizeof(int32_t) * n));
        if (wmin) MCG_TRY(d2h(ctx, wmin, ctx->wmin.p, sizeof(double) * n));
    }
    const int rc_status = mcg_check_status(ctx);
    MCG_MARK("results on host");
    return rc_status;
}

extern "C" int mcg_featurise_score(mcg_ctx *ctx, const uint8_t *bases, const int64_t *offsets,
                                   int64_t n, int encoding, const double *cov, int n_samples,
                                   const double *cen_prof, const double *cen_cov, int64_t n_bins,
                                   int32_t *argmin, double *wmin) {
    MCG_ENTER(ctx);
    if (n <= 0) return mcg_fail(ctx, MCG_ERR_ARG, "no contigs");
    MCG_MARK("enter");
    const bool waves = offsets && offsets[n] >= (64 << 20) && !getenv("MCG_NO_WAVES");
    // The coverage table either goes first, in one copy (round 1), or its rows travel with their
    // wave on a side stream.  Measured on the B200 box (A/B in one process, scratch/e2e_ab.py):
    // per wave wins when the table is at least as large as the 2-bit contigs (5 M contigs x 100
    // samples: 4 GB next to 3.5 GB -- e2e 337 -> 282 ms, the scans no longer wait for it) and
    // loses below that (1 M x 50: 27.7 -> 36 ms; 100 k x 10: 6.5 -> 8.3 ms), so that is the switch.
    // MCG_COV_WAVES=0 / 1 forces one or the other.
    bool cov_waves = waves && 8.0 * static_cast<double>(n) * n_samples >= 0.25 * static_cast<double>(offsets[n]);
    if (const char *env = getenv("MCG_COV_WAVES")) cov_waves = waves && atoi(env) != 0;
    MCG_TRY(load_coverage_locked(ctx, cov, n, n_samples, !cov_waves));
    MCG_TRY(set_centroids_locked(ctx, cen_prof, cen_cov, n_bins, n_samples));
    MCG_MARK("cov+centroids queued");
    if (waves) {
        // Large loads (ASCII or 2 bit): every contig range is scanned, normalised and scored as
        // soon as its packed words are complete in stream order, while the rest is still
        // uploading; after the last copy only the last range is left to do.
        // the side stream starts behind whatever the compute stream still has queued (the last
        // call's readers of the coverage tables)
        MCG_CUDA(ctx, cudaEventRecord(ctx->copy_start, ctx->stream));
        MCG_CUDA(ctx, cudaStreamWaitEvent(ctx->cov_stream, ctx->copy_start, 0));
        if (!cov_waves) MCG_TRY(cov_terms_locked(ctx));
        MCG_TRY(mcg_reserve(ctx, ctx->counts, sizeof(uint32_t) * MCG_NPROF * n));
        MCG_TRY(mcg_reserve(ctx, ctx->prof, sizeof(double) * MCG_NPROF * n));
        MCG_TRY(mcg_reserve(ctx, ctx->prof_norm, sizeof(double) * n));
        MCG_TRY(mcg_reserve(ctx, ctx->argmin, sizeof(int32_t) * n));
        MCG_TRY(mcg_reserve(ctx, ctx->wmin, sizeof(double) * n));
        MCG_CUDA(ctx, cudaMemsetAsync(ctx->counts.p, 0, sizeof(uint32_t) * MCG_NPROF * n, ctx->stream));
        ctx->n_prof = n;
        const bool was_profiling = ctx->profiling;
        const WaveFactory factory = [&](const std::vector<ContigMeta> &meta, int64_t n_seg) -> WaveFn {
            return [ctx, &meta, n_seg, n, cov, cov_waves](int64_t c0, int64_t c1) -> int {
                if (c1 <= c0) return MCG_OK;
                const bool prof = ctx->profiling;
                ctx->profiling = false;   // one event pair per timer: the waves share T_H2D
                const int64_t seg0 = meta[c0].seg_first;
                const int64_t seg1 = c1 < n ? meta[c1].seg_first : n_seg;
                int rc = mcg_k_scan_range(ctx, ctx->packed.as<uint32_t>(), ctx->meta.as<ContigMeta>(),
                                          ctx->seg_contig.as<int32_t>(), seg0, seg1,
                                          ctx->counts.as<uint32_t>());
                if (rc == MCG_OK)
                    rc = mcg_k_normalise_norms(ctx, ctx->counts.as<uint32_t>() + c0 * MCG_NPROF,
                                               ctx->meta.as<ContigMeta>() + c0, c1 - c0,
                                               ctx->prof.as<double>() + c0 * MCG_NPROF,
                                               ctx->prof_norm.as<double>() + c0);
                // coverage rows of the range: queued on the side stream now (they travel and are
                // featurised under the kernels of the waves in front); the compute stream waits
                // for them here, behind this range's scan
                if (rc == MCG_OK && cov_waves) rc = wave_coverage(ctx, cov, c0, c1);
                if (rc == MCG_OK) {
                    ScoreSide a = contig_side(ctx, nullptr, c1 - c0);
                    a.first = c0;
                    ScoreOut out;
                    out.argmin = ctx->argmin.as<int32_t>() + c0;
                    out.wmin = ctx->wmin.as<double>() + c0;
                    // the kernels the whole set would take in one call (two-kernel path from
                    // 2048 contigs), whatever the size of the range: the weights must not
                    // depend on where the upload happened to cut the waves
                    ctx->force_split = n >= 2048 || ctx->force_split_calls;
                    rc = mcg_k_pairs(ctx, a, centroid_side(ctx), ctx->n_samples, out,
                                     ctx->status.as<int32_t>());
                    ctx->force_split = false;
                }
                ctx->profiling = prof;
                return rc;
            };
        };
        const int rc = load_contigs_locked(ctx, bases, offsets, n, encoding, &factory);
        ctx->profiling = was_profiling;
        MCG_TRY(rc);
        MCG_MARK("load+waves done");
    } else {
        MCG_TRY(load_contigs_locked(ctx, bases, offsets, n, encoding));
        MCG_TRY(step_locked(ctx, nullptr, nullptr));
    }
    {
        TimerScope t(ctx, T_D2H);
        if (argmin) MCG_TRY(d2h(ctx, argmin, ctx->argmin.p, sizeof(int32_t) * n));
        if (wmin) MCG_TRY(d2h(ctx, wmin, ctx->wmin.p, sizeof(double) * n));
    }
    return mcg_check_status(ctx);
}

// ---- timing -------------------------------------------------------------------------------------------------------------------
extern "C" int mcg_profile(mcg_ctx *ctx, int on) {
    MCG_ENTER(ctx);
    ctx->profiling = on != 0;
    ctx->launches = 0;
    for (int i = 0; i < MCG_N_TIMERS; ++i) { ctx->ev_used[i] = false; ctx->ms[i] = 0.f; }
    return MCG_OK;
}

extern "C" int mcg_timings(mcg_ctx *ctx, float *ms, int64_t *launches) {
    MCG_ENTER(ctx);
    MCG_TRY(sync(ctx));
    for (int i = 0; i < MCG_N_TIMERS; ++i) {
        float t = 0.f;
        if (ctx->ev_used[i]) cudaEventElapsedTime(&t, ctx->ev[2 * i], ctx->ev[2 * i + 1]);
        if (ms) ms[i] = t;
        ctx->ev_used[i] = false;
    }
    if (launches) *launches = ctx->launches;
    ctx->launches = 0;
    return MCG_OK;
}

// ---- synthetic data / peak -----------------------------------------------------------------------------------------------------
extern "C" int mcg_synth_bases(mcg_ctx *ctx, const int64_t *offsets, int64_t n,
                               const int32_t *genome_of, const float *trans, int n_genomes,
                               uint64_t seed, double n_frac, double lower_frac, uint8_t *bases) {
    MCG_ENTER(ctx);
    if (n <= 0 || !offsets || !genome_of || !trans || n_genomes <= 0 || !bases || offsets[0] != 0)
        return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    for (int64_t i = 0; i < n; ++i)
        if (genome_of[i] < 0 || genome_of[i] >= n_genomes)
            return mcg_fail(ctx, MCG_ERR_ARG, "genome_of[%lld] out of range", (long long)i);
    const int64_t n_bases = offsets[n];
    MCG_TRY(mcg_reserve(ctx, ctx->ascii, static_cast<size_t>(n_bases) + 256));
    MCG_TRY(mcg_reserve(ctx, ctx->offsets, sizeof(int64_t) * (n + 1)));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], sizeof(int32_t) * n));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[1], sizeof(float) * 16 * n_genomes));
    MCG_TRY(h2d(ctx, ctx->offsets.p, offsets, sizeof(int64_t) * (n + 1)));
    MCG_TRY(h2d(ctx, ctx->scratch[0].p, genome_of, sizeof(int32_t) * n));
    MCG_TRY(h2d(ctx, ctx->scratch[1].p, trans, sizeof(float) * 16 * n_genomes));
    MCG_TRY(mcg_k_synth(ctx, ctx->offsets.as<int64_t>(), n, ctx->scratch[0].as<int32_t>(),
                        ctx->scratch[1].as<float>(), seed, n_frac, lower_frac,
                        ctx->ascii.as<uint8_t>(), n_bases));
    MCG_TRY(d2h(ctx, bases, ctx->ascii.p, static_cast<size_t>(n_bases)));
    return sync(ctx);
}

extern "C" int mcg_pack_2bit_host(const uint8_t *bases, const int64_t *offsets, int64_t n,
                                  uint8_t *packed, int level) {
    if (n < 0 || (n > 0 && (!offsets || !packed))) return MCG_ERR_ARG;
    std::vector<int64_t> first(static_cast<size_t>(n) + 1, 0);
    for (int64_t c = 0; c < n; ++c) {
        const int64_t len = offsets[c + 1] - offsets[c];
        if (len < 0) return MCG_ERR_ARG;
        first[c + 1] = first[c] + 4 * ((len + 63) / 64);
    }
    if (n > 0 && offsets[n] > offsets[0] && !bases) return MCG_ERR_ARG;
    mcg_hostpack_range(bases, offsets, 0, n, first.data(), sizeof(int64_t), packed, level);
    return MCG_OK;
}

// The same with `threads` host threads over contig ranges of equal bases (<= 0: all hardware
// threads): what a FASTA reader calls once at parse time so that later uploads carry 2 bits per base.
extern "C" int mcg_pack_2bit_host_mt(const uint8_t *bases, const int64_t *offsets, int64_t n,
                                     uint8_t *packed, int threads) {
    if (n < 0 || (n > 0 && (!offsets || !packed))) return MCG_ERR_ARG;
    if (n == 0) return MCG_OK;
    std::vector<int64_t> first(static_cast<size_t>(n) + 1, 0);
    for (int64_t c = 0; c < n; ++c) {
        const int64_t len = offsets[c + 1] - offsets[c];
        if (len < 0) return MCG_ERR_ARG;
        first[c + 1] = first[c] + 4 * ((len + 63) / 64);
    }
    if (offsets[n] > offsets[0] && !bases) return MCG_ERR_ARG;
    if (threads <= 0) threads = static_cast<int>(std::thread::hardware_concurrency());
    const int64_t total = offsets[n] - offsets[0];
    if (threads > 1 && total < (8 << 20)) threads = 1;
    std::vector<int64_t> cut(static_cast<size_t>(threads) + 1, n);
    cut[0] = 0;
    for (int t = 1; t < threads; ++t)
        cut[t] = std::lower_bound(offsets, offsets + n + 1, offsets[0] + total / threads * t) - offsets;
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t)
        pool.emplace_back([&, t]() {
            if (cut[t + 1] > cut[t])
                mcg_hostpack_range(bases, offsets, cut[t], cut[t + 1], first.data(), sizeof(int64_t),
                                   packed, -1);
        });
    if (cut[1] > cut[0])
        mcg_hostpack_range(bases, offsets, cut[0], cut[1], first.data(), sizeof(int64_t), packed, -1);
    for (std::thread &t : pool) t.join();
    return MCG_OK;
}

extern "C" int mcg_set_host_pack_threads(mcg_ctx *ctx, int threads) {
    MCG_ENTER(ctx);
    ctx->host_pack_threads = threads;
    return MCG_OK;
}

// ---- multi-GPU exchange over peer memory ---------------------------------------------------------
// The two exchange steps of the path (bin tables from rank 0, per-contig results to
// every rank) are small against NVLink 5 and latency-bound as NCCL collectives.  With every
// rank's buffer mapped into every process (symmetric memory: CUDA VMM / IPC handles, here
// obtained by the caller), they are plain stores: one kernel on the compute stream writes the
// local block into the same place of every peer's buffer, through NVLink / NVSwitch.
struct PeerDst {
    void *p[MCG_MAX_PEERS];
};

template <typename T>
__global__ void push_to_peers_kernel(const T *__restrict__ src, PeerDst dst, int world,
                                     size_t count) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < count;
         i += stride) {
        const T v = src[i];
        for (int r = 0; r < world; ++r) static_cast<T *>(dst.p[r])[i] = v;
    }
}

extern "C" int mcg_push_to_peers(mcg_ctx *ctx, const void *d_src, size_t bytes,
                                 void *const *peer_dst, int world, void *stream) {
    MCG_ENTER(ctx);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    if (!d_src || !peer_dst || world < 1 || world > MCG_MAX_PEERS || (bytes & 3))
        return mcg_fail(ctx, MCG_ERR_ARG, "bad peer push arguments");
    if (bytes == 0) return MCG_OK;
    PeerDst dst;
    bool aligned = (reinterpret_cast<uintptr_t>(d_src) & 15) == 0 && (bytes & 15) == 0;
    for (int r = 0; r < world; ++r) {
        if (!peer_dst[r]) return mcg_fail(ctx, MCG_ERR_ARG, "peer %d has no buffer", r);
        dst.p[r] = peer_dst[r];
        aligned = aligned && (reinterpret_cast<uintptr_t>(peer_dst[r]) & 15) == 0;
    }
    const int threads = 256;
    const size_t count = aligned ? bytes / 16 : bytes / 4;   // 16-byte stores when everything allows it
    size_t blocks = (count + threads - 1) / threads;
    if (blocks > static_cast<size_t>(4 * ctx->sm_count)) blocks = 4 * ctx->sm_count;
    if (aligned)
        push_to_peers_kernel<uint4><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
            static_cast<const uint4 *>(d_src), dst, world, count);
    else
        push_to_peers_kernel<uint32_t><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
            static_cast<const uint32_t *>(d_src), dst, world, count);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

extern "C" int mcg_set_reserved_sms(mcg_ctx *ctx, int sms) {
    MCG_ENTER(ctx);
    if (sms < 0 || sms >= ctx->sm_count) return mcg_fail(ctx, MCG_ERR_ARG, "bad SM count");
    ctx->reserved_sms = sms;
    return MCG_OK;
}

extern "C" int mcg_set_prune(mcg_ctx *ctx, int enabled) {
    MCG_ENTER(ctx);
    ctx->prune = enabled ? 1 : 0;
    return MCG_OK;
}

extern "C" int mcg_set_tc(mcg_ctx *ctx, int enabled) {
    MCG_ENTER(ctx);
    ctx->tc = enabled ? 1 : 0;
    return MCG_OK;
}

// Diagnostic / parity form of the tensor-core distance bound: d2a [nq, n_bins] (fp32) of the
// queries against the resident centroids and the margin [nq] within which every d2a is
// guaranteed to lie around the float64 squared distance.
extern "C" int mcg_tc_distances(mcg_ctx *ctx, const int64_t *queries, int64_t nq, float *d2a,
                                float *margin) {
    MCG_ENTER(ctx);
    MCG_TRY(need_features(ctx));
    if (ctx->n_bins == 0) return mcg_fail(ctx, MCG_ERR_STATE, "no centroids resident");
    if (nq < 0 || (nq > 0 && (!d2a || !margin))) return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    if (nq == 0) return MCG_OK;
    const int64_t *d_queries = nullptr;
    if (queries) {
        MCG_TRY(check_ids(ctx, queries, nq, ctx->n_prof, "queries"));
        MCG_TRY(mcg_reserve(ctx, ctx->scratch[0], sizeof(int64_t) * nq));
        MCG_TRY(h2d(ctx, ctx->scratch[0].p, queries, sizeof(int64_t) * nq));
        d_queries = ctx->scratch[0].as<int64_t>();
    } else if (nq > ctx->n_prof) {
        return mcg_fail(ctx, MCG_ERR_ARG, "nq exceeds the resident contig table");
    }
    const int64_t B = ctx->n_bins;
    const long long ldf = mcg_tc_ldf(B);
    const long long rows_pad = (nq + 127) / 128 * 128;
    ctx->cand_valid = false;
    ctx->redo_valid = false;
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[5], sizeof(float) * rows_pad * ldf));
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[1], sizeof(float) * nq));
    MCG_TRY(mcg_k_tc_prep_bins(ctx, centroid_side(ctx)));
    {
        TimerScope t(ctx, T_DIST);
        MCG_TRY(mcg_k_tc_dist(ctx, contig_side(ctx, d_queries, nq), B, ctx->scratch[5].as<float>(), ldf,
                              ctx->scratch[1].as<float>(), nullptr, 0));
    }
    MCG_CUDA(ctx, cudaMemcpy2DAsync(d2a, sizeof(float) * B, ctx->scratch[5].p, sizeof(float) * ldf,
                                    sizeof(float) * B, nq, cudaMemcpyDeviceToHost, ctx->stream));
    MCG_TRY(d2h(ctx, margin, ctx->scratch[1].p, sizeof(float) * nq));
    return sync(ctx);
}

extern "C" int mcg_prune_stats(mcg_ctx *ctx, int64_t stats[4]) {
    MCG_ENTER(ctx);
    if (!stats) return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    return mcg_k_prune_stats(ctx, stats);
}

extern "C" int mcg_ingest_stats(mcg_ctx *ctx, int64_t stats[4]) {
    MCG_ENTER(ctx);
    if (!stats) return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    for (int i = 0; i < 4; ++i) stats[i] = ctx->ingest_stats[i];
    return MCG_OK;
}

extern "C" int mcg_last_redo_count(mcg_ctx *ctx, int64_t *count) {
    MCG_ENTER(ctx);
    if (!count) return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    *count = 0;
    if (!ctx->redo_valid) return MCG_OK;
    int32_t c = 0;
    MCG_CUDA(ctx, cudaMemcpyAsync(&c, ctx->scratch[4].p, sizeof c, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    MCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *count = c;
    return MCG_OK;
}

extern "C" int mcg_fp64_peak(mcg_ctx *ctx, double *tflops) {
    MCG_ENTER(ctx);
    if (!tflops) return mcg_fail(ctx, MCG_ERR_ARG, "bad arguments");
    return mcg_k_fp64_peak(ctx, tflops);
}
