// Tetramer featurisation kernels.
//
//   reference: feature_utils.py:17-70 (nt_bits / mer2bits / compute_kmer_inds /
//   count_kmers).  count_kmers slides a 4-base window over the contig, looks the
//   window code up in the canonical table and does profile[index] += 1.
//
// Data layout in HBM
//   packed store   2 bits per base (A0 C1 G2 T3, anything else 0 -- feature_utils.py:21),
//                  base i of a contig in bits [2(i%16), 2(i%16)+2) of 32-bit word i/16;
//                  every contig starts on a 16-byte boundary.
//   window code    x = bits [2i, 2i+8) of that stream = b_i | b_{i+1}<<2 | b_{i+2}<<4 |
//                  b_{i+3}<<6, i.e. the reference's code with the four bases in reversed
//                  field order; g_canon_x[x] is the reference slot of that window.
//   work unit      a *segment* = 256 consecutive windows of one contig = 16 packed words
//                  (+1 halo word); one lane owns one segment, a warp takes 32
//                  consecutive segments (a "tile") from a global atomic counter, so the
//                  work is balanced by bases, not by contigs (Flye-style 5 Mb contigs and
//                  1 kb contigs mix freely).
//
// Counting
//   Shared-memory atomics cost ~2 cycles per lane on this part, i.e. ~0.5 base/clk/SM.
//   Instead every lane owns a private column of 256 one-byte counters in shared memory,
//   laid out [code>>2][lane][code&3] so that the bank is the lane: the read-modify-write
//   (LDS.U8, IADD, STS.U8) never conflicts and needs no atomic.  A byte holds at most
//   255, which is why a lane takes 255 windows into its column and sends the 256th
//   straight to global memory.  After the 32 segments of a tile are counted the warp
//   sums runs of columns that belong to the same contig (rotated so that bank = column),
//   folds the 256 codes onto the 136 canonical slots and adds them to the contig's row
//   with global reductions.
#include "mcg_internal.cuh"

namespace {

constexpr int kScanWarps = 27;      // 27 x 8 KB tables + alignment slack = 223 KB: one CTA per SM
constexpr int kScanThreads = kScanWarps * 32;
constexpr int kScanCtasPerSm = 1;
constexpr uint32_t kTableBytes = 8192;   // 256 codes x 32 lanes x 1 byte

__device__ uint8_t g_canon_x[256];   // window code x (reversed field order) -> canonical slot

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
}
// (a & b) | c in one LOP3 (the compiler otherwise splits the address build in three)
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint4 ldg_u4(const uint32_t *p) {
    return __ldg(reinterpret_cast<const uint4 *>(p));
}

// One window whose code starts at bit 2*I of the 64-bit value hi:lo.
//   z: the window at bit 0 (one funnel shift); its low two bits select the byte in the word
//   y = z << 5: code bits [2,8) at address bits [7,13)
// Every LOP3 / SHF / IADD shares the ALU pipe at one warp instruction per two cycles, which is
// what bounds this loop; the shift by 5 and the "+ 1" are therefore written as multiply-adds
// (`one` is a register the compiler cannot fold), which issue to the FMA pipe instead.
template <int I>
__device__ __forceinline__ void count_window(uint32_t lo, uint32_t hi, uint32_t lanebase,
                                             uint32_t one) {
    uint32_t z;
    if constexpr (I == 0)
        z = lo;
    else
        z = __funnelshift_r(lo, hi, static_cast<uint32_t>(2 * I));
    uint32_t y;
    asm("mul.lo.u32 %0, %1, 32;" : "=r"(y) : "r"(z));
    const uint32_t addr = and_or(z, 3u, and_or(y, 0x1F80u, lanebase));
    uint32_t v = lds_u8(addr);
    asm("mad.lo.u32 %0, %0, %1, %1;" : "+r"(v) : "r"(one));
    sts_u8(addr, v);
}

// 16 windows starting in word `lo` (the last three reach into `hi`).  When `divert` is
// set the 16th window goes to global memory instead of the byte counter.
__device__ __forceinline__ void count_word(uint32_t lo, uint32_t hi, uint32_t lanebase,
                                           uint32_t one, bool divert, uint32_t *row) {
    count_window<0>(lo, hi, lanebase, one);
    count_window<1>(lo, hi, lanebase, one);
    count_window<2>(lo, hi, lanebase, one);
    count_window<3>(lo, hi, lanebase, one);
    count_window<4>(lo, hi, lanebase, one);
    count_window<5>(lo, hi, lanebase, one);
    count_window<6>(lo, hi, lanebase, one);
    count_window<7>(lo, hi, lanebase, one);
    count_window<8>(lo, hi, lanebase, one);
    count_window<9>(lo, hi, lanebase, one);
    count_window<10>(lo, hi, lanebase, one);
    count_window<11>(lo, hi, lanebase, one);
    count_window<12>(lo, hi, lanebase, one);
    count_window<13>(lo, hi, lanebase, one);
    count_window<14>(lo, hi, lanebase, one);
    if (!divert) {
        count_window<15>(lo, hi, lanebase, one);
    } else {
        const uint32_t x = __funnelshift_r(lo, hi, 30) & 0xFFu;
        atomicAdd(row + g_canon_x[x], 1u);
    }
}

// The last, partly filled word of a contig: `rem` (1..15) windows.
__device__ __forceinline__ void count_partial(uint32_t lo, uint32_t hi, uint32_t lanebase,
                                              int rem) {
    for (int i = 0; i < rem; ++i) {
        const uint32_t x = __funnelshift_r(lo, hi, 2 * i) & 0xFFu;
        const uint32_t addr = lanebase | ((x >> 2) << 7) | (x & 3u);
        sts_u8(addr, lds_u8(addr) + 1u);
    }
}

// The lane's eight code totals (table rows lane and lane + 32, bytes 0..3) of one contig go to
// the contig's row of `counts`, folded onto the canonical slots.
__device__ __forceinline__ void emit_held(uint32_t *__restrict__ counts, int contig,
                                          uint32_t (&acc)[8], uint32_t canA, uint32_t canB) {
    if (contig >= 0) {
        uint32_t *row = counts + static_cast<int64_t>(contig) * MCG_NPROF;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (acc[b]) atomicAdd(row + ((canA >> (8 * b)) & 0xFFu), acc[b]);
            if (acc[4 + b]) atomicAdd(row + ((canB >> (8 * b)) & 0xFFu), acc[4 + b]);
        }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b] = 0;
}

__global__ void __launch_bounds__(kScanThreads, kScanCtasPerSm)
scan_kernel(const uint32_t *__restrict__ packed, const ContigMeta *__restrict__ meta,
            const int32_t *__restrict__ seg_contig, long long seg_base, long long n_seg,
            unsigned long long *__restrict__ tile_counter, uint32_t *__restrict__ counts) {
    // segments [seg_base, n_seg) of the store (a contig range while the rest is still uploading)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    // the (y & 0x1F80) | lanebase address form needs every warp's table 8 KB-aligned
    const uint32_t tab = ((smem_u32(smem_raw) + kTableBytes - 1u) & ~(kTableBytes - 1u)) +
                         warp * kTableBytes;
    const uint32_t lanebase = tab + lane * 4u;
    const uint32_t one = n_seg > 0 ? 1u : 0u;   // a 1 the compiler has to keep in a register
    {   // the alignment slack assumes the dynamic window starts >= 1 KB into shared memory
        uint32_t dyn_size;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
        if (tab + kTableBytes > smem_u32(smem_raw) + dyn_size) __trap();
    }

#pragma unroll 8
    for (int r = 0; r < 64; ++r) sts_u32(lanebase + r * 128u, 0u);
    __syncwarp();

    // canonical slots of the 8 codes this lane folds at flush time
    uint32_t canA = 0, canB = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        canA |= static_cast<uint32_t>(g_canon_x[4 * lane + b]) << (8 * b);
        canB |= static_cast<uint32_t>(g_canon_x[128 + 4 * lane + b]) << (8 * b);
    }

    // A warp takes strips of kStrip consecutive tiles and keeps the totals of the contig it is
    // in (`held`) in registers across tiles: a long contig costs one set of global reductions
    // per strip instead of one per tile, which otherwise serialise on its 136 addresses.
    constexpr int kStrip = 4;
    uint32_t acc[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b] = 0;
    int held = -1;
    const long long n_tiles = (n_seg - seg_base + 31) >> 5;
    unsigned long long tile = 0;
    int in_strip = kStrip;
    for (;;) {
        if (in_strip == kStrip) {
            // the totals do not travel between strips: the next strip is somewhere else
            emit_held(counts, held, acc, canA, canB);
            held = -1;
            if (lane == 0) tile = atomicAdd(tile_counter, 1ULL) * kStrip;
            tile = __shfl_sync(0xffffffffu, tile, 0);
            in_strip = 0;
        } else {
            ++tile;
        }
        ++in_strip;
        if (static_cast<long long>(tile) >= n_tiles) break;

        const long long g = seg_base + static_cast<long long>(tile) * 32 + lane;
        int contig = -1, nwin = 0, k = 0;
        const uint32_t *wp = packed;
        if (g < n_seg) {
            contig = __ldg(seg_contig + g);
            const int4 raw = __ldg(reinterpret_cast<const int4 *>(meta + contig));
            const long long pk_word =
                (static_cast<long long>(static_cast<uint32_t>(raw.y)) << 32) |
                static_cast<uint32_t>(raw.x);
            k = static_cast<int>(g - raw.z);
            nwin = min(256, raw.w - 256 * k);
            wp = packed + pk_word + 16ll * k;
        }
        uint32_t *row = counts + static_cast<int64_t>(max(contig, 0)) * MCG_NPROF;

        // 16 words of this lane's segment; the halo word comes from the next lane when
        // that lane holds the next segment of the same contig, else from memory.
        uint4 cur = ldg_u4(wp);
        uint4 nxt = ldg_u4(wp + 4);
        const uint32_t nb_word = __shfl_down_sync(0xffffffffu, cur.x, 1);
        const int nb_contig = __shfl_down_sync(0xffffffffu, contig, 1);
        const int nb_k = __shfl_down_sync(0xffffffffu, k, 1);
        uint32_t halo;
        if (lane < 31 && nb_contig == contig && nb_k == k + 1)
            halo = nb_word;
        else
            halo = __ldg(wp + 16);

        const int nfull = nwin >> 4;
        const int rem = nwin & 15;
#pragma unroll 1
        for (int gi = 0; gi < 4; ++gi) {
            uint4 nn = cur;
            if (gi < 2) nn = ldg_u4(wp + 8 + 4 * gi);
            const uint32_t w0 = cur.x, w1 = cur.y, w2 = cur.z, w3 = cur.w;
            const uint32_t w4 = (gi < 3) ? nxt.x : halo;
            const int j = 4 * gi;
            if (j + 0 < nfull) count_word(w0, w1, lanebase, one, false, row);
            else if (j + 0 == nfull) count_partial(w0, w1, lanebase, rem);
            if (j + 1 < nfull) count_word(w1, w2, lanebase, one, false, row);
            else if (j + 1 == nfull) count_partial(w1, w2, lanebase, rem);
            if (j + 2 < nfull) count_word(w2, w3, lanebase, one, false, row);
            else if (j + 2 == nfull) count_partial(w2, w3, lanebase, rem);
            if (j + 3 < nfull) count_word(w3, w4, lanebase, one, gi == 3, row);
            else if (j + 3 == nfull) count_partial(w3, w4, lanebase, rem);
            cur = nxt;
            nxt = nn;
        }
        __syncwarp();

        // ---- flush: lane j sums table rows j and j+32 over the 32 columns ---------
        // Columns of one contig are adjacent lanes; the runs are warp-uniform, so all lanes
        // walk the same run at the same time and emit together.  Inside a run lane j starts
        // at column c0 + j mod len, which keeps bank = column distinct across lanes when
        // the run spans the warp (the common case).
        const uint32_t rowA = tab + lane * 128u;
        const int prev = __shfl_up_sync(0xffffffffu, contig, 1);
        uint32_t heads = __ballot_sync(0xffffffffu, lane == 0 || contig != prev);
        while (heads) {
            const int c0 = __ffs(heads) - 1;
            heads &= heads - 1;
            const int c1 = heads ? (__ffs(heads) - 1) : 32;
            const int len = c1 - c0;
            const int run_contig = __shfl_sync(0xffffffffu, contig, c0);
            if (run_contig < 0) continue;   // idle lanes: their columns were never touched
            uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 0;
            int col = c0 + static_cast<int>(lane) % len;
#pragma unroll 4
            for (int t = 0; t < len; ++t) {
                const uint32_t addrA = rowA + col * 4u;
                const uint32_t wA = lds_u32(addrA);
                const uint32_t wB = lds_u32(addrA + 4096u);
                sts_u32(addrA, 0u);
                sts_u32(addrA + 4096u, 0u);
                a0 += wA & 0x00FF00FFu;
                a1 += (wA >> 8) & 0x00FF00FFu;
                b0 += wB & 0x00FF00FFu;
                b1 += (wB >> 8) & 0x00FF00FFu;
                if (++col == c1) col = c0;
            }
            if (run_contig != held) {
                emit_held(counts, held, acc, canA, canB);
                held = run_contig;
            }
            // word byte k holds code 4*row + k; a0 = bytes 0|2, a1 = bytes 1|3 (16-bit fields)
            acc[0] += a0 & 0xFFFFu; acc[1] += a1 & 0xFFFFu; acc[2] += a0 >> 16; acc[3] += a1 >> 16;
            acc[4] += b0 & 0xFFFFu; acc[5] += b1 & 0xFFFFu; acc[6] += b0 >> 16; acc[7] += b1 >> 16;
        }
        __syncwarp();
    }
    emit_held(counts, held, acc, canA, canB);
}

// freqs = counts / max(1, sum(counts)) with sum(counts) = windows  (feature_utils.py:70)
__global__ void normalise_kernel(const uint32_t *__restrict__ counts,
                                 const ContigMeta *__restrict__ meta, long long n_contigs,
                                 double *__restrict__ prof) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n_contigs * MCG_NPROF) return;
    const long long c = i / MCG_NPROF;
    const int nwin = meta[c].nwin;
    const double denom = static_cast<double>(nwin > 1 ? nwin : 1);
    prof[i] = static_cast<double>(counts[i]) / denom;
}

// seg_contig[g] = the contig that owns scan segment g.
__global__ void fill_seg_contig_kernel(const ContigMeta *__restrict__ meta, long long n_contigs,
                                       long long n_seg, int32_t *__restrict__ seg_contig) {
    const long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (g >= n_seg) return;
    long long lo = 0, hi = n_contigs;   // largest c with seg_first[c] <= g
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (meta[mid].seg_first <= g) lo = mid; else hi = mid;
    }
    seg_contig[g] = static_cast<int32_t>(lo);
}

// ---- ASCII -> 2 bit ---------------------------------------------------------------
// 0x80 in every byte of x that is zero (exact, no cross-byte borrow).
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) {
    const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | x | 0x7F7F7F7Fu);
}
// Four ASCII bases -> 8 bits.  Only 'C', 'G', 'T' are non-zero (feature_utils.py:21,31-35).
__device__ __forceinline__ uint32_t pack4(uint32_t w) {
    const uint32_t isC = zero_bytes(w ^ 0x43434343u);
    const uint32_t isG = zero_bytes(w ^ 0x47474747u);
    const uint32_t isT = zero_bytes(w ^ 0x54545454u);
    const uint32_t code = ((isC | isT) >> 7) | ((isG | isT) >> 6);   // 2 bits per byte
    return (code * 0x01041040u) >> 24;   // gather the four 2-bit fields into one byte
}

// One thread packs 64 bases (16 output bytes).
__global__ void pack_ascii_kernel(const uint8_t *__restrict__ ascii,
                                  const int64_t *__restrict__ offsets,
                                  const ContigMeta *__restrict__ meta, long long n_contigs,
                                  long long unit0, long long n_units,
                                  uint32_t *__restrict__ packed) {
    const long long u = unit0 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (u >= unit0 + n_units) return;
    const long long word = 4 * u;
    long long lo = 0, hi = n_contigs;   // largest c with pk_word[c] <= word
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (meta[mid].pk_word <= word) lo = mid; else hi = mid;
    }
    const long long c = lo;
    const long long base0 = (word - meta[c].pk_word) * 16;
    const long long len = offsets[c + 1] - offsets[c];
    const long long left = len - base0;
    const int nb = left < 0 ? 0 : (left > 64 ? 64 : static_cast<int>(left));
    const uint8_t *src = ascii + offsets[c] + base0;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
    const uint32_t *src4 = reinterpret_cast<const uint32_t *>(addr & ~static_cast<uintptr_t>(3));
    const uint32_t sh = static_cast<uint32_t>(addr & 3) * 8;
    uint32_t w[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) w[i] = __ldg(src4 + i);
    uint32_t out[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        uint32_t r = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t v = __funnelshift_r(w[4 * o + q], w[4 * o + q + 1], sh);
            r |= pack4(v) << (8 * q);
        }
        const int valid = nb - 16 * o;
        if (valid <= 0) r = 0;
        else if (valid < 16) r &= (1u << (2 * valid)) - 1u;
        out[o] = r;
    }
    reinterpret_cast<uint4 *>(packed)[u] = make_uint4(out[0], out[1], out[2], out[3]);
}

// ---- synthetic contigs ----------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// One thread writes 256 consecutive bases of the blob.
__global__ void synth_kernel(const int64_t *__restrict__ offsets, long long n_contigs,
                             const int32_t *__restrict__ genome_of,
                             const float *__restrict__ trans, unsigned long long seed,
                             float n_run_prob, float lower_frac, uint8_t *__restrict__ bases,
                             long long n_bases) {
    const long long chunk = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long p0 = chunk * 256;
    if (p0 >= n_bases) return;
    const long long p1 = p0 + 256 < n_bases ? p0 + 256 : n_bases;
    long long lo = 0, hi = n_contigs;   // largest c with offsets[c] <= p0
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (offsets[mid] <= p0) lo = mid; else hi = mid;
    }
    long long c = lo;
    long long cend = offsets[c + 1];
    const uint64_t key = splitmix64(seed ^ (static_cast<uint64_t>(chunk) * 0xD6E8FEB86659FD93ull));
    // an N run somewhere in this chunk
    int n_lo = 1 << 30, n_hi = 0;
    if (static_cast<float>(key & 0xFFFFFF) * (1.0f / 16777216.0f) < n_run_prob) {
        n_lo = static_cast<int>((key >> 24) & 0xFF);
        n_hi = n_lo + 1 + static_cast<int>((key >> 32) % 50);
    }
    const char alphabet[4] = {'A', 'C', 'G', 'T'};
    uint32_t prev = static_cast<uint32_t>(key >> 40) & 3u;
    uint64_t bits = 0;
    int have = 0;
    uint64_t ctr = 0;
    for (long long p = p0; p < p1; ++p) {
        while (p >= cend) { ++c; cend = offsets[c + 1]; }
        const int g = genome_of[c];
        if (have == 0) { bits = splitmix64(key + (++ctr)); have = 4; }
        const float uni = static_cast<float>(bits & 0xFFFF) * (1.0f / 65536.0f);
        bits >>= 16;
        --have;
        const float *t = trans + (g * 4 + prev) * 4;
        uint32_t b = 3;
        const float c0 = t[0], c1 = c0 + t[1], c2 = c1 + t[2];
        if (uni < c0) b = 0; else if (uni < c1) b = 1; else if (uni < c2) b = 2;
        prev = b;
        uint8_t ch = static_cast<uint8_t>(alphabet[b]);
        const int rel = static_cast<int>(p - p0);
        if (rel >= n_lo && rel < n_hi) ch = 'N';
        const uint64_t ck = splitmix64(seed * 31ull + static_cast<uint64_t>(c));
        if (static_cast<float>(ck & 0xFFFFFF) * (1.0f / 16777216.0f) < lower_frac) ch |= 0x20;
        bases[p] = ch;
    }
}

}  // namespace

// ---- launchers --------------------------------------------------------------------------
static bool g_canon_ready[64] = {};

static int ensure_canon_table(mcg_ctx *ctx) {
    if (ctx->device >= 0 && ctx->device < 64 && g_canon_ready[ctx->device]) return MCG_OK;
    uint8_t ref[256], x_table[256];
    mcg_kmer_table(ref);
    for (int x = 0; x < 256; ++x) {
        // x = b0 | b1<<2 | b2<<4 | b3<<6  ->  reference code b0<<6 | b1<<4 | b2<<2 | b3
        const int code = ((x & 3) << 6) | (((x >> 2) & 3) << 4) | (((x >> 4) & 3) << 2) |
                         ((x >> 6) & 3);
        x_table[x] = ref[code];
    }
    MCG_CUDA(ctx, cudaMemcpyToSymbolAsync(g_canon_x, x_table, 256, 0, cudaMemcpyHostToDevice,
                                          ctx->stream));
    MCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->device >= 0 && ctx->device < 64) g_canon_ready[ctx->device] = true;
    return MCG_OK;
}

int mcg_k_pack_ascii(mcg_ctx *ctx, const uint8_t *d_ascii, const int64_t *d_offsets,
                     const ContigMeta *d_meta, int64_t n_contigs, int64_t word0, int64_t n_words,
                     uint32_t *d_packed) {
    const int64_t unit0 = word0 / 4, n_units = n_words / 4;
    if (n_units == 0 || n_contigs == 0) return MCG_OK;
    const int threads = 256;
    const int64_t blocks = (n_units + threads - 1) / threads;
    pack_ascii_kernel<<<static_cast<unsigned>(blocks), threads, 0, ctx->stream>>>(
        d_ascii, d_offsets, d_meta, n_contigs, unit0, n_units, d_packed);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_fill_seg_contig(mcg_ctx *ctx, const ContigMeta *d_meta, int64_t n_contigs,
                          int64_t n_seg, int32_t *d_seg_contig) {
    if (n_seg == 0) return MCG_OK;
    const int threads = 256;
    const int64_t blocks = (n_seg + threads - 1) / threads;
    fill_seg_contig_kernel<<<static_cast<unsigned>(blocks), threads, 0, ctx->stream>>>(
        d_meta, n_contigs, n_seg, d_seg_contig);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

// Segments [seg0, seg1) into counts (which the caller has zeroed).
int mcg_k_scan_range(mcg_ctx *ctx, const uint32_t *d_packed, const ContigMeta *d_meta,
                     const int32_t *d_seg_contig, int64_t seg0, int64_t seg1, uint32_t *d_counts) {
    MCG_TRY(ensure_canon_table(ctx));
    if (seg1 <= seg0) return MCG_OK;
    MCG_TRY(mcg_reserve(ctx, ctx->tile_counter, sizeof(unsigned long long)));
    MCG_CUDA(ctx, cudaMemsetAsync(ctx->tile_counter.p, 0, sizeof(unsigned long long),
                                  ctx->stream));
    const size_t smem = kScanWarps * kTableBytes + kTableBytes - 1024;   // + alignment slack
    static bool attr_set[64] = {};
    if (!(ctx->device < 64 && attr_set[ctx->device])) {
        MCG_CUDA(ctx, cudaFuncSetAttribute(scan_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
        if (ctx->device < 64) attr_set[ctx->device] = true;
    }
    const int64_t n_tiles = (seg1 - seg0 + 31) / 32;
    // the scan CTAs are persistent and fill an SM each (shared memory): SMs left free let a
    // collective on another stream start under the scan instead of behind it
    int sms = ctx->sm_count - ctx->reserved_sms;
    if (sms < 1) sms = 1;
    int64_t grid = static_cast<int64_t>(sms) * kScanCtasPerSm;
    const int64_t need = (n_tiles + kScanWarps - 1) / kScanWarps;
    if (grid > need) grid = need;
    scan_kernel<<<static_cast<unsigned>(grid), kScanThreads, smem, ctx->stream>>>(
        d_packed, d_meta, d_seg_contig, seg0, seg1, ctx->tile_counter.as<unsigned long long>(),
        d_counts);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_scan(mcg_ctx *ctx, const uint32_t *d_packed, const ContigMeta *d_meta,
               const int32_t *d_seg_contig, int64_t n_seg, int64_t n_contigs,
               uint32_t *d_counts) {
    MCG_CUDA(ctx, cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * MCG_NPROF * n_contigs,
                                  ctx->stream));
    return mcg_k_scan_range(ctx, d_packed, d_meta, d_seg_contig, 0, n_seg, d_counts);
}

int mcg_k_normalise(mcg_ctx *ctx, const uint32_t *d_counts, const ContigMeta *d_meta,
                    int64_t n_contigs, double *d_prof) {
    if (n_contigs == 0) return MCG_OK;
    const int threads = 256;
    const int64_t total = n_contigs * MCG_NPROF;
    const int64_t blocks = (total + threads - 1) / threads;
    normalise_kernel<<<static_cast<unsigned>(blocks), threads, 0, ctx->stream>>>(
        d_counts, d_meta, n_contigs, d_prof);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_synth(mcg_ctx *ctx, const int64_t *d_offsets, int64_t n, const int32_t *d_genome_of,
                const float *d_trans, uint64_t seed, double n_frac, double lower_frac,
                uint8_t *d_bases, int64_t n_bases) {
    if (n_bases == 0) return MCG_OK;
    const int64_t chunks = (n_bases + 255) / 256;
    const int threads = 128;
    const int64_t blocks = (chunks + threads - 1) / threads;
    // expected N bases per chunk = prob * 25.5 (mean run length) = n_frac * 256
    const float run_prob = static_cast<float>(n_frac * 256.0 / 25.5);
    synth_kernel<<<static_cast<unsigned>(blocks), threads, 0, ctx->stream>>>(
        d_offsets, n, d_genome_of, d_trans, seed, run_prob, static_cast<float>(lower_frac),
        d_bases, n_bases);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}
