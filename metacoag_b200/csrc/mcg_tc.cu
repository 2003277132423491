// Tetramer distances on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), as a BOUND.
//
//   reference arithmetic: matching_utils.py:28-39 -- d = ||u - v||_2 of two 136-slot profiles,
//   then prob_comp(d).  Rule C (label_prop_utils.py:308-345) needs d for every (contig, bin)
//   pair only to rule bins OUT: the branch and bound of mcg_score.cu lists the few bins whose
//   weight can still be the minimum, and only those pairs are scored in float64.
//
// So the all-pairs distance does not have to be float64.  This file computes
//     d2a(u, v) = |x_u|^2 + |x_v|^2 - 2 x_u.x_v,   x = profile - c0,
// with x_u.x_v on tcgen05.mma (kind::f16, fp32 accumulators in tensor memory): every x is split
// into two fp16 numbers, x * 2^12 = hi + lo (22 significant bits), and the three products
// hi.hi + hi.lo + lo.hi are ONE GEMM over a concatenated K (A' = [hi | hi | lo],
// B' = [hi | lo | hi], 3 x 144 = 432 of 448 fp16 per row).  c0 is the profile of uniform base
// composition (1/256 per tetramer code, so 1/256 or 2/256 per canonical slot): centring keeps
// |x| at a few 10^-2 and with it the absolute error of the product.  The result is used with a
// rigorous margin  |d2a - d^2| <= kTcRel |x_u| |x_v| + kTcAbs (|x_u| + |x_v|) + kTcFloor
// (tests/test_tc.py measures the actual error against float64 on adversarial rows): a bin is
// ruled out only when even d2a - margin is beyond the threshold, and every listed pair gets its
// exact float64 d^2 again (mcg_score.cu, exact_d2), so argmin and weights keep their bits.
//
// Kernel shape (persistent: one CTA per SM walks 128-contig row tiles; 6 warps, warp-specialised):
//   warp 0   TMA producer: the row tile's A' (7 k-blocks of 128 rows x 128 bytes, SWIZZLE_128B),
//            resident while the tile's bin blocks run, one full / empty mbarrier pair per
//            k-block so that the NEXT row tile streams in under the MMAs of the last bin block;
//            the B' tiles of every 128-bin block through a 4-stage ring
//            (cp.async.bulk.tensor.2d + mbarrier complete_tx);
//   warp 1   MMA issuer: one elected lane issues 4 tcgen05.mma (M 128, N 128, K 16) per
//            k-block, tcgen05.commit frees the ring slot (and, in the last bin block, the A'
//            k-block) / publishes the accumulator; two accumulator stages (2 x 128 TMEM
//            columns) so that the epilogue of one block runs under the MMAs of the next, across
//            row tiles as well;
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> d2a in fp32 ->
//            128-byte stores of the thread's row, plus the smallest lower bound per 64-bin tile.
// Bound: HBM (896 B of A' in and 4 B per pair out per contig; the MMAs are ~1 us per row tile).
#include <cuda.h>
#include <cuda_fp16.h>

#include "mcg_internal.cuh"

namespace {

constexpr int kTcSeg = 144;                    // one of the three K segments (136 slots + 8 zeros)
constexpr int kTcK = 448;                      // fp16 per operand row (3 x 144 + 16 zeros) = 896 B
constexpr int kTcM = 128, kTcN = 128;          // MMA tile
constexpr int kTcKB = 64;                      // k-block: 64 fp16 = 128 bytes = the swizzle span
constexpr int kTcKBlocks = kTcK / kTcKB;       // 7
constexpr int kTcStages = 4;                   // B' ring
constexpr int kTcTileBytes = kTcM * kTcKB * 2; // 16 KB per k-block of either operand
constexpr int kTcThreads = 192;
constexpr double kTcScale = 4096.0;            // x * 2^12 = hi + lo
constexpr size_t kTcSmem = 1024 + static_cast<size_t>(kTcKBlocks + kTcStages) * kTcTileBytes + 512;

// ---- PTX wrappers (sm_100a) -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst,
                                            int crd0, int crd1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(crd0),
          "r"(crd1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
                 : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, both operands K-major; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// the mbarrier gets one arrival when every tcgen05 operation issued so far by this thread is done
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base lane + t)
__device__ __forceinline__ void tmem_ld32(uint32_t addr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor of a K-major tile whose rows are 128 bytes, laid out by TMA
// with SWIZZLE_128B (8-row groups of 1024 bytes): start address >> 4, leading byte offset
// (unused for swizzled K-major layouts) 1, stride byte offset 1024 >> 4 between 8-row groups,
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// Instruction descriptor: D fp32 (bits 4-5 = 1), A and B fp16 (formats 0), both K-major (bits 15,
// 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t kTcIdesc = (1u << 4) | (static_cast<uint32_t>(kTcN >> 3) << 17) |
                              (static_cast<uint32_t>(kTcM >> 4) << 24);

struct TcArgs {
    const float *nu;        // [rows] |x_u|^2
    const float *nv;        // [bins] |x_v|^2
    const float *mrow;      // [rows] margin of the row against any bin
    float *d2a;             // [rows padded to 128][ldf]
    long long ldf;          // multiple of 128
    unsigned long long *tilemin;   // [rows][nbt64] ordered bits of the smallest lower bound, or null
    long long nbt64;
    long long rows, bins;
};

__global__ void __launch_bounds__(kTcThreads, 1)
tc_dist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               TcArgs a) {
    extern __shared__ uint8_t tc_smem_raw[];
    // SWIZZLE_128B tiles want 1024-byte alignment
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) &
                                                ~static_cast<uintptr_t>(1023));
    uint8_t *sm_a = base;                                   // [7][16 KB] resident
    uint8_t *sm_b = base + kTcKBlocks * kTcTileBytes;       // [4][16 KB] ring
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm_b + kTcStages * kTcTileBytes);
    uint64_t *a_full = bars;                                // [k-blocks]
    uint64_t *a_empty = a_full + kTcKBlocks;                // [k-blocks]
    uint64_t *b_full = a_empty + kTcKBlocks;                // [stages]
    uint64_t *b_empty = b_full + kTcStages;                 // [stages]
    uint64_t *acc_full = b_empty + kTcStages;               // [2]
    uint64_t *acc_empty = acc_full + 2;                     // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = static_cast<int>((a.bins + kTcN - 1) / kTcN);
    const long long m_tiles = (a.rows + kTcM - 1) / kTcM;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < kTcKBlocks; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < kTcStages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * kTcN);         // 256 columns: two accumulator stages
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, it = 0;
            for (long long mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it) {
                // k-block kb of A' is free again when the MMAs of the previous row tile's last
                // bin tile have read it: the next row tile streams in under those MMAs
                for (int kb = 0; kb < kTcKBlocks; ++kb) {
                    mbar_wait(a_empty + kb, (it & 1) ^ 1);
                    mbar_expect_tx(a_full + kb, kTcTileBytes);
                    tma_load_2d(&map_a, a_full + kb, sm_a + kb * kTcTileBytes, kb * kTcKB,
                                static_cast<int>(mt * kTcM));
                }
                for (int j = 0; j < n_tiles; ++j)
                    for (int kb = 0; kb < kTcKBlocks; ++kb) {
                        mbar_wait(b_empty + stage, phase ^ 1);
                        mbar_expect_tx(b_full + stage, kTcTileBytes);
                        tma_load_2d(&map_b, b_full + stage, sm_b + stage * kTcTileBytes, kb * kTcKB,
                                    j * kTcN);
                        if (++stage == kTcStages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, it = 0, tile = 0;
            for (long long mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it)
            for (int j = 0; j < n_tiles; ++j, ++tile) {
                const uint32_t acc = tile & 1;
                mbar_wait(acc_empty + acc, ((tile >> 1) & 1) ^ 1);   // the epilogue has drained it
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kTcN;
                for (int kb = 0; kb < kTcKBlocks; ++kb) {
                    if (j == 0) mbar_wait(a_full + kb, it & 1);
                    mbar_wait(b_full + stage, phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sm_a + kb * kTcTileBytes);
                    const uint32_t b_addr = smem_u32(sm_b + stage * kTcTileBytes);
#pragma unroll
                    for (int k = 0; k < kTcKB / 16; ++k)     // K = 16 per MMA: 32 bytes along the row
                        umma_f16(tmem_d, umma_desc_sw128(a_addr + 32 * k),
                                 umma_desc_sw128(b_addr + 32 * k), kTcIdesc,
                                 (kb | k) != 0 ? 1u : 0u);
                    umma_commit(b_empty + stage);            // slot free when these MMAs are done
                    if (j == n_tiles - 1) umma_commit(a_empty + kb);   // A' k-block free as well
                    if (++stage == kTcStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(acc_full + acc);                 // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lanes 32 * (warp % 4) .. + 31 =====
        const int quarter = warp & 3;
        const float inv_s2 = static_cast<float>(-2.0 / (kTcScale * kTcScale));
        uint32_t tile = 0;
        for (long long mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        const long long row = mt * kTcM + 32 * quarter + lane;
        const bool row_ok = row < a.rows;
        const float nu = row_ok ? a.nu[row] : 0.f;
        const float mrow = row_ok ? a.mrow[row] : 0.f;
        float *out_row = a.d2a + row * a.ldf;
        for (int j = 0; j < n_tiles; ++j, ++tile) {
            const uint32_t acc = tile & 1;
            mbar_wait(acc_full + acc, (tile >> 1) & 1);
            tc_fence_after();
            // |x_v|^2 is padded with +inf up to ldf (tc_prep_kernel), so a column beyond the bins
            // comes out as +inf without a test per element: it never is a minimum, and its
            // store lands in the padding of d2a
            float tmin0 = INFINITY, tmin1 = INFINITY;
#pragma unroll 1
            for (int c = 0; c < kTcN / 32; ++c) {
                float v[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * quarter) << 16) +
                              (acc * kTcN + 32u * c), v);
                const int col0 = j * kTcN + 32 * c;
                const float4 *nv4 = reinterpret_cast<const float4 *>(a.nv + col0);
                float lo = INFINITY;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 n = __ldg(nv4 + (i >> 2));
                    v[i] = fmaf(v[i], inv_s2, nu + n.x);
                    v[i + 1] = fmaf(v[i + 1], inv_s2, nu + n.y);
                    v[i + 2] = fmaf(v[i + 2], inv_s2, nu + n.z);
                    v[i + 3] = fmaf(v[i + 3], inv_s2, nu + n.w);
                    lo = fminf(fminf(lo, fminf(v[i], v[i + 1])), fminf(v[i + 2], v[i + 3]));
                }
                if (c < 2) tmin0 = fminf(tmin0, lo); else tmin1 = fminf(tmin1, lo);
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4 *>(out_row + col0 + i) =
                            make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + acc);
            if (a.tilemin && row_ok) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const long long t64 = 2LL * j + h;
                    if (t64 < a.nbt64) {
                        const double lb = fmax(static_cast<double>(h ? tmin1 : tmin0) - static_cast<double>(mrow), 0.0);
                        a.tilemin[row * a.nbt64 + t64] =
                            static_cast<unsigned long long>(__double_as_longlong(lb));
                    }
                }
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * kTcN);
    }
}

// ---- operand preparation ----------------------------------------------------------------------
// profile of uniform base composition: every tetramer code 1/256; a canonical slot collects one
// code (the 16 palindromes) or two
__device__ uint8_t g_slot_codes[MCG_NPROF];

// One warp per row: x = p - c0, fp16 split of x * 2^12 into the three K segments, |x|^2 and |x|.
// B_SIDE selects the segment order ([hi | lo | hi] instead of [hi | hi | lo]).
template <bool B_SIDE>
__global__ void __launch_bounds__(256)
tc_prep_kernel(const double *__restrict__ prof, const int64_t *__restrict__ index, long long first,
               long long rows, __half *__restrict__ out, float *__restrict__ norm2,
               float *__restrict__ norm1, long long pad_to) {
    const long long r = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) {
        if (r < pad_to && lane == 0) norm2[r] = INFINITY;   // see the epilogue of tc_dist_kernel
        return;
    }
    const long long src = index ? index[first + r] : (first + r);
    const double *p = prof + src * MCG_NPROF;
    __half *o = out + r * kTcK;
    double acc = 0.0;
    // all of the lane's slots in flight before the first is used: the kernel waits for HBM
    // (ncu r02h: long-scoreboard stalls 16.7 per issue, 42 % of the DRAM peak)
    double pv[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int k = lane + 32 * i;
        pv[i] = k < MCG_NPROF ? __ldg(p + k) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int k = lane + 32 * i;
        if (k >= kTcSeg) break;
        __half hi = __float2half(0.f), lo = hi;
        if (k < MCG_NPROF) {
            const double x = pv[i] - static_cast<double>(g_slot_codes[k]) * (1.0 / 256.0);
            acc = fma(x, x, acc);
            const double xs = x * kTcScale;
            hi = __double2half(xs);
            lo = __double2half(xs - static_cast<double>(__half2float(hi)));
        }
        o[k] = hi;
        o[kTcSeg + k] = B_SIDE ? lo : hi;
        o[2 * kTcSeg + k] = B_SIDE ? hi : lo;
    }
    if (lane < kTcK - 3 * kTcSeg) o[3 * kTcSeg + lane] = __float2half(0.f);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (lane == 0) {
        norm2[r] = static_cast<float>(acc);
        norm1[r] = static_cast<float>(sqrt(acc) * (1.0 + 1e-6));   // rounded up
    }
}

// margin of a row against any bin: kTcRel |x_u| max|x_v| + kTcAbs (|x_u| + max|x_v|) + kTcFloor
constexpr float kTcRel = 4e-5f, kTcAbs = 4e-7f, kTcFloor = 1e-12f;

__global__ void tc_bmax_kernel(const float *__restrict__ norm1, long long bins, float *out) {
    float m = 0.f;
    for (long long i = threadIdx.x; i < bins; i += blockDim.x) m = fmaxf(m, norm1[i]);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    __shared__ float part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) m = fmaxf(m, part[w]);
        out[0] = m;
    }
}

__global__ void tc_margin_kernel(const float *__restrict__ norm1, const float *__restrict__ bmax,
                                 long long rows, float *__restrict__ mrow) {
    const long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (r >= rows) return;
    const float b = bmax[0], u = norm1[r];
    mrow[r] = (kTcRel * u * b + kTcAbs * (u + b) + kTcFloor) * 1.0001f;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows][448] fp16, row-major: boxes of 64 columns (128 bytes) x 128 rows, 128-byte swizzle;
// rows beyond `rows` read as zeros
int make_map(mcg_ctx *ctx, CUtensorMap *map, const __half *ptr, long long rows) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return mcg_fail(ctx, MCG_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kTcK), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kTcK) * 2};
    const cuuint32_t box[2] = {kTcKB, kTcM};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half *>(ptr), dims,
                           strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS)
        return mcg_fail(ctx, MCG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(rc));
    return MCG_OK;
}

int ensure_tc(mcg_ctx *ctx) {
    static bool ready[64] = {};
    if (ctx->device < 64 && ready[ctx->device]) return MCG_OK;
    uint8_t table[256], codes[MCG_NPROF] = {};
    mcg_kmer_table(table);
    for (int c = 0; c < 256; ++c) codes[table[c]]++;
    MCG_CUDA(ctx, cudaMemcpyToSymbolAsync(g_slot_codes, codes, sizeof codes, 0, cudaMemcpyHostToDevice,
                                          ctx->stream));
    MCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MCG_CUDA(ctx, cudaFuncSetAttribute(tc_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kTcSmem)));
    if (ctx->device < 64) ready[ctx->device] = true;
    return MCG_OK;
}

}  // namespace

long long mcg_tc_ldf(long long bins) { return (bins + kTcN - 1) / kTcN * kTcN; }

// fp16 operands, norms and margin tables of the B side (bin centroids); valid until the
// centroids change (the caller re-runs it per rule-C call: B x 896 bytes)
int mcg_k_tc_prep_bins(mcg_ctx *ctx, const ScoreSide &b) {
    MCG_TRY(ensure_tc(ctx));
    MCG_TRY(mcg_reserve(ctx, ctx->tc_b, sizeof(__half) * kTcK * b.n));
    // [|x_v|^2, padded with +inf to ldf] [|x_v| bins] [max |x_v|]
    const long long ldf = mcg_tc_ldf(b.n);
    MCG_TRY(mcg_reserve(ctx, ctx->tc_baux, sizeof(float) * (ldf + b.n + 4)));
    float *nv = ctx->tc_baux.as<float>();
    float *n1 = nv + ldf;
    tc_prep_kernel<true><<<static_cast<unsigned>((ldf * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        b.prof, b.index, b.first, b.n, ctx->tc_b.as<__half>(), nv, n1, ldf);
    MCG_KERNEL_CHECK(ctx);
    tc_bmax_kernel<<<1, 256, 0, ctx->stream>>>(n1, b.n, n1 + b.n);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

// d2a [rows padded][ldf] (fp32) of the rows of `a` against the prepared bins, the per-row margin
// (mrow_out, fp32 [rows]) and, when tilemin is given, the smallest LOWER bound of d^2 per row and
// 64-bin tile as ordered double bits (what bound_kernel reads).
int mcg_k_tc_dist(mcg_ctx *ctx, const ScoreSide &a, long long bins, float *d_d2a, long long ldf,
                  float *d_mrow, unsigned long long *d_tilemin, long long nbt64) {
    MCG_TRY(ensure_tc(ctx));
    if (a.n == 0) return MCG_OK;
    const long long tiles = (a.n + kTcM - 1) / kTcM;
    MCG_TRY(mcg_reserve(ctx, ctx->tc_a, sizeof(__half) * kTcK * tiles * kTcM));
    MCG_TRY(mcg_reserve(ctx, ctx->tc_aaux, sizeof(float) * 2 * a.n));
    float *nu = ctx->tc_aaux.as<float>();
    float *n1 = nu + a.n;
    tc_prep_kernel<false><<<static_cast<unsigned>((a.n * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        a.prof, a.index, a.first, a.n, ctx->tc_a.as<__half>(), nu, n1, 0);
    MCG_KERNEL_CHECK(ctx);
    const float *bn = ctx->tc_baux.as<float>();
    tc_margin_kernel<<<static_cast<unsigned>((a.n + 255) / 256), 256, 0, ctx->stream>>>(
        n1, bn + mcg_tc_ldf(bins) + bins, a.n, d_mrow);
    MCG_KERNEL_CHECK(ctx);
    CUtensorMap map_a, map_b;
    MCG_TRY(make_map(ctx, &map_a, ctx->tc_a.as<__half>(), a.n));
    MCG_TRY(make_map(ctx, &map_b, ctx->tc_b.as<__half>(), bins));
    TcArgs args;
    args.nu = nu;
    args.nv = bn;
    args.mrow = d_mrow;
    args.d2a = d_d2a;
    args.ldf = ldf;
    args.tilemin = d_tilemin;
    args.nbt64 = nbt64;
    args.rows = a.n;
    args.bins = bins;
    // persistent: one CTA per SM walks row tiles blockIdx.x, blockIdx.x + grid, ...
    const long long grid = tiles < ctx->sm_count ? tiles : ctx->sm_count;
    tc_dist_kernel<<<static_cast<unsigned>(grid), kTcThreads, kTcSmem, ctx->stream>>>(map_a, map_b,
                                                                                       args);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}
