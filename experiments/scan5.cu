// Scratch: throughput model of a 5-mer / stride-2 private byte-counter scan (two tetramer
// windows per shared-memory read-modify-write), 32 KB table per warp, 7 warps per SM.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

#ifndef ALIGNED
#define ALIGNED 1
#endif
template <int OFF> __device__ __forceinline__ uint32_t addr5(uint32_t lo, uint32_t hi, uint32_t lanebase) {
    uint32_t z = OFF == 0 ? lo : __funnelshift_r(lo, hi, OFF);
    uint32_t y; asm("mul.lo.u32 %0, %1, 32;" : "=r"(y) : "r"(z));
#if ALIGNED
    return and_or(z, 3u, and_or(y, 0x7F80u, lanebase));
#else
    return and_or(z, 3u, y & 0x7F80u) + lanebase;
#endif
}
template <int OFF> __device__ __forceinline__ void rmw1(uint32_t lo, uint32_t hi, uint32_t lanebase, uint32_t one) {
    uint32_t a = addr5<OFF>(lo, hi, lanebase);
    uint32_t v = lds_u8(a);
    asm("mad.lo.u32 %0, %0, %1, %1;" : "+r"(v) : "r"(one));
    sts_u8(a, v);
}
template <int OFF> __device__ __forceinline__ void rmw2(uint32_t lo, uint32_t hi, uint32_t lanebase, uint32_t one) {
    uint32_t a1 = addr5<OFF>(lo, hi, lanebase), a2 = addr5<OFF + 4>(lo, hi, lanebase);
    uint32_t v1 = lds_u8(a1), v2 = lds_u8(a2);
    asm("mad.lo.u32 %0, %0, %1, %1;" : "+r"(v1) : "r"(one));
    asm("mad.lo.u32 %0, %0, %1, %1;" : "+r"(v2) : "r"(one));
    if (a1 == a2) v2 += 1;
    sts_u8(a1, v1);
    sts_u8(a2, v2);
}

// each lane: `words` consecutive words (16 bases each); tile = 32 lanes
#ifndef NW
#define NW 6
#endif
template <int ILP>
__global__ void __launch_bounds__(NW * 32, 1) scan5(const uint32_t *__restrict__ packed, long n_words, int words,
                                                 unsigned long long *counter, unsigned int *out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tab = (ALIGNED ? ((smem_u32(smem_raw) + 32767u) & ~32767u) : ((smem_u32(smem_raw) + 127u) & ~127u)) + warp * 32768u;
    const uint32_t lanebase = tab + lane * 4;
    const uint32_t one = n_words > 0 ? 1u : 0u;
    for (int r = 0; r < 256; ++r) asm volatile("st.shared.u32 [%0], %1;" ::"r"(lanebase + r * 128), "r"(0u));
    __syncwarp();
    const long n_tiles = n_words / (32L * words);
    unsigned int check = 0;
    for (;;) {
        unsigned long long tile = 0;
        if (lane == 0) tile = atomicAdd(counter, 1ULL);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if ((long)tile >= n_tiles) break;
        const uint32_t *wp = packed + ((long)tile * 32 + lane) * words;
        uint4 cur = __ldg((const uint4 *)wp);
        for (int g = 0; g < words / 4; ++g) {
            uint4 nxt = __ldg((const uint4 *)(wp + 4 * g + 4));
            uint32_t w[5] = {cur.x, cur.y, cur.z, cur.w, nxt.x};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (ILP == 1) {
                    rmw1<0>(w[k], w[k + 1], lanebase, one); rmw1<4>(w[k], w[k + 1], lanebase, one);
                    rmw1<8>(w[k], w[k + 1], lanebase, one); rmw1<12>(w[k], w[k + 1], lanebase, one);
                    rmw1<16>(w[k], w[k + 1], lanebase, one); rmw1<20>(w[k], w[k + 1], lanebase, one);
                    rmw1<24>(w[k], w[k + 1], lanebase, one); rmw1<28>(w[k], w[k + 1], lanebase, one);
                } else {
                    rmw2<0>(w[k], w[k + 1], lanebase, one); rmw2<8>(w[k], w[k + 1], lanebase, one);
                    rmw2<16>(w[k], w[k + 1], lanebase, one); rmw2<24>(w[k], w[k + 1], lanebase, one);
                }
            }
            cur = nxt;
        }
        // stand-in for the flush: every 255 steps the table would be summed and cleared
        __syncwarp();
    }
    for (int r = 0; r < 256; ++r) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(lanebase + r * 128)); check += v * (r + 1); }
    atomicAdd(out, check);
}

int main() {
    long n_words = 78L << 20;   // 1.31 Gbases
    uint32_t *d; unsigned long long *ctr; unsigned int *out;
    cudaMalloc(&d, (n_words + 64) * 4); cudaMalloc(&ctr, 8); cudaMalloc(&out, 4);
    std::vector<uint32_t> h(n_words + 64);
    uint64_t s = 88172645463325252ull;
    for (auto &x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (uint32_t)(s >> 16); }
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    const int smem = 7 * 32768 + 2048;   // alignment slack: the dynamic window starts ~1 KB in
    cudaFuncSetAttribute(scan5<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    cudaFuncSetAttribute(scan5<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    (void)smem;
    for (int ilp = 1; ilp <= 2; ++ilp)
        for (int words : {32, 64}) {
            float best = 1e9; unsigned int chk = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaMemset(ctr, 0, 8); cudaMemset(out, 0, 4);
                cudaEventRecord(e0);
                if (ilp == 1) scan5<1><<<148, NW * 32, 232448>>>(d, n_words, words, ctr, out);
                else scan5<2><<<148, NW * 32, 232448>>>(d, n_words, words, ctr, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
                cudaMemcpy(&chk, out, 4, cudaMemcpyDeviceToHost);
            }
            double bases = (double)(n_words / (32L * words)) * 32 * words * 16;
            printf("ilp %d words/lane %d: %.3f ms  %.2f Tbase/s  %.1f windows/clk/SM  check %u  err %s\n", ilp, words, best,
                   bases / (best * 1e-3) / 1e12, bases / (best * 1e-3) / 148 / 1.965e9, chk, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
