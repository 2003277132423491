// Experiment: "prefix/suffix" tetramer counting.  A 3-mer m = (b1 b2 b3) names a row of eight
// byte counters per lane: four for the window (b0 m) and four for the window (m b4).  One
// LDS.64 + STS.64 serves two windows, i.e. one LSU instruction per window instead of two --
// the shipped kernel sits at the LSU issue floor (~0.55 instructions/clk/SM).
// 16 KB table per warp, 13 warps per SM; ILP = windows pairs in flight per warp (alias fix-up).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#ifndef NW
#define NW 13
#endif
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(v.x), "r"(v.y)); }
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

// bases b0..b4 at bits [OFF, OFF+10) of hi:lo; m = bits [OFF+2, OFF+8)
struct Upd { uint32_t addr, ilo, ihi; };
template <int OFF> __device__ __forceinline__ Upd upd(uint32_t lo, uint32_t hi, uint32_t lanebase) {
    const uint32_t z = OFF == 0 ? lo : __funnelshift_r(lo, hi, OFF);   // b0 at bits 0..1
    Upd u;
    uint32_t y; asm("mul.lo.u32 %0, %1, 64;" : "=r"(y) : "r"(z));      // m (bits 2..7) -> bits 8..13: rows of 256 B
    u.addr = and_or(y, 0x3F00u, lanebase);
    u.ilo = 1u << ((z & 3u) * 8u);                                      // prefix counter b0
    u.ihi = 1u << (((z >> 8) & 3u) * 8u);                               // suffix counter b4
    return u;
}
template <int OFF> __device__ __forceinline__ void rmw1(uint32_t lo, uint32_t hi, uint32_t lanebase) {
    const Upd u = upd<OFF>(lo, hi, lanebase);
    uint2 v = lds64(u.addr);
    v.x += u.ilo; v.y += u.ihi;
    sts64(u.addr, v);
}
template <int OFF> __device__ __forceinline__ void rmw2(uint32_t lo, uint32_t hi, uint32_t lanebase) {
    const Upd u1 = upd<OFF>(lo, hi, lanebase), u2 = upd<OFF + 4>(lo, hi, lanebase);
    uint2 v1 = lds64(u1.addr), v2 = lds64(u2.addr);
    v1.x += u1.ilo; v1.y += u1.ihi;
    v2.x += u2.ilo; v2.y += u2.ihi;
    if (u1.addr == u2.addr) { v2.x += u1.ilo; v2.y += u1.ihi; }
    sts64(u1.addr, v1);
    sts64(u2.addr, v2);
}

template <int ILP>
__global__ void __launch_bounds__(NW * 32, 1) scan_ps(const uint32_t *__restrict__ packed, long n_words, int words,
                                                       unsigned long long *counter, unsigned int *out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tab = ((smem_u32(smem_raw) + 16383u) & ~16383u) + warp * 16384u;
    const uint32_t lanebase = tab + lane * 8;
    for (int r = 0; r < 64; ++r) sts64(lanebase + r * 256, make_uint2(0, 0));
    __syncwarp();
    const long n_tiles = n_words / (32L * words);
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
                    rmw1<0>(w[k], w[k + 1], lanebase); rmw1<4>(w[k], w[k + 1], lanebase);
                    rmw1<8>(w[k], w[k + 1], lanebase); rmw1<12>(w[k], w[k + 1], lanebase);
                    rmw1<16>(w[k], w[k + 1], lanebase); rmw1<20>(w[k], w[k + 1], lanebase);
                    rmw1<24>(w[k], w[k + 1], lanebase); rmw1<28>(w[k], w[k + 1], lanebase);
                } else {
                    rmw2<0>(w[k], w[k + 1], lanebase); rmw2<8>(w[k], w[k + 1], lanebase);
                    rmw2<16>(w[k], w[k + 1], lanebase); rmw2<24>(w[k], w[k + 1], lanebase);
                }
            }
            cur = nxt;
        }
        __syncwarp();
    }
    unsigned int check = 0;
    for (int r = 0; r < 64; ++r) { uint2 v = lds64(lanebase + r * 256); check += (v.x + 3 * v.y) * (r + 1); }
    atomicAdd(out, check);
}

int main() {
    long n_words = 78L << 20;
    uint32_t *d; unsigned long long *ctr; unsigned int *out;
    cudaMalloc(&d, (n_words + 64) * 4); cudaMalloc(&ctr, 8); cudaMalloc(&out, 4);
    std::vector<uint32_t> h(n_words + 64);
    uint64_t s = 88172645463325252ull;
    for (auto &x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (uint32_t)(s >> 16); }
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    cudaFuncSetAttribute(scan_ps<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    cudaFuncSetAttribute(scan_ps<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    for (int ilp = 1; ilp <= 2; ++ilp)
        for (int words : {32, 64}) {
            float best = 1e9; unsigned int chk = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaMemset(ctr, 0, 8); cudaMemset(out, 0, 4);
                cudaEventRecord(e0);
                if (ilp == 1) scan_ps<1><<<148, NW * 32, 232448>>>(d, n_words, words, ctr, out);
                else scan_ps<2><<<148, NW * 32, 232448>>>(d, n_words, words, ctr, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
                cudaMemcpy(&chk, out, 4, cudaMemcpyDeviceToHost);
            }
            double bases = (double)(n_words / (32L * words)) * 32 * words * 16;
            printf("NW %d ilp %d words/lane %d: %.3f ms  %.2f Tbase/s  %.1f windows/clk/SM  check %u  err %s\n", NW, ilp, words, best,
                   bases / (best * 1e-3) / 1e12, bases / (best * 1e-3) / 148 / 1.965e9, chk, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
