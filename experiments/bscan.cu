// Scratch: (1) throughput of mma.sync m16n8k256 b1 and.popc on sm_100a,
// (2) prototype of the bit-plane tetramer scan on binary MMA, checked against a CPU count.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void bmma(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                     uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) bmma_peak(int *out, int iters, uint32_t seed) {
    int c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0;
    uint32_t a = seed * (threadIdx.x + 1), b = seed ^ threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) bmma(c[i], a, a ^ 1, a ^ 2, a ^ 3, b, b ^ 5);
    }
    int s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 0x7fffffff) out[0] = s;
}

// ---------------------------------------------------------------------------------------
// Prototype.  Data: block s = 256 positions = 8 x {L word, H word}; contigs start on
// block boundaries.  seg_owner[s] = contig of block s.  nwin[c] = windows, seg_first[c].
__global__ void __launch_bounds__(256) bscan_proto(const uint4 *__restrict__ data,  // 4 uint4 per block
                                                   const int *__restrict__ seg_owner,
                                                   const int *__restrict__ seg_first,
                                                   const int *__restrict__ nwin, long n_seg,
                                                   int strip, unsigned int *__restrict__ counts256) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const uint32_t cL = (g & 1) ? 0u : ~0u;         // L ^ cL = (L == yl)
    const uint32_t cH = (g & 2) ? 0u : ~0u;
    const uint32_t cX = (g & 4) ? 0u : ~0u;
    const uint2 *data2 = reinterpret_cast<const uint2 *>(data);

    for (long s0 = warp * strip; s0 < n_seg; s0 += nwarps * strip) {
        long s1 = s0 + strip < n_seg ? s0 + strip : n_seg;
        int cur = -1;
        int acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
        for (long s = s0; s < s1; ++s) {
            int c = __ldg(seg_owner + s);
            if (c != cur) {
                if (cur >= 0) {
                    unsigned int *dst = counts256 + (long)cur * 256;
                    atomicAdd(dst + g * 16 + 2 * t, acc0[0]);
                    atomicAdd(dst + g * 16 + 2 * t + 1, acc0[1]);
                    atomicAdd(dst + (g + 8) * 16 + 2 * t, acc0[2]);
                    atomicAdd(dst + (g + 8) * 16 + 2 * t + 1, acc0[3]);
                    atomicAdd(dst + g * 16 + 8 + 2 * t, acc1[0]);
                    atomicAdd(dst + g * 16 + 8 + 2 * t + 1, acc1[1]);
                    atomicAdd(dst + (g + 8) * 16 + 8 + 2 * t, acc1[2]);
                    atomicAdd(dst + (g + 8) * 16 + 8 + 2 * t + 1, acc1[3]);
                }
                for (int i = 0; i < 4; ++i) acc0[i] = acc1[i] = 0;
                cur = c;
            }
            uint4 w = __ldg(data + s * 4 + t);             // {L0,H0,L1,H1}: segs 2t, 2t+1
            uint2 nx = __ldg(data2 + s * 8 + 2 * t + 2);   // seg 2t+2 (first of next thread / block)
            uint32_t q0 = (w.y ^ cH) & (w.x ^ cL);
            uint32_t q1 = (w.w ^ cH) & (w.z ^ cL);
            uint32_t q2 = (nx.y ^ cH) & (nx.x ^ cL);
            uint32_t qs0 = __funnelshift_r(q0, q1, 1);
            uint32_t qs1 = __funnelshift_r(q1, q2, 1);
            uint32_t qs2 = q2 >> 1;                         // only bits 0..1 of D2 are used
            uint32_t v0 = (w.x ^ cX) & qs0, v1 = (w.z ^ cX) & qs1, v2 = (nx.x ^ cX) & qs2;
            uint32_t dl0 = ~w.y & v0, dh0 = w.y & v0;
            uint32_t dl1 = ~w.w & v1, dh1 = w.w & v1;
            uint32_t dl2 = ~nx.y & v2, dh2 = nx.y & v2;
            uint32_t bl0 = __funnelshift_r(dl0, dl1, 2), bh0 = __funnelshift_r(dh0, dh1, 2);
            uint32_t bl1 = __funnelshift_r(dl1, dl2, 2), bh1 = __funnelshift_r(dh1, dh2, 2);
            int wleft = nwin[c] - (int)(s - seg_first[c]) * 256;   // valid windows from this block on
            if (wleft < 256) {
                int n0 = wleft - 64 * t, n1 = n0 - 32;
                uint32_t m0 = n0 >= 32 ? ~0u : (n0 <= 0 ? 0u : ((1u << n0) - 1u));
                uint32_t m1 = n1 >= 32 ? ~0u : (n1 <= 0 ? 0u : ((1u << n1) - 1u));
                dl0 &= m0; dh0 &= m0; dl1 &= m1; dh1 &= m1;
            }
            bmma(acc0, dl0, dh0, dl1, dh1, bl0, bl1);
            bmma(acc1, dl0, dh0, dl1, dh1, bh0, bh1);
        }
        if (cur >= 0) {
            unsigned int *dst = counts256 + (long)cur * 256;
            atomicAdd(dst + g * 16 + 2 * t, acc0[0]);
            atomicAdd(dst + g * 16 + 2 * t + 1, acc0[1]);
            atomicAdd(dst + (g + 8) * 16 + 2 * t, acc0[2]);
            atomicAdd(dst + (g + 8) * 16 + 2 * t + 1, acc0[3]);
            atomicAdd(dst + g * 16 + 8 + 2 * t, acc1[0]);
            atomicAdd(dst + g * 16 + 8 + 2 * t + 1, acc1[1]);
            atomicAdd(dst + (g + 8) * 16 + 8 + 2 * t, acc1[2]);
            atomicAdd(dst + (g + 8) * 16 + 8 + 2 * t + 1, acc1[3]);
        }
    }
}

static uint64_t rng_state = 88172645463325252ull;
static inline uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 16); }

int main(int argc, char **argv) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    int *dout; cudaMalloc(&dout, 64);
    for (int wpb = 4; wpb <= 32; wpb *= 2) {
        int blocks = 148 * (32 / wpb) * 2, iters = 2048;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); bmma_peak<<<blocks, wpb * 32>>>(dout, iters, 12345u + rep); cudaEventRecord(e1);
            cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        double mmas = 8.0 * iters * (double)blocks * wpb;
        printf("bmma m16n8k256: %d warps/block %d blocks: %.3f ms, %.1f Gmma/s, %.3f mma/clk/SM @1.965GHz, %.1f Tbit-op/s\n",
               wpb, blocks, ms, mmas / ms / 1e6, mmas / (ms * 1e-3) / 148 / 1.965e9, mmas * 65536 / (ms * 1e-3) / 1e12);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));

    // ---- prototype ---------------------------------------------------------------------
    long n_contigs = argc > 1 ? atol(argv[1]) : 100000;
    int lmin = 1000, lmax = 24000;
    std::vector<int> len(n_contigs), segfirst(n_contigs), nwin(n_contigs);
    long nseg = 0, nbases = 0;
    for (long i = 0; i < n_contigs; ++i) {
        int L = lmin + rnd() % (lmax - lmin);
        if (i < 600) L = (int)i;        // edge lengths 0..599
        len[i] = L; segfirst[i] = (int)nseg; nwin[i] = L > 3 ? L - 3 : 0;
        nseg += (L + 255) / 256; nbases += L;
    }
    std::vector<uint32_t> data((nseg + 1) * 16, 0);
    std::vector<int> owner(nseg);
    std::vector<uint8_t> check_bases;   // bases of the first 2000 contigs for the CPU check
    std::vector<long> check_off;
    const long ncheck = 2000 < n_contigs ? 2000 : n_contigs;
    for (long i = 0; i < n_contigs; ++i) {
        long blk = segfirst[i];
        if (i < ncheck) check_off.push_back((long)check_bases.size());
        for (int p = 0; p < len[i]; ++p) {
            uint32_t b = rnd() & 3;
            if (i < ncheck) check_bases.push_back((uint8_t)b);
            long word = blk * 16 + (p / 32) * 2;
            data[word] |= (b & 1) << (p & 31);
            data[word + 1] |= (b >> 1) << (p & 31);
        }
        for (int k = 0; k < (len[i] + 255) / 256; ++k) owner[blk + k] = (int)i;
    }
    check_off.push_back((long)check_bases.size());
    uint32_t *d_data; int *d_owner, *d_first, *d_nwin; unsigned int *d_counts;
    cudaMalloc(&d_data, data.size() * 4); cudaMalloc(&d_owner, nseg * 4); cudaMalloc(&d_first, n_contigs * 4);
    cudaMalloc(&d_nwin, n_contigs * 4); cudaMalloc(&d_counts, n_contigs * 256 * 4);
    cudaMemcpy(d_data, data.data(), data.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_owner, owner.data(), nseg * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_first, segfirst.data(), n_contigs * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_nwin, nwin.data(), n_contigs * 4, cudaMemcpyHostToDevice);
    printf("contigs %ld, bases %ld, blocks %ld (%.1f MB)\n", n_contigs, nbases, nseg, nseg * 64 / 1e6);
    for (int wpb = 8; wpb <= 32; wpb *= 2)
        for (int strip = 16; strip <= 64; strip *= 2) {
            int blocks = 148 * (64 / wpb);
            float best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                cudaMemsetAsync(d_counts, 0, n_contigs * 256 * 4);
                cudaEventRecord(e0);
                bscan_proto<<<blocks, wpb * 32>>>((const uint4 *)d_data, d_owner, d_first, d_nwin, nseg, strip, d_counts);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("proto wpb=%d strip=%d: %.3f ms, %.2f Tbase/s, %.0f GB/s\n", wpb, strip, best,
                   nbases / (best * 1e-3) / 1e12, nseg * 64.0 / (best * 1e-3) / 1e9);
        }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    std::vector<unsigned int> got(ncheck * 256);
    cudaMemcpy(got.data(), d_counts, ncheck * 256 * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (long i = 0; i < ncheck; ++i) {
        unsigned int ref[256] = {0};
        const uint8_t *b = check_bases.data() + check_off[i];
        for (int p = 0; p + 3 < len[i]; ++p) ref[(b[p] << 6) | (b[p + 1] << 4) | (b[p + 2] << 2) | b[p + 3]]++;
        for (int k = 0; k < 256; ++k) if (ref[k] != got[i * 256 + k]) { if (bad < 5) printf("mismatch contig %ld len %d code %d ref %u got %u\n", i, len[i], k, ref[k], got[i * 256 + k]); ++bad; }
    }
    printf("check: %ld mismatches over %ld contigs\n", bad, ncheck);
    return bad != 0;
}
