// Is mma.sync.m8n8k4.f64 (DMMA) bit-identical to a chain of FP64 FMAs over k in ascending order?
// (round 2: the exact d^2 of a listed (contig, bin) pair has to come out with the bits
// dist2_kernel / row_norms_kernel produce; if DMMA is an FMA chain, one thread can redo a pair.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_fma_chain dmma_fma_chain.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double* A, const double* B, double* C, double* F, double* G, int K) {
    int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double* Ab = A + blockIdx.x * 8 * K; const double* Bb = B + blockIdx.x * 8 * K;
    double c0 = 0, c1 = 0;
    for (int k0 = 0; k0 < K; k0 += 4) {
        double a = Ab[g * K + k0 + t], b = Bb[g * K + k0 + t];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    C[blockIdx.x * 64 + g * 8 + 2 * t] = c0; C[blockIdx.x * 64 + g * 8 + 2 * t + 1] = c1;
    // chain of FMAs, ascending k, for the two outputs of this lane
    for (int h = 0; h < 2; ++h) {
        int i = g, j = 2 * t + h;
        double acc = 0;
        for (int kk = 0; kk < K; ++kk) acc = fma(Ab[i * K + kk], Bb[j * K + kk], acc);
        F[blockIdx.x * 64 + i * 8 + j] = acc;
        // per k-step of 4: exact products summed pairwise then added (another plausible datapath)
        double acc2 = 0;
        for (int k0 = 0; k0 < K; k0 += 4) {
            double p = 0;
            for (int q = 3; q >= 0; --q) p = fma(Ab[i * K + k0 + q], Bb[j * K + k0 + q], p);
            acc2 += p;
        }
        G[blockIdx.x * 64 + i * 8 + j] = acc2;
    }
}
int main() {
    const int K = 136, NB = 4096;
    size_t n = (size_t)NB * 8 * K;
    double *hA = (double*)malloc(n * 8), *hB = (double*)malloc(n * 8);
    srand(1);
    for (size_t i = 0; i < n; ++i) { hA[i] = (rand() % 100000 + 1) / 1.3e7; hB[i] = (rand() % 100000 + 1) / 0.9e7; }
    double *dA, *dB, *dC, *dF, *dG; cudaMalloc(&dA, n * 8); cudaMalloc(&dB, n * 8);
    cudaMalloc(&dC, NB * 64 * 8); cudaMalloc(&dF, NB * 64 * 8); cudaMalloc(&dG, NB * 64 * 8);
    cudaMemcpy(dA, hA, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, n * 8, cudaMemcpyHostToDevice);
    k<<<NB, 32>>>(dA, dB, dC, dF, dG, K);
    double *hC = (double*)malloc(NB * 64 * 8), *hF = (double*)malloc(NB * 64 * 8), *hG = (double*)malloc(NB * 64 * 8);
    cudaMemcpy(hC, dC, NB * 64 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hF, dF, NB * 64 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hG, dG, NB * 64 * 8, cudaMemcpyDeviceToHost);
    long same_chain = 0, same_group = 0;
    for (int i = 0; i < NB * 64; ++i) { same_chain += hC[i] == hF[i]; same_group += hC[i] == hG[i]; }
    printf("outputs %d: DMMA == ascending FMA chain %ld, == grouped-by-4 %ld, err=%s\n", NB * 64, same_chain, same_group, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
