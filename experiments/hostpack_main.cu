#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <chrono>
#include <cuda_runtime.h>
extern "C" void pack_avx512(const uint8_t*, uint64_t, uint8_t*);
extern "C" void pack_avx2(const uint8_t*, uint64_t, uint8_t*);
extern "C" int cpu_has_avx512(); extern "C" int cpu_has_avx2();
static double now(){ return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(){
    size_t n=1ull<<30; uint8_t *src,*dst,*d; 
    cudaHostAlloc(&src,n,cudaHostAllocDefault); cudaHostAlloc(&dst,n/4,cudaHostAllocDefault); cudaMalloc(&d,n+n/4);
    const char al[4]={'A','C','G','T'}; 
    { std::vector<std::thread> th; for(int t=0;t<8;++t) th.emplace_back([&,t]{ uint64_t s=1234567+t; for(size_t i=t*(n/8);i<(t+1)*(n/8);++i){ s=s*6364136223846793005ull+1442695040888963407ull; src[i]=al[(s>>33)&3]; }}); for(auto&x:th)x.join(); }
    memset(dst,0,n/4);
    printf("hw threads %u avx512 %d avx2 %d\n", std::thread::hardware_concurrency(), cpu_has_avx512(), cpu_has_avx2());
    auto pack=[&](int nt,size_t bytes,int use512){ std::vector<std::thread> th; size_t per=(bytes/64)/nt;
        for(int t=0;t<nt;++t) th.emplace_back([=]{ if(use512) pack_avx512(src+64*per*t,per,dst+16*per*t); else pack_avx2(src+64*per*t,per,dst+16*per*t); }); for(auto&x:th)x.join(); };
    int use512=cpu_has_avx512();
    for(int nt: {1,4,8,12,15,16,24,32}){ double best=1e9; for(int r=0;r<3;++r){ double t0=now(); pack(nt,n,use512); double dt=now()-t0; if(dt<best)best=dt;} printf("pack %2d threads: %.2f ms  %.1f GB/s\n",nt,best*1e3,n/best/1e9); }
    if(use512){ double best=1e9; for(int r=0;r<3;++r){ double t0=now(); pack(16,n,0); double dt=now()-t0; if(dt<best)best=dt;} printf("pack avx2 16 threads: %.2f ms %.1f GB/s\n",best*1e3,n/best/1e9);} 
    cudaStream_t s; cudaStreamCreate(&s);
    for(int r=0;r<3;++r){ double t0=now(); cudaMemcpyAsync(d,src,n,cudaMemcpyHostToDevice,s); cudaStreamSynchronize(s); double dt=now()-t0; printf("h2d 1 GiB: %.2f ms %.1f GB/s\n",dt*1e3,n/dt/1e9); }
    // concurrent: DMA of the first half while threads pack the second half
    for(int nt: {8,15,16}){ for(int r=0;r<2;++r){ double t0=now(); cudaMemcpyAsync(d,src,n/2,cudaMemcpyHostToDevice,s);
        std::vector<std::thread> th; size_t per=(n/2/64)/nt; double tp0=now();
        for(int t=0;t<nt;++t) th.emplace_back([=]{ const uint8_t* a=src+n/2+64*per*t; uint8_t* o=dst+16*per*t; if(use512) pack_avx512(a,per,o); else pack_avx2(a,per,o); }); for(auto&x:th)x.join(); double tp=now()-tp0;
        cudaStreamSynchronize(s); double tc=now()-t0; printf("concurrent %d thr: pack half %.2f ms (%.1f GB/s), dma half done at %.2f ms (%.1f GB/s)\n",nt,tp*1e3,n/2/tp/1e9,tc*1e3,n/2/tc/1e9);} }
    return 0; }
