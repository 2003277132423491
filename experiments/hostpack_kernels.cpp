// scratch: host ASCII -> 2-bit pack (64 bases -> 16 bytes)
#include <immintrin.h>
#include <cstdint>
#include <cstring>
extern "C" __attribute__((target("avx512bw,avx512vl,bmi2")))
void pack_avx512(const uint8_t* src, uint64_t n64, uint8_t* dst) {
    const __m512i vC=_mm512_set1_epi8('C'), vG=_mm512_set1_epi8('G'), vT=_mm512_set1_epi8('T');
    for (uint64_t i=0;i<n64;++i){
        __m512i v=_mm512_loadu_si512(src+64*i);
        uint64_t mc=_mm512_cmpeq_epi8_mask(v,vC), mg=_mm512_cmpeq_epi8_mask(v,vG), mt=_mm512_cmpeq_epi8_mask(v,vT);
        uint64_t lo=mc|mt, hi=mg|mt;
        uint64_t w0=_pdep_u64(lo,0x5555555555555555ull)|_pdep_u64(hi,0xAAAAAAAAAAAAAAAAull);
        uint64_t w1=_pdep_u64(lo>>32,0x5555555555555555ull)|_pdep_u64(hi>>32,0xAAAAAAAAAAAAAAAAull);
        memcpy(dst+16*i,&w0,8); memcpy(dst+16*i+8,&w1,8);
    }
}
extern "C" __attribute__((target("avx2,bmi2")))
void pack_avx2(const uint8_t* src, uint64_t n64, uint8_t* dst) {
    const __m256i vC=_mm256_set1_epi8('C'), vG=_mm256_set1_epi8('G'), vT=_mm256_set1_epi8('T');
    for (uint64_t i=0;i<2*n64;++i){
        __m256i v=_mm256_loadu_si256((const __m256i*)(src+32*i));
        uint32_t mc=_mm256_movemask_epi8(_mm256_cmpeq_epi8(v,vC)), mg=_mm256_movemask_epi8(_mm256_cmpeq_epi8(v,vG)), mt=_mm256_movemask_epi8(_mm256_cmpeq_epi8(v,vT));
        uint64_t lo=mc|mt, hi=mg|mt;
        uint64_t w=_pdep_u64(lo,0x5555555555555555ull)|_pdep_u64(hi,0xAAAAAAAAAAAAAAAAull);
        memcpy(dst+8*i,&w,8);
    }
}
extern "C" int cpu_has_avx512() { return __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("bmi2"); }
extern "C" int cpu_has_avx2() { return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2"); }
