#include <immintrin.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <omp.h>
#include <cstring>
static void widen(const uint16_t* s, int32_t* d, size_t n){
  size_t i=0;
  const __m256i ffff=_mm256_set1_epi32(0xFFFF), inf=_mm256_set1_epi32(0x7fffffff);
  for(;i+8<=n;i+=8){
    __m256i v=_mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i*)(s+i)));
    __m256i m=_mm256_cmpeq_epi32(v,ffff);
    v=_mm256_blendv_epi8(v,inf,m);
    _mm256_stream_si256((__m256i*)(d+i),v);
  }
  for(;i<n;++i) d[i]= s[i]==0xFFFF?0x7fffffff:s[i];
}
int main(int argc,char**argv){
  size_t n=(size_t)512<<20; // 512M cells = 1GB in, 2GB out
  int T=argc>1?atoi(argv[1]):omp_get_max_threads();
  uint16_t* s=(uint16_t*)aligned_alloc(64,n*2); int32_t* d=(int32_t*)aligned_alloc(64,n*4);
  memset(s,1,n*2); memset(d,0,n*4);
  for(int r=0;r<3;r++){
  auto t0=std::chrono::steady_clock::now();
  #pragma omp parallel for num_threads(T) schedule(static)
  for(size_t c=0;c<n;c+=(1<<20)) widen(s+c,d+c,1<<20);
  double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
  printf("T=%d %.3f s  out %.1f GB/s\n",T,dt,n*4/dt/1e9);}
}
