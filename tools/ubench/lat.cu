// Micro-benchmark: dependent-chain latencies of shared-memory / warp ops (one warp, one SM)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, int iters, long long* cyc, int nwarps_busy) {
  __shared__ uint32_t s[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = (i * 7 + 3) & 1023;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
  uint32_t v = lane;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t a = sa + 4u * ((v & 1023) + (threadIdx.x >> 5) * 0);
      if (MODE == 0) {  // LDS chain
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
      } else if (MODE == 1) {  // ATOMS.AND chain (mask keeps value)
        asm volatile("atom.shared.and.b32 %0, [%1], %2;" : "=r"(v) : "r"(a), "r"(0xffffffffu) : "memory");
      } else if (MODE == 2) {  // ATOMS.ADD 0 chain
        asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(v) : "r"(a), "r"(0u) : "memory");
      } else if (MODE == 3) {  // SHFL chain
        v = __shfl_sync(0xffffffffu, v, (lane + 1) & 31);
      } else if (MODE == 4) {  // VOTE+POPC chain
        v = __popc(__ballot_sync(0xffffffffu, v & 1)) + lane;
      } else if (MODE == 5) {  // ALU chain (IMAD)
        v = v * 3 + 1;
      } else if (MODE == 6) {  // BAR chain
        __syncthreads();
      } else if (MODE == 7) {  // LOP3/shift chain
        v = ((v << 1) ^ (v >> 3)) + 1;
      } else if (MODE == 8) {  // STS then LDS same address (store-load)
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
      } else if (MODE == 9) {  // ATOMS.AND lane 0 only (others predicated off) chain via shfl
        uint32_t o = 0;
        if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(sa), "r"(1u) : "memory");
        v = __shfl_sync(0xffffffffu, o, 0);
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[threadIdx.x] = v;
}

template <int MODE>
void run(const char* name, int warps) {
  uint32_t* out; long long* cyc; long long h;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  const int iters = 1000;
  k<MODE><<<1, warps * 32>>>(out, iters, cyc, warps);
  k<MODE><<<1, warps * 32>>>(out, iters, cyc, warps);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps=%2d : %.1f cycles per dependent op\n", name, warps, (double)h / (iters * 8.0));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 32}) {
    run<0>("LDS", w);
    run<1>("ATOMS.AND (32 lanes)", w);
    run<2>("ATOMS.ADD (32 lanes)", w);
    run<9>("ATOMS.ADD lane0 + SHFL", w);
    run<3>("SHFL", w);
    run<4>("VOTE+POPC+IADD", w);
    run<5>("IMAD", w);
    run<7>("SHF,SHF,LOP3,IADD", w);
    run<6>("BAR.SYNC", w);
    run<8>("STS+LDS", w);
  }
  return 0;
}
