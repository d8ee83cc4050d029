// Micro-benchmark: shared-memory atomic throughput on one SM (cycles per warp-instruction)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atoms atoms.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, int iters, int activeLanes, int stride, long long* cyc) {
  __shared__ uint32_t s[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = 0xffffffffu;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool act = lane < activeLanes;
  uint32_t acc = 0;
  uint32_t idx = (warp * 257 + lane * stride) & 8191;
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t a = sa + 4u * ((idx + u * 33) & 8191);
      const uint32_t m = ~(1u << ((it + u) & 31));
      if (MODE == 0) {  // atom with return
        uint32_t old = 0;
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %3, 0;\n @p atom.shared.and.b32 %0, [%1], %2;\n}"
                     : "+r"(old) : "r"(a), "r"(m), "r"((uint32_t)act) : "memory");
        acc += old;
      } else if (MODE == 1) {  // red (no return)
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p red.shared.and.b32 [%0], %1;\n}"
                     :: "r"(a), "r"(m), "r"((uint32_t)act) : "memory");
      } else if (MODE == 2) {  // plain load + store
        uint32_t v = 0;
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p ld.shared.u32 %0, [%1];\n}"
                     : "+r"(v) : "r"(a), "r"((uint32_t)act) : "memory");
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p st.shared.u32 [%0], %1;\n}"
                     :: "r"(a), "r"(v & m), "r"((uint32_t)act) : "memory");
        acc += v;
      } else if (MODE == 3) {  // plain load only
        uint32_t v = 0;
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p ld.shared.u32 %0, [%1];\n}"
                     : "+r"(v) : "r"(a), "r"((uint32_t)act) : "memory");
        acc += v;
      } else if (MODE == 4) {  // atom.or
        uint32_t old = 0;
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %3, 0;\n @p atom.shared.or.b32 %0, [%1], %2;\n}"
                     : "+r"(old) : "r"(a), "r"(~m), "r"((uint32_t)act) : "memory");
        acc += old;
      } else if (MODE == 5) {  // atom.exch
        uint32_t old = 0;
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %3, 0;\n @p atom.shared.exch.b32 %0, [%1], %2;\n}"
                     : "+r"(old) : "r"(a), "r"(m), "r"((uint32_t)act) : "memory");
        acc += old;
      } else if (MODE == 6) {  // atom.add
        uint32_t old = 0;
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %3, 0;\n @p atom.shared.add.u32 %0, [%1], %2;\n}"
                     : "+r"(old) : "r"(a), "r"(m), "r"((uint32_t)act) : "memory");
        acc += old;
      }
    }
    idx = (idx + 7) & 8191;
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int warps, int activeLanes, int stride) {
  uint32_t* out; long long* cyc; long long h;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<MODE><<<1, warps * 32>>>(out, iters, activeLanes, stride, cyc);
  k<MODE><<<1, warps * 32>>>(out, iters, activeLanes, stride, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double perInstr = (double)h / (iters * 8.0 * warps);
  printf("%-10s warps=%2d lanes=%2d stride=%2d : %.2f cyc per warp-instr (SM-wide), %.2f cyc/lane\n", name, warps,
         activeLanes, stride, perInstr, perInstr / activeLanes);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int lanes : {32, 12, 4, 1}) {
    run<0>("atom.and", 16, lanes, 1);
    run<1>("red.and", 16, lanes, 1);
    run<4>("atom.or", 16, lanes, 1);
    run<5>("atom.exch", 16, lanes, 1);
    run<6>("atom.add", 16, lanes, 1);
    run<2>("ld+st", 16, lanes, 1);
    run<3>("ld", 16, lanes, 1);
  }
  run<0>("atom.and", 16, 12, 33);
  run<0>("atom.and", 32, 12, 1);
  run<0>("atom.and", 4, 12, 1);
  run<1>("red.and", 32, 12, 1);
  return 0;
}
