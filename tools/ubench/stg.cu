// Micro-benchmark: scattered 4-byte global stores, cycles per warp-instruction (per SM, all SMs busy)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(int32_t* out, size_t strideWords, int iters, int activeLanes, long long* cyc, size_t span) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t* base = out + (size_t)blockIdx.x * span;
  size_t idx = (size_t)warp * 40961 + (size_t)lane * strideWords;
  long long t0 = clock64();
  if (lane < activeLanes) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        base[(idx + (size_t)u * 12347) % span] = it;
      }
      idx += 98765;
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  const size_t span = 1 << 20;  // 4 MB per CTA
  int32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * span * 4); cudaMalloc(&cyc, 148 * 8);
  for (int warps : {8, 32}) for (int lanes : {32, 12, 1}) for (size_t stride : {(size_t)1, (size_t)8, (size_t)1024}) {
    const int iters = 500;
    k<<<148, warps * 32>>>(out, stride, iters, lanes, cyc, span);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<148, warps * 32>>>(out, stride, iters, lanes, cyc, span);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 8 * warps;
    printf("warps=%2d lanes=%2d strideWords=%4zu : %.2f cyc per STG instr per SM (issue side), kernel %.3f ms => %.2f cyc/instr wall, %.2f cyc/sector\n",
           warps, lanes, stride, h[0] / n, ms, ms * 1e-3 * 1.92e9 / n,
           ms * 1e-3 * 1.92e9 / n / (stride >= 8 ? lanes : (lanes + 7) / 8));
  }
  return 0;
}
