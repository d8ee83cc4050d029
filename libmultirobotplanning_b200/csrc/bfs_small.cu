// bfs_small.cu — distance fields on maps that fit one 32x32 tile (8x8, 32x32).
//
// Replaces ShortestPathHeuristic (example/shortest_path_heuristic.hpp:12-62 of
// the reference: Floyd–Warshall over every cell) by one BFS per goal, the
// layout of the reference's disabled computeHeuristic (example/cbs.cpp:445-557).
//
// One warp per (map, goal) job.  The whole map lives in registers: lane y holds
// row y of the free mask as one 32-bit word.  A BFS level is
//     cand = (f<<1 | f>>1 | shfl_up(f) | shfl_down(f)) & open
// i.e. two shifts for the horizontal neighbours and two warp shuffles for the
// vertical ones; the wavefront ends when __any_sync(cand) is false.  Distances
// are staged in a padded shared-memory tile and written out coalesced.
#include "common.cuh"

namespace mrp {

constexpr int kSmallWarps = 8;           // warps (jobs) per CTA
constexpr int kPad = kTile + 1;          // bank-conflict-free row stride

__global__ void __launch_bounds__(kSmallWarps * 32)
bfs_small_kernel(const uint32_t* __restrict__ rows,   // [n_maps][32]
                 const int32_t* __restrict__ dims,    // [n_maps][2]
                 const int4* __restrict__ jobs,       // map, goal cell, off lo/hi
                 int n_jobs, int32_t* __restrict__ out) {
  __shared__ int32_t sdist[kSmallWarps][kTile * kPad];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int32_t* sd = sdist[warp];
  for (int job = blockIdx.x * kSmallWarps + warp; job < n_jobs;
       job += gridDim.x * kSmallWarps) {
    const int4 jb = jobs[job];
    const int map = jb.x, goal = jb.y;
    const int dimx = dims[2 * map], dimy = dims[2 * map + 1];
    int32_t* o = out + (((long long)jb.w << 32) | (unsigned)jb.z);
    const uint32_t freeRow = rows[map * kTile + lane];
    const int gx = goal % dimx, gy = goal / dimx;

#pragma unroll
    for (int k = 0; k < kPad; ++k) sd[k * 32 + lane] = MRP_INF;
    __syncwarp();

    uint32_t open = freeRow;
    uint32_t f = (lane == gy) ? ((1u << gx) & freeRow) : 0u;
    open &= ~f;
    if (lane == gy) sd[gy * kPad + gx] = 0;  // d[v][v] = 0 even on an obstacle
    int level = 0;
    while (true) {
      uint32_t up = __shfl_up_sync(0xffffffffu, f, 1);
      uint32_t dn = __shfl_down_sync(0xffffffffu, f, 1);
      if (lane == 0) up = 0;
      if (lane == 31) dn = 0;
      uint32_t cand = ((f << 1) | (f >> 1) | up | dn) & open;
      if (!__any_sync(0xffffffffu, cand)) break;
      ++level;
      open &= ~cand;
      f = cand;
      uint32_t m = cand;
      while (m) {
        int b = __ffs(m) - 1;
        m &= m - 1;
        sd[lane * kPad + b] = level;
      }
    }
    __syncwarp();
    const int cells = dimx * dimy;
    if ((dimx & (dimx - 1)) == 0) {
      const int sh = __ffs(dimx) - 1;
      for (int idx = lane; idx < cells; idx += 32)
        o[idx] = sd[(idx >> sh) * kPad + (idx & (dimx - 1))];
    } else {
      for (int idx = lane; idx < cells; idx += 32)
        o[idx] = sd[(idx / dimx) * kPad + (idx % dimx)];
    }
    __syncwarp();
  }
}

int launchBfsSmall(const uint32_t* d_rows, const int32_t* d_dims,
                   const int4* d_jobs, int n_jobs, int32_t* d_out,
                   cudaStream_t st) {
  if (n_jobs <= 0) return 0;
  int blocks = (n_jobs + kSmallWarps - 1) / kSmallWarps;
  const int maxBlocks = ctx().smCount * 8;
  if (blocks > maxBlocks) blocks = maxBlocks;
  bfs_small_kernel<<<blocks, kSmallWarps * 32, 0, st>>>(d_rows, d_dims, d_jobs,
                                                        n_jobs, d_out);
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mrp
