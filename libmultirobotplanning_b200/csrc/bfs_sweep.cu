// bfs_sweep.cu — distance fields by DETOUR LEVELS on maps up to 1024 columns
// (the headline case: the synthetic 1024x1024 map, config C5).
//
// Replaces ShortestPathHeuristic (example/shortest_path_heuristic.hpp:12-62 of
// the reference) by one BFS distance field per goal, the layout of the
// reference's disabled computeHeuristic (example/cbs.cpp:445-557).
//
// Idea.  On a 4-connected grid every move changes the Manhattan distance M to
// the goal by +1 ("outward") or -1 ("inward"), so the BFS distance of a cell
// is d = M + 2*delta, where delta is the smallest number of inward moves on
// any path from the goal.  A BFS by d needs ~1500 tiny levels on the C5 map;
// a Dijkstra by delta (outward moves cost 0, inward moves cost 1) needs ~80,
// and the cells of one delta level are the closure of its seeds under outward
// moves, which is a regular sweep over the rows away from the goal row:
//
//   S_k[y] = fill_x( ( in_x(S_{k-1}[y]) | S_{k-1}[y+1] | S_k[y-1] ) & open[y] )
//
// (upper half, rows y >= gy; the lower half is the mirror image; the goal row
// takes the inward moves of both halves).  in_x shifts a row mask one cell
// towards the goal column, fill_x floods it away from the goal column through
// open cells: with a row as 32 words in the 32 lanes of a warp, the flood is
// the carry chain of one integer addition per word (a seed added to a run of
// ones ripples to the end of the run) plus a carry resolution across the
// lanes by the same trick on two ballots.  Level k needs level k-1 of the row
// one step further out, so step tau of the sweep handles row r = tau - 2k of
// every level k at once: a systolic pipeline over (level, half) pairs, one
// __syncthreads per step, ~maxrow + 2*levels steps per goal instead of ~1500
// BFS levels, and no atomics.
//
// One CTA per goal (persistent over goals):
//   * dense levels 0..kDense-1: one warp per (level, half), always running;
//     their cells are collected as bit planes of delta in a ring of rows;
//   * two writer warps (one per half) turn a row that left the last dense
//     level into int32 distances (M + 2*delta, MRP_INF for obstacles and for
//     cells not reached yet) and write it as whole 16-byte stores: every
//     sector of the field is written in full exactly once here;
//   * sparse levels kDense..kLmax-1: the remaining warps; a (level, half)
//     whose three input rows are empty is skipped, the others store their few
//     cells directly into the row the writer put out a few steps earlier (the
//     row is still in L2);
//   * a goal whose field needs more than kLmax levels (mazes) is handed to
//     the queue kernel (bfs_queue.cu).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace mrp {

constexpr int kSwLmax = 96;    // detour levels per goal (C5: max 83 over 512 goals)
constexpr int kSwDense = 8;    // levels with a warp of their own per half
constexpr int kSwRO = 200;     // rows of the open ring per half (>= 2*kSwLmax - 1)
constexpr int kSwRP = 20;      // rows of the plane ring per half (>= 2*kSwDense + 1)
constexpr int kSwPlanes = 3;   // bits of a dense level
constexpr int kSwThreads = 1024;
constexpr int kSwWarps = kSwThreads / 32;
constexpr int kSwSparseWarps = kSwWarps - 2 * kSwDense - 2;
static_assert((1 << kSwPlanes) >= kSwDense, "planes");
static_assert(kSwRO >= 2 * kSwLmax - 1 && kSwRP >= 2 * kSwDense + 1, "rings");

struct BfsSweepParams {
  const uint32_t* __restrict__ rowbits;  // bordered free mask (see mrp_map_s)
  const int32_t* __restrict__ goals;
  int32_t* __restrict__ out;
  uint32_t* ws;  // [0] goal counter, [2] number of goals handed over, [64..] their indices
  int n_goals, dimx, dimy, WPR;
  int dbg;  // MRP_SWEEP_DBG (timing experiments only): bit 0 = no sparse levels
};

#ifdef MRP_SWEEP_TIMING
// busy cycles per warp of CTA 0, summed over its goals; [32] = steps, [33] = goals (tools/sweep_timing.py)
__device__ unsigned long long g_swBusy[40];
extern "C" int mrp_debug_sweep_busy(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_swBusy, sizeof(unsigned long long) * 40);
}
#endif

constexpr uint32_t kFull = 0xffffffffu;

// word `lane` of row y of the free mask in natural layout (bit b = cell 32*lane + b)
__device__ __forceinline__ uint32_t sweepFreeWord(const uint32_t* __restrict__ rowbits, int WPR, int dimy, int y,
                                                  int lane) {
  if (y < 0 || y >= dimy) return 0u;
  const uint32_t* row = rowbits + (size_t)(y + 1) * WPR;
  const uint32_t lo = lane < WPR ? __ldg(row + lane) : 0u;
  const uint32_t hi = lane + 1 < WPR ? __ldg(row + lane + 1) : 0u;
  return __funnelshift_r(lo, hi, 1);
}

__device__ __forceinline__ void addCarry(uint32_t a, uint32_t b, uint32_t& sum, uint32_t& carry) {
  asm("{\n\tadd.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;\n\t}" : "=r"(sum), "=r"(carry) : "r"(a), "r"(b));
}

// carries into the lanes: c[j] = g[j-1] | (p[j-1] & c[j-1]), as one addition
__device__ __forceinline__ uint32_t resolveCarries(uint32_t g, uint32_t p) {
  const uint32_t t = g << 1, m2 = (p << 1) | t;
  return (((m2 + t) ^ m2) & m2) | t;
}

// One (level, row) step.  sp = S_{k-1}[row], spu = the inward vertical sources
// (S_{k-1} of the next row out), npv = S_k of the previous row (outward
// vertical), op = open cells of the row.  mR / mL: cells right / left of the
// goal column, cb: the goal column.
__device__ __forceinline__ uint32_t sweepRowStep(uint32_t sp, uint32_t spu, uint32_t npv, uint32_t op,
                                                 uint32_t mR, uint32_t mL, uint32_t cb, int lane) {
  const uint32_t spR = sp & mR, spL = sp & mL;
  uint32_t hi = __shfl_down_sync(kFull, spR, 1);
  uint32_t lo = __shfl_up_sync(kFull, spL, 1);
  if (lane == 31) hi = 0u;
  if (lane == 0) lo = 0u;
  const uint32_t seed = (__funnelshift_r(spR, hi, 1) | __funnelshift_l(lo, spL, 1) | spu | npv) & op;
  if (__ballot_sync(kFull, seed != 0u) == 0u) return 0u;
  const uint32_t sc = seed & cb;
  const uint32_t mU = (op & mR) | sc, sU = seed & (mR | cb);
  const uint32_t mD = __brev((op & mL) | sc), sD = __brev(seed & (mL | cb));
  uint32_t sumU, cU, sumD, cD;
  addCarry(mU, sU, sumU, cU);
  addCarry(mD, sD, sumD, cD);
  const uint32_t gU = __ballot_sync(kFull, cU != 0u), gD = __ballot_sync(kFull, cD != 0u);
  if (gU | gD) {
    const uint32_t qU = __ballot_sync(kFull, sumU == kFull), qD = __ballot_sync(kFull, sumD == kFull);
    const uint32_t u = resolveCarries(gU, qU), d = resolveCarries(__brev(gD), __brev(qD));
    sumU += (u >> lane) & 1u;
    sumD += (d >> (31 - lane)) & 1u;
  }
  return (((sumU ^ mU) & mU) | sU) | __brev(((sumD ^ mD) & mD) | sD);
}

struct SweepGoal {
  int gx, gy, dimx, dimy;
  uint32_t mR, mL, cb;
  int32_t* out;
};

// stores the cells of `nw` (row y, offset r, level k) one word at a time, lane = cell
__device__ __forceinline__ void sweepStoreCells(const SweepGoal& g, uint32_t nzb, uint32_t nw, int y, int r, int k,
                                                int lane) {
  int32_t* row = g.out + (size_t)y * g.dimx;
  while (nzb) {
    const int j = __ffs(nzb) - 1;
    nzb &= nzb - 1u;
    const uint32_t w = __shfl_sync(kFull, nw, j);
    if ((w >> lane) & 1u) {
      const int x = 32 * j + lane;
      row[x] = abs(x - g.gx) + r + 2 * k;
    }
  }
}

template <bool kVec>
__global__ void __launch_bounds__(kSwThreads, 1) bfs_sweep_kernel(BfsSweepParams p) {
  extern __shared__ uint32_t smem[];
  uint32_t* const S = smem;                                    // [kSwLmax][2][3][32]
  uint32_t* const ring = S + kSwLmax * 2 * 3 * 32;             // [2][kSwRO][32]
  uint32_t* const planes = ring + 2 * kSwRO * 32;              // [2][kSwRP][kSwPlanes][32]
  uint32_t* const nz = planes + 2 * kSwRP * kSwPlanes * 32;    // [kSwLmax][2][4]
  __shared__ int sGoal, sHandOver;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dimx = p.dimx, dimy = p.dimy;

  while (true) {
    __syncthreads();
    if (tid == 0) {
      sGoal = (int)atomicAdd(p.ws, 1u);
      sHandOver = 0;
    }
    for (int i = tid; i < kSwLmax * 2 * 3 * 32; i += kSwThreads) S[i] = 0u;
    for (int i = tid; i < 2 * kSwRP * kSwPlanes * 32; i += kSwThreads) planes[i] = 0u;
    for (int i = tid; i < kSwLmax * 2 * 4; i += kSwThreads) nz[i] = 0u;
    __syncthreads();
    const int gidx = sGoal;
    if (gidx >= p.n_goals) break;
    const int goal = p.goals[gidx];
    SweepGoal g;
    g.dimx = dimx;
    g.dimy = dimy;
    g.gy = goal / dimx;
    g.gx = goal - g.gy * dimx;
    g.out = p.out + (size_t)gidx * dimx * dimy;
    const int gx = g.gx, gy = g.gy;
    const int jg = gx >> 5, bg = gx & 31;
    g.mR = lane > jg ? kFull : (lane == jg ? (bg == 31 ? 0u : kFull << (bg + 1)) : 0u);
    g.mL = lane < jg ? kFull : (lane == jg ? (1u << bg) - 1u : 0u);
    g.cb = lane == jg ? 1u << bg : 0u;
    const int maxr = max(gy, dimy - 1 - gy);
    const int writerEnd = maxr + 2 * kSwDense;  // the writer puts out row maxr at step writerEnd - 1
    const bool goalFree = (sweepFreeWord(p.rowbits, p.WPR, dimy, gy, jg) >> bg) & 1u;

    // ---- roles ----
    const bool isDense = warp < 2 * kSwDense;
    const bool isWriter = !isDense && warp < 2 * kSwDense + 2;
    const int dk = warp >> 1, dh = warp & 1;          // dense: level, half
    const int wh = warp - 2 * kSwDense;                // writer: half
    const int sw = warp - 2 * kSwDense - 2;            // sparse warp index
    uint32_t pf = 0u;                                  // prefetched free word (level 0, writer)
    if (isDense && dk == 0) pf = dh == 0 ? sweepFreeWord(p.rowbits, p.WPR, dimy, gy, lane) : 0u;

    int idle = 0;
#ifdef MRP_SWEEP_TIMING
    long long busy = 0;
    const long long tGoal = clock64();
#endif
    int tau;
    for (tau = 0;; ++tau) {
      bool act = false;
#ifdef MRP_SWEEP_TIMING
      const long long t0 = clock64();
#endif
      if (isDense) {
        const int k = dk, h = dh, r = tau - 2 * k;
        if (r >= 0 && !(h == 1 && r == 0)) {
          const int y = h ? gy - r : gy + r;
          const int slot = r % 3;
          uint32_t* const Sk = S + (k * 2 + h) * 3 * 32;
          if (y >= 0 && y < dimy) {
            uint32_t op, sp = 0u, spu = 0u, npv = 0u;
            uint32_t* const ringRow = ring + (h * kSwRO + r % kSwRO) * 32;
            if (k == 0) {
              op = pf;
              if (r == 0) npv = g.cb;
            } else {
              op = ringRow[lane];
              const uint32_t* Sp = S + ((k - 1) * 2 + h) * 3 * 32;
              if (r == 0) {
                sp = Sp[lane];
                spu = Sp[32 + lane] | Sp[3 * 32 + 32 + lane];  // row 1 of both halves
              } else {
                sp = Sp[slot * 32 + lane];
                spu = Sp[((r + 1) % 3) * 32 + lane];
              }
            }
            if (r == 1)
              npv = S[(k * 2) * 3 * 32 + lane];  // the goal row of this level (half 0, slot 0)
            else if (r > 1)
              npv = Sk[((r - 1) % 3) * 32 + lane];
            const uint32_t nw = sweepRowStep(sp, spu, npv, op, g.mR, g.mL, g.cb, lane);
            Sk[slot * 32 + lane] = nw;
            if (k == 0 || nw) ringRow[lane] = op & ~nw;
            const uint32_t nzb = __ballot_sync(kFull, nw != 0u);
            if (lane == 0) nz[(k * 2 + h) * 4 + slot] = nzb;
            if (nw) {
              uint32_t* pl = planes + ((h * kSwRP + r % kSwRP) * kSwPlanes) * 32 + lane;
#pragma unroll
              for (int b = 0; b < kSwPlanes; ++b)
                if ((k >> b) & 1) pl[b * 32] |= nw;
            }
            act = nzb != 0u;
          } else {
            Sk[slot * 32 + lane] = 0u;
            if (lane == 0) nz[(k * 2 + h) * 4 + slot] = 0u;
          }
        }
        // free word of the next row of this half (level 0 only)
        if (k == 0) pf = sweepFreeWord(p.rowbits, p.WPR, dimy, h ? gy - (r + 1) : gy + (r + 1), lane);
      } else if (isWriter) {
        const int h = wh, r = tau - 2 * kSwDense + 1;
        const int y = h ? gy - r : gy + r;
        if (r >= 0 && !(h == 1 && r == 0) && y >= 0 && y < dimy) {
          const uint32_t fr = pf;
          const uint32_t op = ring[(h * kSwRO + r % kSwRO) * 32 + lane];
          uint32_t* pl = planes + ((h * kSwRP + r % kSwRP) * kSwPlanes) * 32 + lane;
          uint32_t n0 = pl[0], n1 = pl[32], n2 = pl[64], n3 = fr & ~op;
          pl[0] = 0u;
          pl[32] = 0u;
          pl[64] = 0u;
          // 4x4 bit transposition: nibble i of n_w = (visited, delta bits 2..0) of cell 4i + w
          uint32_t t;
          t = ((n0 >> 2) ^ n2) & 0x33333333u; n2 ^= t; n0 ^= t << 2;
          t = ((n1 >> 2) ^ n3) & 0x33333333u; n3 ^= t; n1 ^= t << 2;
          t = ((n0 >> 1) ^ n1) & 0x55555555u; n1 ^= t; n0 ^= t << 1;
          t = ((n2 >> 1) ^ n3) & 0x55555555u; n3 ^= t; n2 ^= t << 1;
          int32_t* row = g.out + (size_t)y * dimx;
          const int xb = 32 * lane - gx;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int x0 = 32 * lane + 4 * i;
            int v[4];
            const uint32_t nn[4] = {n0, n1, n2, n3};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const uint32_t nib = nn[w] >> (4 * i);
              const int d = abs(xb + 4 * i + w) + r + (int)((nib & 7u) << 1);
              v[w] = (nib & 8u) ? d : MRP_INF;
            }
            if (kVec) {
              if (x0 < dimx) *reinterpret_cast<int4*>(row + x0) = make_int4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int w = 0; w < 4; ++w)
                if (x0 + w < dimx) row[x0 + w] = v[w];
            }
          }
        }
        // free word of the next row
        {
          const int rn = r + 1;
          pf = (rn >= 0 && !(h == 1 && rn == 0)) ? sweepFreeWord(p.rowbits, p.WPR, dimy, h ? gy - rn : gy + rn, lane)
                                                 : 0u;
        }
      } else if (!(p.dbg & 1)) {
        // sparse levels: (level, half) pairs dealt round-robin to the sparse warps
        for (int q = sw; q < (kSwLmax - kSwDense) * 2; q += kSwSparseWarps) {
          const int k = kSwDense + (q >> 1), h = q & 1, r = tau - 2 * k;
          if (r < 0) break;  // deeper levels have not started either
          if (h == 1 && r == 0) continue;
          const int y = h ? gy - r : gy + r;
          const int slot = r % 3;
          uint32_t* const Sk = S + (k * 2 + h) * 3 * 32;
          uint32_t nw = 0u, nzb = 0u;
          if (y >= 0 && y < dimy) {
            const int bp = ((k - 1) * 2 + h) * 4, bk = (k * 2 + h) * 4;
            uint32_t live;
            if (r == 0)
              live = nz[bp] | nz[bp + 1] | nz[bp + 4 + 1];
            else
              live = nz[bp + slot] | nz[bp + (r + 1) % 3] | (r == 1 ? nz[(k * 2) * 4] : nz[bk + (r - 1) % 3]);
            if (live) {
              uint32_t* const ringRow = ring + (h * kSwRO + r % kSwRO) * 32;
              const uint32_t op = ringRow[lane];
              const uint32_t* Sp = S + ((k - 1) * 2 + h) * 3 * 32;
              uint32_t sp, spu, npv = 0u;
              if (r == 0) {
                sp = Sp[lane];
                spu = Sp[32 + lane] | Sp[3 * 32 + 32 + lane];
              } else {
                sp = Sp[slot * 32 + lane];
                spu = Sp[((r + 1) % 3) * 32 + lane];
              }
              if (r == 1)
                npv = S[(k * 2) * 3 * 32 + lane];
              else if (r > 1)
                npv = Sk[((r - 1) % 3) * 32 + lane];
              nw = sweepRowStep(sp, spu, npv, op, g.mR, g.mL, g.cb, lane);
              nzb = __ballot_sync(kFull, nw != 0u);
              if (nzb) {
                if (nw) ringRow[lane] = op & ~nw;
                sweepStoreCells(g, nzb, nw, y, r, k, lane);
                act = true;
                if (k == kSwLmax - 1) sHandOver = 1;
              }
            }
          }
          Sk[slot * 32 + lane] = nw;
          if (lane == 0) nz[(k * 2 + h) * 4 + slot] = nzb;
        }
      }
#ifdef MRP_SWEEP_TIMING
      busy += clock64() - t0;
#endif
      const int any = __syncthreads_or(act ? 1 : 0);
      idle = any ? 0 : idle + 1;
      if (tau >= writerEnd && idle >= 2) break;
    }
#ifdef MRP_SWEEP_TIMING
    if (blockIdx.x == 0 && lane == 0) {
      atomicAdd(&g_swBusy[warp], (unsigned long long)busy);
      if (warp == 0) {
        atomicAdd(&g_swBusy[32], (unsigned long long)(tau + 1));
        atomicAdd(&g_swBusy[33], 1ull);
        atomicAdd(&g_swBusy[34], (unsigned long long)(clock64() - tGoal));
      }
    }
#endif
    if (sHandOver) {
      // more levels than this kernel holds: redone by the queue kernel
      if (tid == 0) p.ws[64 + atomicAdd(&p.ws[2], 1u)] = (uint32_t)gidx;
    } else if (!goalFree && tid == 0) {
      g.out[goal] = 0;  // Floyd–Warshall row of an obstacle
    }
  }
}

constexpr size_t kSwSmemBytes =
    4 * ((size_t)kSwLmax * 2 * 3 * 32 + 2 * kSwRO * 32 + 2 * kSwRP * kSwPlanes * 32 + kSwLmax * 2 * 4);

bool bfsSweepFits(const mrp_map_s* map) {
  if (getenv("MRP_BFS_NOSWEEP")) return false;
  return map->dimx <= 1024 && map->dimy < 32768 && kSwSmemBytes + 1024 <= ctx().smemOptin;
}

size_t bfsSweepWorkspaceWords(int n_goals) { return ((size_t)64 + (size_t)std::max(n_goals, 1) + 63) & ~(size_t)63; }

int launchBfsSweep(const mrp_map_s* map, const int32_t* d_goal_cell, int n_goals, int32_t* d_out, void* d_ws,
                   cudaStream_t st) {
  BfsSweepParams p;
  p.rowbits = map->d_rowbits;
  p.goals = d_goal_cell;
  p.out = d_out;
  p.ws = static_cast<uint32_t*>(d_ws);
  p.n_goals = n_goals;
  p.dimx = map->dimx;
  p.dimy = map->dimy;
  p.WPR = ((map->dimx + 2 + 31) / 32) | 1;
  p.dbg = getenv("MRP_SWEEP_DBG") ? atoi(getenv("MRP_SWEEP_DBG")) : 0;
  MRP_CUDA(cudaMemsetAsync(d_ws, 0, 64 * 4, st));
  const bool vec = (map->dimx & 3) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0;
  auto fn = vec ? bfs_sweep_kernel<true> : bfs_sweep_kernel<false>;
  MRP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSwSmemBytes));
  const int blocks = std::min(ctx().smCount, n_goals);
  fn<<<blocks, kSwThreads, kSwSmemBytes, st>>>(p);
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mrp
