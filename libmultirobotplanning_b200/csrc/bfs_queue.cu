// bfs_queue.cu — distance fields on maps larger than one 32x32 tile whose
// 1-bit-per-cell state fits shared memory (up to ~1.3 M cells; the headline
// case is the synthetic 1024x1024 map, config C5).
//
// Replaces ShortestPathHeuristic (example/shortest_path_heuristic.hpp:12-62 of
// the reference; Floyd–Warshall is O(V^3) and infeasible at V = 2^20) by one
// BFS distance field per goal, the layout of the reference's disabled
// computeHeuristic (example/cbs.cpp:445-557).
//
// Design (one CTA per goal, persistent over goals):
//   * `open` = free & not-yet-visited, 1 bit per cell, row-major with a one
//     cell border of zeros (cell (x,y) = bit x+1 of row y+1), so neighbour
//     look-ups need no bounds checks (two more zero rows below the map hold a
//     dummy cell for idle lanes).  Row stride is an odd number of words:
//     the cells of a diagonal wavefront fall into distinct banks.
//   * A level is a compacted queue of frontier cells, one thread per cell.
//     A thread tests its 4 neighbours in `open`, claims the open ones with a
//     shared-memory atomicAnd (exactly one winner per cell), stores the level
//     into the int32 field and appends the cell to the next queue; queue
//     slots are handed out with one atomic per warp (ballot ranking).
//   * One barrier per level: queues are double-buffered and the three level
//     counters rotate, so nothing has to be reset between two barriers.
//   * MRP_INF (obstacles, other components) is written AHEAD of the wavefront
//     as whole 32-byte sectors: the sectors whose nearest cell lies at
//     Manhattan distance R from the goal get 8 x MRP_INF at level R - kLead,
//     i.e. always before any cell of the sector can be visited (BFS distance
//     >= Manhattan distance).  The 4-byte level stores that follow then hit
//     sectors that are resident and dirty in L2: no DRAM fill for the partial
//     write, one write-back per sector, no final sweep, no second pass over
//     the map.
//   * A level that does not fit the shared-memory queues marks the goal as
//     overflowed; those goals are redone by the tiled kernel (bfs_large.cu),
//     whose lists are bounded by the tile count.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace mrp {

struct BfsQueueParams {
  const uint32_t* __restrict__ rowbits;  // [(dimy+2)*WPR] free mask with border
  const int32_t* __restrict__ goals;     // goal cells
  int32_t* __restrict__ out;             // [n_goals][cells]
  uint32_t* ws;  // [0] goal counter, [2] number of overflowed goals, [64..] their indices
  int n_goals;
  int dimx, dimy;
  int WPR;       // words per bitmap row (odd)
  int nOpenWords;
  int cap;       // queue capacity (entries)
  int dbg;       // debug: bit0 = skip the field stores of the level loop
};

#ifdef MRP_BFS_TIMING
// per-level log of CTA 0: frontier size and clock (tools/bfsq_timing.py)
__device__ unsigned int g_bfsqLevels[4096][2];
__device__ unsigned long long g_bfsqBlocks[1024][4];
extern "C" int mrp_debug_bfsq_blocks(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_bfsqBlocks, sizeof(unsigned long long) * 1024 * 4);
}
__device__ __forceinline__ unsigned long long globalTimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
extern "C" int mrp_debug_bfsq_levels(unsigned int* out) {
  return (int)cudaMemcpyFromSymbol(out, g_bfsqLevels, sizeof(unsigned int) * 4096 * 2);
}
#endif

// ---- shared-memory accesses by 32-bit shared-window address ----
// (inline PTX: the addresses stay in registers and the four claims of a cell
// are issued back to back instead of behind compiler-generated branches)
__device__ __forceinline__ uint32_t ldShared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void stShared(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void stSharedIf(uint32_t addr, uint32_t v, uint32_t pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}" ::"r"(addr),
               "r"(v), "r"(pred)
               : "memory");
}
__device__ __forceinline__ void stGlobalIf(void* ptr, int32_t v, uint32_t pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u32 [%0], %1;\n\t}" ::"l"(ptr),
               "r"(v), "r"(pred)
               : "memory");
}
__device__ __forceinline__ uint32_t atomAddShared(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t atomAnd(uint32_t addr, uint32_t mask) {
  uint32_t old;
  asm volatile("atom.shared.and.b32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(mask) : "memory");
  return old;
}
// keeps a value in a register (no re-derivation from special registers inside the level loop)
__device__ __forceinline__ uint32_t pin(uint32_t v) {
  asm volatile("" : "+r"(v));
  return v;
}

constexpr int kLead = 2;  // the template runs this many levels ahead of the wavefront

// Ring R of the template: every 8-cell sector (row y, columns 8s..8s+7) whose
// nearest cell is at Manhattan distance exactly R from the goal is filled with
// MRP_INF.  Sector column s is at x-distance dx(s) from the goal column, so the
// ring meets it in the rows gy +- (R - dx(s)).  Threads stride over s.
__device__ __forceinline__ void templateRing(int32_t* __restrict__ out, int R, int s0, int sStride, int gx,
                                             int gy, int dimx, int dimy, bool vec) {
  const int nS = (dimx + 7) >> 3, gs = gx >> 3;
  for (int s = s0; s < nS; s += sStride) {
    const int dxs = s == gs ? 0 : (s > gs ? 8 * s - gx : gx - (8 * s + 7));
    const int dy = R - dxs;
    if (dy < 0) continue;
    const int n = min(8, dimx - 8 * s);
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const int y = side ? gy - dy : gy + dy;
      if (y < 0 || y >= dimy || (side && dy == 0)) continue;
      int32_t* q = out + (size_t)y * dimx + 8 * s;
      if (vec) {
        const int4 inf4 = make_int4(MRP_INF, MRP_INF, MRP_INF, MRP_INF);
        reinterpret_cast<int4*>(q)[0] = inf4;
        if (n > 4) reinterpret_cast<int4*>(q)[1] = inf4;
      } else {
        for (int k = 0; k < n; ++k) q[k] = MRP_INF;
      }
    }
  }
}

__global__ void __launch_bounds__(1024, 1) bfs_queue_kernel(BfsQueueParams p) {
  extern __shared__ uint32_t smem[];
  __shared__ int sCount[3];
  __shared__ int sGoal;
  __shared__ uint32_t sScratch[32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int nThreads = blockDim.x;
  const uint32_t ltMask = (1u << lane) - 1u;
  const int WPR = p.WPR, dimx = p.dimx, cap = p.cap;
  const int cells = p.dimx * p.dimy;
  uint32_t* open = smem;
  uint32_t* q0 = smem + ((p.nOpenWords + 3) & ~3);
  const uint32_t openS = pin((uint32_t)__cvta_generic_to_shared(open));
  const uint32_t q0S = pin((uint32_t)__cvta_generic_to_shared(q0)), q1S = q0S + 4u * (uint32_t)cap;
  const uint32_t cntS = pin((uint32_t)__cvta_generic_to_shared(sCount));
  // slot allocation: lane 0 adds to the level counter, the other lanes add 0
  // to private scratch words (one ATOMS for the warp, and no uniform address
  // for ptxas to wrap into its vote/elect aggregation sequence)
  const uint32_t scrS = pin((uint32_t)__cvta_generic_to_shared(sScratch) + 4u * (uint32_t)lane);
  // cells of the zero rows below the map: all four neighbours are closed (one
  // word per lane, so the idle lanes of a warp do not collide)
  const uint32_t dummy = ((uint32_t)(p.dimy + 2) << 16) | (32u * (uint32_t)(lane % WPR) + 1u);
  const uint32_t rowB = 4u * (uint32_t)WPR;
  const size_t outRowB = 4 * (size_t)dimx;
  const uint32_t nodbg = (p.dbg & 1) ? 0u : 1u;
  // 16-byte stores for the template when rows and the buffer allow it
  const bool vec = (dimx & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
  const int ringFirst = nThreads >= 256 ? nThreads - 128 : nThreads - 32;
  const uint32_t firstOff = 4u * (uint32_t)min(tid, cap - 1);
  if (tid < 32) sScratch[tid] = 0;
#ifdef MRP_BFS_TIMING
  if (tid == 0 && blockIdx.x < 1024) g_bfsqBlocks[blockIdx.x][0] = globalTimer();
#endif

  while (true) {
    __syncthreads();
    if (tid == 0) sGoal = (int)atomicAdd(p.ws, 1u);
    __syncthreads();
    const int gidx = sGoal;
    if (gidx >= p.n_goals) break;
#ifdef MRP_BFS_TIMING
    if (blockIdx.x == 0 && tid == 0) g_bfsqLevels[0][0] = (unsigned)clock64();
#endif
    for (int i = tid; i < p.nOpenWords; i += nThreads) open[i] = __ldg(&p.rowbits[i]);
    if (tid == 0) {
      sCount[0] = 0;
      sCount[1] = 0;
      sCount[2] = 0;
    }
    const int goal = p.goals[gidx];
    int32_t* out = p.out + (size_t)gidx * cells;
    const int gy = goal / dimx, gx = goal - gy * dimx;
    const int gw = (gy + 1) * WPR + ((gx + 1) >> 5);
    const uint32_t gbit = 1u << ((gx + 1) & 31);
    __syncthreads();
    const bool goalFree = (open[gw] & gbit) != 0;
    // largest Manhattan distance from the goal to a cell of the map
    const int maxR = max(gx, dimx - 1 - gx) + max(gy, p.dimy - 1 - gy);
    // the first rings of the template (everything levels 0 and 1 can touch and
    // what the level loop expects to be in place), by the whole CTA
    for (int R = tid >> 5; R <= (goalFree ? kLead + 1 : maxR); R += nThreads >> 5)
      templateRing(out, R, lane, 32, gx, gy, dimx, p.dimy, vec);
    __syncthreads();
    if (goalFree) {
      if (tid == 0) {
        // levels 0 and 1 by one thread: the goal and its (up to 4) open
        // neighbours.  From level 2 on every frontier cell has a visited
        // neighbour, so a cell appends at most 3 cells (2 ballots rank them).
        open[gw] &= ~gbit;
        out[goal] = 0;
        const uint32_t ge = ((uint32_t)(gy + 1) << 16) | (uint32_t)(gx + 1);
        const int dwi[4] = {-WPR, WPR, ((gx + 1) & 31) == 0 ? -1 : 0, ((gx + 1) & 31) == 31 ? 1 : 0};
        const int dbit[4] = {0, 0, -1, 1};
        const uint32_t de[4] = {0xffff0000u, 0x10000u, 0xffffffffu, 1u};
        const int dout[4] = {-dimx, dimx, -1, 1};
        int n = 0;
        for (int k = 0; k < 4; ++k) {
          const uint32_t nb = 1u << ((gx + 1 + dbit[k]) & 31);
          if (open[gw + dwi[k]] & nb) {
            open[gw + dwi[k]] &= ~nb;
            q0[n++] = ge + de[k];
            out[goal + dout[k]] = 1;
          }
        }
        sCount[2] = n;
      }
      __syncthreads();
#ifdef MRP_BFS_TIMING
      if (tid == 0 && blockIdx.x < 1024) g_bfsqBlocks[blockIdx.x][2] = globalTimer();
#endif
      // counters of the current level, the next level, and the one to reset
      uint32_t aCi = cntS + 8u, aNi = cntS, aRi = cntS + 4u;
      // field address of bitmap cell (X, Y) = outb + 4 * (Y * dimx + X)
      char* const outb = reinterpret_cast<char*>(out) - 4 * (size_t)(dimx + 1);
      int count, level;
      for (level = 2;; ++level) {
        const uint32_t qcS = (level & 1) ? q1S : q0S, qnS = (level & 1) ? q0S : q1S;
        // the first entry of this thread is fetched together with the counter
        uint32_t e;
        asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%3];"
                     : "=r"(count), "=r"(e)
                     : "r"(aCi), "r"(qcS + firstOff)
                     : "memory");
        // a level that outgrew the queue ends the attempt (the counter keeps
        // counting past the capacity, the entries are dropped)
        if (count == 0 || count > cap) break;
        if (tid == 0) stShared(aRi, 0u);
        // template ring of this level, by the last warps of the CTA (the
        // frontier is served from warp 0 up, so these are the idle ones)
        if (tid >= ringFirst) templateRing(out, level + kLead, tid - ringFirst, nThreads - ringFirst, gx, gy, dimx, p.dimy, vec);
#ifdef MRP_BFS_TIMING
        if (blockIdx.x == 0 && tid == 0 && level < 4096) {
          g_bfsqLevels[level][0] = (unsigned)count;
          g_bfsqLevels[level][1] = (unsigned)clock64();
        }
#endif
        for (int i = tid; i - lane < count;) {
          // lanes past the end expand a dummy cell whose neighbourhood is closed
          if (i >= count) e = dummy;
          const uint32_t X = e & 0xffffu, Y = e >> 16;
          const uint32_t b = X & 31u;
          const uint32_t aC = openS + 4u * (Y * (uint32_t)WPR + (X >> 5));
          const uint32_t aL = b == 0u ? aC - 4u : aC, aR = b == 31u ? aC + 4u : aC;
          const uint32_t bit = 1u << b;
          const uint32_t bitL = __funnelshift_r(bit, bit, 1), bitR = __funnelshift_l(bit, bit, 1);
          // claims: four independent atomics in flight, no pre-check (an ATOMS
          // costs the same LSU time whatever the number of active lanes, and a
          // bit only ever goes 1 -> 0)
          const uint32_t oU = atomAnd(aC - rowB, ~bit), oD = atomAnd(aC + rowB, ~bit);
          const uint32_t oL = atomAnd(aL, ~bitL), oR = atomAnd(aR, ~bitR);
          char* const pc = outb + 4 * (size_t)(Y * (uint32_t)dimx + X);  // this cell in the field
          const uint32_t tU = (oU >> b) & 1u, tD = (oD >> b) & 1u;
          const uint32_t tL = (oL & bitL) != 0u, tR = (oR & bitR) != 0u;
          // queue slots: a cell appends c <= 3 cells; two ballots (the bits of
          // c) rank them, one atomic per warp
          const uint32_t c = tU + tD + tL + tR;
          const uint32_t B0 = __ballot_sync(0xffffffffu, c & 1u);
          const uint32_t B1 = __ballot_sync(0xffffffffu, c & 2u);
          const uint32_t total = __popc(B0) + 2 * __popc(B1);
          if (total != 0) {
            uint32_t slot = atomAddShared(lane == 0 ? aNi : scrS, lane == 0 ? total : 0u);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot + total <= (uint32_t)cap) {
              const uint32_t qU = qnS + 4u * (slot + __popc(B0 & ltMask) + 2 * __popc(B1 & ltMask));
              const uint32_t qD = qU + 4u * tU, qL = qD + 4u * tD, qR = qL + 4u * tL;
              stSharedIf(qU, e - 0x10000u, tU);
              stSharedIf(qD, e + 0x10000u, tD);
              stSharedIf(qL, e - 1u, tL);
              stSharedIf(qR, e + 1u, tR);
              stGlobalIf(pc - outRowB, level, tU & nodbg);
              stGlobalIf(pc + outRowB, level, tD & nodbg);
              stGlobalIf(pc - 4, level, tL & nodbg);
              stGlobalIf(pc + 4, level, tR & nodbg);
            }
          }
          i += nThreads;
          if (i - lane >= count) break;
          e = ldShared(qcS + 4u * (uint32_t)min(i, count - 1));
        }
        const uint32_t t = aRi;
        aRi = aCi;
        aCi = aNi;
        aNi = t;
        __syncthreads();
      }
      if (count > cap) {
        // redone by the tiled kernel; whatever this CTA wrote is overwritten
        if (tid == 0) p.ws[64 + atomicAdd(&p.ws[2], 1u)] = (uint32_t)gidx;
        continue;
      }
      // the rings the wavefront did not get to (cells farther away than the
      // last level: other components, obstacles at the far end of the map)
      for (int R = level + kLead + (tid >> 5); R <= maxR; R += nThreads >> 5)
        templateRing(out, R, lane, 32, gx, gy, dimx, p.dimy, vec);
    }

#ifdef MRP_BFS_TIMING
    if (blockIdx.x == 0 && tid == 0) g_bfsqLevels[1][0] = (unsigned)clock64();
    if (tid == 0 && blockIdx.x < 1024) g_bfsqBlocks[blockIdx.x][3] = globalTimer();
#endif
    if (!goalFree) {
      __syncthreads();
      if (tid == 0) out[goal] = 0;  // Floyd–Warshall row of an obstacle
    }
#ifdef MRP_BFS_TIMING
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) g_bfsqLevels[1][1] = (unsigned)clock64();
#endif
  }
#ifdef MRP_BFS_TIMING
  if (tid == 0 && blockIdx.x < 1024) g_bfsqBlocks[blockIdx.x][1] = globalTimer();
#endif
}

struct QueueGeom {
  int WPR, nOpenWords, cap, threads;
  size_t smemBytes;
  bool fits;
};

static QueueGeom queueGeometry(const mrp_map_s* map) {
  QueueGeom q;
  q.WPR = ((map->dimx + 2 + 31) / 32) | 1;
  q.nOpenWords = (map->dimy + 4) * q.WPR;  // border rows + two zero rows (dummy cell)
  const int span = map->dimx + map->dimy;
  int th = 64;
  while (th < 1024 && th < span / 2) th <<= 1;
  if (const char* e = getenv("MRP_BFS_THREADS")) th = atoi(e);
  q.threads = th;
  // a wavefront on an open grid holds < 2*(dimx+dimy) cells; leave 4x room
  size_t cap = std::min<size_t>(8192, std::max<size_t>(256, 4 * (size_t)span));
  const size_t openBytes = (size_t)((q.nOpenWords + 3) & ~3) * 4;
  const size_t limit = ctx().smemOptin - 1024;
  while (cap > 256 && openBytes + 2 * cap * 4 > limit) cap /= 2;
  if (const char* e = getenv("MRP_BFS_QCAP")) cap = std::max(32, atoi(e));
  q.cap = (int)cap;
  q.smemBytes = openBytes + 2 * cap * 4;
  q.fits = q.smemBytes <= limit && map->dimx < 65534 && map->dimy < 65534;
  return q;
}

bool bfsQueueFits(const mrp_map_s* map) {
  if (getenv("MRP_BFS_TILES")) return false;
  return queueGeometry(map).fits;
}

static int queueBlocks(const QueueGeom& q) {
  cudaFuncSetAttribute(bfs_queue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)q.smemBytes);
  int perSm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, bfs_queue_kernel, q.threads,
                                                    q.smemBytes) != cudaSuccess ||
      perSm < 1)
    perSm = 1;
  if (perSm > 8) perSm = 8;
  return ctx().smCount * perSm;
}

// workspace words in front of the tiled kernel's own workspace: header (64) +
// one slot per goal for the overflow list
size_t bfsQueueWorkspaceWords(int n_goals) {
  return ((size_t)64 + (size_t)std::max(n_goals, 1) + 63) & ~(size_t)63;
}

int launchBfsQueue(const mrp_map_s* map, const int32_t* d_goal_cell, int n_goals,
                   int32_t* d_out, void* d_ws, cudaStream_t st) {
  const QueueGeom q = queueGeometry(map);
  BfsQueueParams p;
  p.rowbits = map->d_rowbits;
  p.goals = d_goal_cell;
  p.out = d_out;
  p.ws = static_cast<uint32_t*>(d_ws);
  p.n_goals = n_goals;
  p.dimx = map->dimx;
  p.dimy = map->dimy;
  p.WPR = q.WPR;
  p.nOpenWords = q.nOpenWords;
  p.cap = q.cap;
  p.dbg = getenv("MRP_BFS_DBG") ? atoi(getenv("MRP_BFS_DBG")) : 0;
  MRP_CUDA(cudaMemsetAsync(d_ws, 0, 64 * 4, st));
  int blocks = queueBlocks(q);
  if (blocks > n_goals) blocks = n_goals;
  bfs_queue_kernel<<<blocks, q.threads, q.smemBytes, st>>>(p);
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mrp
