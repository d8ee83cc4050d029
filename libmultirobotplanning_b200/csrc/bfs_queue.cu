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
//     shared-memory atomicAnd (exactly one winner per cell), stores its own
//     level into the int32 field and appends the winners to the next queue; queue
//     slots are handed out with one atomic per warp (one ballot per direction
//     ranks the winners).  A level costs about as many cycles as the
//     instructions on a warp's path through it (a dependent instruction
//     issues every ~4 cycles), so the body is kept short: bit-position
//     entries, row offsets as immediates, no bounds check on the slot.
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
  // goals handed over by the sweep kernel (bfs_sweep.cu): indices into goals[] /
  // out[]; NULL = all n_goals goals
  const uint32_t* goalList;
  const uint32_t* goalListCount;
  int n_goals;
  int dimx, dimy;
  int WPR;       // words per bitmap row (odd)
  int nOpenWords;
  int cap;       // queue capacity (entries, a power of two)
  uint32_t wprMagic;  // row of bitmap word w = umulhi(w, wprMagic) >> wprShift  (= w / WPR)
  int wprShift;
  int tma;  // 1: the free mask of a goal arrives by one bulk-async copy (MRP_BFS_TMA=0 for the load loop)
};

// ---- bulk-async (TMA, 1-D) copy global -> shared, completion on an mbarrier ----
__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bulkLoad(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  // earlier generic-proxy accesses to the destination (the claims of the last goal) before the async-proxy write
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t}" ::"r"(bar),
      "r"(phase)
      : "memory");
}

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

// Claims with the row offset as an immediate when the row stride is a
// compile-time constant.
template <int kOff>
__device__ __forceinline__ uint32_t atomAndOff(uint32_t addr, uint32_t mask) {
  uint32_t old;
  asm volatile("atom.shared.and.b32 %0, [%1+%3], %2;" : "=r"(old) : "r"(addr), "r"(mask), "n"(kOff) : "memory");
  return old;
}
// Tail of a cell expansion: win tests, ranking, slot allocation and queue
// appends (see the call site).
__device__ __forceinline__ void appendWinners(uint32_t oU, uint32_t oD, uint32_t oL, uint32_t oR, uint32_t bit,
                                              uint32_t bitL, uint32_t bitR, uint32_t ltMask, uint32_t slotA,
                                              uint32_t lane0Mask, uint32_t qnS, uint32_t capMask, uint32_t eU,
                                              uint32_t eD, uint32_t e) {
  asm volatile(
      "{\n\t"
      ".reg .pred pu, pd, pl, pr;\n\t"
      ".reg .b32 x, bu, bd, bl, br, nu, nd, nl, nr, nud, nudl, tot, slot, ru, rd, rl, rr, qb, a, v;\n\t"
      "and.b32 x, %0, %4;\n\tsetp.ne.u32 pu, x, 0;\n\t"
      "and.b32 x, %1, %4;\n\tsetp.ne.u32 pd, x, 0;\n\t"
      "and.b32 x, %2, %5;\n\tsetp.ne.u32 pl, x, 0;\n\t"
      "and.b32 x, %3, %6;\n\tsetp.ne.u32 pr, x, 0;\n\t"
      "vote.sync.ballot.b32 bu, pu, 0xffffffff;\n\t"
      "vote.sync.ballot.b32 bd, pd, 0xffffffff;\n\t"
      "vote.sync.ballot.b32 bl, pl, 0xffffffff;\n\t"
      "vote.sync.ballot.b32 br, pr, 0xffffffff;\n\t"
      "popc.b32 nu, bu;\n\tpopc.b32 nd, bd;\n\tpopc.b32 nl, bl;\n\tpopc.b32 nr, br;\n\t"
      "add.u32 nud, nu, nd;\n\tadd.u32 nudl, nud, nl;\n\tadd.u32 tot, nudl, nr;\n\t"
      "and.b32 x, tot, %9;\n\t"
      "atom.shared.add.u32 slot, [%8], x;\n\t"
      "and.b32 x, bu, %7;\n\tpopc.b32 ru, x;\n\t"
      "and.b32 x, bd, %7;\n\tpopc.b32 rd, x;\n\tadd.u32 rd, rd, nu;\n\t"
      "and.b32 x, bl, %7;\n\tpopc.b32 rl, x;\n\tadd.u32 rl, rl, nud;\n\t"
      "and.b32 x, br, %7;\n\tpopc.b32 rr, x;\n\tadd.u32 rr, rr, nudl;\n\t"
      "shfl.sync.idx.b32 slot, slot, 0, 0x1f, 0xffffffff;\n\t"
      "and.b32 slot, slot, %11;\n\t"
      "mad.lo.u32 qb, slot, 4, %10;\n\t"
      "mad.lo.u32 a, ru, 4, qb;\n\t@pu st.shared.u32 [a], %12;\n\t"
      "mad.lo.u32 a, rd, 4, qb;\n\t@pd st.shared.u32 [a], %13;\n\t"
      "mad.lo.u32 a, rl, 4, qb;\n\tadd.u32 v, %14, -1;\n\t@pl st.shared.u32 [a], v;\n\t"
      "mad.lo.u32 a, rr, 4, qb;\n\tadd.u32 v, %14, 1;\n\t@pr st.shared.u32 [a], v;\n\t"
      "}" ::"r"(oU),
      "r"(oD), "r"(oL), "r"(oR), "r"(bit), "r"(bitL), "r"(bitR), "r"(ltMask), "r"(slotA), "r"(lane0Mask), "r"(qnS),
      "r"(capMask), "r"(eU), "r"(eD), "r"(e)
      : "memory");
}

// kWPR / kDimX: words per bitmap row and cells per map row as compile-time
// constants (0 = read them from the parameters).  The 1024-column map of the
// headline configuration gets its own instance: row offsets become immediates.
//
// A queue entry is the BIT POSITION of the cell in the bordered bitmap,
// e = (y + 1) * 32 * WPR + (x + 1): word = e >> 5, bit = e & 31, the four
// neighbours are e -+ 32*WPR and e -+ 1, and the field index follows from
// e - (e / (32*WPR)) * (32*WPR - dimx).
template <int kWPR, int kDimX, bool kList>
__global__ void __launch_bounds__(1024, 1) bfs_queue_kernel(BfsQueueParams p) {
  extern __shared__ uint32_t smem[];
  __shared__ int sCount[3];
  __shared__ int sGoal;
  __shared__ uint32_t sScratch[32];
  __shared__ __align__(8) unsigned long long sBar;
  const int tid = threadIdx.x, lane = tid & 31;
  const int nThreads = blockDim.x;
  const uint32_t ltMask = (1u << lane) - 1u;
  const int WPR = kWPR ? kWPR : p.WPR, dimx = kDimX ? kDimX : p.dimx, cap = p.cap;
  const uint32_t S = 32u * (uint32_t)WPR;          // bitmap bits per row
  const uint32_t padBits = S - (uint32_t)dimx;     // e - Y * padBits - (dimx + 1) = field index
  const int cells = dimx * p.dimy;
  uint32_t* open = smem;
  uint32_t* q0 = smem + ((p.nOpenWords + 3) & ~3);
  const uint32_t openS = pin((uint32_t)__cvta_generic_to_shared(open));
  // each queue is followed by 128 words of slack: a warp appends at most 96
  // entries, and the slot index is wrapped instead of bounds-checked
  const uint32_t q0S = pin((uint32_t)__cvta_generic_to_shared(q0)), q1S = q0S + 4u * (uint32_t)(cap + 128);
  const uint32_t capMask = (uint32_t)cap - 1u;     // cap is a power of two
  const uint32_t cntS = pin((uint32_t)__cvta_generic_to_shared(sCount));
  // slot allocation: lane 0 adds to the level counter, the other lanes add 0
  // to private scratch words (one ATOMS for the warp, and no uniform address
  // for ptxas to wrap into its vote/elect aggregation sequence)
  const uint32_t scrS = pin((uint32_t)__cvta_generic_to_shared(sScratch) + 4u * (uint32_t)lane);
  // cells of the zero rows below the map: all four neighbours are closed (one
  // word per lane, so the idle lanes of a warp do not collide)
  const uint32_t dummy = (uint32_t)(p.dimy + 2) * S + 32u * (uint32_t)(lane % WPR) + 1u;
  const uint32_t rowB = 4u * (uint32_t)WPR;
  // 16-byte stores for the template when rows and the buffer allow it
  const bool vec = (dimx & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
  const int ringFirst = nThreads >= 256 ? nThreads - 128 : nThreads - 32;
  const uint32_t firstOff = 4u * (uint32_t)min(tid, cap - 1);
  if (tid < 32) sScratch[tid] = 0;
  const uint32_t barS = (uint32_t)__cvta_generic_to_shared(&sBar);
  uint32_t barPhase = 0;
  if (p.tma && tid == 0) {
    mbarInit(barS, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#ifdef MRP_BFS_TIMING
  if (tid == 0 && blockIdx.x < 1024) g_bfsqBlocks[blockIdx.x][0] = globalTimer();
#endif

  while (true) {
    __syncthreads();
    if (tid == 0) sGoal = (int)atomicAdd(p.ws, 1u);
    __syncthreads();
    // kList: the goals another kernel handed over (indices into goals[] / out[])
    const int nWork = kList ? (int)*p.goalListCount : p.n_goals;
    if (sGoal >= nWork) break;
    const int gidx = kList ? (int)p.goalList[sGoal] : sGoal;
#ifdef MRP_BFS_TIMING
    if (blockIdx.x == 0 && tid == 0) g_bfsqLevels[0][0] = (unsigned)clock64();
#endif
    if (p.tma) {
      // one bulk-async copy of the whole bitmap (16-byte granules; the allocation is padded)
      if (tid == 0) bulkLoad(openS, p.rowbits, (uint32_t)(((p.nOpenWords + 3) & ~3) * 4), barS);
    } else {
      for (int i = tid; i < p.nOpenWords; i += nThreads) open[i] = __ldg(&p.rowbits[i]);
    }
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
    if (p.tma) {
      mbarWait(barS, barPhase);
      barPhase ^= 1u;
    }
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
        // neighbours
        open[gw] &= ~gbit;
        out[goal] = 0;
        const uint32_t ge = (uint32_t)(gy + 1) * S + (uint32_t)(gx + 1);
        const int dwi[4] = {-WPR, WPR, ((gx + 1) & 31) == 0 ? -1 : 0, ((gx + 1) & 31) == 31 ? 1 : 0};
        const int dbit[4] = {0, 0, -1, 1};
        const uint32_t de[4] = {0u - S, S, 0xffffffffu, 1u};
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
      // field address of the cell at bit position e in row Y:
      //   outb + 4 * (e - Y * padBits)
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
        // counting past the capacity, the entries wrap around)
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
        // slot counter of this level for lane 0, a private word for the others
        const uint32_t slotA = lane == 0 ? aNi : scrS;
        const uint32_t lane0Mask = lane == 0 ? 0xffffffffu : 0u;
        for (int i = tid; i - lane < count;) {
          // lanes past the end expand a dummy cell whose neighbourhood is closed
          if (i >= count) e = dummy;
          const uint32_t b = e & 31u, w = e >> 5;
          const uint32_t aC = openS + 4u * w;
          const uint32_t aL = b == 0u ? aC - 4u : aC, aR = b == 31u ? aC + 4u : aC;
          const uint32_t bit = 1u << b;
          const uint32_t bitL = __funnelshift_r(bit, bit, 1), bitR = __funnelshift_l(bit, bit, 1);
          // claims: four independent atomics in flight, no pre-check (an ATOMS
          // costs the same LSU time whatever the number of active lanes, and a
          // bit only ever goes 1 -> 0)
          uint32_t oU, oD;
          if constexpr (kWPR != 0) {
            oU = atomAndOff<-4 * kWPR>(aC, ~bit);
            oD = atomAndOff<4 * kWPR>(aC, ~bit);
          } else {
            oU = atomAnd(aC - rowB, ~bit);
            oD = atomAnd(aC + rowB, ~bit);
          }
          const uint32_t oL = atomAnd(aL, ~bitL), oR = atomAnd(aR, ~bitR);
          // The distance of a cell is stored when the cell is expanded (one full
          // store per warp pass instead of four sparse ones for the cells it
          // discovers); row Y of the bitmap = word / WPR.
          const uint32_t Y = kWPR ? w / (uint32_t)(kWPR ? kWPR : 1) : __umulhi(w, p.wprMagic) >> p.wprShift;
          if (i < count) *reinterpret_cast<int32_t*>(outb + 4 * (size_t)(e - Y * padBits)) = level - 1;
          // winners and queue slots in one PTX block (the four win predicates
          // stay in predicate registers from the test to the stores).  Slots are
          // direction-major inside the warp's block: one ballot per direction
          // ranks the winners, one atomic per warp (lane 0 adds the total, the
          // other lanes add 0 to private words) allocates, and the slot index
          // wraps at the capacity instead of being checked.
          appendWinners(oU, oD, oL, oR, bit, bitL, bitR, ltMask, slotA, lane0Mask, qnS, capMask, e - S, e + S, e);
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

static thread_local int t_bfsBlockCap = 0;
void setBfsBlockCap(int blocks) { t_bfsBlockCap = blocks; }
int bfsBlockCap() { return t_bfsBlockCap; }

struct QueueGeom {
  int WPR, nOpenWords, cap, threads;
  size_t smemBytes;
  bool fits;
};

constexpr int kQueueSlack = 128;  // words behind each queue (see the kernel)

static QueueGeom queueGeometry(const mrp_map_s* map) {
  QueueGeom q;
  q.WPR = bitmapRowWords(map->dimx);
  q.nOpenWords = (map->dimy + 4) * q.WPR;  // border rows + two zero rows (dummy cell)
  const int span = map->dimx + map->dimy;
  // 12 warps serve a 1024x1024 map as fast as 32 (the level loop is bound by
  // the shared-memory atomics and the scattered stores, not by warp count;
  // measured 1014 us per goal at 384 threads, 1037 at 512, 1066 at 768)
  int th = std::min(384, std::max(64, (span / 4 + 31) & ~31));
  if (const char* e = getenv("MRP_BFS_THREADS")) th = atoi(e);
  q.threads = th;
  // a wavefront on an open grid holds < 2*(dimx+dimy) cells; leave 4x room
  // (rounded up to a power of two: the slot index is wrapped, not checked)
  size_t cap = 256;
  while (cap < 8192 && cap < 4 * (size_t)span) cap *= 2;
  const size_t openBytes = (size_t)((q.nOpenWords + 3) & ~3) * 4;
  const size_t limit = ctx().smemOptin - 1024;
  while (cap > 256 && openBytes + 2 * (cap + kQueueSlack) * 4 > limit) cap /= 2;
  if (const char* e = getenv("MRP_BFS_QCAP")) {
    cap = 32;
    while ((int)cap < atoi(e)) cap *= 2;
  }
  q.cap = (int)cap;
  q.smemBytes = openBytes + 2 * (cap + kQueueSlack) * 4;
  q.fits = q.smemBytes <= limit && map->dimx < 65534 && map->dimy < 65534;
  return q;
}

bool bfsQueueFits(const mrp_map_s* map) {
  if (getenv("MRP_BFS_TILES")) return false;
  return queueGeometry(map).fits;
}

typedef void (*QueueKernel)(BfsQueueParams);

// the instance for this map: row strides as immediates for the power-of-two
// widths of the synthetic configurations, run-time strides for everything else
static QueueKernel queueKernelFor(const mrp_map_s* map, const QueueGeom& q, bool list) {
  if (list) return bfs_queue_kernel<0, 0, true>;  // hand-over path: run-time strides
  if (!getenv("MRP_BFS_GENERIC")) {
    if (map->dimx == 1024 && q.WPR == 33) return bfs_queue_kernel<33, 1024, false>;
    if (map->dimx == 512 && q.WPR == 17) return bfs_queue_kernel<17, 512, false>;
    if (map->dimx == 256 && q.WPR == 9) return bfs_queue_kernel<9, 256, false>;
    if (map->dimx == 2048 && q.WPR == 65) return bfs_queue_kernel<65, 2048, false>;
  }
  return bfs_queue_kernel<0, 0, false>;
}

static int queueBlocks(QueueKernel fn, const QueueGeom& q) {
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q.smemBytes);
  int perSm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, q.threads, q.smemBytes) !=
          cudaSuccess ||
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
                   int32_t* d_out, void* d_ws, cudaStream_t st, const uint32_t* d_goalList,
                   const uint32_t* d_goalListCount) {
  const QueueGeom q = queueGeometry(map);
  BfsQueueParams p;
  p.rowbits = map->d_rowbits;
  p.goals = d_goal_cell;
  p.out = d_out;
  p.ws = static_cast<uint32_t*>(d_ws);
  p.n_goals = n_goals;
  p.goalList = d_goalList;
  p.goalListCount = d_goalListCount;
  p.dimx = map->dimx;
  p.dimy = map->dimy;
  p.WPR = q.WPR;
  p.nOpenWords = q.nOpenWords;
  p.cap = q.cap;
  // w / WPR for w < 2^27 (exact): magic = ceil(2^(32+s) / WPR), s = floor(log2 WPR)
  bitmapRowDivision(q.WPR, &p.wprMagic, &p.wprShift);
  const char* tma = getenv("MRP_BFS_TMA");
  p.tma = tma ? atoi(tma) : 1;
  if ((reinterpret_cast<uintptr_t>(map->d_rowbits) & 15) != 0) p.tma = 0;
  MRP_CUDA(cudaMemsetAsync(d_ws, 0, 64 * 4, st));
  const QueueKernel fn = queueKernelFor(map, q, d_goalList != nullptr);
  int blocks = queueBlocks(fn, q);
  if (blocks > n_goals) blocks = n_goals;
  if (bfsBlockCap() > 0 && blocks > bfsBlockCap()) blocks = bfsBlockCap();
  fn<<<blocks, q.threads, q.smemBytes, st>>>(p);
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mrp
