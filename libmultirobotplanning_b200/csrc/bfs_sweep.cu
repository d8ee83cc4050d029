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
constexpr int kSwDense = 4;    // levels swept row by row, one warp per (level, half)
constexpr int kSwRO = 200;     // rows of the open ring per half (>= 2*kSwLmax - 1)
constexpr int kSwRP = 12;      // rows of the plane ring per half (>= 2*kSwDense + 1)
constexpr int kSwPlanes = 2;   // bits of a dense level
constexpr int kSwThreads = 1024;
constexpr int kSwWarps = kSwThreads / 32;
constexpr int kSwWriterWarps = 4;  // (half, part of the row)
constexpr int kSwSparseWarps = kSwWarps - 2 * kSwDense - kSwWriterWarps;
constexpr int kSwSparseLevels = kSwLmax - kSwDense;
constexpr int kSwListCap = 2048;   // word items per step
static_assert((1 << kSwPlanes) >= kSwDense && kSwPlanes <= 3, "planes");
static_assert(kSwRO >= 2 * kSwLmax - 1 && kSwRP >= 2 * kSwDense + 1, "rings");
static_assert(kSwSparseLevels <= 128, "item encoding");

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

// One (level, row) step of a dense level, a row = the 32 lanes of the warp.
// sp = S_{k-1}[row], spu = the inward vertical sources (S_{k-1} of the next row
// out), npv = S_k of the previous row (outward vertical), op = open cells of
// the row.  mR / mL: cells right / left of the goal column, cb: the goal column.
__device__ __forceinline__ uint32_t sweepRowStep(uint32_t sp, uint32_t spu, uint32_t npv, uint32_t op,
                                                 uint32_t mR, uint32_t mL, uint32_t cb, int lane) {
  const uint32_t spR = sp & mR, spL = sp & mL;
  uint32_t hi = __shfl_down_sync(kFull, spR, 1);
  uint32_t lo = __shfl_up_sync(kFull, spL, 1);
  if (lane == 31) hi = 0u;
  if (lane == 0) lo = 0u;
  const uint32_t seed = (__funnelshift_r(spR, hi, 1) | __funnelshift_l(lo, spL, 1) | spu | npv) & op;
  const uint32_t sc = seed & cb;
  const uint32_t mU = (op & mR) | sc, sU = seed & (mR | cb);
  const uint32_t mD = __brev((op & mL) | sc), sD = __brev(seed & (mL | cb));
  uint32_t sumU, cU, sumD, cD;
  addCarry(mU, sU, sumU, cU);
  addCarry(mD, sD, sumD, cD);
  const uint32_t gU = __ballot_sync(kFull, cU != 0u), gD = __ballot_sync(kFull, cD != 0u);
  const uint32_t qU = __ballot_sync(kFull, sumU == kFull), qD = __ballot_sync(kFull, sumD == kFull);
  const uint32_t u = resolveCarries(gU, qU), d = resolveCarries(__brev(gD), __brev(qD));
  sumU += (u >> lane) & 1u;
  sumD += (d >> (31 - lane)) & 1u;
  return (((sumU ^ mU) & mU) | sU) | __brev(((sumD ^ mD) & mD) | sD);
}

// ---- sparse levels: word items ------------------------------------------------
// A sparse level is not swept: the cells a step reaches in one word of a row
// post their consequences as seed bits for the words they can reach next
//   outward vertical   -> same level, next row out, same word      (step + 1)
//   inward vertical    -> next level, next row in, same word       (step + 1)
//   inward horizontal  -> next level, same row, shifted one cell   (step + 2)
// into an accumulator word per (level, half, word) and step; whoever finds the
// accumulator empty also appends the word to the item list of that step.  A
// thread per item takes the seeds (exchange with 0), floods them through the
// open cells of its word (the addition trick inside one word; a carry makes the
// thread go on with the next word), claims the cells with an atomic AND on the
// open ring (exactly one winner per cell), stores their distances and posts
// their consequences.  Idle levels cost nothing.
struct SweepShared {
  uint32_t* ring;    // [2][kSwRO][32] open cells of the rows in flight
  uint32_t* acc;     // [kSwSparseLevels][2][3][32] seeds per step % 3
  uint16_t* list;    // [4][kSwListCap] items of step % 4: (level - kSwDense) << 6 | half << 5 | word
  uint32_t* cnt;     // [4]
  int* handOver;
};

struct SweepGoal {
  int gx, gy, jg, dimx, dimy;
  uint32_t hiG, loG, cbG;  // masks of the goal word: right of / left of / the goal column
  int32_t* out;
};

// One seed word for (level k, half h, word j) at `step`.  The accumulator update
// is issued at once; whether this was the first seed of the word (and the word
// therefore needs an item) is told by the returned old value.
struct SweepPost {
  uint32_t old;   // accumulator before the update (valid if code != 0xffff)
  uint32_t code;  // item code, 0xffff = nothing posted
};
__device__ __forceinline__ SweepPost sweepPost(const SweepShared& sh, int step3, int k, int h, int j, uint32_t bits,
                                               bool cond) {
  SweepPost r;
  r.old = 1u;
  r.code = 0xffffu;
  if (cond && bits != 0u) {
    if (k >= kSwLmax) {
      *sh.handOver = 1;
    } else {
      const int kk = k - kSwDense;
      r.old = atomicOr(&sh.acc[((kk * 2 + h) * 3 + step3) * 32 + j], bits);
      r.code = (uint32_t)((kk << 6) | (h << 5) | j);
    }
  }
  return r;
}
// appends the words of `a`, `b`, `c` that had no seed before to the item list
__device__ __forceinline__ void sweepAppend(const SweepShared& sh, int step, const SweepPost& a, const SweepPost& b,
                                            const SweepPost& c) {
  const uint32_t fa = a.old == 0u, fb = b.old == 0u, fc = c.old == 0u;
  const uint32_t n = fa + fb + fc;
  if (n) {
    uint32_t pos = atomicAdd(&sh.cnt[step & 3], n);
    if (pos + n > (uint32_t)kSwListCap) {
      *sh.handOver = 1;
      return;
    }
    uint16_t* l = sh.list + (step & 3) * kSwListCap;
    if (fa) l[pos++] = (uint16_t)a.code;
    if (fb) l[pos++] = (uint16_t)b.code;
    if (fc) l[pos] = (uint16_t)c.code;
  }
}

// consequences of the cells `won` reached in word j of row (h, r) at level k during step tau
__device__ __forceinline__ void sweepPostAll(const SweepShared& sh, const SweepGoal& g, int tau, int k, int h, int r,
                                             int j, uint32_t won, uint32_t mR, uint32_t mL, bool outward) {
  const int s1 = (tau + 1) % 3, s2 = (tau + 2) % 3;
  // step + 1: outward vertical (same level, next row out; the goal row feeds both halves), inward vertical
  const int yOut = h ? g.gy - (r + 1) : g.gy + (r + 1);
  const SweepPost o0 = sweepPost(sh, s1, k, r == 0 ? 0 : h, j, won,
                                 outward && (r == 0 ? g.gy + 1 < g.dimy : (yOut >= 0 && yOut < g.dimy)));
  const SweepPost o1 = sweepPost(sh, s1, k, 1, j, won, outward && r == 0 && g.gy >= 1);
  const SweepPost iv = sweepPost(sh, s1, k + 1, r == 1 ? 0 : h, j, won, r >= 1);
  // step + 2: inward horizontal (next level, same row, one cell towards the goal column)
  const uint32_t wR = won & mR, wL = won & mL;
  const SweepPost h0 = sweepPost(sh, s2, k + 1, h, j, (wR >> 1) | (wL << 1), true);
  const SweepPost h1 = sweepPost(sh, s2, k + 1, h, j - 1, 0x80000000u, (wR & 1u) != 0u);  // j > jg >= 0 here
  const SweepPost h2 = sweepPost(sh, s2, k + 1, h, j + 1, 1u, (wL >> 31) != 0u);          // j < jg <= 31 here
  sweepAppend(sh, tau + 1, o0, o1, iv);
  sweepAppend(sh, tau + 2, h0, h1, h2);
}

__device__ __forceinline__ void sweepItem(const SweepShared& sh, const SweepGoal& g, int tau, uint32_t item) {
  const int kk = item >> 6, h = (item >> 5) & 1, k = kk + kSwDense;
  int j = item & 31;
  const int r = tau - 2 * k;
  const int y = h ? g.gy - r : g.gy + r;
  uint32_t* const ringRow = sh.ring + (h * kSwRO + r % kSwRO) * 32;
  uint32_t* const accRow = sh.acc + ((kk * 2 + h) * 3 + tau % 3) * 32;
  int32_t* const row = g.out + (size_t)y * g.dimx;
  const int base = r + 2 * k;
  const int j0 = j;
  uint32_t extra = 0u, leftCarry0 = 0u;
  int phase = 0;  // 0: the item's own word, 1: following a carry to the right, 2: to the left
  while (true) {
    const uint32_t seed = atomicExch(&accRow[j], 0u) | extra;
    const uint32_t op = ringRow[j];
    const uint32_t s = seed & op;
    uint32_t cU = 0u, cD = 0u;
    if (s) {
      const uint32_t mR = j > g.jg ? kFull : (j == g.jg ? g.hiG : 0u);
      const uint32_t mL = j < g.jg ? kFull : (j == g.jg ? g.loG : 0u);
      const uint32_t cb = j == g.jg ? g.cbG : 0u;
      const uint32_t sc = s & cb;
      const uint32_t mU = (op & mR) | sc, sU = s & (mR | cb);
      const uint32_t mD = __brev((op & mL) | sc), sD = __brev(s & (mL | cb));
      uint32_t sumU, sumD;
      addCarry(mU, sU, sumU, cU);
      addCarry(mD, sD, sumD, cD);
      const uint32_t cand = (((sumU ^ mU) & mU) | sU) | __brev(((sumD ^ mD) & mD) | sD);
      const uint32_t won = atomicAnd(&ringRow[j], ~cand) & cand;
      if (won) {
        uint32_t w = won;
        while (w) {
          const int b = __ffs(w) - 1;
          w &= w - 1u;
          const int x = 32 * j + b;
          row[x] = abs(x - g.gx) + base;
        }
        sweepPostAll(sh, g, tau, k, h, r, j, won, mR, mL, true);
      }
    }
    // a flood that reached the end of the word goes on in the next one: first
    // to the right, then (only the goal word can do both) to the left
    if (phase == 0) {
      leftCarry0 = cD;
      phase = 1;
    }
    if (phase == 1) {
      if (cU && j < 31) {
        ++j;
        extra = 1u;
        continue;
      }
      phase = 2;
      j = j0;
      cD = leftCarry0;
    }
    if (cD && j > 0) {
      --j;
      extra = 0x80000000u;
      continue;
    }
    break;
  }
}

template <bool kVec>
__global__ void __launch_bounds__(kSwThreads, 1) bfs_sweep_kernel(BfsSweepParams p) {
  extern __shared__ uint32_t smem[];
  uint32_t* const Sd = smem;                                    // [kSwDense][2][3][32]
  uint32_t* const ring = Sd + kSwDense * 2 * 3 * 32;            // [2][kSwRO][32]
  uint32_t* const planes = ring + 2 * kSwRO * 32;               // [2][kSwRP][kSwPlanes][32]
  uint32_t* const acc = planes + 2 * kSwRP * kSwPlanes * 32;    // [kSwSparseLevels][2][3][32]
  uint32_t* const cnt = acc + kSwSparseLevels * 2 * 3 * 32;     // [4] (+ padding to 8)
  uint16_t* const list = reinterpret_cast<uint16_t*>(cnt + 8);  // [4][kSwListCap]
  __shared__ int sGoal, sHandOver;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dimx = p.dimx, dimy = p.dimy;
  SweepShared sh;
  sh.ring = ring;
  sh.acc = acc;
  sh.list = list;
  sh.cnt = cnt;
  sh.handOver = &sHandOver;

  while (true) {
    __syncthreads();
    if (tid == 0) {
      sGoal = (int)atomicAdd(p.ws, 1u);
      sHandOver = 0;
    }
    for (int i = tid; i < kSwDense * 2 * 3 * 32; i += kSwThreads) Sd[i] = 0u;
    for (int i = tid; i < kSwSparseLevels * 2 * 3 * 32; i += kSwThreads) acc[i] = 0u;
    if (tid < 8) cnt[tid] = 0u;
    __syncthreads();
    const int gidx = sGoal;
    if (gidx >= p.n_goals) break;
    const int goal = p.goals[gidx];
    SweepGoal g;
    g.dimx = dimx;
    g.dimy = dimy;
    g.gy = goal / dimx;
    g.gx = goal - g.gy * dimx;
    g.jg = g.gx >> 5;
    g.out = p.out + (size_t)gidx * dimx * dimy;
    const int gx = g.gx, gy = g.gy, jg = g.jg, bg = gx & 31;
    g.hiG = bg == 31 ? 0u : kFull << (bg + 1);
    g.loG = (1u << bg) - 1u;
    g.cbG = 1u << bg;
    const uint32_t mR = lane > jg ? kFull : (lane == jg ? g.hiG : 0u);
    const uint32_t mL = lane < jg ? kFull : (lane == jg ? g.loG : 0u);
    const uint32_t cb = lane == jg ? g.cbG : 0u;
    const int maxr = max(gy, dimy - 1 - gy);
    const int writerEnd = maxr + 2 * kSwDense;  // the writers put out row maxr at step writerEnd - 1
    const bool goalFree = (sweepFreeWord(p.rowbits, p.WPR, dimy, gy, jg) >> bg) & 1u;

    // ---- roles ----
    const bool isDense = warp < 2 * kSwDense;
    const bool isWriter = !isDense && warp < 2 * kSwDense + kSwWriterWarps;
    const int dk = warp >> 1, dh = warp & 1;                       // dense: level, half
    const int wh = (warp - 2 * kSwDense) & 1, wp = (warp - 2 * kSwDense) >> 1;  // writer: half, part
    const int st = tid - (2 * kSwDense + kSwWriterWarps) * 32;     // sparse thread index
    uint32_t pf = 0u;      // prefetched free word (level 0, writers)
    uint32_t prevNw = 0u;  // dense: this level's cells of the previous row
    if (isDense && dk == 0 && dh == 0) pf = sweepFreeWord(p.rowbits, p.WPR, dimy, gy, lane);

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
          const bool inRange = y >= 0 && y < dimy;
          const int slot = r % 3;
          uint32_t* const Sk = Sd + (k * 2 + h) * 3 * 32;
          uint32_t* const ringRow = ring + (h * kSwRO + r % kSwRO) * 32;
          uint32_t op, sp = 0u, spu = 0u, npv = prevNw;
          if (k == 0) {
            op = pf;
            if (r == 0) npv = cb;
          } else {
            op = inRange ? ringRow[lane] : 0u;
            const uint32_t* Sp = Sd + ((k - 1) * 2 + h) * 3 * 32;
            if (r == 0) {
              sp = Sp[lane];
              spu = Sp[32 + lane] | Sp[3 * 32 + 32 + lane];  // row 1 of both halves
            } else {
              sp = Sp[slot * 32 + lane];
              spu = Sp[((r + 1) % 3) * 32 + lane];
            }
          }
          if (r == 1 && h == 1) npv = Sd[(k * 2) * 3 * 32 + lane];  // the goal row of this level (half 0, slot 0)
          const uint32_t nw = sweepRowStep(sp, spu, npv, op, mR, mL, cb, lane);
          prevNw = nw;
          Sk[slot * 32 + lane] = nw;
          if (inRange) {
            if (k == 0 || nw) ringRow[lane] = op & ~nw;
            uint32_t* pl = planes + ((h * kSwRP + r % kSwRP) * kSwPlanes) * 32 + lane;
            if (k == 0) {
#pragma unroll
              for (int b = 0; b < kSwPlanes; ++b) pl[b * 32] = 0u;
            } else if (nw) {
#pragma unroll
              for (int b = 0; b < kSwPlanes; ++b)
                if ((k >> b) & 1) pl[b * 32] |= nw;
            }
            if (k == kSwDense - 1 && nw && !(p.dbg & 1)) sweepPostAll(sh, g, tau, k, h, r, lane, nw, mR, mL, false);
          }
          act = nw != 0u;
        }
        // free word of the next row of this half (level 0 only)
        if (k == 0) pf = sweepFreeWord(p.rowbits, p.WPR, dimy, h ? gy - (r + 1) : gy + (r + 1), lane);
      } else if (isWriter) {
        const int h = wh, r = tau - 2 * kSwDense + 1;
        const int y = h ? gy - r : gy + r;
        if (r >= 0 && !(h == 1 && r == 0) && y >= 0 && y < dimy) {
          const uint32_t fr = pf;
          const uint32_t op = ring[(h * kSwRO + r % kSwRO) * 32 + lane];
          const uint32_t* pl = planes + ((h * kSwRP + r % kSwRP) * kSwPlanes) * 32 + lane;
          uint32_t n0 = pl[0], n1 = kSwPlanes > 1 ? pl[32] : 0u, n2 = kSwPlanes > 2 ? pl[64] : 0u, n3 = fr & ~op;
          // 4x4 bit transposition: nibble i of n_w = (visited, delta bits 2..0) of cell 4i + w
          uint32_t t;
          t = ((n0 >> 2) ^ n2) & 0x33333333u; n2 ^= t; n0 ^= t << 2;
          t = ((n1 >> 2) ^ n3) & 0x33333333u; n3 ^= t; n1 ^= t << 2;
          t = ((n0 >> 1) ^ n1) & 0x55555555u; n1 ^= t; n0 ^= t << 1;
          t = ((n2 >> 1) ^ n3) & 0x55555555u; n3 ^= t; n2 ^= t << 1;
          int32_t* row = g.out + (size_t)y * dimx;
          const int xb = 32 * lane - gx;
          const uint32_t nn[4] = {n0, n1, n2, n3};
#pragma unroll
          for (int ii = 0; ii < 8 / (kSwWriterWarps / 2); ++ii) {
            const int i = wp * (8 / (kSwWriterWarps / 2)) + ii;
            const int x0 = 32 * lane + 4 * i;
            int v[4];
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
      } else {
        // sparse levels: one thread per word item of this step
        const int n = min((int)cnt[tau & 3], kSwListCap);
        act = n > 0;
        // item i goes to lane i / #warps of sparse warp i % #warps: few items per warp, short divergent paths
        for (int i = (st >> 5) + kSwSparseWarps * (st & 31); i < n; i += kSwSparseWarps * 32)
          sweepItem(sh, g, tau, list[(tau & 3) * kSwListCap + i]);
        if (st == 0) cnt[(tau + 3) & 3] = 0u;  // the list of the previous step
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

constexpr size_t kSwSmemBytes = 4 * ((size_t)kSwDense * 2 * 3 * 32 + 2 * kSwRO * 32 + 2 * kSwRP * kSwPlanes * 32 +
                                     kSwSparseLevels * 2 * 3 * 32 + 8) +
                                2 * (size_t)4 * kSwListCap;

bool bfsSweepFits(const mrp_map_s* map) {
  // Opt-in: bit-exact, but measured 2.2-2.6x slower than the queue kernel on the
  // C5 map (profiles/README.md, round 2): the ~900 steps of a goal are a chain of
  // CTA-wide barriers, and a word item of a sparse level is a chain of four
  // shared-memory atomics.
  if (!getenv("MRP_BFS_SWEEP")) return false;
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
  p.WPR = bitmapRowWords(map->dimx);
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
