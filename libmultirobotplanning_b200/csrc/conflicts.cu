// conflicts.cu — vertex / edge-swap conflict detection and counting over packed
// path tables cell[N][Tpad] (+ len[N]).
//
// Replaces, on the reference side,
//   Environment::getFirstConflict   example/cbs.cpp:335-386 (bound size-1),
//                                   example/cbs_ta.cpp:369-420 (bound size)
//   Environment::focalHeuristic     example/ecbs.cpp:315-350
//   Environment::focalStateHeuristic / focalTransitionHeuristic
//                                   example/ecbs.cpp:282-312
// with getState()'s clamp to the last state (example/cbs.cpp:420-429).
//
// all-pairs kernel: a CTA owns a 64x64 block of agent pairs (upper triangle
// only) and a chunk of 64 timesteps.  Both 64-agent slabs of the table are
// staged in shared memory ([agent][t], odd stride), every thread keeps a 4x4
// register tile of pairs and walks the chunk in time, re-using the t+1 column
// as the next t column.  Per pair-step: 1 vertex compare and 2 edge compares.
// The first conflict is the minimum of the packed key (t, type, i, j) — the
// exact iteration order of the reference loops — reduced with a 64-bit
// atomicMin; counts are reduced per CTA and added with one atomic.
//
// hashed kernels (256 < N <= 4096, O(N*T) instead of O(N^2*T)): the table is
// transposed and clamped once, then one CTA per timestep; two kernels with
// identical answers, the second one is the default:
//   conflict_hash2_kernel  one table of (cell, agent) entries, one CAS per probe
//   conflict_sieve_kernel  hashed occupancy bitmaps pick the few agents that
//                          can be in a conflict; exact pass over those only
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace mrp {

constexpr int kPB = 64;        // agents per block side
constexpr int kTC = 64;        // timesteps per chunk
constexpr int kStride = kTC + 3;  // 67: odd => conflict-light column reads

__device__ __forceinline__ unsigned long long warpMin64Key(unsigned long long v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// result layout per table (4 x u64): [0] min key, [1] count, [2] max len, [3] -
__global__ void conflict_prep_kernel(const int32_t* __restrict__ len, int N,
                                     unsigned long long* __restrict__ result) {
  const int b = blockIdx.x;
  const int32_t* l = len + (size_t)b * N;
  __shared__ int smax[32];
  int m = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) m = max(m, l[i]);
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? smax[threadIdx.x] : 0;
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) {
      result[4 * b + 0] = kNoConflict;
      result[4 * b + 1] = 0ull;
      result[4 * b + 2] = (unsigned long long)m;
      result[4 * b + 3] = 0ull;
    }
  }
}

template <bool kFirst, bool kCount>
__global__ void __launch_bounds__(256)
conflict_pairs_kernel(const int32_t* __restrict__ cellAll,
                      const int32_t* __restrict__ lenAll, int N, int Tpad,
                      int mode, int nb, unsigned long long* __restrict__ resultAll,
                      int shard, int nShards) {
  __shared__ int32_t sA[kPB * kStride];
  __shared__ int32_t sB[kPB * kStride];
  __shared__ int sCount[8];

  const int table = blockIdx.z;
  const int32_t* cell = cellAll + (size_t)table * N * Tpad;
  const int32_t* len = lenAll + (size_t)table * N;
  unsigned long long* result = resultAll + 4 * (size_t)table;

  const int maxLen = (int)result[2];
  const int max_t = maxLen - (mode == 0 ? 1 : 0);
  const int t0 = blockIdx.y * kTC;
  if (t0 >= max_t) return;
  if (kFirst && !kCount) {
    // a conflict at an earlier chunk makes this chunk irrelevant
    const unsigned long long best = *(volatile unsigned long long*)&result[0];
    if (best != kNoConflict && (int)(best >> 41) < t0) return;
  }
  const int tEnd = min(t0 + kTC, max_t);  // timesteps [t0, tEnd)

  // upper-triangular block index -> (bi, bj), bi <= bj; with nShards > 1 the
  // agent-pair blocks are dealt round-robin and this launch sweeps every
  // nShards-th one (multi-GPU: conflict checks by agent-pair block)
  int bi = 0, rem = blockIdx.x * nShards + shard;
  if (rem >= nb * (nb + 1) / 2) return;
  while (rem >= nb - bi) {
    rem -= nb - bi;
    ++bi;
  }
  const int bj = bi + rem;

  // stage positions t0 .. tEnd (inclusive: the edge test needs t+1)
  const int nT = tEnd - t0 + 1;
  for (int idx = threadIdx.x; idx < kPB * (kTC + 1); idx += blockDim.x) {
    const int a = idx / (kTC + 1), k = idx - a * (kTC + 1);
    if (k >= nT) continue;
    const int t = t0 + k;
    int ia = bi * kPB + a, ib = bj * kPB + a;
    int va = -2 - a, vb = -2 - kPB - a;  // padding agents never match anything
    if (ia < N) {
      const int L = len[ia];
      if (L > 0) va = cell[(size_t)ia * Tpad + min(t, L - 1)];
    }
    if (ib < N) {
      const int L = len[ib];
      if (L > 0) vb = cell[(size_t)ib * Tpad + min(t, L - 1)];
    }
    sA[a * kStride + k] = va;
    sB[a * kStride + k] = vb;
  }
  __syncthreads();

  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const int i0 = bi * kPB + ti * 4, j0 = bj * kPB + tj * 4;
  const int32_t* pa = sA + (ti * 4) * kStride;
  const int32_t* pb = sB + (tj * 4) * kStride;

  // pair (k,l) is live iff i < j (only matters on diagonal blocks) and in range
  unsigned live = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int l = 0; l < 4; ++l)
      if (i0 + k < j0 + l && i0 + k < N && j0 + l < N) live |= 1u << (k * 4 + l);

  int cnt = 0;
  unsigned long long best = kNoConflict;
  if (live) {
    int a0[4], b0[4], a1[4], b1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a0[k] = pa[k * kStride];
      b0[k] = pb[k * kStride];
    }
    for (int k = 0; k < tEnd - t0; ++k) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a1[q] = pa[q * kStride + k + 1];
        b1[q] = pb[q * kStride + k + 1];
      }
      unsigned vhit = 0, ehit = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const unsigned bit = 1u << (q * 4 + l);
          if (a0[q] == b0[l]) vhit |= bit;
          if (a0[q] == b1[l] && a1[q] == b0[l]) ehit |= bit;
        }
      vhit &= live;
      ehit &= live;
      if (kCount) cnt += __popc(vhit) + __popc(ehit);
      if (kFirst && (vhit | ehit) && best == kNoConflict) {
        // first hit of this thread in time; vertex before edge, then (i, j)
        const unsigned hit = vhit ? vhit : ehit;
        const int p = __ffs(hit) - 1;  // bit order q*4+l == (i, j) ascending
        best = conflictKey(t0 + k, vhit ? 0 : 1, i0 + (p >> 2), j0 + (p & 3));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a0[q] = a1[q];
        b0[q] = b1[q];
      }
    }
  }
  if (kFirst && best != kNoConflict) atomicMin(&result[0], best);
  if (kCount) {
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) sCount[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += sCount[w];
      if (tot) atomicAdd(&result[1], (unsigned long long)tot);
    }
  }
}

// ---------------------------------------------------------------------------
// Large N: per-timestep hashing, O(N*T) instead of O(N^2*T).
//
// Same answers as the pair loops (example/cbs.cpp:343-383, ecbs.cpp:315-350):
//   vertex conflicts at t  = sum over cells of C(k, 2), k agents on the cell;
//   edge conflicts at t    = pairs {i, j} whose moves are reverse to each other
//                            (two agents resting on one cell count, as in the
//                            reference's test state1a==state2b && state1b==state2a);
//   first conflict         = min packed key (t, type, i, j): per cell the two
//                            smallest agent ids, per move the smallest reverse mover.
// One CTA per timestep; the N positions of that timestep come from a
// transposed, already clamped copy of the table (coalesced), and are inserted
// into an open-addressing table in shared memory (64-bit keys, atomicCAS).
// ---------------------------------------------------------------------------
constexpr int kHashThreads = 1024;
constexpr int kHashMaxN = 4096;          // table of 2*N slots x 16 B <= 128 KB
constexpr int kHashMinN = 257;
constexpr int kHashPerThread = kHashMaxN / kHashThreads;

// Row stride of the transposed table: rows start on 16-byte boundaries (the
// sieve kernel reads four agents per thread with one load); the pad entries
// are agents without a path.
__host__ __device__ __forceinline__ int rowStride(int N) { return (N + 3) & ~3; }

// Transposes and clamps the whole table (rows t = 0 .. Tpad; the sweep kernels
// stop at max_t themselves) and, in its first block, does the work of
// conflict_prep_kernel (result = {no conflict, 0, max len, 0}): one launch and
// no dependency of the transposition on the maximum.
__global__ void conflict_transpose_kernel(const int32_t* __restrict__ cell,
                                          const int32_t* __restrict__ len, int N, int Tpad,
                                          unsigned long long* __restrict__ result,
                                          int32_t* __restrict__ posT,
                                          unsigned char* __restrict__ todo, int tBase, int tEnd,
                                          int firstOnly) {
  // tile: 32 agents x 64 timesteps, 256 threads, eight elements per thread
  // (all eight loads are issued before the first one is used)
  __shared__ int32_t tile[32][65];
  __shared__ int smax[32];
  // rows tBase .. tEnd of the transposed table (a first-conflict-only sweep
  // goes through the table in windows of growing size: once a window has
  // found a conflict, the blocks of the later ones have nothing to do)
  if (firstOnly && tBase > 0) {
    const unsigned long long b = *(volatile unsigned long long*)&result[0];
    if (b != kNoConflict && (int)(b >> 41) < tBase) return;
  }
  const int t0 = tBase + blockIdx.y * 64, i0 = blockIdx.x * 32;
  const int ld = rowStride(N);
  {
    int L[4], v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + threadIdx.y + 8 * k;
      L[k] = i < N ? len[i] : 0;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + threadIdx.y + 8 * k;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = t0 + 32 * h + threadIdx.x;
        // agents without a path never match anything
        v[2 * k + h] = L[k] > 0 ? cell[(size_t)i * Tpad + min(t, L[k] - 1)] : -2 - i;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      tile[threadIdx.y + 8 * k][threadIdx.x] = v[2 * k];
      tile[threadIdx.y + 8 * k][threadIdx.x + 32] = v[2 * k + 1];
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = threadIdx.y + 8 * k;
    const int t = t0 + r, i = i0 + threadIdx.x;
    if (t <= tEnd && i < ld) posT[(size_t)t * ld + i] = tile[threadIdx.x][r];
  }
  // the hand-over flags of the sieve kernel start out clear
  if (todo && blockIdx.x == 0 && threadIdx.y < 2 && t0 + 32 * threadIdx.y + threadIdx.x < Tpad)
    todo[t0 + 32 * threadIdx.y + threadIdx.x] = 0;
  if (tBase == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
    const int tid = threadIdx.y * 32 + threadIdx.x;
    int m = 0;
    for (int i = tid; i < N; i += 256) m = max(m, len[i]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) smax[threadIdx.y] = m;
    __syncthreads();
    if (threadIdx.y == 0) {
      m = threadIdx.x < 8 ? smax[threadIdx.x] : 0;
      for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (threadIdx.x == 0) {
        result[0] = kNoConflict;
        result[1] = 0ull;
        result[2] = (unsigned long long)m;
        result[3] = 0ull;
      }
    }
  }
}

__device__ __forceinline__ uint32_t hash64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (uint32_t)k;
}

// ---- one table, one atomic per agent ----------------------------------------
// (A first generation spent three atomics per agent and phase on 16-byte slots
// (key, count, smallest id) and built a second table for the moves: 0.45 ms on
// the C5 table; removed in round 2.)  ONE table of 64-bit entries
// (cell << 32 | agent) per timestep:
//   * an agent is inserted with a single atomicCAS per probe; every entry of
//     its own cell that it walks past on the way to its free slot is an agent
//     that got there first, i.e. one vertex-conflict pair (i, j), seen exactly
//     once: by whichever of the two was inserted later;
//   * a swap partner of agent i (cells a -> b) stood on b at time t, so it is
//     one of the entries of cell b in the same table; its position at t + 1
//     comes from the transposed table (rowB[j] == a).  Two agents resting on
//     one cell match each other here as in the reference's test
//     (cbs.cpp:363-382).  Every pair is seen from both sides.
// 8 bytes per slot instead of 16 and 512 threads: three CTAs per SM at N = 4096.
constexpr int kHash2Threads = 512;
constexpr int kHash2PerThread = kHashMaxN / kHash2Threads;

__device__ __forceinline__ uint32_t hash32(uint32_t k) {
  k ^= k >> 16;
  k *= 0x85ebca6bu;
  k ^= k >> 13;
  k *= 0xc2b2ae35u;
  k ^= k >> 16;
  return k;
}

template <bool kFirst, bool kCount>
__global__ void __launch_bounds__(kHash2Threads, 3)
conflict_hash2_kernel(const int32_t* __restrict__ posT, int N, int mode, int H,
                      unsigned long long* __restrict__ result,
                      const unsigned char* __restrict__ todo) {
  extern __shared__ unsigned long long tab[];  // [H]
  __shared__ unsigned long long sBest[kHash2Threads / 32];
  __shared__ unsigned long long sSum[kHash2Threads / 32];
  const int maxLen = (int)result[2];
  const int max_t = maxLen - (mode == 0 ? 1 : 0);
  // One CTA per timestep, or (behind the sieve kernel) a small grid whose CTAs
  // stride over the timesteps and redo the ones that were handed over.
  if (todo && *(volatile unsigned long long*)&result[3] == 0ull) return;  // nothing was handed over
  for (int t = blockIdx.x; t < max_t; t += gridDim.x) {
  if (todo && !todo[t]) continue;
  if (kFirst && !kCount) {
    // an earlier conflict is already known: skip the step.  One thread looks, so that the
    // whole block takes the same branch around the barriers below.
    __shared__ int sSkip;
    if (threadIdx.x == 0) {
      const unsigned long long b = *(volatile unsigned long long*)&result[0];
      sSkip = b != kNoConflict && (int)(b >> 41) < t;
    }
    __syncthreads();
    const int skip = sSkip;
    __syncthreads();  // the next iteration writes sSkip again
    if (skip) continue;
  }
  const uint32_t mask = (uint32_t)H - 1u;
  const int32_t* rowA = posT + (size_t)t * rowStride(N);
  const int32_t* rowB = rowA + rowStride(N);
  const int tid = threadIdx.x;
  constexpr unsigned long long kEmpty = ~0ull;  // no agent has id 2^32 - 1

  int a[kHash2PerThread], b[kHash2PerThread];
#pragma unroll
  for (int k = 0; k < kHash2PerThread; ++k) {
    const int i = tid + k * kHash2Threads;
    a[k] = i < N ? rowA[i] : 0;
    b[k] = i < N ? rowB[i] : 0;
  }
  {
    ulonglong2* tab2 = reinterpret_cast<ulonglong2*>(tab);
    for (int s = tid; s < H / 2; s += kHash2Threads) tab2[s] = make_ulonglong2(kEmpty, kEmpty);
  }
  __syncthreads();
  unsigned long long best = kNoConflict;
  unsigned int pairs2 = 0;  // 2 * vertex pairs + swap pairs seen from this side
#pragma unroll
  for (int k = 0; k < kHash2PerThread; ++k) {
    const int i = tid + k * kHash2Threads;
    if (i >= N) continue;
    const unsigned long long entry = ((unsigned long long)(uint32_t)a[k] << 32) | (uint32_t)i;
    uint32_t s = hash32((uint32_t)a[k]) & mask;
    while (true) {
      const unsigned long long prev = atomicCAS(&tab[s], kEmpty, entry);
      if (prev == kEmpty) break;
      if ((uint32_t)(prev >> 32) == (uint32_t)a[k]) {
        const int j = (int)(uint32_t)prev;
        pairs2 += 2;
        if (kFirst) best = min(best, conflictKey(t, 0, min(i, j), max(i, j)));
      }
      s = (s + 1) & mask;
    }
  }
  // the CAS results above were consumed (returning atomics): every entry is
  // in the table before the barrier publishes it
  __threadfence_block();
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kHash2PerThread; ++k) {
    const int i = tid + k * kHash2Threads;
    if (i >= N) continue;
    uint32_t s = hash32((uint32_t)b[k]) & mask;
    while (true) {
      const unsigned long long cur = tab[s];
      if (cur == kEmpty) break;
      if ((uint32_t)(cur >> 32) == (uint32_t)b[k]) {
        const int j = (int)(uint32_t)cur;
        if (j != i && __ldg(rowB + j) == a[k]) {
          pairs2 += 1;
          if (kFirst) best = min(best, conflictKey(t, 1, min(i, j), max(i, j)));
        }
      }
      s = (s + 1) & mask;
    }
  }
  if (kFirst) {
    best = warpMin64Key(best);
    if ((tid & 31) == 0) sBest[tid >> 5] = best;
  }
  if (kCount) {
    unsigned long long sum = pairs2;
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((tid & 31) == 0) sSum[tid >> 5] = sum;
  }
  __syncthreads();
  if (tid == 0) {
    if (kFirst) {
      unsigned long long bb = kNoConflict;
      for (int w = 0; w < kHash2Threads / 32; ++w) bb = min(bb, sBest[w]);
      if (bb != kNoConflict) atomicMin(&result[0], bb);
    }
    if (kCount) {
      unsigned long long tot = 0;
      for (int w = 0; w < kHash2Threads / 32; ++w) tot += sSum[w];
      if (tot) atomicAdd(&result[1], tot / 2);
    }
  }
  __syncthreads();  // the table and the reduction slots are reused by the next timestep
  }
}

// ---- sieve: most agents of a timestep cannot be in any conflict ----------------
// Two hashed bitmaps per timestep (128 Ki bits each) sort the agents out before
// anything exact happens: `occ` has the bit of every occupied cell, `multi`
// the bits that were set twice.  An agent can only be part of a vertex
// conflict if the `multi` bit of its cell is set, and of a swap if it moves and
// the `occ` bit of its target cell is set (its partner stands there); both
// partners of a real conflict pass their test, so the exact single-table
// procedure of conflict_hash2_kernel restricted to these candidates (~6 % of
// the agents at N = 4096: hash collisions plus the real ones) finds every
// pair.  Setting a bit is a plain ATOMS.OR (4 wavefronts per warp instruction;
// an ATOMS.CAS.64 takes one per lane) and the probing loops, whose length is
// the maximum over the lanes of a warp, run over a compacted candidate list.
// 41 KB of shared memory per CTA instead of 64 KB.  A timestep with more than
// kSieveCand candidates (dense pile-ups) is handed to conflict_hash2_kernel
// through `todo`.
#ifndef MRP_SIEVE_PER
#define MRP_SIEVE_PER 8
#endif
constexpr int kSievePer = MRP_SIEVE_PER;                // agents per thread (4 or 8)
constexpr int kSieveThreads = kHashMaxN / kSievePer;    // 512 threads, four CTAs per SM (4 per thread: 0.068 ms, 8: 0.064 ms)
constexpr int kSieveBits = 1 << 17;  // 3 % of the cells of a timestep collide at N = 4096
constexpr int kSieveCand = 512;
constexpr int kFirstWindow = 128;  // first-conflict-only sweeps: see launchConflicts
constexpr int kSieveSlots = 1024;

template <bool kFirst, bool kCount>
__global__ void __launch_bounds__(kSieveThreads, 2048 / kSieveThreads)
conflict_sieve_kernel(const int32_t* __restrict__ posT, int N, int mode,
                      unsigned long long* __restrict__ result, unsigned char* __restrict__ todo,
                      int tBase) {
  __shared__ __align__(16) uint32_t occ[kSieveBits / 32];
  __shared__ __align__(16) uint32_t multi[kSieveBits / 32];
  __shared__ __align__(16) unsigned long long tab[kSieveSlots];
  __shared__ uint16_t cand[kSieveCand];
  __shared__ int sN;
  __shared__ unsigned int sPairs2;
  const int t = tBase + blockIdx.x;
  const int maxLen = (int)result[2];
  const int max_t = maxLen - (mode == 0 ? 1 : 0);
  if (t >= max_t) return;
  if (kFirst && !kCount) {
    // block-uniform early exit: one thread reads the best key so far
    __shared__ int sSkip;
    if (threadIdx.x == 0) {
      const unsigned long long b = *(volatile unsigned long long*)&result[0];
      sSkip = b != kNoConflict && (int)(b >> 41) < t;
    }
    __syncthreads();
    if (sSkip) return;
  }
  const int ld = rowStride(N);
  const int32_t* rowA = posT + (size_t)t * ld;
  const int32_t* rowB = rowA + ld;
  const int tid = threadIdx.x, lane = tid & 31;
  constexpr unsigned long long kEmpty = ~0ull;
  constexpr uint32_t kSlotMask = kSieveSlots - 1;
  // multiplicative hashes: the top 17 bits of one product for the bitmaps, the
  // top bits of another for the exact table
  auto bitOf = [](int c) { return ((uint32_t)c * 0x9E3779B1u) >> 15; };
  auto slotOf = [](int c) { return (((uint32_t)c * 0x85EBCA6Bu) >> 20) & kSlotMask; };
  static_assert(kSieveBits == 1 << 17, "bitOf keeps 17 bits");

  // thread tid owns kSievePer consecutive agents of the (padded) row: 16-byte
  // loads, no loop (N <= kSievePer * kSieveThreads)
  int a[kSievePer], b[kSievePer];
#pragma unroll
  for (int q = 0; q < kSievePer / 4; ++q) {
    int4 A = make_int4(-1, -1, -1, -1), B = A;
    if (kSievePer * tid + 4 * q < ld) {
      A = reinterpret_cast<const int4*>(rowA)[(kSievePer / 4) * tid + q];
      B = reinterpret_cast<const int4*>(rowB)[(kSievePer / 4) * tid + q];
    }
    a[4 * q] = A.x, a[4 * q + 1] = A.y, a[4 * q + 2] = A.z, a[4 * q + 3] = A.w;
    b[4 * q] = B.x, b[4 * q + 1] = B.y, b[4 * q + 2] = B.z, b[4 * q + 3] = B.w;
  }
  {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u), e = make_uint4(~0u, ~0u, ~0u, ~0u);
    for (int s = tid; s < kSieveBits / 128; s += kSieveThreads) {
      reinterpret_cast<uint4*>(occ)[s] = z;
      reinterpret_cast<uint4*>(multi)[s] = z;
    }
    for (int s = tid; s < kSieveSlots / 2; s += kSieveThreads) reinterpret_cast<uint4*>(tab)[s] = e;
  }
  if (tid == 0) {
    sN = 0;
    sPairs2 = 0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSievePer; ++k) {
    if (a[k] < 0) continue;  // an agent without a path (or a pad entry)
    const uint32_t h = bitOf(a[k]), bit = 1u << (h & 31);
    if (atomicOr(&occ[h >> 5], bit) & bit) atomicOr(&multi[h >> 5], bit);
  }
  __syncthreads();
  {
    uint32_t cm = 0;  // candidates among this thread's agents
#pragma unroll
    for (int k = 0; k < kSievePer; ++k) {
      if (a[k] < 0) continue;
      const uint32_t ha = bitOf(a[k]);
      uint32_t c = (multi[ha >> 5] >> (ha & 31)) & 1u;
      if (!c && a[k] != b[k]) {
        const uint32_t hb = bitOf(b[k]);
        c = (occ[hb >> 5] >> (hb & 31)) & 1u;
      }
      cm |= c << k;
    }
    if (__any_sync(0xffffffffu, cm != 0)) {
      const int cnt = __popc(cm);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      int base = 0;
      if (lane == 31) base = atomicAdd(&sN, incl);
      base = __shfl_sync(0xffffffffu, base, 31);
      int pos = base + incl - cnt;
#pragma unroll
      for (int k = 0; k < kSievePer; ++k)
        if ((cm >> k) & 1u) {
          if (pos < kSieveCand) cand[pos] = (uint16_t)(kSievePer * tid + k);
          ++pos;
        }
    }
  }
  __syncthreads();
  const int nc = sN;
  if (nc > kSieveCand) {
    if (tid == 0) {
      todo[t] = 1;
      atomicAdd(&result[3], 1ull);  // number of timesteps handed over (0: the next kernel exits at once)
    }
    return;
  }
  unsigned long long best = kNoConflict;
  unsigned int pairs2 = 0;  // 2 * vertex pairs + swap pairs seen from this side
  for (int c = tid; c < nc; c += kSieveThreads) {
    const int i = cand[c];
    const int a = rowA[i];
    const unsigned long long entry = ((unsigned long long)(uint32_t)a << 32) | (uint32_t)i;
    uint32_t s = slotOf(a);
    while (true) {
      const unsigned long long prev = atomicCAS(&tab[s], kEmpty, entry);
      if (prev == kEmpty) break;
      if ((uint32_t)(prev >> 32) == (uint32_t)a) {
        const int j = (int)(uint32_t)prev;
        pairs2 += 2;
        if (kFirst) best = min(best, conflictKey(t, 0, min(i, j), max(i, j)));
      }
      s = (s + 1) & kSlotMask;
    }
  }
  __threadfence_block();
  __syncthreads();
  for (int c = tid; c < nc; c += kSieveThreads) {
    const int i = cand[c];
    const int a = rowA[i], b = rowB[i];
    uint32_t s = slotOf(b);
    while (true) {
      const unsigned long long cur = tab[s];
      if (cur == kEmpty) break;
      if ((uint32_t)(cur >> 32) == (uint32_t)b) {
        const int j = (int)(uint32_t)cur;
        if (j != i && __ldg(rowB + j) == a) {
          pairs2 += 1;
          if (kFirst) best = min(best, conflictKey(t, 1, min(i, j), max(i, j)));
        }
      }
      s = (s + 1) & kSlotMask;
    }
  }
  // only the warps that held candidates have anything to report; the doubled
  // pair counts of a timestep meet in shared memory before they are halved
  if ((tid & ~31) < nc) {
    if (kFirst) {
      best = warpMin64Key(best);
      if (lane == 0 && best != kNoConflict) atomicMin(&result[0], best);
    }
    if (kCount) {
      const unsigned int s2 = __reduce_add_sync(0xffffffffu, pairs2);
      if (lane == 0 && s2) atomicAdd(&sPairs2, s2);
    }
  }
  if (kCount) {
    __syncthreads();
    if (tid == 0 && sPairs2) atomicAdd(&result[1], (unsigned long long)(sPairs2 / 2));
  }
}

// focal counts: one warp per candidate move, lanes stride over the agents
__global__ void focal_counts_kernel(const int32_t* __restrict__ cell,
                                    const int32_t* __restrict__ len, int N,
                                    int Tpad, int self,
                                    const int32_t* __restrict__ ct,
                                    const int32_t* __restrict__ cfrom,
                                    const int32_t* __restrict__ cto, int n_cand,
                                    int32_t* __restrict__ stateCnt,
                                    int32_t* __restrict__ transCnt) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= n_cand) return;
  const int t = ct[k], from = cfrom[k], to = cto[k];
  int s = 0, tr = 0;
  for (int i = lane; i < N; i += 32) {
    const int L = len[i];
    if (i == self || L <= 0) continue;
    const int pa = cell[(size_t)i * Tpad + min(t, L - 1)];
    const int pb = cell[(size_t)i * Tpad + min(t + 1, L - 1)];
    s += (pb == to);
    tr += (from == pb && to == pa);
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    tr += __shfl_xor_sync(0xffffffffu, tr, o);
  }
  if (lane == 0) {
    stateCnt[k] = s;
    transCnt[k] = tr;
  }
}

static int launchPairs(const int32_t* d_cell, const int32_t* d_len, int B, int N,
                       int Tpad, int mode, bool wantFirst, bool wantCount,
                       unsigned long long* d_result, cudaStream_t st, int shard = 0, int nShards = 1) {
  conflict_prep_kernel<<<B, 256, 0, st>>>(d_len, N, d_result);
  countLaunch();
  const int nb = (N + kPB - 1) / kPB;
  const int nPairBlocks = (nb * (nb + 1) / 2 + nShards - 1) / nShards;
  const int chunks = (Tpad + kTC - 1) / kTC;
  dim3 grid(nPairBlocks, chunks, B);
  if (wantFirst && wantCount)
    conflict_pairs_kernel<true, true><<<grid, 256, 0, st>>>(d_cell, d_len, N, Tpad,
                                                            mode, nb, d_result, shard, nShards);
  else if (wantFirst)
    conflict_pairs_kernel<true, false><<<grid, 256, 0, st>>>(d_cell, d_len, N, Tpad,
                                                             mode, nb, d_result, shard, nShards);
  else
    conflict_pairs_kernel<false, true><<<grid, 256, 0, st>>>(d_cell, d_len, N, Tpad,
                                                             mode, nb, d_result, shard, nShards);
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  return 0;
}

size_t conflictsWorkspaceBytes(int N, int Tpad) {
  if (N < kHashMinN || N > kHashMaxN) return 0;
  // transposed table, then one hand-over flag per timestep
  return (size_t)(Tpad + 1) * rowStride(N) * 4 + (((size_t)Tpad + 255) & ~(size_t)255);
}

int launchConflicts(const int32_t* d_cell, const int32_t* d_len, int N, int Tpad,
                    int mode, bool wantFirst, bool wantCount,
                    unsigned long long* d_result, void* d_ws, size_t wsBytes,
                    cudaStream_t st) {
  const size_t need = conflictsWorkspaceBytes(N, Tpad);
  if (need == 0 || d_ws == nullptr || wsBytes < need || getenv("MRP_CONFLICTS_ALLPAIRS"))
    return launchPairs(d_cell, d_len, 1, N, Tpad, mode, wantFirst, wantCount, d_result, st);
  // hashed path: prep (max len) -> transpose + clamp -> one CTA per timestep
  int32_t* posT = static_cast<int32_t*>(d_ws);
  // hand-over flags of the sieve kernel, cleared by the transposition
  unsigned char* todoAll = static_cast<unsigned char*>(d_ws) + (size_t)(Tpad + 1) * rowStride(N) * 4;
  const bool sieve = !getenv("MRP_CONFLICTS_HASH2");
  auto transpose = [&](int tBase, int tEnd, int firstOnly) {
    dim3 tg((rowStride(N) + 31) / 32, (tEnd - tBase + 1 + 63) / 64);
    conflict_transpose_kernel<<<tg, dim3(32, 8), 0, st>>>(d_cell, d_len, N, Tpad, d_result, posT,
                                                            todoAll, tBase, tEnd, firstOnly);
  };
  int H = 512;
  while (H < 2 * N) H <<= 1;
  {
    // default: sieve kernel, then the single-table kernel on the timesteps it
    // handed over (MRP_CONFLICTS_HASH2: single-table kernel on every timestep)
    unsigned char* todo = nullptr;
    if (sieve) {
      todo = todoAll;
      auto sweep = [&](int tBase, int tEnd) {
        const int grid = tEnd - tBase;
        if (wantFirst && wantCount)
          conflict_sieve_kernel<true, true><<<grid, kSieveThreads, 0, st>>>(posT, N, mode, d_result, todo, tBase);
        else if (wantFirst)
          conflict_sieve_kernel<true, false><<<grid, kSieveThreads, 0, st>>>(posT, N, mode, d_result, todo, tBase);
        else
          conflict_sieve_kernel<false, true><<<grid, kSieveThreads, 0, st>>>(posT, N, mode, d_result, todo, tBase);
        countLaunch(2);
      };
      if (wantFirst && !wantCount) {
        // first conflict only: the first kFirstWindow timesteps, then the rest
        // (whose blocks exit at once if the first window found a conflict; more
        // windows cost more in launches than they save: four windows took
        // 0.055 ms on the C5 table against 0.045-0.050 ms for a single sweep)
        for (int t0 = 0, w = kFirstWindow; t0 < Tpad; t0 += w, w = Tpad) {
          const int t1 = std::min(Tpad, t0 + w);
          transpose(t0, t1, 1);
          sweep(t0, t1);
        }
      } else {
        transpose(0, Tpad, 0);
        sweep(0, Tpad);
      }
    } else {
      transpose(0, Tpad, 0);
      countLaunch();
    }
    const size_t smem2 = (size_t)H * 8;
    auto run2 = [&](auto kern) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      // behind the sieve kernel only a few timesteps (usually none) are left
      const int grid = todo ? std::min(Tpad, 2 * 148) : Tpad;
      kern<<<grid, kHash2Threads, smem2, st>>>(posT, N, mode, H, d_result, todo);
    };
    if (wantFirst && wantCount)
      run2(conflict_hash2_kernel<true, true>);
    else if (wantFirst)
      run2(conflict_hash2_kernel<true, false>);
    else
      run2(conflict_hash2_kernel<false, true>);
    countLaunch();
    MRP_CUDA(cudaGetLastError());
    return 0;
  }
}

// this rank's share of the agent-pair blocks of one table (all-pairs kernel);
// the ranks' results meet in an all-reduce (multi.cu)
int launchConflictsPairShard(const int32_t* d_cell, const int32_t* d_len, int N, int Tpad, int mode,
                             bool wantFirst, bool wantCount, unsigned long long* d_result, int shard,
                             int nShards, cudaStream_t st) {
  return launchPairs(d_cell, d_len, 1, N, Tpad, mode, wantFirst, wantCount, d_result, st, shard, nShards);
}

int launchConflictsBatch(const int32_t* d_cell, const int32_t* d_len, int B,
                         int N, int Tpad, int mode,
                         unsigned long long* d_result, cudaStream_t st) {
  return launchPairs(d_cell, d_len, B, N, Tpad, mode, true, true, d_result, st);
}

int launchFocalCounts(const int32_t* d_cell, const int32_t* d_len, int N,
                      int Tpad, int self, const int32_t* d_t,
                      const int32_t* d_from, const int32_t* d_to, int n_cand,
                      int32_t* d_state, int32_t* d_trans, cudaStream_t st) {
  if (n_cand <= 0) return 0;
  const int threads = 256;
  const int blocks = (n_cand * 32 + threads - 1) / threads;
  focal_counts_kernel<<<blocks, threads, 0, st>>>(d_cell, d_len, N, Tpad, self,
                                                  d_t, d_from, d_to, n_cand,
                                                  d_state, d_trans);
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mrp
