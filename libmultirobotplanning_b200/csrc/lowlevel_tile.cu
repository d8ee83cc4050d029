// lowlevel_tile.cu — low-latency replans on single-tile maps (<= 32x32 cells:
// the benchmark sets of the reference, configs C1-C4) with cbs / ecbs moves.
//
// Same searches as lowlevel.cu (AStar::search, a_star.hpp:63-161;
// AStarEpsilon::search, a_star_epsilon.hpp:86-285; the Environment callbacks of
// example/cbs.cpp:266-333,431-444 and example/ecbs.cpp:282-312), same selection
// rule, same tie-breaking, same node numbering: results are identical, node for
// node.  What changes is where an expansion finds its data.  A batch of replans
// lasts as long as its longest search, and a search is one dependent chain of
// expansions, so the time of ONE expansion is what a lock-step iteration of the
// high-level drivers pays.  The general kernel spends ~6000 cycles per expansion:
// ~440 instructions issued one every ~5 cycles, plus the global round trips of
// the hash probe, the heuristic and the other agents' positions
// (profiles/lowlevel_r01.txt).  Here, per warp (one warp per CTA, everything in
// shared memory):
//   * closed/open membership = one bit per space-time state, vis[t][row] (a
//     32-cell row is one word): the probe is one shared-memory atomicOr.  With
//     unit move costs g == t, a state seen before can never be improved, so no
//     node index has to be found (a_star.hpp:133-143 never fires);
//   * vertex constraints are marked in vis[] before the search starts (a
//     constrained state is never generated: stateValid, cbs.cpp:431-436); edge
//     constraints go through a 1024-bit Bloom filter held in one register per
//     lane, the exact list is walked only on a hit (transitionValid, :438-444);
//   * the distance field of the goal (the admissible heuristic) is staged once
//     per job as 16-bit values; the 32 row masks of the map likewise;
//   * the first 1024 nodes keep (state, parent) in shared memory, OPEN its first
//     768 entries; nothing else is stored per node: f, g and the focal value
//     travel in the packed OPEN key;
//   * focal values (focalStateHeuristic / focalTransitionHeuristic,
//     ecbs.cpp:282-312): a small kernel in front builds, per path table, two
//     occupancy planes per time step (cells holding >= 1 and >= 2 agents).  The
//     vertex count of a successor is then two bit tests minus this agent's own
//     old position; the swap count can only be non-zero when both cells of the
//     move are occupied at the right times, and only then (a few percent of the
//     expansions) the exact pass over the other agents' rows runs.
// A search that leaves what shared memory holds (time steps beyond the bitmap,
// f >= 512, more than 96 constraints, no goal) ends with kTileStatusRedo and is
// redone by the general kernel.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "lowlevel.cuh"

namespace mrp {

constexpr int kOpenT = 768;   // OPEN entries in shared memory
constexpr int kNodeT = 1024;  // nodes with (state, parent) in shared memory
constexpr int kFMaxT = 512;   // f values tracked by the histogram
constexpr int kOccSmemMax = 160 * 1024;  // shared memory of focal_occ_kernel (Tpad * 256 B)
constexpr int kSelfT = 256;   // time steps of this agent's old path kept in shared memory

struct TileLayout {
  int openKey, vis, openState, nodeKey, field, nodePar, hist, rows, cons, selfRow, stage, total;
};
__host__ __device__ inline TileLayout tileLayout(int TB) {
  TileLayout L;
  int o = 0;
  L.openKey = o;   o += kOpenT * 8;
  L.vis = o;       o += TB * 32 * 4;
  L.openState = o; o += kOpenT * 4;
  L.nodeKey = o;   o += kNodeT * 4;
  L.hist = o;      o += kFMaxT * 2;
  L.field = o;     o += 1024 * 2;
  L.nodePar = o;   o += kNodeT * 2;
  L.rows = o;      o += 32 * 4;
  L.cons = o;      o += kConsCache * 3 * 4;
  L.selfRow = o;   o += kSelfT * 2;
  L.stage = o;     o += 32 * 32;  // 3 x 8 B of occupancy words per lane, in 32-byte slots
  L.total = (o + 15) & ~15;
  return L;
}

// Occupancy planes of one path table: occ[t][y] = (cells of row y holding >= 1
// agent at time t, cells holding >= 2); every agent with a path counts, parked
// agents stay on their last cell (getState, example/cbs.cpp:420-429).
__global__ void __launch_bounds__(256) focal_occ_kernel(const int32_t* __restrict__ tables,
                                                        const int32_t* __restrict__ tableLen, int N,
                                                        int Tpad, int dimx, uint2* __restrict__ occ,
                                                        int32_t* __restrict__ occMany) {
  extern __shared__ uint32_t sOcc[];  // [Tpad*32] plane 1, [Tpad*32] plane 2
  __shared__ int sMany;
  const int words = Tpad * 32;
  uint32_t* o1 = sOcc;
  uint32_t* o2 = sOcc + words;
  for (int i = threadIdx.x; i < 2 * words; i += blockDim.x) sOcc[i] = 0;
  if (threadIdx.x == 0) sMany = 0;
  __syncthreads();
  const int32_t* tab = tables + (size_t)blockIdx.x * N * Tpad;
  const int32_t* len = tableLen + (size_t)blockIdx.x * N;
  for (int idx = threadIdx.x; idx < N * Tpad; idx += blockDim.x) {
    const int a = idx / Tpad, t = idx - a * Tpad;
    const int L = len[a];
    if (L <= 0) continue;
    const int c = tab[(size_t)a * Tpad + min(t, L - 1)];
    const int x = c % dimx, y = c / dimx;
    const uint32_t bit = 1u << x;
    const uint32_t old1 = atomicOr(&o1[t * 32 + y], bit);
    if (old1 & bit) {
      const uint32_t old2 = atomicOr(&o2[t * 32 + y], bit);
      if (old2 & bit) sMany = 1;
    }
  }
  __syncthreads();
  uint2* dst = occ + (size_t)blockIdx.x * words;
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = make_uint2(o1[i], o2[i]);
  if (threadIdx.x == 0) occMany[blockIdx.x] = sMany;
}

// ---- shared-memory accesses by 32-bit shared-window address ----
// (the search loop is one dependent chain: explicit addresses keep every access a
// single LDS / STS / ATOMS without generic-address arithmetic in front of it; the
// statements are volatile, so they keep their order among themselves)
__device__ __forceinline__ uint32_t sLd32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t sLdU16(uint32_t a) {
  uint32_t v;
  asm volatile("{\n\t.reg .u16 h;\n\tld.shared.u16 h, [%1];\n\tcvt.u32.u16 %0, h;\n\t}" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned long long sLd64(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sSt32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sStU16(uint32_t a, uint32_t v) {
  asm volatile("{\n\t.reg .u16 h;\n\tcvt.u16.u32 h, %1;\n\tst.shared.u16 [%0], h;\n\t}" ::"r"(a), "r"(v));
}
__device__ __forceinline__ void sSt64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v)); }
__device__ __forceinline__ uint32_t sAtomOr(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v));
  return old;
}
__device__ __forceinline__ void sAtomAdd(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v)); }
// 8-byte asynchronous copy global -> shared (completion tracked by the copy group, not by
// a scoreboard: instructions in between never wait for it)
__device__ __forceinline__ void cpAsync8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// lane numbers of the set bits of a 5-bit mask, ascending, 3 bits each
__constant__ unsigned short kNthSet[32] = {
    0,     0,     1,     0 | 1 << 3, 2,     0 | 2 << 3, 1 | 2 << 3, 0 | 1 << 3 | 2 << 6,
    3,     0 | 3 << 3, 1 | 3 << 3, 0 | 1 << 3 | 3 << 6, 2 | 3 << 3, 0 | 2 << 3 | 3 << 6, 1 | 2 << 3 | 3 << 6,
    0 | 1 << 3 | 2 << 6 | 3 << 9,
    4,     0 | 4 << 3, 1 | 4 << 3, 0 | 1 << 3 | 4 << 6, 2 | 4 << 3, 0 | 2 << 3 | 4 << 6, 1 | 2 << 3 | 4 << 6,
    0 | 1 << 3 | 2 << 6 | 4 << 9,
    3 | 4 << 3, 0 | 3 << 3 | 4 << 6, 1 | 3 << 3 | 4 << 6, 0 | 1 << 3 | 3 << 6 | 4 << 9, 2 | 3 << 3 | 4 << 6,
    0 | 2 << 3 | 3 << 6 | 4 << 9, 1 | 2 << 3 | 3 << 6 | 4 << 9, 0 | 1 << 3 | 2 << 6 | 3 << 9 | 4 << 12};

__device__ __forceinline__ uint32_t edgeHash(int t, int from, int to) {
  return hashState((uint32_t)t * 0x9E3779B1u ^ ((uint32_t)from << 11) ^ (uint32_t)to);
}

__global__ void __launch_bounds__(32) lowlevel_tile_kernel(LLParams p) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const TileLayout lay = tileLayout(p.TB);
  uint32_t* vis = reinterpret_cast<uint32_t*>(smemRaw + lay.vis);
  uint32_t* hist32 = reinterpret_cast<uint32_t*>(smemRaw + lay.hist);
  unsigned short* fieldS = reinterpret_cast<unsigned short*>(smemRaw + lay.field);
  uint32_t* rowsS = reinterpret_cast<uint32_t*>(smemRaw + lay.rows);
  int32_t* cons = reinterpret_cast<int32_t*>(smemRaw + lay.cons);
  unsigned short* selfRow = reinterpret_cast<unsigned short*>(smemRaw + lay.selfRow);
  // the same regions as 32-bit shared-window addresses (search loop)
  const uint32_t sBase = (uint32_t)__cvta_generic_to_shared(smemRaw);
  const uint32_t aOpen = sBase + lay.openKey, aVis = sBase + lay.vis, aOpenState = sBase + lay.openState,
                 aNodeKey = sBase + lay.nodeKey, aHist = sBase + lay.hist, aField = sBase + lay.field,
                 aNodePar = sBase + lay.nodePar, aRows = sBase + lay.rows, aCons = sBase + lay.cons,
                 aSelf = sBase + lay.selfRow, aStage = sBase + lay.stage;

  const int lane = threadIdx.x;
  const int slot = blockIdx.x;
  const int TB = p.TB;
  const int dimx = p.dimx, dimy = p.dimy;
  const int cells = dimx * dimy;
  // what does not fit shared memory: per warp slot, or in the job's state blob
  uint32_t* const nodeKeySlot = p.nodeKey + (size_t)slot * p.maxNodes;
  int32_t* const nodeParSlot = p.nodeParent + (size_t)slot * p.maxNodes;
  unsigned long long* const openSlot = p.openKey + (size_t)slot * p.maxNodes;
  const int imageBytes = lay.total;

  // a cell of the map (x + dimx*y) as tile position (y*32 + x) and back
  auto toTile = [&](int c) -> int { return dimx == 32 ? c : ((c / dimx) << 5) | (c % dimx); };
  auto toCell = [&](int tile) -> int { return dimx == 32 ? tile : (tile & 31) + dimx * (tile >> 5); };
  int visTop = TB;  // rows of vis[] that may hold bits of the previous job

  while (true) {
    int idx = 0;
    if (lane == 0) idx = (int)atomicAdd(p.counter, 1u);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= p.n_jobs) break;
    const int job = p.jobList ? p.jobList[idx] : idx;
    const mrp_job jb = p.jobs[job];
    const uint32_t* bits = p.mapBits[jb.map];
    const int goal = jb.goal_cell;
    const int32_t* field = (jb.field >= 0) ? p.fields + (size_t)jb.field * cells : nullptr;
    const int nVc = jb.vc_end - jb.vc_begin, nEc = jb.ec_end - jb.ec_begin;
    const int32_t* vc = p.vc + 2 * (size_t)jb.vc_begin;
    const int32_t* ec = p.ec + 3 * (size_t)jb.ec_begin;
    const int32_t* tab = (jb.table >= 0 && p.tables) ? p.tables + (size_t)jb.table * p.N * p.Tpad : nullptr;
    const int32_t* tlen = tab ? p.tableLen + (size_t)jb.table * p.N : nullptr;
    const bool wantFocal = p.focalMode && tab;

    int status = -1;  // running
    if (goal < 0 || nVc + nEc > kConsCache) status = kTileStatusRedo;
    // sliced search: blob = [64-byte header][image of the shared-memory regions][spill arrays]
    const int stateId = p.jobState ? p.jobState[job] : -1;
    unsigned char* blob = stateId >= 0 ? p.stateChunks[stateId >> kStateChunkBits] +
                                             (size_t)(stateId & (kStateChunk - 1)) * p.blobBytes
                                       : nullptr;
    const bool resume = blob && p.jobResume[job] != 0;
    uint32_t* nodeKeyG = nodeKeySlot;
    int32_t* nodeParG = nodeParSlot;
    unsigned long long* openG = openSlot;
    if (blob) {
      unsigned char* sp = blob + 64 + imageBytes;
      openG = reinterpret_cast<unsigned long long*>(sp);
      nodeKeyG = reinterpret_cast<uint32_t*>(sp + (size_t)p.maxNodes * 8);
      nodeParG = reinterpret_cast<int32_t*>(sp + (size_t)p.maxNodes * 12);
    }

    // ---- per-job setup ----
    __syncwarp();
    rowsS[lane] = bits[lane];
    {
      const uint4 z = make_uint4(0, 0, 0, 0);
      uint4* v4 = reinterpret_cast<uint4*>(vis);
      for (int i = lane; i < visTop * 8; i += 32) v4[i] = z;
      uint4* h4 = reinterpret_cast<uint4*>(hist32);
      for (int i = lane; i < kFMaxT * 2 / 16; i += 32) h4[i] = z;
    }
    if (field)
      for (int c = lane; c < cells; c += 32) {
        const int v = field[c];
        fieldS[toTile(c)] = v == MRP_INF ? (unsigned short)0xFFFF : (unsigned short)v;
      }
    // this agent's old path: its positions are part of the occupancy planes
    int selfLen = 0;
    bool occOK = false;
    const uint2* occ = nullptr;
    if (wantFocal) {
      if (jb.self >= 0 && jb.self < p.N) selfLen = tlen[jb.self];
      occOK = p.occ != nullptr && p.occMany[jb.table] == 0 && selfLen <= kSelfT;
      occ = p.occ + (size_t)jb.table * p.Tpad * 32;
      if (occOK)
        for (int t = lane; t < selfLen; t += 32)
          selfRow[t] = (unsigned short)toTile(tab[(size_t)jb.self * p.Tpad + t]);
    }
    __syncwarp();
    // constraints: cached, vertex constraints marked as visited, edge constraints
    // hashed into the Bloom register; lastGoalConstraint (example/cbs.cpp:266-276)
    // and the time after which no constraint can apply any more
    int lastGoal = -1, tFree = 0, tMark = 0;
    uint32_t bloom = 0;
    if (status == -1) {
      for (int i = lane; i < 2 * nVc; i += 32) cons[i] = vc[i];
      for (int i = lane; i < 3 * nEc; i += 32) cons[2 * nVc + i] = ec[i];
      __syncwarp();
      for (int i = lane; i < nVc; i += 32) {
        const int t = cons[2 * i], c = cons[2 * i + 1];
        if (c == goal) lastGoal = max(lastGoal, t);
        tFree = max(tFree, t);
        if (t >= 0 && t < TB && c >= 0 && c < cells) {
          const int tl = toTile(c);
          atomicOr(&vis[t * 32 + (tl >> 5)], 1u << (tl & 31));
          tMark = max(tMark, t);
        }
      }
      for (int i = 0; i < nEc; ++i) {
        const int t = cons[2 * nVc + 3 * i];
        tFree = max(tFree, t + 1);
        const uint32_t h = edgeHash(t, cons[2 * nVc + 3 * i + 1], cons[2 * nVc + 3 * i + 2]) & 1023u;
        if ((int)(h >> 5) == lane) bloom |= 1u << (h & 31u);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        lastGoal = max(lastGoal, __shfl_xor_sync(0xffffffffu, lastGoal, o));
        tFree = max(tFree, __shfl_xor_sync(0xffffffffu, tFree, o));
        tMark = max(tMark, __shfl_xor_sync(0xffffffffu, tMark, o));
      }
    }
    __syncwarp();
    const int goalTile = goal >= 0 ? toTile(goal) : -1;
    const int goalX = goalTile & 31, goalY = goalTile >> 5;
    const uint32_t aEc = aCons + 8u * (uint32_t)nVc;  // edge constraints: (t, from, to)

    // admissible heuristic: the reference's value (Manhattan, example/cbs.cpp:278-284,
    // or the distance field) raised to the time bound of the goal test (see lowlevel.cu)
    auto heur = [&](int tile, int t) -> int {
      int h;
      if (field) {
        const int v = (int)sLdU16(aField + 2u * (uint32_t)tile);
        h = v == 0xFFFF ? MRP_INF : v;
      } else {
        h = abs((tile & 31) - goalX) + abs((tile >> 5) - goalY);
      }
      if (h != MRP_INF) h = max(h, lastGoal + 1 - t);
      return h;
    };
    const bool exactTail = !p.focalMode && field != nullptr;
    auto selfAt = [&](int t) -> int {
      return selfLen > 0 ? (int)sLdU16(aSelf + 2u * (uint32_t)min(t, selfLen - 1)) : -1;
    };
    auto openGet = [&](int i) -> unsigned long long {
      if (i < kOpenT) return sLd64(aOpen + 8u * (uint32_t)i);
      return openG[i];
    };
    auto openSet = [&](int i, unsigned long long v) {
      if (i < kOpenT) sSt64(aOpen + 8u * (uint32_t)i, v); else openG[i] = v;
    };
    auto nodeKeyOf = [&](int n) -> uint32_t {
      if (n < kNodeT) return sLd32(aNodeKey + 4u * (uint32_t)n);
      return nodeKeyG[n];
    };

    // ---- root ----
    int nNodes = 0, nOpen = 0, expanded = 0, tTop = tMark;
    int goalNode = -1, tailFrom = -1, goalF = 0;
    const int startTile = toTile(jb.start_cell);
    int bestF = 0;
    // warp copy between the shared-memory regions and their image in the blob
    auto copyRegion = [&](int off, int bytes, bool toShared) {
      uint4* sm = reinterpret_cast<uint4*>(smemRaw + off);
      uint4* gl = reinterpret_cast<uint4*>(blob + 64 + off);
      const int n = (bytes + 15) >> 4;
      if (toShared)
        for (int i = lane; i < n; i += 32) sm[i] = gl[i];
      else
        for (int i = lane; i < n; i += 32) gl[i] = sm[i];
    };
    auto copyState = [&](bool toShared) {
      copyRegion(lay.openKey, min(nOpen, kOpenT) * 8, toShared);
      copyRegion(lay.openState, min(nOpen, kOpenT) * 4, toShared);
      copyRegion(lay.vis, (tTop + 1) * 128, toShared);
      copyRegion(lay.nodeKey, min(nNodes, kNodeT) * 4, toShared);
      copyRegion(lay.nodePar, min(nNodes, kNodeT) * 2, toShared);
      copyRegion(lay.hist, kFMaxT * 2, toShared);
    };
    int expandedBefore = 0;
    if (status == -1 && resume) {
      // the setup above rebuilt everything that does not change during a search; the
      // search itself continues from the blob
      const int* hdr = reinterpret_cast<const int*>(blob);
      nNodes = hdr[0];
      nOpen = hdr[1];
      expanded = hdr[2];
      bestF = hdr[3];
      tTop = max(tTop, hdr[4]);
      expandedBefore = expanded;
      __syncwarp();
      copyState(true);
    } else if (status == -1) {
      const int h0 = heur(startTile, 0);
      if (h0 == MRP_INF) {
        status = 1;
      } else if (h0 >= kFMaxT) {
        status = kTileStatusRedo;
      } else {
        if (lane == 0) {
          sSt32(aNodeKey, (uint32_t)startTile);
          sStU16(aNodePar, 0xFFFFu);
          sSt64(aOpen, packOpenKey(0, h0, 0, 0));
          sSt32(aOpenState, (uint32_t)startTile);
          sAtomOr(aVis + 4u * (uint32_t)(startTile >> 5), 1u << (startTile & 31));
          sStU16(aHist + 2u * (uint32_t)h0, 1u);
        }
        nNodes = 1;
        nOpen = 1;
        bestF = h0;
      }
    }
    __syncwarp();
    // per-lane minimum of the eligible OPEN entries this lane owns (positions == lane mod 32)
    unsigned long long cbest = ~0ull;
    int cpos = 0;
    int cachedBound = -1;
    auto rescanLane = [&](int L, int n, int fb) {
      unsigned long long b = ~0ull;
      int bp = 0;
      for (int i = L + 32 * lane; i < n; i += 1024) {
        const unsigned long long e = openGet(i);
        const int f = (int)((e >> 38) & 0xfffull);
        if (f <= fb && e < b) {
          b = e;
          bp = i;
        }
      }
      const unsigned long long m = warpMin64(b);
      const uint32_t w = __ballot_sync(0xffffffffu, b == m);
      const int p2 = __shfl_sync(0xffffffffu, bp, __ffs(w) - 1);
      if (lane == L) {
        cbest = m;
        cpos = p2;
      }
    };
    auto afterRemoval = [&](int pos, int nAfter, int fb) {
      const int ownerL = pos & 31, lastL = nAfter & 31;
      const bool lastWasBest =
          __shfl_sync(0xffffffffu, (int)(cpos == nAfter && cbest != ~0ull), lastL) != 0;
      rescanLane(ownerL, nAfter, fb);
      if (lastL != ownerL && lastWasBest) rescanLane(lastL, nAfter, fb);
    };

    while (status == -1) {
      if (nOpen == 0) {
        status = 1;
        break;
      }
      if (blob && p.sliceCap > 0 && expanded - expandedBefore >= p.sliceCap) {
        status = kTileStatusSuspended;  // the rest of this search belongs to a later launch
        break;
      }
      while (bestF < kFMaxT && sLdU16(aHist + 2u * (uint32_t)bestF) == 0) ++bestF;
      if (bestF >= kFMaxT) {
        status = kTileStatusRedo;
        break;
      }
      const float bound = p.focalMode ? (float)bestF * p.w : (float)bestF;
      const int fBound = min((int)bound, 4095);
      if (fBound != cachedBound) {
        unsigned long long best = ~0ull;
        int bestPos = 0;
        const int nS = min(nOpen, kOpenT);
        for (int i = lane; i < nS; i += 32) {
          const unsigned long long e = sLd64(aOpen + 8u * (uint32_t)i);
          const int f = (int)((e >> 38) & 0xfffull);
          if (f <= fBound && e < best) {
            best = e;
            bestPos = i;
          }
        }
        for (int i0 = kOpenT; i0 < nOpen; i0 += 128) {
          unsigned long long e[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + lane + 32 * u;
            e[u] = i < nOpen ? openG[i] : ~0ull;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int f = (int)((e[u] >> 38) & 0xfffull);
            if (f <= fBound && e[u] < best) {
              best = e[u];
              bestPos = i0 + lane + 32 * u;
            }
          }
        }
        cbest = best;
        cpos = bestPos;
        cachedBound = fBound;
      }
      // ---- select ----
      const unsigned long long bestAll = warpMin64(cbest);
      const uint32_t who = __ballot_sync(0xffffffffu, cbest == bestAll);
      const int owner = __ffs(who) - 1;
      const int pos = __shfl_sync(0xffffffffu, cpos, owner);
      const int cur = (int)(bestAll & 0x3ffffffull);
      const uint32_t ckey = pos < kOpenT ? sLd32(aOpenState + 4u * (uint32_t)pos) : nodeKeyOf(cur);
      const int ct = (int)(ckey >> 10), ctile = (int)(ckey & 1023u);
      const int cg = 4095 - (int)((bestAll >> 26) & 0xfffull);
      const int cfo = (int)(bestAll >> 50);
      const int cf = (int)((bestAll >> 38) & 0xfffull);
      ++expanded;  // onExpandLowLevelNode, a_star.hpp:87

      const bool atGoal = ctile == goalTile;
      if (atGoal && ct > lastGoal) {  // isSolution, example/cbs.cpp:286-289
        goalNode = cur;
        goalF = cf;
        status = 0;
        break;
      }
      if (exactTail && ct >= tFree) {
        goalNode = cur;
        tailFrom = cur;
        goalF = cf;
        status = 0;
        break;
      }
      if (expanded > p.maxExpanded) {
        status = 2;
        break;
      }
      const int nt = ct + 1;
      if (nt >= TB) {  // beyond the visited bitmap
        status = kTileStatusRedo;
        break;
      }
      tTop = max(tTop, nt);
      // ---- successors: lanes 0..4 = Wait, Left, Right, Up, Down (cbs.cpp:299-332) ----
      const int cx = ctile & 31, cy = ctile >> 5;
      int nx = cx, ny = cy;
      if (lane == 1) nx = cx - 1;
      if (lane == 2) nx = cx + 1;
      if (lane == 3) ny = cy + 1;
      if (lane == 4) ny = cy - 1;
      const bool inMap = lane < 5 && nx >= 0 && ny >= 0 && nx < dimx && ny < dimy;
      const int ntile = inMap ? (ny << 5) | nx : ctile;
      // occupancy of the successor's cell at both times and of this cell at the next step
      // (consumed by the focal values below): asynchronous copies into the lane's staging
      // words, so that nothing in between waits for them
      const bool useOcc = wantFocal && occOK;
      if (useOcc) {
        if (inMap) {
          const int r1 = min(nt, p.Tpad - 1) * 32, r0 = min(ct, p.Tpad - 1) * 32;
          const uint32_t st = aStage + 32u * (uint32_t)lane;
          cpAsync8(st, occ + r1 + (ntile >> 5));
          cpAsync8(st + 8u, occ + r0 + (ntile >> 5));
          cpAsync8(st + 16u, occ + r1 + cy);
        }
        cpAsyncCommit();
      }
      // remove from OPEN (swap with last) and the histogram
      __syncwarp();
      if (lane == 0) {
        const unsigned long long last = openGet(nOpen - 1);
        openSet(pos, last);
        if (pos < kOpenT)
          sSt32(aOpenState + 4u * (uint32_t)pos,
                nOpen - 1 < kOpenT ? sLd32(aOpenState + 4u * (uint32_t)(nOpen - 1))
                                   : nodeKeyOf((int)(last & 0x3ffffffull)));
        sStU16(aHist + 2u * (uint32_t)cf, sLdU16(aHist + 2u * (uint32_t)cf) - 1u);
      }
      --nOpen;
      __syncwarp();
      afterRemoval(pos, nOpen, fBound);

      bool ok = inMap && ((sLd32(aRows + 4u * (uint32_t)(ny & 31)) >> (nx & 31)) & 1u);
      const int cc = toCell(ctile), nc = toCell(ntile);
      // transitionValid (example/cbs.cpp:438-444): Bloom filter first
      if (nEc) {
        const uint32_t h = edgeHash(ct, cc, nc) & 1023u;
        const uint32_t word = __shfl_sync(0xffffffffu, bloom, (int)(h >> 5));
        if (ok && ((word >> (h & 31u)) & 1u))
          for (int i = 0; i < nEc; ++i)
            if ((int)sLd32(aEc + 12u * (uint32_t)i) == ct && (int)sLd32(aEc + 12u * (uint32_t)i + 4u) == cc &&
                (int)sLd32(aEc + 12u * (uint32_t)i + 8u) == nc)
              ok = false;
      }
      // closed/open membership; a vertex-constrained state reads as seen
      bool isNew = false;
      if (ok) {
        const uint32_t bit = 1u << nx;
        isNew = (sAtomOr(aVis + 4u * (uint32_t)(nt * 32 + ny), bit) & bit) == 0;
      }
      int nh = 0;
      if (ok) nh = heur(ntile, nt);
      if (ok && nh == MRP_INF) isNew = false;  // unreachable goal component
      const int ng = cg + 1;
      const int nf = ng + nh;
      const uint32_t newMask = __ballot_sync(0xffffffffu, isNew);
      const int nNew = __popc(newMask);
      if (nNodes + nNew > p.maxNodes) {
        status = 2;
        break;
      }
      if (__any_sync(0xffffffffu, isNew && nf >= kFMaxT)) {
        status = kTileStatusRedo;
        break;
      }
      const int myNode = nNodes + __popc(newMask & ((1u << lane) - 1u));
      // ---- focal values of the new nodes (ecbs.cpp:282-312) ----
      int focalAdd = 0;
      if (useOcc) {
        cpAsyncWaitAll();
        __syncwarp();
      }
      if (wantFocal && newMask) {
        bool needExact = !occOK;
        if (occOK) {
          const uint32_t st = aStage + 32u * (uint32_t)lane;
          const unsigned long long oA = inMap ? sLd64(st) : 0ull, oB = inMap ? sLd64(st + 8u) : 0ull,
                                   oC = inMap ? sLd64(st + 16u) : 0ull;
          const int sN = selfAt(nt), sC = selfAt(ct);
          // (low word: cells with >= 1 agent, high word: cells with >= 2)
          const int vN = (int)((oA >> nx) & 1ull) + (int)((oA >> (32 + nx)) & 1ull) - (sN == ntile ? 1 : 0);
          const int othersHere = (int)((oC >> cx) & 1ull) + (int)((oC >> (32 + cx)) & 1ull) - (sN == ctile ? 1 : 0);
          const int othersFrom = (int)((oB >> nx) & 1ull) + (int)((oB >> (32 + nx)) & 1ull) - (sC == ntile ? 1 : 0);
          focalAdd = vN;
          // a swap needs another agent on this cell at nt that stood on the successor's cell at ct
          needExact = __any_sync(0xffffffffu, isNew && othersHere > 0 && othersFrom > 0);
        }
        if (needExact) {
          int kc[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) kc[k] = __shfl_sync(0xffffffffu, nc, k);
          uint32_t w01 = 0, w23 = 0, w4 = 0;
          for (int a0 = 0; a0 < p.N; a0 += 128) {
            int pa[4], pb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int a = a0 + lane + 32 * u;
              const int L = a < p.N ? tlen[a] : 0;
              const bool use = a < p.N && a != jb.self && L > 0;
              pa[u] = use ? tab[(size_t)a * p.Tpad + min(ct, L - 1)] : -1;
              pb[u] = use ? tab[(size_t)a * p.Tpad + min(nt, L - 1)] : -1;
            }
            bool arrives[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) arrives[u] = cc == pb[u];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
              if (!((newMask >> k) & 1u)) continue;
              uint32_t c = 0;
#pragma unroll
              for (int u = 0; u < 4; ++u)
                c += (uint32_t)(pb[u] == kc[k]) + (uint32_t)(arrives[u] && kc[k] == pa[u]);
              if (k == 0) w01 += c;
              if (k == 1) w01 += c << 16;
              if (k == 2) w23 += c;
              if (k == 3) w23 += c << 16;
              if (k == 4) w4 += c;
            }
          }
          w01 = __reduce_add_sync(0xffffffffu, w01);
          w23 = __reduce_add_sync(0xffffffffu, w23);
          w4 = __reduce_add_sync(0xffffffffu, w4);
          if (lane == 0) focalAdd = (int)(w01 & 0xffffu);
          if (lane == 1) focalAdd = (int)(w01 >> 16);
          if (lane == 2) focalAdd = (int)(w23 & 0xffffu);
          if (lane == 3) focalAdd = (int)(w23 >> 16);
          if (lane == 4) focalAdd = (int)w4;
        }
      }
      unsigned long long newKey = ~0ull;
      if (isNew) {
        const uint32_t nkey = ((uint32_t)nt << 10) | (uint32_t)ntile;
        if (myNode < kNodeT) {
          sSt32(aNodeKey + 4u * (uint32_t)myNode, nkey);
          sStU16(aNodePar + 2u * (uint32_t)myNode, (uint32_t)cur);
        } else {
          nodeKeyG[myNode] = nkey;
          nodeParG[myNode] = cur;
        }
        newKey = packOpenKey(cfo + focalAdd, nf, ng, myNode);
        const int at = nOpen + myNode - nNodes;
        openSet(at, newKey);
        if (at < kOpenT) sSt32(aOpenState + 4u * (uint32_t)at, nkey);
        // two 16-bit counters per word: same-bin successors add up in the atomic
        sAtomAdd(aHist + 4u * (uint32_t)(nf >> 1), (nf & 1) ? 0x10000u : 1u);
      }
      bestF = min(bestF, (int)__reduce_min_sync(0xffffffffu, isNew ? (uint32_t)nf : 0x7fffffffu));
      // the new OPEN entries (positions nOpen .. nOpen + nNew - 1) join the caches of their
      // lanes: lane (nOpen + r) & 31 takes the r-th new successor
      {
        const int r = (lane - nOpen) & 31;
        const int src = (int)((kNthSet[newMask & 31u] >> (3 * min(r, 4))) & 7u);
        const unsigned long long kkey = __shfl_sync(0xffffffffu, newKey, src);
        if (r < nNew) {
          const int kf = (int)((kkey >> 38) & 0xfffull);
          if (kf <= fBound && kkey < cbest) {
            cbest = kkey;
            cpos = nOpen + r;
          }
        }
      }
      nNodes += nNew;
      nOpen += nNew;
      __syncwarp();
    }
    __syncwarp();

    if (status == kTileStatusSuspended) {
      copyState(false);
      if (lane == 0) {
        int* hdr = reinterpret_cast<int*>(blob);
        hdr[0] = nNodes;
        hdr[1] = nOpen;
        hdr[2] = expanded;
        hdr[3] = bestF;
        hdr[4] = tTop;
      }
    }
    // ---- result ----
    mrp_path_info pi;
    pi.status = status;
    pi.cost = 0;
    pi.fmin = 0;
    pi.length = 0;
    pi.expanded = expanded;
    if (status == 0) {
      int32_t *oc, *og;
      int outCap;
      pathOutput(p, job, oc, og, outCap);
      const uint32_t gk = nodeKeyOf(goalNode);
      const int gt = (int)(gk >> 10);  // depth of the goal node = its time = its g
      int len = gt + 1;
      int cost = gt;
      const int tailTile = (int)(gk & 1023u);
      if (tailFrom >= 0) {
        const int rest = (int)sLdU16(aField + 2u * (uint32_t)tailTile);
        len += rest;
        cost += rest;
      }
      pi.cost = cost;
      // fmin: A* returns f of the goal node (a_star.hpp:106); A*-epsilon the
      // minimum f in OPEN at termination (a_star_epsilon.hpp:210)
      pi.fmin = p.focalMode ? bestF : goalF;
      pi.length = len;
      if (len > outCap) {
        pi.status = 2;
      } else if (lane == 0) {
        pathLength(p, job, len);
        int n = goalNode;
        for (int t = gt; t >= 0; --t) {
          oc[t] = toCell((int)(nodeKeyOf(n) & 1023u));
          if (og) og[t] = t;
          n = n < kNodeT ? (int)sLdU16(aNodePar + 2u * (uint32_t)n) : nodeParG[n];
        }
        if (tailFrom >= 0) {
          // follow the field's gradient: Left, Right, Up, Down (any optimum)
          int c = tailTile, t = gt;
          auto fieldAt = [&](int tile) -> int { return (int)sLdU16(aField + 2u * (uint32_t)tile); };
          auto rowAt = [&](int y) -> uint32_t { return sLd32(aRows + 4u * (uint32_t)y); };
          while (fieldAt(c) > 0) {
            const int x = c & 31, y = c >> 5, d = fieldAt(c);
            int nxt;
            if (x > 0 && ((rowAt(y) >> (x - 1)) & 1u) && fieldAt(c - 1) == d - 1) nxt = c - 1;
            else if (x + 1 < dimx && ((rowAt(y) >> (x + 1)) & 1u) && fieldAt(c + 1) == d - 1) nxt = c + 1;
            else if (y + 1 < dimy && ((rowAt(y + 1) >> x) & 1u) && fieldAt(c + 32) == d - 1) nxt = c + 32;
            else nxt = c - 32;
            c = nxt;
            ++t;
            oc[t] = toCell(c);
            if (og) og[t] = t;
          }
        }
      }
    }
    if (lane == 0) p.info[job] = pi;
    visTop = min(TB, tTop + 1);
    __syncwarp();
  }
}

// ---- host side ----
size_t lowlevelTileBlobBytes(int TB, int maxNodes) {
  return (64 + (size_t)tileLayout(TB).total + (size_t)maxNodes * 16 + 255) & ~(size_t)255;
}

static int tileRows(const LLParams& p) {
  if (p.TB > 0) return p.TB;  // sliced searches: the layout the state blobs were made for
  if (const char* e = getenv("MRP_LL_TILE_TB")) return std::max(32, std::min(512, atoi(e) & ~31));
  if (p.dimx * p.dimy <= 64) return 64;
  return p.Tpad <= 88 ? 128 : (p.Tpad <= 150 ? 192 : 256);
}

bool lowlevelTileEligible(const LLParams& p, int n_tables) {
  if (getenv("MRP_LL_GENERIC")) return false;  // A/B switch, read per call (tests flip it)
  if (p.variant != 0 || p.W != 1 || p.dimy > 32 || p.dimx > 32) return false;
  if (n_tables > 0 && p.tables && (size_t)p.Tpad * 256 > (size_t)kOccSmemMax) return false;
  return true;
}

size_t lowlevelTileOccBytes(int n_tables, int Tpad) {
  return (size_t)std::max(n_tables, 0) * Tpad * 32 * sizeof(uint2) + (size_t)std::max(n_tables, 1) * 4;
}

// The shared-memory limits of the two kernels are set once, to the largest size
// any launch asks for: the attribute belongs to the function, and lanes launch
// concurrently with different sizes.
static void setSmemLimits() {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(focal_occ_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kOccSmemMax);
    cudaFuncSetAttribute(lowlevel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tileLayout(512).total);
  });
}

int lowlevelTileSlots(const LLParams& pIn) {
  LLParams p = pIn;
  p.TB = tileRows(p);
  const TileLayout lay = tileLayout(p.TB);
  setSmemLimits();
  int perSm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, lowlevel_tile_kernel, 32, lay.total) != cudaSuccess ||
      perSm < 1)
    perSm = 1;
  return ctx().smCount * perSm;
}

int launchLowlevelTile(const LLParams& pIn, int n_tables, uint2* d_occ, int32_t* d_occMany, int slots,
                       cudaStream_t st) {
  LLParams p = pIn;
  p.TB = tileRows(p);
  p.occ = nullptr;
  p.occMany = nullptr;
  const bool noOcc = getenv("MRP_LL_TILE_NOOCC") != nullptr;
  if (p.focalMode && p.tables && n_tables > 0 && d_occ && !noOcc) {
    const size_t smem = (size_t)p.Tpad * 32 * 8;
    setSmemLimits();
    focal_occ_kernel<<<n_tables, 256, smem, st>>>(p.tables, p.tableLen, p.N, p.Tpad, p.dimx, d_occ, d_occMany);
    countLaunch();
    MRP_CUDA(cudaGetLastError());
    p.occ = d_occ;
    p.occMany = d_occMany;
  }
  const TileLayout lay = tileLayout(p.TB);
  lowlevel_tile_kernel<<<slots, 32, lay.total, st>>>(p);
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mrp
