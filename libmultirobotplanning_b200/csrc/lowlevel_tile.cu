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

#include "lowlevel.cuh"

namespace mrp {

constexpr int kOpenT = 768;   // OPEN entries in shared memory
constexpr int kNodeT = 1024;  // nodes with (state, parent) in shared memory
constexpr int kFMaxT = 512;   // f values tracked by the histogram
constexpr int kSelfT = 256;   // time steps of this agent's old path kept in shared memory

struct TileLayout {
  int openKey, vis, openState, nodeKey, field, nodePar, hist, rows, cons, selfRow, total;
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

__device__ __forceinline__ uint32_t edgeHash(int t, int from, int to) {
  return hashState((uint32_t)t * 0x9E3779B1u ^ ((uint32_t)from << 11) ^ (uint32_t)to);
}

__global__ void __launch_bounds__(32) lowlevel_tile_kernel(LLParams p) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const TileLayout lay = tileLayout(p.TB);
  unsigned long long* openS = reinterpret_cast<unsigned long long*>(smemRaw + lay.openKey);
  uint32_t* vis = reinterpret_cast<uint32_t*>(smemRaw + lay.vis);
  uint32_t* openStateS = reinterpret_cast<uint32_t*>(smemRaw + lay.openState);
  uint32_t* nodeKeyS = reinterpret_cast<uint32_t*>(smemRaw + lay.nodeKey);
  uint32_t* hist32 = reinterpret_cast<uint32_t*>(smemRaw + lay.hist);
  unsigned short* hist = reinterpret_cast<unsigned short*>(smemRaw + lay.hist);
  unsigned short* fieldS = reinterpret_cast<unsigned short*>(smemRaw + lay.field);
  unsigned short* nodeParS = reinterpret_cast<unsigned short*>(smemRaw + lay.nodePar);
  uint32_t* rowsS = reinterpret_cast<uint32_t*>(smemRaw + lay.rows);
  int32_t* cons = reinterpret_cast<int32_t*>(smemRaw + lay.cons);
  unsigned short* selfRow = reinterpret_cast<unsigned short*>(smemRaw + lay.selfRow);

  const int lane = threadIdx.x;
  const int slot = blockIdx.x;
  const int TB = p.TB;
  const int dimx = p.dimx, dimy = p.dimy;
  const int cells = dimx * dimy;
  uint32_t* nodeKeyG = p.nodeKey + (size_t)slot * p.maxNodes;
  int32_t* nodeParG = p.nodeParent + (size_t)slot * p.maxNodes;
  unsigned long long* openG = p.openKey + (size_t)slot * p.maxNodes;

  // a cell of the map (x + dimx*y) as tile position (y*32 + x) and back
  auto toTile = [&](int c) -> int { return dimx == 32 ? c : ((c / dimx) << 5) | (c % dimx); };
  auto toCell = [&](int tile) -> int { return dimx == 32 ? tile : (tile & 31) + dimx * (tile >> 5); };
  auto openGet = [&](int i) -> unsigned long long { return i < kOpenT ? openS[i] : openG[i]; };
  auto openSet = [&](int i, unsigned long long v) {
    if (i < kOpenT) openS[i] = v; else openG[i] = v;
  };
  auto nodeKeyOf = [&](int n) -> uint32_t { return n < kNodeT ? nodeKeyS[n] : nodeKeyG[n]; };

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
    const int32_t* ecp = cons + 2 * nVc;
    const int goalTile = goal >= 0 ? toTile(goal) : -1;
    const int goalX = goalTile & 31, goalY = goalTile >> 5;

    // admissible heuristic: the reference's value (Manhattan, example/cbs.cpp:278-284,
    // or the distance field) raised to the time bound of the goal test (see lowlevel.cu)
    auto heur = [&](int tile, int t) -> int {
      int h;
      if (field) {
        const int v = fieldS[tile];
        h = v == 0xFFFF ? MRP_INF : v;
      } else {
        h = abs((tile & 31) - goalX) + abs((tile >> 5) - goalY);
      }
      if (h != MRP_INF) h = max(h, lastGoal + 1 - t);
      return h;
    };
    const bool exactTail = !p.focalMode && field != nullptr;
    auto selfAt = [&](int t) -> int { return selfLen > 0 ? (int)selfRow[min(t, selfLen - 1)] : -1; };

    // ---- root ----
    int nNodes = 0, nOpen = 0, expanded = 0, tTop = tMark;
    int goalNode = -1, tailFrom = -1, goalF = 0;
    const int startTile = toTile(jb.start_cell);
    if (status == -1) {
      const int h0 = heur(startTile, 0);
      if (h0 == MRP_INF) {
        status = 1;
      } else if (h0 >= kFMaxT) {
        status = kTileStatusRedo;
      } else {
        if (lane == 0) {
          nodeKeyS[0] = (uint32_t)startTile;
          nodeParS[0] = 0xFFFF;
          openS[0] = packOpenKey(0, h0, 0, 0);
          openStateS[0] = (uint32_t)startTile;
          vis[startTile >> 5] |= 1u << (startTile & 31);
          hist[h0] = 1;
        }
        nNodes = 1;
        nOpen = 1;
      }
    }
    __syncwarp();
    int bestF = (status == -1) ? (int)(openS[0] >> 38) & 0xfff : 0;
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
      while (bestF < kFMaxT && hist[bestF] == 0) ++bestF;
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
          const unsigned long long e = openS[i];
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
      const uint32_t ckey = pos < kOpenT ? openStateS[pos] : nodeKeyOf(cur);
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
      // occupancy of the successor's cell at both times and of this cell at the next
      // step (consumed by the focal values below; issued here to overlap the rest)
      uint2 oA = make_uint2(0, 0), oB = oA, oC = oA;
      if (occOK && inMap) {
        const int r1 = min(nt, p.Tpad - 1) * 32, r0 = min(ct, p.Tpad - 1) * 32;
        oA = occ[r1 + (ntile >> 5)];
        oB = occ[r0 + (ntile >> 5)];
        oC = occ[r1 + cy];
      }
      // remove from OPEN (swap with last) and the histogram
      __syncwarp();
      if (lane == 0) {
        openSet(pos, openGet(nOpen - 1));
        if (pos < kOpenT)
          openStateS[pos] = nOpen - 1 < kOpenT ? openStateS[nOpen - 1]
                                               : nodeKeyOf((int)(openGet(nOpen - 1) & 0x3ffffffull));
        hist[cf] -= 1;
      }
      --nOpen;
      __syncwarp();
      afterRemoval(pos, nOpen, fBound);

      bool ok = inMap && ((rowsS[ny & 31] >> (nx & 31)) & 1u);
      const int cc = toCell(ctile), nc = toCell(ntile);
      // transitionValid (example/cbs.cpp:438-444): Bloom filter first
      if (nEc) {
        const uint32_t h = edgeHash(ct, cc, nc) & 1023u;
        const uint32_t word = __shfl_sync(0xffffffffu, bloom, (int)(h >> 5));
        if (ok && ((word >> (h & 31u)) & 1u))
          for (int i = 0; i < nEc; ++i)
            if (ecp[3 * i] == ct && ecp[3 * i + 1] == cc && ecp[3 * i + 2] == nc) ok = false;
      }
      // closed/open membership; a vertex-constrained state reads as seen
      bool isNew = false;
      if (ok) {
        const uint32_t bit = 1u << nx;
        isNew = (atomicOr(&vis[nt * 32 + ny], bit) & bit) == 0;
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
      if (wantFocal && newMask) {
        bool needExact = !occOK;
        if (occOK) {
          const int sN = selfAt(nt), sC = selfAt(ct);
          const int vN = (int)((oA.x >> nx) & 1u) + (int)((oA.y >> nx) & 1u) - (sN == ntile ? 1 : 0);
          const int othersHere = (int)((oC.x >> cx) & 1u) + (int)((oC.y >> cx) & 1u) - (sN == ctile ? 1 : 0);
          const int othersFrom = (int)((oB.x >> nx) & 1u) + (int)((oB.y >> nx) & 1u) - (sC == ntile ? 1 : 0);
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
          nodeKeyS[myNode] = nkey;
          nodeParS[myNode] = (unsigned short)cur;
        } else {
          nodeKeyG[myNode] = nkey;
          nodeParG[myNode] = cur;
        }
        newKey = packOpenKey(cfo + focalAdd, nf, ng, myNode);
        const int at = nOpen + myNode - nNodes;
        openSet(at, newKey);
        if (at < kOpenT) openStateS[at] = nkey;
        // two 16-bit counters per word: same-bin successors add up in the atomic
        atomicAdd(&hist32[nf >> 1], (nf & 1) ? 0x10000u : 1u);
      }
      bestF = min(bestF, (int)__reduce_min_sync(0xffffffffu, isNew ? (uint32_t)nf : 0x7fffffffu));
      __syncwarp();
      // the new OPEN entries join the caches of their lanes
      for (int k = 0; k < 5; ++k) {
        if (!((newMask >> k) & 1u)) continue;
        const int kf = __shfl_sync(0xffffffffu, nf, k);
        const int kpos = nOpen + __popc(newMask & ((1u << k) - 1u));
        const unsigned long long kkey = __shfl_sync(0xffffffffu, newKey, k);
        if (lane == (kpos & 31) && kf <= fBound && kkey < cbest) {
          cbest = kkey;
          cpos = kpos;
        }
      }
      nNodes += nNew;
      nOpen += nNew;
      __syncwarp();
    }

    // ---- result ----
    mrp_path_info pi;
    pi.status = status;
    pi.cost = 0;
    pi.fmin = 0;
    pi.length = 0;
    pi.expanded = expanded;
    if (status == 0) {
      int32_t* oc = p.outCells + (size_t)job * p.pathCap;
      int32_t* og = p.outG + (size_t)job * p.pathCap;
      const uint32_t gk = nodeKeyOf(goalNode);
      const int gt = (int)(gk >> 10);  // depth of the goal node = its time = its g
      int len = gt + 1;
      int cost = gt;
      const int tailTile = (int)(gk & 1023u);
      if (tailFrom >= 0) {
        const int rest = fieldS[tailTile];
        len += rest;
        cost += rest;
      }
      pi.cost = cost;
      // fmin: A* returns f of the goal node (a_star.hpp:106); A*-epsilon the
      // minimum f in OPEN at termination (a_star_epsilon.hpp:210)
      pi.fmin = p.focalMode ? bestF : goalF;
      pi.length = len;
      if (len > p.pathCap) {
        pi.status = 2;
      } else if (lane == 0) {
        int n = goalNode;
        for (int t = gt; t >= 0; --t) {
          oc[t] = toCell((int)(nodeKeyOf(n) & 1023u));
          og[t] = t;
          n = n < kNodeT ? (int)nodeParS[n] : nodeParG[n];
        }
        if (tailFrom >= 0) {
          // follow the field's gradient: Left, Right, Up, Down (any optimum)
          int c = tailTile, t = gt;
          while (fieldS[c] > 0) {
            const int x = c & 31, y = c >> 5, d = fieldS[c];
            int nxt;
            if (x > 0 && ((rowsS[y] >> (x - 1)) & 1u) && fieldS[c - 1] == d - 1) nxt = c - 1;
            else if (x + 1 < dimx && ((rowsS[y] >> (x + 1)) & 1u) && fieldS[c + 1] == d - 1) nxt = c + 1;
            else if (y + 1 < dimy && ((rowsS[y + 1] >> x) & 1u) && fieldS[c + 32] == d - 1) nxt = c + 32;
            else nxt = c - 32;
            c = nxt;
            ++t;
            oc[t] = toCell(c);
            og[t] = t;
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
static int tileRows(const LLParams& p) {
  if (const char* e = getenv("MRP_LL_TILE_TB")) return std::max(32, std::min(512, atoi(e) & ~31));
  if (p.dimx * p.dimy <= 64) return 64;
  return p.Tpad <= 88 ? 128 : (p.Tpad <= 150 ? 192 : 256);
}

bool lowlevelTileEligible(const LLParams& p, int n_tables) {
  if (getenv("MRP_LL_GENERIC")) return false;  // A/B switch, read per call (tests flip it)
  if (p.variant != 0 || p.W != 1 || p.dimy > 32 || p.dimx > 32) return false;
  if (n_tables > 0 && p.tables && (size_t)p.Tpad * 256 > 160 * 1024) return false;
  return true;
}

size_t lowlevelTileOccBytes(int n_tables, int Tpad) {
  return (size_t)std::max(n_tables, 0) * Tpad * 32 * sizeof(uint2) + (size_t)std::max(n_tables, 1) * 4;
}

int lowlevelTileSlots(const LLParams& pIn) {
  LLParams p = pIn;
  p.TB = tileRows(p);
  const TileLayout lay = tileLayout(p.TB);
  cudaFuncSetAttribute(lowlevel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total);
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
    cudaFuncSetAttribute(focal_occ_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    focal_occ_kernel<<<n_tables, 256, smem, st>>>(p.tables, p.tableLen, p.N, p.Tpad, p.dimx, d_occ, d_occMany);
    countLaunch();
    MRP_CUDA(cudaGetLastError());
    p.occ = d_occ;
    p.occMany = d_occMany;
  }
  const TileLayout lay = tileLayout(p.TB);
  cudaFuncSetAttribute(lowlevel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total);
  lowlevel_tile_kernel<<<slots, 32, lay.total, st>>>(p);
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mrp
