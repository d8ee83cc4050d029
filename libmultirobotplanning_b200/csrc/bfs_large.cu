// bfs_large.cu — distance fields on large maps (anything bigger than one
// 32x32 tile; the headline case is the synthetic 1024x1024 map, config C5).
//
// Replaces ShortestPathHeuristic (example/shortest_path_heuristic.hpp:12-62 of
// the reference; Floyd–Warshall is O(V^3) and infeasible at V = 2^20) by one
// BFS distance field per goal, the layout of the reference's disabled
// computeHeuristic (example/cbs.cpp:445-557).
//
// Design (one CTA per goal, persistent over goals):
//   * The map is cut into 8x4-cell tiles, one 32-bit word each (bit = 8*(y&3)
//     + (x&7)).  Per goal the only per-cell state is the `vis` mask (1 bit per
//     cell, 128 KB for 1024x1024) and it lives in shared memory; the static
//     `free` mask is read through the read-only path.
//   * A level touches only ACTIVE tiles, kept as a compacted list: one thread
//     per active tile.  A tile step is pure bit arithmetic:
//        spread = in-tile shifts of vis  |  edge rows/columns of the 4
//                 neighbouring tiles' vis
//        cand   = spread & free & ~vis
//   * Each level runs in two phases separated by a barrier.  Phase A only
//     reads `vis` and records cand; phase B commits vis |= cand, stores the
//     level into the int32 field for every new cell and builds the next active
//     list.  Because phase A is read-only and phase B writes values that are a
//     pure function of the pre-level state, processing a tile twice is
//     harmless, so the de-duplication of the next list may be approximate: a
//     tile is appended by whichever proposer last wrote its 8-bit claim.
//     List slots are handed out with one warp-aggregated atomic per warp.
//   * Obstacle cells of a tile get MRP_INF when the tile gains its first cell,
//     so that all partial sector writes of one tile meet in L2 before they are
//     evicted; tiles never reached and cells of other components are written
//     in a final sweep.  DRAM traffic stays close to 4 B per cell.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace mrp {

struct BfsTileParams {
  const uint32_t* __restrict__ free84;  // [TH*TW] free mask, 8x4 tiles
  const int32_t* __restrict__ goals;    // goal cells
  int32_t* __restrict__ out;            // [n_goals][cells]
  uint32_t* ws;                         // workspace (header + per-CTA state)
  size_t wsWordsPerCta;
  int n_goals;
  int dimx, dimy, TW, TH;
  int TWv;  // row stride of vis (== 1 mod 32: vertical runs hit distinct banks)
  int TWc;  // row stride of the byte claims (TWc/4 == 1 mod 32)
  int cap;  // list entries kept in shared memory (multiple of the block size)
  int spillCap;  // capacity of the global fallback lists (same multiple)
  int nCompute;  // threads that run the level logic; the rest only store
  int dbg;  // debug: bit0 = skip level stores, bit1 = skip obstacle stores
  // optional indirection (goals the queue kernel handed over, bfs_queue.cu):
  // goalList[0..*goalListCount) are indices into goals / out
  const uint32_t* goalList;
  const uint32_t* goalListCount;
};

constexpr int kWsHeaderWords = 64;  // [0] goal counter, [1] overflow flag

#ifdef MRP_BFS_TIMING
__device__ unsigned long long g_bfsTiming[32][12];
#define TICK(k)                                              \
  do {                                                       \
    const long long _n = clock64();                          \
    if (blockIdx.x == 0 && lane == 0) tacc[k] += _n - tlast; \
    tlast = _n;                                              \
  } while (0)
#else
#define TICK(k)
#endif

__device__ __forceinline__ uint32_t tileInBounds(int tx, int ty, int dimx, int dimy) {
  const int nx = min(8, dimx - 8 * tx), ny = min(4, dimy - 4 * ty);
  const uint32_t row = (1u << nx) - 1u;  // nx in 1..8
  uint32_t m = row | (row << 8) | (row << 16) | (row << 24);
  if (ny < 4) m &= (1u << (8 * ny)) - 1u;
  return m;
}

template <bool kSmemState>
__global__ void __launch_bounds__(1024, 1)
bfs_tiles_kernel(BfsTileParams p) {
  extern __shared__ uint32_t smem[];
  __shared__ int sGoal;
  __shared__ int sCount[2];
  __shared__ int sOverflow;
  const int nTiles = p.TW * p.TH;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t ltMask = (1u << lane) - 1u;
  const int cells = p.dimx * p.dimy;
  const int TW = p.TW, TH = p.TH, dimx = p.dimx, TWv = p.TWv, TWc = p.TWc;
  const int nVis = TH * TWv, nClaimWords = (TH * TWc + 3) / 4;
  const int nThreads = blockDim.x;
  // role split: the upper half of the block only issues the field stores
  const int nCompute = p.nCompute;
  const bool isStore = tid >= nCompute;

  // ---- storage ----
  // lists live in shared memory (capacity p.cap).  If a level ever needs more
  // the goal is restarted with the lists in the global workspace (capacity
  // 2*nTiles >= any de-duplicated list): one pointer swap, no per-access branch.
  uint32_t* g = p.ws + kWsHeaderWords + (size_t)blockIdx.x * p.wsWordsPerCta;
  const int spillCap = p.spillCap;
  uint32_t* vis = kSmemState ? smem + 5 * p.cap + (p.cap + 3) / 4
                             : g + 5 * spillCap + (spillCap + 3) / 4;
  uint8_t* claim = reinterpret_cast<uint8_t*>(vis + nVis);

  bool spilled = false;
  while (true) {
    __syncthreads();
    if (!spilled) {
      if (tid == 0) sGoal = (int)atomicAdd(p.ws, 1u);
      __syncthreads();
    }
    int gidx = sGoal;
    if (gidx >= (p.goalList ? (int)*p.goalListCount : p.n_goals)) break;
    if (p.goalList) gidx = (int)p.goalList[gidx];
    const int cap = spilled ? spillCap : p.cap;
    // two lists of `cap` (tile, free word) entries, then cand[cap], pm[cap]
    uint2* list0 = reinterpret_cast<uint2*>(spilled ? g : smem);
    uint32_t* candArr = reinterpret_cast<uint32_t*>(list0 + 2 * cap);
    uint8_t* pmArr = reinterpret_cast<uint8_t*>(candArr + cap);

    const int goal = p.goals[gidx];
    int32_t* out = p.out + (size_t)gidx * cells;
    const int gx = goal % dimx, gy = goal / dimx;
    const int gtx = gx >> 3, gty = gy >> 2;
    const int gT = gty * TW + gtx;
    const uint32_t gbit = 1u << (((gy & 3) << 3) | (gx & 7));
    const bool goalFree = (p.free84[gT] & gbit) != 0;

    for (int i = tid; i < nVis; i += blockDim.x) vis[i] = 0;
    for (int i = tid; i < nClaimWords; i += blockDim.x)
      reinterpret_cast<uint32_t*>(claim)[i] = 0xffffffffu;
    if (tid == 0) sOverflow = 0;
    __syncthreads();

    if (goalFree) {
      if (tid == 0) {
        // level 0: the goal cell; its tile's obstacles get MRP_INF right away
        vis[gty * TWv + gtx] = gbit;
        out[goal] = 0;
        uint32_t ob = ~p.free84[gT] & tileInBounds(gtx, gty, dimx, p.dimy);
        while (ob) {
          const int b = __ffs(ob) - 1;
          ob &= ob - 1;
          out[(4 * gty + (b >> 3)) * dimx + 8 * gtx + (b & 7)] = MRP_INF;
        }
        int n = 0;
        const uint32_t ge = ((uint32_t)gty << 16) | (uint32_t)gtx;
        uint2* l1 = list0 + cap;
        l1[n++] = make_uint2(ge, p.free84[gT]);
        if (gty > 0) l1[n++] = make_uint2(ge - 0x10000u, p.free84[gT - TW]);
        if (gty < TH - 1) l1[n++] = make_uint2(ge + 0x10000u, p.free84[gT + TW]);
        if (gtx > 0) l1[n++] = make_uint2(ge - 1u, p.free84[gT - 1]);
        if (gtx < TW - 1) l1[n++] = make_uint2(ge + 1u, p.free84[gT + 1]);
        sCount[1] = n;
        sCount[0] = 0;
      }
      __syncthreads();

#ifdef MRP_BFS_TIMING
      long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      long long tlast = clock64();
#endif
      for (int level = 1;; ++level) {
        const int cur = level & 1, nxt = cur ^ 1;
        const int count = sCount[cur];
        if (count == 0 || sOverflow) break;
        const uint2* lcur = list0 + cur * cap;
        uint2* lnxt = list0 + nxt * cap;
        TICK(0);
        // ---------------- phase A: read-only, shared memory only ----------------
        for (int idx = tid; idx < (isStore ? 0 : count); idx += nCompute) {
          const uint2 ent = lcur[idx];  // x = ty<<16 | tx, y = free word of the tile
          const int ty = (int)(ent.x >> 16), tx = (int)(ent.x & 0xffffu);
          const int V = ty * TWv + tx, C = ty * TWc + tx;
          const uint32_t v = vis[V];
          const uint32_t vU = ty > 0 ? vis[V - TWv] : 0u;
          const uint32_t vD = ty < TH - 1 ? vis[V + TWv] : 0u;
          const uint32_t vL = tx > 0 ? vis[V - 1] : 0u;
          const uint32_t vR = tx < TW - 1 ? vis[V + 1] : 0u;
          const uint32_t spread =
              ((v << 1) & 0xfefefefeu) | ((v >> 1) & 0x7f7f7f7fu) | (v << 8) | (v >> 8) |
              (vU >> 24) | (vD << 24) | ((vL >> 7) & 0x01010101u) |
              ((vR & 0x01010101u) << 7);
          const uint32_t cand = spread & ent.y & ~v;
          uint32_t pm = 0;
          if (cand) {
            const uint8_t id = (uint8_t)idx;
            pm = v ? 1u : 0x21u;  // bit 5: first gain of this tile
            claim[C] = id;
            if (ty > 0 && (((cand & 0xffu) << 24) & ~vU)) {
              pm |= 2;
              claim[C - TWc] = id;
            }
            if (ty < TH - 1 && ((cand >> 24) & ~vD)) {
              pm |= 4;
              claim[C + TWc] = id;
            }
            if (tx > 0 && (((cand & 0x01010101u) << 7) & ~vL)) {
              pm |= 8;
              claim[C - 1] = id;
            }
            if (tx < TW - 1 && (((cand >> 7) & 0x01010101u) & ~vR)) {
              pm |= 16;
              claim[C + 1] = id;
            }
          }
          candArr[idx] = cand;
          pmArr[idx] = (uint8_t)pm;
        }
        if (tid == 0) sCount[nxt] = 0;
        TICK(1);
        __syncthreads();
        TICK(2);
        // ---------------- phase B: commit + next list + field stores ----------------
        // the field stores of this level: every new cell gets its level; the
        // obstacle cells of a tile that gains its first cell get MRP_INF at
        // the same time
        auto fieldStores = [&](int first, int stride) {
          for (int idx = first; idx < count; idx += stride) {
            const uint32_t cand = candArr[idx];
            if (!cand) continue;
            const uint32_t pm = pmArr[idx];
            const uint2 ent = lcur[idx];
            const int ty = (int)(ent.x >> 16), tx = (int)(ent.x & 0xffffu);
            int32_t* obase = out + (4 * ty) * dimx + 8 * tx;
            uint32_t inf = 0;
            if ((pm & 0x20u) && !(p.dbg & 2)) inf = ~ent.y & tileInBounds(tx, ty, dimx, p.dimy);
            uint32_t m = ((p.dbg & 1) ? 0u : cand) | inf;
            while (m) {
              const int b = __ffs(m) - 1;
              m &= m - 1;
              if (p.dbg & 4)
                out[(idx & 31)] = level;  // debug: same store count, one line per warp
              else
                obase[(b >> 3) * dimx + (b & 7)] = ((inf >> b) & 1u) ? MRP_INF : level;
            }
          }
        };
        if (isStore) {
          // store warps run concurrently with the compute warps' commit
          fieldStores(tid - nCompute, nThreads - nCompute);
        } else {
          const int rounds = (count + nCompute - 1) / nCompute;
          for (int r = 0; r < rounds; ++r) {
            const int idx = r * nCompute + tid;
            if (idx - lane >= count) break;  // warp-uniform: nothing left for this warp
            const bool valid = idx < count;
            const uint32_t cand = valid ? candArr[idx] : 0u;
            const uint32_t pm = valid ? pmArr[idx] : 0u;
            const uint2 ent = lcur[valid ? idx : 0];
            const uint32_t e = ent.x;
            const int ty = (int)(e >> 16), tx = (int)(e & 0xffffu);
            const int C = ty * TWc + tx;
            const int F = ty * TW + tx;
            uint32_t wmask = 0;  // bit k: this entry appends neighbour k (0 = itself)
            if (cand) {
              const uint8_t id = (uint8_t)idx;
              if (claim[C] == id) wmask |= 1;
              if ((pm & 2) && claim[C - TWc] == id) wmask |= 2;
              if ((pm & 4) && claim[C + TWc] == id) wmask |= 4;
              if ((pm & 8) && claim[C - 1] == id) wmask |= 8;
              if ((pm & 16) && claim[C + 1] == id) wmask |= 16;
              vis[ty * TWv + tx] |= cand;
            }
            // free words of the tiles this entry appends (the next level then
            // runs out of shared memory only); the loads complete behind the
            // slot allocation
            uint32_t f1 = 0, f2 = 0, f3 = 0, f4 = 0;
            if (wmask & 2u) f1 = __ldg(&p.free84[F - TW]);
            if (wmask & 4u) f2 = __ldg(&p.free84[F + TW]);
            if (wmask & 8u) f3 = __ldg(&p.free84[F - 1]);
            if (wmask & 16u) f4 = __ldg(&p.free84[F + 1]);
            TICK(6);
            // slot allocation: ballots rank the winners, one atomic per warp
            const uint32_t b0 = __ballot_sync(0xffffffffu, wmask & 1u);
            const uint32_t b1 = __ballot_sync(0xffffffffu, wmask & 2u);
            const uint32_t b2 = __ballot_sync(0xffffffffu, wmask & 4u);
            const uint32_t b3 = __ballot_sync(0xffffffffu, wmask & 8u);
            const uint32_t b4 = __ballot_sync(0xffffffffu, wmask & 16u);
            const int n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3);
            const int total = n0 + n1 + n2 + n3 + __popc(b4);
            int base = 0;
            if (total) {
              if (lane == 0) base = atomicAdd(&sCount[nxt], total);
              base = __shfl_sync(0xffffffffu, base, 0);
              if (base + total > cap) {
                if (lane == 0) sOverflow = 1;
                wmask = 0;
              }
            }
            TICK(7);
            if (wmask & 1u) lnxt[base + __popc(b0 & ltMask)] = make_uint2(e, ent.y);
            if (wmask & 2u) lnxt[base + n0 + __popc(b1 & ltMask)] = make_uint2(e - 0x10000u, f1);
            if (wmask & 4u)
              lnxt[base + n0 + n1 + __popc(b2 & ltMask)] = make_uint2(e + 0x10000u, f2);
            if (wmask & 8u)
              lnxt[base + n0 + n1 + n2 + __popc(b3 & ltMask)] = make_uint2(e - 1u, f3);
            if (wmask & 16u)
              lnxt[base + n0 + n1 + n2 + n3 + __popc(b4 & ltMask)] = make_uint2(e + 1u, f4);
          }
          if (nCompute == nThreads) fieldStores(tid, nThreads);  // small blocks: no split
        }
        TICK(8);
        TICK(3);
        __syncthreads();
        TICK(4);
#ifdef MRP_BFS_TIMING
        if (blockIdx.x == 0 && lane == 0) tacc[5] += count;
#endif
      }
#ifdef MRP_BFS_TIMING
      if (blockIdx.x == 0 && lane == 0)
        for (int k = 0; k < 12; ++k) g_bfsTiming[tid >> 5][k] = tacc[k];
#endif
      __syncthreads();
      if (sOverflow) {
        // the shared-memory lists were too small for this goal: redo it with
        // the lists in the global workspace
        if (spilled) {
          if (tid == 0) p.ws[1] = 1u;  // cannot happen: capacity 2*nTiles
        } else {
          spilled = true;
          continue;
        }
      }
    }
    spilled = false;

    // ---- MRP_INF for everything the wavefront did not reach ----
    __syncthreads();
    for (int T = tid; T < nTiles; T += blockDim.x) {
      const int ty = T / TW, tx = T - ty * TW;
      const uint32_t v = vis[ty * TWv + tx];
      const uint32_t inb = tileInBounds(tx, ty, dimx, p.dimy);
      // gained tiles already hold their obstacles; only free cells of other
      // components are left.  Tiles that never gained get every cell.
      uint32_t un = (v ? (__ldg(&p.free84[T]) & ~v) : 0xffffffffu) & inb;
      int32_t* obase = out + (4 * ty) * dimx + 8 * tx;
      while (un) {
        const int b = __ffs(un) - 1;
        un &= un - 1;
        obase[(b >> 3) * dimx + (b & 7)] = MRP_INF;
      }
    }
    __syncthreads();
    if (!goalFree && tid == 0) out[goal] = 0;  // Floyd–Warshall row of an obstacle
  }
}

struct TileGeom {
  int TW, TH, TWv, TWc, nTiles, cap, spillCap, threads;
  size_t smemBytes, wsWordsPerCta;
  bool smemState;
};

static TileGeom tileGeometry(const mrp_map_s* map) {
  TileGeom t;
  t.TW = (map->dimx + 7) / 8;
  t.TH = (map->dimy + 3) / 4;
  t.nTiles = t.TW * t.TH;
  // padded row strides: vertical / diagonal runs of tiles (the shape of a
  // wavefront) must fall into distinct shared-memory banks
  t.TWv = ((t.TW + 31) & ~31) + 1;
  t.TWc = ((t.TW + 127) & ~127) + 4;
  if (t.TW < 32) t.TWv = t.TW | 1;
  if (t.TW < 124) t.TWc = (t.TW + 3) & ~3;
  // block size: a power of two (the striped list addressing uses shifts)
  int th = 128;
  while (th < 1024 && th * 32 < t.nTiles) th <<= 1;
  if (const char* e = getenv("MRP_BFS_THREADS")) th = atoi(e);
  t.threads = th;
  auto roundUp = [th](size_t n) { return (n + th - 1) / th * th; };
  t.cap = (int)std::max<size_t>(th, std::min<size_t>(2048, roundUp(t.nTiles)));
  if (const char* e = getenv("MRP_BFS_CAP")) t.cap = (int)roundUp(std::min(t.cap, atoi(e)));
  t.spillCap = (int)roundUp(2 * (size_t)t.nTiles);
  const size_t listBytes = (size_t)5 * t.cap * 4 + (size_t)((t.cap + 3) / 4) * 4;
  const size_t stateBytes = (size_t)t.TH * t.TWv * 4 + (size_t)((t.TH * t.TWc + 3) / 4) * 4;
  t.smemState = listBytes + stateBytes + 2048 <= ctx().smemOptin;
  if (getenv("MRP_BFS_GLOBAL")) t.smemState = false;
  t.smemBytes = listBytes + (t.smemState ? stateBytes : 0);
  t.wsWordsPerCta = 5 * (size_t)t.spillCap + ((size_t)t.spillCap + 3) / 4 +
                    (t.smemState ? 0 : stateBytes / 4) + 16;
  return t;
}

static int tileBlocks(const TileGeom& t, bool smemState) {
  auto kern = smemState ? bfs_tiles_kernel<true> : bfs_tiles_kernel<false>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smemBytes);
  int perSm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, t.threads, t.smemBytes) !=
          cudaSuccess || perSm < 1)
    perSm = 1;
  if (perSm > 16) perSm = 16;
  return ctx().smCount * perSm;
}

#ifdef MRP_BFS_TIMING
extern "C" int mrp_debug_bfs_timing(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_bfsTiming, sizeof(unsigned long long) * 32 * 12);
}
#endif

static size_t tilesWorkspaceBytes(const mrp_map_s* map) {
  const TileGeom t = tileGeometry(map);
  const int blocks = tileBlocks(t, t.smemState);
  return ((size_t)kWsHeaderWords + t.wsWordsPerCta * (size_t)blocks) * 4;
}

size_t bfsLargeWorkspaceBytes(const mrp_map_s* map, int n_goals) {
  size_t bytes = tilesWorkspaceBytes(map);
  if (bfsQueueFits(map)) bytes += bfsQueueWorkspaceWords(n_goals) * 4;
  if (bfsSweepFits(map)) bytes += bfsSweepWorkspaceWords(n_goals) * 4;
  return bytes;
}

static int launchBfsTiles(const mrp_map_s* map, const int32_t* d_goal_cell, int n_goals,
                          int32_t* d_out, void* d_ws, const uint32_t* d_goalList,
                          const uint32_t* d_goalListCount, cudaStream_t st);

// Maps of up to 1024 columns run the detour-level sweep (bfs_sweep.cu); the
// goals it hands over (more detour levels than it holds: mazes) run the queue
// kernel if the bitmap fits shared memory; the goals that one hands over (a
// level larger than its queues) and all larger maps run the tiled kernel.
int launchBfsLarge(const mrp_map_s* map, const int32_t* d_goal_cell, int n_goals,
                   int32_t* d_out, void* d_ws, cudaStream_t st) {
  if (n_goals <= 0) return 0;
  uint32_t* ws = static_cast<uint32_t*>(d_ws);
  const uint32_t *list = nullptr, *listCount = nullptr;
  if (bfsSweepFits(map)) {
    if (int rc = launchBfsSweep(map, d_goal_cell, n_goals, d_out, ws, st)) return rc;
    list = ws + 64;
    listCount = ws + 2;
    ws += bfsSweepWorkspaceWords(n_goals);
  }
  if (bfsQueueFits(map)) {
    if (int rc = launchBfsQueue(map, d_goal_cell, n_goals, d_out, ws, st, list, listCount)) return rc;
    list = ws + 64;
    listCount = ws + 2;
    ws += bfsQueueWorkspaceWords(n_goals);
  }
  return launchBfsTiles(map, d_goal_cell, n_goals, d_out, ws, list, listCount, st);
}

static int launchBfsTiles(const mrp_map_s* map, const int32_t* d_goal_cell, int n_goals,
                          int32_t* d_out, void* d_ws, const uint32_t* d_goalList,
                          const uint32_t* d_goalListCount, cudaStream_t st) {
  const TileGeom t = tileGeometry(map);
  MRP_CHECK(t.TW < 65536 && t.TH < 65536, MRP_ERR_UNSUPPORTED,
            "map %dx%d exceeds the packed tile coordinates", map->dimx, map->dimy);
  BfsTileParams p;
  p.free84 = map->d_bits84;
  p.goals = d_goal_cell;
  p.out = d_out;
  p.ws = static_cast<uint32_t*>(d_ws);
  p.wsWordsPerCta = t.wsWordsPerCta;
  p.n_goals = n_goals;
  p.dimx = map->dimx;
  p.dimy = map->dimy;
  p.TW = t.TW;
  p.TH = t.TH;
  p.TWv = t.TWv;
  p.TWc = t.TWc;
  p.cap = t.cap;
  p.spillCap = t.spillCap;
  p.nCompute = t.threads >= 256 ? t.threads / 2 : t.threads;
  if (const char* e = getenv("MRP_BFS_COMPUTE")) p.nCompute = atoi(e);
  p.dbg = getenv("MRP_BFS_DBG") ? atoi(getenv("MRP_BFS_DBG")) : 0;
  p.goalList = d_goalList;
  p.goalListCount = d_goalListCount;
  MRP_CUDA(cudaMemsetAsync(d_ws, 0, kWsHeaderWords * 4, st));
  int blocks = tileBlocks(t, t.smemState);
  if (blocks > n_goals) blocks = n_goals;
  if (t.smemState)
    bfs_tiles_kernel<true><<<blocks, t.threads, t.smemBytes, st>>>(p);
  else
    bfs_tiles_kernel<false><<<blocks, t.threads, t.smemBytes, st>>>(p);
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mrp
