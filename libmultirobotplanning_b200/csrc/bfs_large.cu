// bfs_large.cu — distance fields on large maps (anything bigger than one
// 32x32 tile; the headline case is the synthetic 1024x1024 map, config C5).
//
// Replaces ShortestPathHeuristic (example/shortest_path_heuristic.hpp:12-62 of
// the reference; Floyd–Warshall is O(V^3) and infeasible at V = 2^20) by one
// BFS distance field per goal, the layout of the reference's disabled
// computeHeuristic (example/cbs.cpp:445-557).
//
// Design (one CTA per goal, persistent over goals):
//   * The map is cut into 32x32-cell tiles; tile (s, tx) = rows 32s..32s+31,
//     columns 32tx..32tx+31, stored as 32 consecutive words (tile-major), so a
//     warp reads one tile with one conflict-free 128-byte access.
//   * Per goal the only per-cell state is the `open` mask (free and not yet
//     visited): 1 bit per cell, 128 KB for 1024x1024 — it lives in shared
//     memory.  The static `free` mask is shared by all goals and is read
//     through the read-only path (L1/L2 resident).
//   * A BFS level is processed only on ACTIVE tiles (tiles that gained a cell
//     in the previous level, or whose neighbour gained a cell on the shared
//     edge).  Warp w owns tile rows w, w+nWarps, ...; lane r owns row r of the
//     tile:  cand = (vis<<1 | vis>>1 | shfl_up(vis) | shfl_down(vis) | halo)
//     & open, with vis = free & ~open.  Expanding from all visited cells gives
//     the same level sets as expanding from the frontier only, and needs no
//     frontier bitmap.
//   * Cross-tile neighbours come from per-tile edge words (top/bottom row,
//     left/right column) double-buffered by level parity, so a tile never sees
//     cells its neighbour gained in the same level.
//   * One __syncthreads_or per level decides termination.
//   * Newly visited cells store their level straight into the int32 field.
//     Obstacle cells of a tile are written (MRP_INF) when the wavefront first
//     touches the tile, so that the partial sector writes of one tile meet in
//     L2 before eviction; tiles the BFS never touches and cells that stay
//     unreachable are written in a final coalesced sweep.  DRAM traffic is
//     therefore ~4 B per cell, the algorithmic minimum.
#include "common.cuh"

namespace mrp {

struct BfsLargeParams {
  const uint32_t* __restrict__ bits;  // free mask, tile-major
  const int32_t* __restrict__ goals;  // goal cells
  int32_t* __restrict__ out;          // [n_goals][cells]
  uint32_t* ws;                       // workspace: [0] goal counter, then state
  size_t wsWordsPerCta;               // global state words per CTA (0 if smem)
  int n_goals;
  int dimx, dimy, W, S, AW;           // AW = words of the active mask per stripe
};

constexpr int kWsHeaderWords = 64;

__device__ __forceinline__ void markTile(uint32_t* act, int AW, int s, int tx) {
  atomicOr(&act[s * AW + (tx >> 5)], 1u << (tx & 31));
}

template <bool kSmemState>
__global__ void __launch_bounds__(1024, 1)
bfs_large_kernel(BfsLargeParams p) {
  extern __shared__ uint32_t smem[];
  __shared__ int sGoal;
  const int nTiles = p.S * p.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nWarps = blockDim.x >> 5;
  const int cells = p.dimx * p.dimy;

  // state carve-up: act masks always in shared memory
  uint32_t* act = smem;                        // [2][S*AW]
  uint32_t* touched = act + 2 * p.S * p.AW;    // [S*AW] tile ever processed
  uint32_t* state = kSmemState
                        ? touched + p.S * p.AW
                        : p.ws + kWsHeaderWords + (size_t)blockIdx.x * p.wsWordsPerCta;
  uint32_t* open = state;                      // [nTiles*32]
  uint32_t* halo = open + (size_t)nTiles * 32; // [2][4][nTiles]: top,bot,left,right
  const int actWords = p.S * p.AW;

  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) sGoal = (int)atomicAdd(p.ws, 1u);
    __syncthreads();
    const int g = sGoal;
    if (g >= p.n_goals) break;
    const int goal = p.goals[g];
    int32_t* out = p.out + (size_t)g * cells;
    const int gx = goal % p.dimx, gy = goal / p.dimx;
    const int gs = gy >> 5, gtx = gx >> 5, gr = gy & 31, gb = gx & 31;
    const int gT = gs * p.W + gtx;
    const bool goalFree = (p.bits[gT * 32 + gr] >> gb) & 1u;

    // ---- init state ----
    for (int i = threadIdx.x; i < nTiles * 32; i += blockDim.x)
      open[i] = p.bits[i];
    for (int i = threadIdx.x; i < 8 * nTiles; i += blockDim.x) halo[i] = 0;
    for (int i = threadIdx.x; i < 3 * actWords; i += blockDim.x) act[i] = 0;
    __syncthreads();

    if (goalFree && threadIdx.x == 0) {
      // level 0: the goal cell
      open[gT * 32 + gr] &= ~(1u << gb);
      out[goal] = 0;
      uint32_t* h0 = halo;  // buffer 0
      if (gr == 0) h0[0 * nTiles + gT] = 1u << gb;
      if (gr == 31) h0[1 * nTiles + gT] = 1u << gb;
      if (gb == 0) h0[2 * nTiles + gT] = 1u << gr;
      if (gb == 31) h0[3 * nTiles + gT] = 1u << gr;
      uint32_t* a1 = act + actWords;  // buffer 1 (level 1)
      markTile(a1, p.AW, gs, gtx);
      if (gs > 0) markTile(a1, p.AW, gs - 1, gtx);
      if (gs < p.S - 1) markTile(a1, p.AW, gs + 1, gtx);
      if (gtx > 0) markTile(a1, p.AW, gs, gtx - 1);
      if (gtx < p.W - 1) markTile(a1, p.AW, gs, gtx + 1);
    }
    __syncthreads();

    if (goalFree) {
      for (int level = 1;; ++level) {
        const int rb = (level - 1) & 1, wb = level & 1;
        uint32_t* actCur = act + wb * actWords;         // tiles of this level
        uint32_t* actNext = act + (wb ^ 1) * actWords;  // marks for level+1
        const uint32_t* hr = halo + (size_t)rb * 4 * nTiles;
        uint32_t* hw = halo + (size_t)wb * 4 * nTiles;
        int marked = 0;
        for (int s = warp; s < p.S; s += nWarps) {
          for (int wi = 0; wi < p.AW; ++wi) {
            uint32_t mask = actCur[s * p.AW + wi];
            if (mask == 0) continue;
            const uint32_t oldTouched = touched[s * p.AW + wi];
            __syncwarp();
            if (lane == 0) {
              actCur[s * p.AW + wi] = 0;
              touched[s * p.AW + wi] = oldTouched | mask;
            }
            while (mask) {
              const int tx = wi * 32 + __ffs(mask) - 1;
              mask &= mask - 1;
              const int T = s * p.W + tx;
              uint32_t o = open[T * 32 + lane];
              const uint32_t fr = __ldg(&p.bits[T * 32 + lane]);
              const uint32_t vis = fr & ~o;
              int32_t* orow = out + (size_t)(s * 32 + lane) * p.dimx + tx * 32;
              if (!((oldTouched >> (tx & 31)) & 1u)) {
                // first touch of this tile: its obstacle cells get MRP_INF now,
                // while the tile's sectors are about to be filled anyway
                uint32_t ob = ~fr;
                const int xlim = p.dimx - tx * 32;
                if (xlim < 32) ob &= (1u << xlim) - 1u;
                if (s * 32 + lane >= p.dimy) ob = 0;
                while (ob) {
                  const int b = __ffs(ob) - 1;
                  ob &= ob - 1;
                  orow[b] = MRP_INF;
                }
              }
              uint32_t up = __shfl_up_sync(0xffffffffu, vis, 1);
              uint32_t dn = __shfl_down_sync(0xffffffffu, vis, 1);
              if (lane == 0) up = (s > 0) ? hr[1 * nTiles + T - p.W] : 0u;
              if (lane == 31) dn = (s < p.S - 1) ? hr[0 * nTiles + T + p.W] : 0u;
              const uint32_t lcol = (tx > 0) ? hr[3 * nTiles + T - 1] : 0u;
              const uint32_t rcol = (tx < p.W - 1) ? hr[2 * nTiles + T + 1] : 0u;
              const uint32_t cand =
                  ((vis << 1) | (vis >> 1) | up | dn | ((lcol >> lane) & 1u) |
                   (((rcol >> lane) & 1u) << 31)) & o;
              if (!__any_sync(0xffffffffu, cand)) continue;
              o &= ~cand;
              open[T * 32 + lane] = o;
              const uint32_t nvis = vis | cand;
              const uint32_t newL = __ballot_sync(0xffffffffu, cand & 1u);
              const uint32_t newR = __ballot_sync(0xffffffffu, cand >> 31);
              const uint32_t visL = __ballot_sync(0xffffffffu, nvis & 1u);
              const uint32_t visR = __ballot_sync(0xffffffffu, nvis >> 31);
              const uint32_t newTop = __shfl_sync(0xffffffffu, cand, 0);
              const uint32_t newBot = __shfl_sync(0xffffffffu, cand, 31);
              if (lane == 0) {
                hw[0 * nTiles + T] = nvis;
                hw[2 * nTiles + T] = visL;
                hw[3 * nTiles + T] = visR;
                markTile(actNext, p.AW, s, tx);
                if (newL && tx > 0) markTile(actNext, p.AW, s, tx - 1);
                if (newR && tx < p.W - 1) markTile(actNext, p.AW, s, tx + 1);
                if (newTop && s > 0) markTile(actNext, p.AW, s - 1, tx);
                if (newBot && s < p.S - 1) markTile(actNext, p.AW, s + 1, tx);
              }
              if (lane == 31) hw[1 * nTiles + T] = nvis;
              marked = 1;
              // distances of the newly visited cells of this row
              uint32_t m = cand;
              while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                orow[b] = level;
              }
            }
          }
        }
        if (!__syncthreads_or(marked)) break;
      }
    }

    // ---- MRP_INF for everything the wavefront did not reach ----
    // (obstacles, other components, and — for a goal on an obstacle — all
    // cells but the goal itself, which gets 0: the Floyd–Warshall row of an
    // isolated vertex)
    __syncthreads();
    for (int T = warp; T < nTiles; T += nWarps) {
      const int s = T / p.W, tx = T - s * p.W;
      const bool wasTouched = (touched[s * p.AW + (tx >> 5)] >> (tx & 31)) & 1u;
      // touched tiles already hold their obstacles; only cells that stayed
      // open (other components) are left.  Untouched tiles get every cell.
      uint32_t unreached = wasTouched ? open[T * 32 + lane] : 0xffffffffu;
      const int xlim = p.dimx - tx * 32;  // columns of this tile inside the map
      if (xlim < 32) unreached &= (1u << xlim) - 1u;
      if (s * 32 + lane >= p.dimy) unreached = 0;
      if (!__any_sync(0xffffffffu, unreached)) continue;
      // row by row: a fully unreached row is one coalesced 128-byte store
      for (int r = 0; r < 32; ++r) {
        const uint32_t u = __shfl_sync(0xffffffffu, unreached, r);
        if ((u >> lane) & 1u)
          out[(size_t)(s * 32 + r) * p.dimx + tx * 32 + lane] = MRP_INF;
      }
    }
    __syncthreads();
    if (!goalFree && threadIdx.x == 0) out[goal] = 0;
  }
}

static void bfsLargeGeometry(const mrp_map_s* map, int* nWarps, size_t* smemBytes,
                             size_t* stateWords, bool* smemState) {
  const int nTiles = map->S * map->W;
  const int AW = (map->W + 31) / 32;
  *stateWords = (size_t)nTiles * 32 + 8 * (size_t)nTiles;
  const size_t actBytes = 3 * (size_t)map->S * AW * 4;
  const size_t full = actBytes + *stateWords * 4;
  *smemState = full + 1024 <= ctx().smemOptin;
  *smemBytes = *smemState ? full : actBytes;
  *nWarps = map->S < 32 ? map->S : 32;
}

size_t bfsLargeWorkspaceBytes(const mrp_map_s* map, int n_goals) {
  (void)n_goals;
  int nWarps;
  size_t smemBytes, stateWords;
  bool smemState;
  bfsLargeGeometry(map, &nWarps, &smemBytes, &stateWords, &smemState);
  size_t words = kWsHeaderWords;
  if (!smemState) words += stateWords * (size_t)(ctx().smCount * 2);
  return words * 4;
}

int launchBfsLarge(const mrp_map_s* map, const int32_t* d_goal_cell,
                   int n_goals, int32_t* d_out, void* d_ws, cudaStream_t st) {
  if (n_goals <= 0) return 0;
  int nWarps;
  size_t smemBytes, stateWords;
  bool smemState;
  bfsLargeGeometry(map, &nWarps, &smemBytes, &stateWords, &smemState);
  BfsLargeParams p;
  p.bits = map->d_bits;
  p.goals = d_goal_cell;
  p.out = d_out;
  p.ws = static_cast<uint32_t*>(d_ws);
  p.wsWordsPerCta = smemState ? 0 : stateWords;
  p.n_goals = n_goals;
  p.dimx = map->dimx;
  p.dimy = map->dimy;
  p.W = map->W;
  p.S = map->S;
  p.AW = (map->W + 31) / 32;
  MRP_CUDA(cudaMemsetAsync(d_ws, 0, kWsHeaderWords * 4, st));
  auto kern = smemState ? bfs_large_kernel<true> : bfs_large_kernel<false>;
  MRP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smemBytes));
  int perSm = 1;
  MRP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, nWarps * 32,
                                                         smemBytes));
  if (perSm < 1) perSm = 1;
  if (!smemState && perSm > 2) perSm = 2;
  int blocks = ctx().smCount * perSm;
  if (blocks > n_goals) blocks = n_goals;
  kern<<<blocks, nWarps * 32, smemBytes, st>>>(p);
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mrp
