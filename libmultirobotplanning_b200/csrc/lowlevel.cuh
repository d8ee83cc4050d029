// lowlevel.cuh — what the two replan kernels share (lowlevel.cu: any map, both
// variants; lowlevel_tile.cu: single-tile maps, cbs/ecbs moves, everything a
// search touches per expansion in shared memory).
#pragma once

#include "common.cuh"

namespace mrp {

constexpr int kFMax = 2048;     // f values tracked by the histogram (general kernel)
constexpr int kConsCache = 96;  // constraints cached in shared memory per warp
constexpr int kLenS = 256;      // path lengths of the other agents cached per warp

struct LLParams {
  const uint32_t* const* mapBits;  // per map: 32x32-tile free mask (tile-major)
  int dimx, dimy, W;               // common map geometry (W = ceil(dimx/32))
  const int32_t* fields;           // [n_fields][cells]
  const int32_t* vc;               // [n_vc][2]
  const int32_t* ec;               // [n_ec][3]
  const int32_t* tables;           // [n_tables][N][Tpad]
  const int32_t* tableLen;         // [n_tables][N]
  int N, Tpad;
  const mrp_job* jobs;
  int n_jobs;
  // order in which the persistent grid takes the jobs (NULL: 0, 1, 2, ...); n_jobs
  // counts the entries of the list
  const int32_t* jobList;
  int variant;
  float w;
  int focalMode;  // 0: A*, 1: A*-epsilon
  int maxExpanded, pathCap;
  mrp_path_info* info;
  int32_t* outCells;
  int32_t* outG;
  // workspace
  unsigned int* counter;
  int maxNodes, hashCap;  // per warp slot
  uint32_t* nodeKey;
  int32_t* nodeG;
  int32_t* nodeF;
  int32_t* nodeFocal;
  int32_t* nodeParent;
  unsigned long long* openKey;  // packed OPEN entries beyond the shared-memory front
  unsigned long long* hashKey;  // (generation << 32 | state): never cleared per job
  int32_t* hashVal;
  uint32_t genBase;             // generation of job j of this launch = genBase + j + 1
  // tile kernel only: occupancy planes of the path tables, [n_tables][Tpad][32 rows] of
  // (cells holding >= 1 agent, cells holding >= 2 agents) at that time, every agent of
  // the table counted and parked agents repeated to the last row; occMany[table] != 0:
  // some cell held three or more agents (the planes no longer count exactly)
  const uint2* occ;
  const int32_t* occMany;
  int TB;  // rows of the visited bitmap (time steps a search may reach)
  // path pool mode (pathpool.cu): the path of job j goes to pool row outSlots[j]
  // (cells only: with unit move costs the g-score of a state is its time step),
  // nothing is written to outCells / outG
  const int32_t* outSlots;
  int32_t* const* poolCells;  // per chunk: [kPoolChunk][rowCap]
  int32_t* const* poolLen;    // per chunk: [kPoolChunk]
  int rowCap;
  // sliced searches (tile kernel): a job with jobState[j] >= 0 stops after sliceCap
  // expansions of this launch (status kTileStatusSuspended) and leaves its search state
  // in blob jobState[j]; jobResume[j] != 0 continues from that blob
  const int32_t* jobState;
  const int32_t* jobResume;
  unsigned char* const* stateChunks;  // per chunk: [kStateChunk][blobBytes]
  size_t blobBytes;
  int sliceCap;
};

constexpr int kPoolChunkBits = 15;
constexpr int kPoolChunk = 1 << kPoolChunkBits;  // pool rows per device allocation

// where job `job` writes its path: its rows of the call's output arrays, or its pool row
__device__ __forceinline__ void pathOutput(const LLParams& p, int job, int32_t*& oc, int32_t*& og, int& cap) {
  if (p.outSlots) {
    const int slot = p.outSlots[job];
    oc = p.poolCells[slot >> kPoolChunkBits] + (size_t)(slot & (kPoolChunk - 1)) * p.rowCap;
    og = nullptr;
    cap = p.rowCap;
  } else {
    oc = p.outCells + (size_t)job * p.pathCap;
    og = p.outG + (size_t)job * p.pathCap;
    cap = p.pathCap;
  }
}
__device__ __forceinline__ void pathLength(const LLParams& p, int job, int len) {
  if (p.outSlots) {
    const int slot = p.outSlots[job];
    p.poolLen[slot >> kPoolChunkBits][slot & (kPoolChunk - 1)] = len;
  }
}

// pool mode of the host entry point (pathpool.cu); all pointers are device pointers
struct LLPool {
  const int32_t* d_tables = nullptr;  // [n_tables][N][Tpad], gathered from the pool
  const int32_t* d_tlen = nullptr;
  int32_t* const* d_poolCells = nullptr;
  int32_t* const* d_poolLen = nullptr;
  const int32_t* h_outSlots = nullptr;  // host: [n_jobs]
  int rowCap = 0;
  // sliced searches: host arrays [n_jobs] (NULL: every job runs to its end)
  const int32_t* h_jobState = nullptr;
  const int32_t* h_jobResume = nullptr;
  unsigned char* const* d_stateChunks = nullptr;
  size_t blobBytes = 0;
  int sliceCap = 0;
  int tileTB = 0;  // rows of the visited bitmap the blobs were laid out for
};
int lowlevelRun(const mrp_map* maps, int n_maps, const int32_t* fields, const int32_t* d_fields,
                int n_fields, const int32_t* vc, int n_vc, const int32_t* ec, int n_ec,
                const int32_t* tables, const int32_t* table_len, int n_tables, int N, int Tpad,
                const mrp_job* jobs, int n_jobs, const mrp_lowlevel_params* params,
                mrp_path_info* info, int32_t* out_cells, int32_t* out_g, const LLPool* pool);

__device__ __forceinline__ bool cellFree(const uint32_t* bits, int W, int dimx, int dimy,
                                         int x, int y) {
  if (x < 0 || y < 0 || x >= dimx || y >= dimy) return false;
  return (bits[((y >> 5) * W + (x >> 5)) * 32 + (y & 31)] >> (x & 31)) & 1u;
}

__device__ __forceinline__ uint32_t hashState(uint32_t k) {
  k ^= k >> 16;
  k *= 0x7feb352du;
  k ^= k >> 15;
  k *= 0x846ca68bu;
  k ^= k >> 16;
  return k;
}

// 64-bit minimum over the warp: two REDUX instructions (high word, then the low
// word among the lanes that hold the minimal high word) instead of ten shuffles
__device__ __forceinline__ unsigned long long warpMin64(unsigned long long v) {
  const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
  const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
  const uint32_t mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  return ((unsigned long long)mhi << 32) | mlo;
}

// packed OPEN entry: (focal:14 | f:12 | 4095-g:12 | node:26); smaller = better
__device__ __forceinline__ unsigned long long packOpenKey(int focal, int f, int g, int node) {
  return ((unsigned long long)min(focal, 16383) << 50) | ((unsigned long long)f << 38) |
         ((unsigned long long)(4095 - min(g, 4095)) << 26) | (unsigned long long)node;
}

// ---- tile kernel (lowlevel_tile.cu) ----
constexpr int kTileStatusRedo = 3;  // job left to the general kernel
constexpr int kTileStatusSuspended = 4;  // slice used up: continue from the state blob (MRP_SUSPENDED)
constexpr int kStateChunkBits = 6;
constexpr int kStateChunk = 1 << kStateChunkBits;  // state blobs per device allocation
size_t lowlevelTileBlobBytes(int TB, int maxNodes);
bool lowlevelTileEligible(const LLParams& p, int n_tables);
size_t lowlevelTileOccBytes(int n_tables, int Tpad);
int lowlevelTileSlots(const LLParams& p);
int launchLowlevelTile(const LLParams& p, int n_tables, uint2* d_occ, int32_t* d_occMany, int slots,
                       cudaStream_t st);

}  // namespace mrp
