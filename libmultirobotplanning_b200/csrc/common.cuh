// common.cuh — shared helpers of the mrp_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/mrp_b200.h"

namespace mrp {

// Words per row of the bordered row-major bitmap (`d_rowbits`: bit x + 1 of row y + 1 = cell (x, y) is
// free): the columns plus a zero border on both sides, rounded to an ODD number of words (the cells of
// a diagonal wavefront then fall into distinct banks) and never fewer than three: the kernels divide
// word indices by this stride with a 32-bit multiply-shift whose constant 2^32 / stride does not exist
// for a stride of one (narrow maps, dimx <= 30, used to get one word per row and with it a zero
// constant: every cell was filed under row 0).
inline int bitmapRowWords(int dimx) {
  const int w = ((dimx + 2 + 31) / 32) | 1;
  return w < 3 ? 3 : w;
}
// word index -> bitmap row without a division: row = umulhi(word, magic) >> shift, exact for words
// below 2^27 (magic = ceil(2^(32 + shift) / stride), shift = floor(log2 stride); the stride is odd and
// at least 3, so the constant fits 32 bits).  Exported for the CPU tests (mrp_bitmap_row_division).
inline void bitmapRowDivision(int rowWords, uint32_t* magic, int* shift) {
  int sh = 0;
  while ((2 << sh) <= rowWords) ++sh;
  *shift = sh;
  *magic = (uint32_t)((((unsigned long long)1 << (32 + sh)) + (unsigned)rowWords - 1) / (unsigned)rowWords);
}

constexpr int kTile = 32;  // maps are tiled in 32x32-cell tiles, one bit per cell

// ---- error handling (thread-local message, never throws across the ABI) ----
std::string& lastErrorStorage();
int fail(int code, const char* fmt, ...);

#define MRP_CUDA(expr)                                                     \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess)                                                 \
      return ::mrp::fail(MRP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,     \
                         cudaGetErrorString(_e), __FILE__, __LINE__);      \
  } while (0)

#define MRP_CHECK(cond, code, ...)                         \
  do {                                                     \
    if (!(cond)) return ::mrp::fail(code, __VA_ARGS__);    \
  } while (0)

// ---- context ----
struct Context {
  cudaEvent_t doneEvent = nullptr;  // blocking-sync event: waitStream() sleeps instead of spinning
  bool ready = false;
  int device = -1;
  int smCount = 0;
  size_t smemOptin = 0;
  cudaStream_t stream = nullptr;   // compute
  cudaStream_t copyStream = nullptr;
  std::string info;
};
// Every host thread works in a LANE: its own streams, scratch buffers, replan
// arenas and API mutex, so that calls from different lanes overlap on the
// device (mrp_set_lane, include/mrp_b200.h).  ctx() / apiMutex() are the
// calling thread's lane's.
Context& ctx();
int ensureInit();
constexpr int kMaxLanes = 64;
struct LaneLL {  // grow-only device allocations of lowlevel.cu
  void* arena = nullptr;
  size_t arenaCap = 0;
  unsigned long long* hashArena = nullptr;
  size_t hashArenaEntries = 0;
  uint32_t launchSerial = 0;
};
LaneLL& laneLL();
// grow-only page-locked host staging buffers of the calling thread's lane.
// The results of the per-iteration calls come back through them: an
// asynchronous device-to-host copy into PAGEABLE memory makes the driver wait
// for the stream's kernel while it holds a context-wide lock, which serialises
// the lanes (measured: 16 lanes took 3x longer than one before this).
int pinnedScratch(int slot, size_t bytes, void** out);
// grow-only device scratch of the calling thread's lane (slots 16.. belong to pathpool.cu)
int deviceScratch(int slot, size_t bytes, void** out);
// Waits for the lane's compute stream without spinning (the batched drivers
// keep more host threads in flight than there are cores).
cudaError_t waitStream(cudaStream_t st);
extern std::atomic<long long> g_launches;
inline void countLaunch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- device map ----
}  // namespace mrp

struct mrp_map_s {
  int dimx, dimy;
  int W;   // tiles per row  = ceil(dimx/32)
  int S;   // tile rows (stripes) = ceil(dimy/32)
  // free mask, tile-major: bits[(s*W + tx)*32 + r] = row 32*s+r, word tx;
  // bit b of that word = cell x = 32*tx + b.  Bits outside the map are 0.
  uint32_t* d_bits;
  uint32_t* h_bits;  // host copy (CLI / validation)
  // free mask in 8x4-cell tiles (bfs_large.cu): bits84[ty*TW8 + tx], bit
  // 8*(y&3) + (x&7), TW8 = ceil(dimx/8), TH4 = ceil(dimy/4)
  uint32_t* d_bits84;
  // free mask, row-major with a one-cell border of zeros and two more zero
  // rows at the bottom, dimy+4 rows in all (bfs_queue.cu):
  // rowbits[(y+1)*WPR + ((x+1)>>5)] bit (x+1)&31, WPR = ceil((dimx+2)/32) | 1
  uint32_t* d_rowbits;
};

struct mrp_fieldset_s {
  int dimx, dimy, n_fields;
  int32_t* d_fields;  // [n_fields][dimx*dimy]
};

namespace mrp {

std::mutex& apiMutex();

// simple RAII device buffer for the host-pointer entry points
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int alloc(size_t n) {
    if (p) {
      cudaFree(p);
      p = nullptr;
    }
    bytes = n;
    if (n == 0) return 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess)
      return fail(MRP_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", n,
                  cudaGetErrorString(e));
    return 0;
  }
  template <class T>
  T* as() { return static_cast<T*>(p); }
};

// kernels launchers (defined in the .cu files)
int launchBfsSmall(const uint32_t* d_rows, const int32_t* d_dims,
                   const int4* d_jobs, int n_jobs, int32_t* d_out,
                   cudaStream_t st);
size_t bfsLargeWorkspaceBytes(const mrp_map_s* map, int n_goals);
int launchBfsLarge(const mrp_map_s* map, const int32_t* d_goal_cell,
                   int n_goals, int32_t* d_out, void* d_ws, cudaStream_t st);
bool bfsQueueFits(const mrp_map_s* map);
size_t bfsQueueWorkspaceWords(int n_goals);
int launchBfsQueue(const mrp_map_s* map, const int32_t* d_goal_cell,
                   int n_goals, int32_t* d_out, void* d_ws, cudaStream_t st,
                   const uint32_t* d_goalList = nullptr,
                   const uint32_t* d_goalListCount = nullptr);
// caps the CTAs of the next BFS launches of the calling thread (0 = no cap):
// the sharded gather (multi.cu) leaves a few SMs to the collective
void setBfsBlockCap(int blocks);
int bfsBlockCap();
bool bfsSweepFits(const mrp_map_s* map);
size_t bfsSweepWorkspaceWords(int n_goals);
int launchBfsSweep(const mrp_map_s* map, const int32_t* d_goal_cell,
                   int n_goals, int32_t* d_out, void* d_ws, cudaStream_t st);
size_t conflictsWorkspaceBytes(int N, int Tpad);
int launchConflicts(const int32_t* d_cell, const int32_t* d_len, int N,
                    int Tpad, int mode, bool wantFirst, bool wantCount,
                    unsigned long long* d_result, void* d_ws, size_t wsBytes,
                    cudaStream_t st);
int launchConflictsPairShard(const int32_t* d_cell, const int32_t* d_len, int N,
                             int Tpad, int mode, bool wantFirst, bool wantCount,
                             unsigned long long* d_result, int shard, int nShards,
                             cudaStream_t st);
int launchConflictsBatch(const int32_t* d_cell, const int32_t* d_len, int B,
                         int N, int Tpad, int mode,
                         unsigned long long* d_result, cudaStream_t st);
int launchFocalCounts(const int32_t* d_cell, const int32_t* d_len, int N,
                      int Tpad, int self, const int32_t* d_t,
                      const int32_t* d_from, const int32_t* d_to, int n_cand,
                      int32_t* d_state, int32_t* d_trans, cudaStream_t st);

// packed first-conflict key: t<<41 | type<<40 | i<<20 | j   (min = first)
__host__ __device__ inline unsigned long long conflictKey(int t, int type, int i,
                                                         int j) {
  return ((unsigned long long)t << 41) | ((unsigned long long)type << 40) |
         ((unsigned long long)i << 20) | (unsigned long long)j;
}
constexpr unsigned long long kNoConflict = ~0ull;
constexpr int kMaxAgents = 1 << 20;
constexpr int kMaxTime = 1 << 22;

}  // namespace mrp
