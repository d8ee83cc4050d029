// multi.cu — the multi-GPU side of the library: one process (or host thread) per
// GPU, an NCCL communicator behind the C ABI, and the two sharded operations of
// the hot path that end in a collective:
//
//   * distance fields by goal + all-gather (north_star: "an NCCL all-gather over
//     NVLink to make every goal's distance field resident on every GPU"; the
//     consumer that needs every field everywhere is the agent x goal cost matrix
//     of cbs_ta, example/cbs_ta.cpp:272-280).  Rank r computes the fields of its
//     slice of the goal list in chunks of one wave of goals; a finished chunk
//     leaves as ONE BYTE per cell — the detour (distance - Manhattan distance) / 2
//     the obstacles force, 255 = MRP_INF: a quarter of the int32 bytes on the
//     wire — through ncclAllGather on a second stream while the next chunk is
//     being computed, and the bytes of the other ranks are expanded to int32 on
//     the device.  A field whose detours do not fit a byte (mazes) raises a flag
//     that all ranks see (all-reduce MAX); the call then repeats the gather with
//     int32.
//   * conflict checks by agent-pair block: every rank holds the path table and
//     sweeps every n-th 64x64 block of agent pairs with the all-pairs kernel;
//     first-conflict keys meet in an all-reduce MIN, counts in a SUM
//     (example/cbs.cpp:343-383 gives the order the key encodes).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the process already
// holds — e.g. the one PyTorch ships — or the system's), so libmrp_b200.so has no
// link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mrp {
namespace {

struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetVersion) getVersion = nullptr;
  decltype(&ncclGetUniqueId) getUniqueId = nullptr;
  decltype(&ncclCommInitRank) commInitRank = nullptr;
  decltype(&ncclCommDestroy) commDestroy = nullptr;
  decltype(&ncclAllGather) allGather = nullptr;
  decltype(&ncclAllReduce) allReduce = nullptr;
  decltype(&ncclGetErrorString) getErrorString = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    const char* names[] = {getenv("MRP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) return a;
#define MRP_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name))
    MRP_SYM(getVersion, "ncclGetVersion");
    MRP_SYM(getUniqueId, "ncclGetUniqueId");
    MRP_SYM(commInitRank, "ncclCommInitRank");
    MRP_SYM(commDestroy, "ncclCommDestroy");
    MRP_SYM(allGather, "ncclAllGather");
    MRP_SYM(allReduce, "ncclAllReduce");
    MRP_SYM(getErrorString, "ncclGetErrorString");
#undef MRP_SYM
    a.ok = a.getVersion && a.getUniqueId && a.commInitRank && a.commDestroy && a.allGather && a.allReduce &&
           a.getErrorString;
    return a;
  }();
  return api;
}

#define MRP_NCCL(expr)                                                                          \
  do {                                                                                          \
    ncclResult_t _r = (expr);                                                                   \
    if (_r != ncclSuccess)                                                                      \
      return ::mrp::fail(MRP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, nccl().getErrorString(_r), \
                         __FILE__, __LINE__);                                                   \
  } while (0)

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nRanks = 1, device = -1;
  cudaStream_t stream = nullptr;   // collectives
  cudaStream_t xstream = nullptr;  // expansion of received chunks (overlaps the next collective)
  cudaEvent_t packed[2] = {nullptr, nullptr};    // chunk is packed (compute stream)
  cudaEvent_t gathered[2] = {nullptr, nullptr};  // chunk has arrived (collective stream)
  cudaEvent_t expanded[2] = {nullptr, nullptr};  // chunk is expanded (expansion stream)
  static constexpr int kTimed = 64;              // collectives of a call that are timed
  cudaEvent_t tc[2 * kTimed] = {};               // around each collective of the last call
  int nTimed = 0;
  long long lastBytesWire = 0, timedBytesWire = 0;
  int lastFormat = 0;
};
Comm g_comm;

// ---- pack / expand (detour bytes) ------------------------------------------------
// grid: x over the 4-cell groups of a field (one int4 <-> one packed word per
// thread: both sides of the copy are fully coalesced), y over fields; 32-bit
// index math only.
// in: int32 fields [n][cells] of goals goalCell[0 .. n); out: bytes [n][cells].
__global__ void __launch_bounds__(256) pack_u8_kernel(const int32_t* __restrict__ in, uint8_t* __restrict__ out,
                                                      int dimx, int cells, const int32_t* __restrict__ goalCell,
                                                      int* __restrict__ overflow) {
  const int f = blockIdx.y;
  const int g = __ldg(goalCell + f);
  const int gx = g % dimx, gy = g / dimx;
  const int4* src = reinterpret_cast<const int4*>(in + (size_t)f * cells);
  uint32_t* dst = reinterpret_cast<uint32_t*>(out + (size_t)f * cells);
  bool ovf = false;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < cells / 4; k += gridDim.x * blockDim.x) {
    const int cell = k * 4;
    const int y0 = cell / dimx, x0 = cell - y0 * dimx;
    const bool oneRow = x0 + 3 < dimx;
    const int dy0 = abs(y0 - gy);
    const int4 a = __ldcs(src + k);
    const int v[4] = {a.x, a.y, a.z, a.w};
    uint32_t w = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = cell + e;
      const int m = oneRow ? abs(x0 + e - gx) + dy0 : abs(c % dimx - gx) + abs(c / dimx - gy);
      uint32_t b = 255u;
      if (v[e] != MRP_INF) {
        const int h = (v[e] - m) >> 1;
        if (h >= 255) ovf = true;
        b = (uint32_t)h & 255u;
      }
      w |= b << (8 * e);
    }
    dst[k] = w;
  }
  if (__syncthreads_or(ovf) && threadIdx.x == 0) atomicOr(overflow, 1);
}

// in: bytes of `per` fields from each of nRanks ranks ([rank][per][cells]); the
// field at (rank r, slot s) = blockIdx.y is goal r*perRank + chunkFirst + s; the
// own rank's fields are already in place.
__global__ void __launch_bounds__(256) expand_u8_kernel(const uint8_t* __restrict__ in, int32_t* __restrict__ out,
                                                        int dimx, int cells, const int32_t* __restrict__ goalCell,
                                                        int nGoals, int perRank, int chunkFirst, int per, int self) {
  const int fs = blockIdx.y;
  const int r = fs / per, s = fs - r * per;
  if (r == self) return;
  const int goal = r * perRank + chunkFirst + s;
  if (chunkFirst + s >= perRank || goal >= nGoals) return;
  const int g = __ldg(goalCell + goal);
  const int gx = g % dimx, gy = g / dimx;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (size_t)fs * cells);
  int4* dst = reinterpret_cast<int4*>(out + (size_t)goal * cells);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < cells / 4; k += gridDim.x * blockDim.x) {
    const int cell = k * 4;
    const uint32_t w = __ldcs(src + k);
    const int y0 = cell / dimx, x0 = cell - y0 * dimx;
    const bool oneRow = x0 + 3 < dimx;
    const int dy0 = abs(y0 - gy);
    int v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = cell + e;
      const uint32_t h = (w >> (8 * e)) & 255u;
      const int m = oneRow ? abs(x0 + e - gx) + dy0 : abs(c % dimx - gx) + abs(c / dimx - gy);
      v[e] = h == 255u ? MRP_INF : (int)(2u * h) + m;
    }
    __stcs(dst + k, make_int4(v[0], v[1], v[2], v[3]));
  }
}

// int32 fall-back of the gather: copies the received fields of the other ranks into place
__global__ void __launch_bounds__(256) place_i32_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                        int cells, int nGoals, int perRank, int chunkFirst, int per,
                                                        int self) {
  const int fs = blockIdx.y;
  const int r = fs / per, s = fs - r * per;
  if (r == self) return;
  const int goal = r * perRank + chunkFirst + s;
  if (chunkFirst + s >= perRank || goal >= nGoals) return;
  const int4* src = reinterpret_cast<const int4*>(in + (size_t)fs * cells);
  int4* dst = reinterpret_cast<int4*>(out + (size_t)goal * cells);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < cells / 4; k += gridDim.x * blockDim.x)
    __stcs(dst + k, __ldcs(src + k));
}

int gatherSpareSms() {
  const char* e = getenv("MRP_GATHER_SPARE_SMS");
  return e ? std::max(0, atoi(e)) : 16;
}

// goals per chunk: one wave of the queue kernel on the SMs it is given (measured, 4096
// goals over 8 GPUs: one wave / 16 spare SMs 13.4 ms, two waves 14.7 ms, a single chunk
// without overlap 14.4 ms; 2048 goals over 2 GPUs: 11.3 / 11.1 / 11.9 ms)
int chunkGoals() {
  const char* e = getenv("MRP_GATHER_CHUNK");
  return e ? std::max(1, atoi(e)) : std::max(1, ctx().smCount - gatherSpareSms());
}

struct GatherPlan {
  int perRank, chunk, nChunks;
  size_t bfsWs, sendBytes, recvBytes, total;
};

GatherPlan gatherPlan(const mrp_map_s* map, int nGoals, int nRanks) {
  GatherPlan g;
  const size_t cells = (size_t)map->dimx * map->dimy;
  g.perRank = (nGoals + nRanks - 1) / nRanks;
  g.chunk = std::min(chunkGoals(), std::max(g.perRank, 1));
  g.nChunks = (g.perRank + g.chunk - 1) / g.chunk;
  g.bfsWs = (bfsLargeWorkspaceBytes(map, g.chunk) + 255) & ~(size_t)255;
  // sized for the int32 fall-back
  g.sendBytes = (size_t)g.chunk * cells * 4;
  g.recvBytes = g.sendBytes * nRanks;
  g.total = g.bfsWs + 2 * g.sendBytes + 2 * g.recvBytes + 256;
  return g;
}

}  // namespace
}  // namespace mrp

using namespace mrp;

extern "C" {

int mrp_comm_unique_id(void* id128) {
  MRP_CHECK(id128 != nullptr, MRP_ERR_INVALID, "id is NULL");
  MRP_CHECK(nccl().ok, MRP_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded: %s", dlerror());
  static_assert(sizeof(ncclUniqueId) == MRP_COMM_ID_BYTES, "NCCL id size");
  ncclUniqueId id;
  MRP_NCCL(nccl().getUniqueId(&id));
  std::memcpy(id128, &id, sizeof id);
  return 0;
}

int mrp_comm_init_rank(const void* id128, int n_ranks, int rank) {
  std::lock_guard<std::mutex> lk(apiMutex());
  MRP_CHECK(id128 != nullptr && n_ranks >= 1 && rank >= 0 && rank < n_ranks, MRP_ERR_INVALID, "bad rank %d of %d",
            rank, n_ranks);
  MRP_CHECK(nccl().ok, MRP_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded");
  MRP_CHECK(g_comm.comm == nullptr, MRP_ERR_INVALID, "communicator already initialised");
  if (int rc = ensureInit()) return rc;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  MRP_CUDA(cudaSetDevice(ctx().device));
  MRP_NCCL(nccl().commInitRank(&g_comm.comm, n_ranks, id, rank));
  g_comm.rank = rank;
  g_comm.nRanks = n_ranks;
  g_comm.device = ctx().device;
  MRP_CUDA(cudaStreamCreateWithFlags(&g_comm.stream, cudaStreamNonBlocking));
  MRP_CUDA(cudaStreamCreateWithFlags(&g_comm.xstream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    MRP_CUDA(cudaEventCreateWithFlags(&g_comm.gathered[i], cudaEventDisableTiming));
    MRP_CUDA(cudaEventCreateWithFlags(&g_comm.packed[i], cudaEventDisableTiming));
    MRP_CUDA(cudaEventCreateWithFlags(&g_comm.expanded[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < 2 * Comm::kTimed; ++i) MRP_CUDA(cudaEventCreate(&g_comm.tc[i]));
  return 0;
}

int mrp_comm_info(int* rank, int* n_ranks, int* nccl_version) {
  if (rank) *rank = g_comm.comm ? g_comm.rank : 0;
  if (n_ranks) *n_ranks = g_comm.comm ? g_comm.nRanks : 1;
  if (nccl_version) {
    *nccl_version = 0;
    if (nccl().ok) nccl().getVersion(nccl_version);
  }
  return g_comm.comm ? 1 : 0;
}

int mrp_comm_destroy(void) {
  std::lock_guard<std::mutex> lk(apiMutex());
  if (!g_comm.comm) return 0;
  cudaSetDevice(g_comm.device);
  cudaStreamSynchronize(g_comm.stream);
  cudaStreamSynchronize(g_comm.xstream);
  nccl().commDestroy(g_comm.comm);
  cudaStreamDestroy(g_comm.stream);
  cudaStreamDestroy(g_comm.xstream);
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(g_comm.gathered[i]);
    cudaEventDestroy(g_comm.packed[i]);
    cudaEventDestroy(g_comm.expanded[i]);
  }
  for (int i = 0; i < 2 * Comm::kTimed; ++i) cudaEventDestroy(g_comm.tc[i]);
  g_comm = Comm();
  return 0;
}

size_t mrp_bfs_allgather_workspace_bytes(mrp_map map, int n_goals) {
  std::lock_guard<std::mutex> lk(apiMutex());
  if (!map || ensureInit() != 0) return 0;
  return gatherPlan(map, std::max(n_goals, 1), g_comm.comm ? g_comm.nRanks : 1).total;
}

// one pass of the sharded computation + gather in the given wire format (1 = detour bytes, 4 = int32)
static int gatherPass(const mrp_map_s* map, const int32_t* d_goal_cell, int nGoals, int32_t* d_out, char* ws,
                      const GatherPlan& g, int fmt, cudaStream_t st, int* d_flag) {
  Comm& c = g_comm;
  const int cells = map->dimx * map->dimy;
  const int myFirst = c.rank * g.perRank;
  const int myCount = std::max(0, std::min(g.perRank, nGoals - myFirst));
  char* bfsWs = ws;
  char* send[2] = {ws + g.bfsWs, ws + g.bfsWs + g.sendBytes};
  char* recv[2] = {ws + g.bfsWs + 2 * g.sendBytes, ws + g.bfsWs + 2 * g.sendBytes + g.recvBytes};
  const size_t chunkWire = (size_t)g.chunk * cells * fmt;
  // blocks per field of the pack / expand kernels: 4 cells per thread and iteration, 4 iterations
  const int gx16 = std::max(1, std::min(256, cells / 4 / (256 * 4)));
  c.nTimed = 0;
  c.timedBytesWire = 0;
  for (int k = 0; k < g.nChunks; ++k) {
    const int b = k & 1;
    const int first = k * g.chunk;                                  // within the rank's slice
    const int n = std::max(0, std::min(g.chunk, myCount - first));  // goals of this rank in the chunk
    int32_t* mine = d_out + (size_t)(myFirst + first) * cells;
    if (n > 0) {
      // leave a few SMs to the collective and the expansion of the previous chunk
      setBfsBlockCap(std::max(1, ctx().smCount - gatherSpareSms()));
      const int rc = launchBfsLarge(map, d_goal_cell + myFirst + first, n, mine, bfsWs, st);
      setBfsBlockCap(0);
      if (rc) return rc;
    }
    // the buffers of chunk k - 2 must have been gathered and expanded
    if (k >= 2) MRP_CUDA(cudaStreamWaitEvent(st, c.expanded[b], 0));
    const void* src = mine;
    if (fmt == 1) {
      if (n > 0) {
        pack_u8_kernel<<<dim3(gx16, n), 256, 0, st>>>(mine, reinterpret_cast<uint8_t*>(send[b]), map->dimx, cells,
                                                     d_goal_cell + myFirst + first, d_flag);
        countLaunch();
      }
      src = send[b];
    } else if (n < g.chunk) {
      // a short (or empty) last chunk: the collective still moves g.chunk fields per rank
      if (n > 0) MRP_CUDA(cudaMemcpyAsync(send[b], mine, (size_t)n * cells * 4, cudaMemcpyDeviceToDevice, st));
      src = send[b];
    }
    MRP_CUDA(cudaEventRecord(c.packed[b], st));
    MRP_CUDA(cudaStreamWaitEvent(c.stream, c.packed[b], 0));
    // the receive buffer of chunk k - 2 must have been expanded
    if (k >= 2) MRP_CUDA(cudaStreamWaitEvent(c.stream, c.expanded[b], 0));
    const bool timed = c.nTimed < Comm::kTimed;
    if (timed) MRP_CUDA(cudaEventRecord(c.tc[2 * c.nTimed], c.stream));
    MRP_NCCL(nccl().allGather(src, recv[b], chunkWire, ncclUint8, c.comm, c.stream));
    c.lastBytesWire += (long long)chunkWire * (c.nRanks - 1);
    if (timed) {
      MRP_CUDA(cudaEventRecord(c.tc[2 * c.nTimed + 1], c.stream));
      c.timedBytesWire += (long long)chunkWire * (c.nRanks - 1);
      ++c.nTimed;
    }
    MRP_CUDA(cudaEventRecord(c.gathered[b], c.stream));
    MRP_CUDA(cudaStreamWaitEvent(c.xstream, c.gathered[b], 0));
    const dim3 eg(gx16, c.nRanks * g.chunk);
    if (fmt == 1)
      expand_u8_kernel<<<eg, 256, 0, c.xstream>>>(reinterpret_cast<const uint8_t*>(recv[b]), d_out, map->dimx, cells,
                                                 d_goal_cell, nGoals, g.perRank, first, g.chunk, c.rank);
    else
      place_i32_kernel<<<eg, 256, 0, c.xstream>>>(reinterpret_cast<const int32_t*>(recv[b]), d_out, cells, nGoals,
                                                 g.perRank, first, g.chunk, c.rank);
    countLaunch();
    MRP_CUDA(cudaEventRecord(c.expanded[b], c.xstream));
  }
  // the caller's stream continues when everything is in place
  MRP_CUDA(cudaStreamWaitEvent(st, c.expanded[(g.nChunks - 1) & 1], 0));
  if (g.nChunks >= 2) MRP_CUDA(cudaStreamWaitEvent(st, c.expanded[g.nChunks & 1], 0));
  MRP_CUDA(cudaGetLastError());
  return 0;
}

int mrp_bfs_fields_allgather_dev(mrp_map map, const int32_t* d_goal_cell, int n_goals, int32_t* d_out,
                                 void* d_workspace, void* stream) {
  std::lock_guard<std::mutex> lk(apiMutex());
  MRP_CHECK(map != nullptr && n_goals >= 0, MRP_ERR_INVALID, "bad arguments");
  if (n_goals == 0) return 0;
  MRP_CHECK(d_goal_cell && d_out && d_workspace, MRP_ERR_INVALID, "NULL device pointer");
  if (int rc = ensureInit()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Comm& c = g_comm;
  if (!c.comm || c.nRanks == 1)  // a single rank owns every goal: nothing to exchange
    return launchBfsLarge(map, d_goal_cell, n_goals, d_out, d_workspace, st);
  const int cells = map->dimx * map->dimy;
  MRP_CHECK(cells % 16 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0, MRP_ERR_UNSUPPORTED,
            "the gathered layout needs dimx*dimy %% 16 == 0 and a 16-byte aligned output");
  const GatherPlan g = gatherPlan(map, n_goals, c.nRanks);
  char* ws = static_cast<char*>(d_workspace);
  int* d_flag = reinterpret_cast<int*>(ws + g.total - 256);
  c.lastBytesWire = 0;
  c.lastFormat = 1;
  MRP_CUDA(cudaMemsetAsync(d_flag, 0, 8, st));
  const bool bytes = !getenv("MRP_GATHER_I32");
  if (bytes) {
    if (int rc = gatherPass(map, d_goal_cell, n_goals, d_out, ws, g, 1, st, d_flag)) return rc;
    // does any rank hold a field whose detours do not fit a byte?  (one int; all ranks must agree)
    MRP_NCCL(nccl().allReduce(d_flag, d_flag + 1, 1, ncclInt32, ncclMax, c.comm, st));
    int flag = 0;
    MRP_CUDA(cudaMemcpyAsync(&flag, d_flag + 1, 4, cudaMemcpyDeviceToHost, st));
    MRP_CUDA(cudaStreamSynchronize(st));
    if (!flag) return 0;
  }
  c.lastFormat = 4;
  c.lastBytesWire = 0;
  return gatherPass(map, d_goal_cell, n_goals, d_out, ws, g, 4, st, d_flag);
}

int mrp_comm_last_gather(double* collective_ms, long long* timed_bytes_in, long long* wire_bytes_in,
                         int* bytes_per_cell) {
  Comm& c = g_comm;
  double sum = 0;
  for (int i = 0; c.comm && i < c.nTimed; ++i) {
    float ms = 0;
    if (cudaEventSynchronize(c.tc[2 * i + 1]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, c.tc[2 * i], c.tc[2 * i + 1]) == cudaSuccess)
      sum += ms;
  }
  if (collective_ms) *collective_ms = sum;
  if (timed_bytes_in) *timed_bytes_in = c.timedBytesWire;
  if (wire_bytes_in) *wire_bytes_in = c.lastBytesWire;
  if (bytes_per_cell) *bytes_per_cell = c.lastFormat;
  return 0;
}

// Conflict checks by agent-pair block over the ranks.  d_result as in
// mrp_conflicts_dev; every rank gets the global answer.
int mrp_conflicts_sharded_dev(const int32_t* d_cell, const int32_t* d_len, int N, int Tpad, int mode,
                              int want_first, int want_count, unsigned long long* d_result, void* stream) {
  std::lock_guard<std::mutex> lk(apiMutex());
  MRP_CHECK(d_cell && d_len && d_result, MRP_ERR_INVALID, "NULL device pointer");
  MRP_CHECK(N >= 2 && N < kMaxAgents && Tpad > 0 && Tpad < kMaxTime, MRP_ERR_INVALID, "bad table shape N=%d Tpad=%d",
            N, Tpad);
  MRP_CHECK(want_first || want_count, MRP_ERR_INVALID, "nothing requested");
  if (int rc = ensureInit()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Comm& c = g_comm;
  const int n = c.comm ? c.nRanks : 1, r = c.comm ? c.rank : 0;
  if (int rc = launchConflictsPairShard(d_cell, d_len, N, Tpad, mode, want_first != 0, want_count != 0, d_result, r,
                                        n, st))
    return rc;
  if (n > 1) {
    // keys are unsigned and "no conflict" is all ones: MIN picks the first conflict
    MRP_NCCL(nccl().allReduce(d_result, d_result, 1, ncclUint64, ncclMin, c.comm, st));
    MRP_NCCL(nccl().allReduce(d_result + 1, d_result + 1, 1, ncclUint64, ncclSum, c.comm, st));
  }
  return 0;
}

}  // extern "C"
