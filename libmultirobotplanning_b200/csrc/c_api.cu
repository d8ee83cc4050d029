// c_api.cu — the extern "C" surface declared in include/mrp_b200.h: context,
// maps, and the host-pointer convenience entry points (they stage H2D/D2H
// copies around the kernels of bfs_small.cu / bfs_large.cu / conflicts.cu /
// lowlevel.cu).  No CPU fallback anywhere: without a device every call fails.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace mrp {

std::atomic<long long> g_launches{0};

std::string& lastErrorStorage() {
  thread_local std::string s;
  return s;
}

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  lastErrorStorage() = buf;
  return code;
}

// grow-only device scratch slots for the host-pointer entry points
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
  int get(size_t bytes, void** out) {
    if (bytes > cap) {
      if (p) cudaFree(p);
      p = nullptr;
      cap = 0;
      size_t want = std::max(bytes, (size_t)1 << 16);
      want += want / 4;
      cudaError_t e = cudaMalloc(&p, want);
      if (e != cudaSuccess)
        return fail(MRP_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want,
                    cudaGetErrorString(e));
      cap = want;
    }
    *out = p;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct PinnedScratch {
  void* p = nullptr;
  size_t cap = 0;
  int get(size_t bytes, void** out) {
    if (bytes > cap) {
      if (p) cudaFreeHost(p);
      p = nullptr;
      cap = 0;
      size_t want = std::max(bytes, (size_t)1 << 16);
      want += want / 4;
      cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
      if (e != cudaSuccess)
        return fail(MRP_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
      cap = want;
    }
    *out = p;
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct Lane {
  Context c;
  std::mutex m;
  Scratch scratch[24];
  PinnedScratch pinned[10];
  LaneLL ll;
};
static Lane g_lanes[kMaxLanes];
static thread_local int t_lane = 0;
static std::atomic<int> g_device{-1};  // the device of the first initialised lane

static Lane& lane() { return g_lanes[t_lane]; }
Context& ctx() { return lane().c; }
std::mutex& apiMutex() { return lane().m; }
LaneLL& laneLL() { return lane().ll; }
int pinnedScratch(int slot, size_t bytes, void** out) { return lane().pinned[slot].get(bytes, out); }
int deviceScratch(int slot, size_t bytes, void** out) { return lane().scratch[slot].get(bytes, out); }
cudaError_t waitStream(cudaStream_t st) {
  Context& c = ctx();
  if (!c.doneEvent) return cudaStreamSynchronize(st);
  cudaError_t e = cudaEventRecord(c.doneEvent, st);
  if (e != cudaSuccess) return e;
  // MRP_WAIT_SPIN_US=n polls for n microseconds before sleeping on the blocking-sync event
  // (a blocking wait costs a wake-up of 50-100 us).  Off by default: with the lanes' OpenMP
  // teams on the same cores polling made the batch times erratic (ECBS batch 0.91 s without,
  // 0.81-2.35 s with 300-1000 us; CBS 8x8 set 0.97 / 0.85 / 2.01 s at 0 / 300 / 1000 us).
  static const int spinUs = [] {
    const char* v = getenv("MRP_WAIT_SPIN_US");
    return v ? atoi(v) : 0;
  }();
  if (spinUs > 0) {
    const auto t0 = std::chrono::steady_clock::now();
    while (true) {
      e = cudaEventQuery(c.doneEvent);
      if (e != cudaErrorNotReady) return e;
      if (std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() >= spinUs)
        break;
    }
  }
  return cudaEventSynchronize(c.doneEvent);
}
#define g_scratch (lane().scratch)

static int initLocked(int device) {
  Context& c = ctx();
  if (c.ready && (device < 0 || device == c.device)) return 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(MRP_ERR_NO_DEVICE,
                "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0) device = g_device.load();  // lanes follow the first one
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  MRP_CHECK(device < n, MRP_ERR_INVALID, "device %d out of range (%d devices)",
            device, n);
  MRP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MRP_CUDA(cudaGetDeviceProperties(&prop, device));
  if (c.ready) {
    cudaStreamDestroy(c.stream);
    cudaStreamDestroy(c.copyStream);
  }
  c.device = device;
  c.smCount = prop.multiProcessorCount;
  c.smemOptin = prop.sharedMemPerBlockOptin;
  MRP_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  MRP_CUDA(cudaStreamCreateWithFlags(&c.copyStream, cudaStreamNonBlocking));
  if (!c.doneEvent)
    MRP_CUDA(cudaEventCreateWithFlags(&c.doneEvent, cudaEventBlockingSync | cudaEventDisableTiming));
  char buf[256];
  snprintf(buf, sizeof buf, "%d.%d %s sm=%d smem_optin=%zu", prop.major, prop.minor,
           prop.name, prop.multiProcessorCount, (size_t)prop.sharedMemPerBlockOptin);
  c.info = buf;
  c.ready = true;
  g_device.store(device);
  return 0;
}

int ensureInit() {
  if (ctx().ready) {
    // other libraries (torch) may have changed the current device
    cudaSetDevice(ctx().device);
    return 0;
  }
  return initLocked(-1);
}



template <class T>
static int scratch(int slot, size_t count, T** out) {
  void* p = nullptr;
  int rc = g_scratch[slot].get(count * sizeof(T), &p);
  *out = static_cast<T*>(p);
  return rc;
}

static void packMapBits(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                        std::vector<uint32_t>& bits, int* Wout, int* Sout) {
  const int W = (dimx + 31) / 32, S = (dimy + 31) / 32;
  bits.assign((size_t)W * S * 32, 0u);
  for (int y = 0; y < dimy; ++y) {
    const int s = y >> 5, r = y & 31;
    for (int tx = 0; tx < W; ++tx) {
      const int n = std::min(32, dimx - tx * 32);
      bits[((size_t)s * W + tx) * 32 + r] = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
    }
  }
  for (int k = 0; k < n_obst; ++k) {
    const int x = obst_xy[2 * k], y = obst_xy[2 * k + 1];
    // obstacles outside the map are never looked up by the reference
    if (x < 0 || y < 0 || x >= dimx || y >= dimy) continue;
    bits[((size_t)(y >> 5) * W + (x >> 5)) * 32 + (y & 31)] &= ~(1u << (x & 31));
  }
  *Wout = W;
  *Sout = S;
}

static int validateGoals(int dimx, int dimy, const int32_t* goal_xy, int n_goals) {
  for (int k = 0; k < n_goals; ++k) {
    const int x = goal_xy[2 * k], y = goal_xy[2 * k + 1];
    MRP_CHECK(x >= 0 && y >= 0 && x < dimx && y < dimy, MRP_ERR_INVALID,
              "goal %d = (%d,%d) lies outside the %dx%d map", k, x, y, dimx, dimy);
  }
  return 0;
}

}  // namespace mrp

using namespace mrp;

extern "C" {

int mrp_init(int device) {
  std::lock_guard<std::mutex> lk(apiMutex());
  return initLocked(device);
}

int mrp_shutdown(void) {
  for (Lane& L : g_lanes) {
    std::lock_guard<std::mutex> lk(L.m);
    Context& c = L.c;
    if (!c.ready) continue;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    for (auto& s : L.scratch) s.release();
    for (auto& s : L.pinned) s.release();
    if (L.ll.arena) cudaFree(L.ll.arena);
    if (L.ll.hashArena) cudaFree(L.ll.hashArena);
    L.ll = LaneLL();
    cudaStreamDestroy(c.stream);
    cudaStreamDestroy(c.copyStream);
    if (c.doneEvent) cudaEventDestroy(c.doneEvent);
    c.doneEvent = nullptr;
    c.ready = false;
  }
  g_device.store(-1);
  return 0;
}

int mrp_set_lane(int lane_index) {
  if (lane_index < 0 || lane_index >= kMaxLanes)
    return fail(MRP_ERR_INVALID, "lane %d out of range (0..%d)", lane_index, kMaxLanes - 1);
  t_lane = lane_index;
  return 0;
}

int mrp_max_lanes(void) { return kMaxLanes; }

int mrp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* mrp_last_error(void) { return lastErrorStorage().c_str(); }

const char* mrp_device_info(void) {
  std::lock_guard<std::mutex> lk(apiMutex());
  if (ensureInit() != 0) return "";
  return ctx().info.c_str();
}

long long mrp_launch_count(void) { return g_launches.load(); }

int mrp_bitmap_row_division(int dimx, int32_t* row_words, uint32_t* magic, int32_t* shift) {
  MRP_CHECK(dimx > 0 && row_words && magic && shift, MRP_ERR_INVALID, "bad arguments");
  *row_words = bitmapRowWords(dimx);
  int sh = 0;
  bitmapRowDivision(*row_words, magic, &sh);
  *shift = sh;
  return 0;
}

// ---- maps ------------------------------------------------------------------
int mrp_map_create(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                   mrp_map* out) {
  std::lock_guard<std::mutex> lk(apiMutex());
  MRP_CHECK(out != nullptr, MRP_ERR_INVALID, "out is NULL");
  *out = nullptr;
  MRP_CHECK(dimx > 0 && dimy > 0, MRP_ERR_INVALID, "bad dimensions %dx%d", dimx, dimy);
  MRP_CHECK((long long)dimx * dimy < (1ll << 30), MRP_ERR_UNSUPPORTED,
            "map of %dx%d cells is too large", dimx, dimy);
  MRP_CHECK(n_obst == 0 || obst_xy != nullptr, MRP_ERR_INVALID, "obst_xy is NULL");
  if (int rc = ensureInit()) return rc;
  std::vector<uint32_t> bits;
  int W, S;
  packMapBits(dimx, dimy, obst_xy, n_obst, bits, &W, &S);
  mrp_map_s* m = new mrp_map_s();
  m->dimx = dimx;
  m->dimy = dimy;
  m->W = W;
  m->S = S;
  m->h_bits = new uint32_t[bits.size()];
  std::memcpy(m->h_bits, bits.data(), bits.size() * 4);
  m->d_bits = nullptr;
  m->d_bits84 = nullptr;
  m->d_rowbits = nullptr;
  cudaError_t e = cudaMalloc(&m->d_bits, bits.size() * 4);
  if (e == cudaSuccess)
    e = cudaMemcpy(m->d_bits, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice);
  if (W > 1 || S > 1) {
    // maps larger than one 32x32 tile: the layouts of the large-map BFS kernels.
    // second layout: 8x4-cell tiles for the tiled BFS kernel
    const int TW8 = (dimx + 7) / 8, TH4 = (dimy + 3) / 4;
    std::vector<uint32_t> bits84((size_t)TW8 * TH4, 0u);
    // third layout: row-major with a zero border for the queue BFS kernel
    const int WPR = bitmapRowWords(dimx);
    // (+ 4 words: the queue kernel's bulk copy moves whole 16-byte granules)
    std::vector<uint32_t> rowbits((size_t)(dimy + 4) * WPR + 4, 0u);
    for (int y = 0; y < dimy; ++y)
      for (int x = 0; x < dimx; ++x)
        if ((bits[((size_t)(y >> 5) * W + (x >> 5)) * 32 + (y & 31)] >> (x & 31)) & 1u) {
          bits84[(size_t)(y >> 2) * TW8 + (x >> 3)] |= 1u << (((y & 3) << 3) | (x & 7));
          rowbits[(size_t)(y + 1) * WPR + ((x + 1) >> 5)] |= 1u << ((x + 1) & 31);
        }
    if (e == cudaSuccess) e = cudaMalloc(&m->d_bits84, bits84.size() * 4);
    if (e == cudaSuccess)
      e = cudaMemcpy(m->d_bits84, bits84.data(), bits84.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_rowbits, rowbits.size() * 4);
    if (e == cudaSuccess)
      e = cudaMemcpy(m->d_rowbits, rowbits.data(), rowbits.size() * 4, cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    cudaFree(m->d_bits);
    cudaFree(m->d_bits84);
    cudaFree(m->d_rowbits);
    delete[] m->h_bits;
    delete m;
    return fail(MRP_ERR_CUDA, "map upload failed: %s", cudaGetErrorString(e));
  }
  *out = m;
  return 0;
}

int mrp_map_destroy(mrp_map map) {
  std::lock_guard<std::mutex> lk(apiMutex());
  if (!map) return 0;
  if (ctx().ready) cudaSetDevice(ctx().device);
  cudaFree(map->d_bits);
  cudaFree(map->d_bits84);
  cudaFree(map->d_rowbits);
  delete[] map->h_bits;
  delete map;
  return 0;
}

// ---- (1) distance fields -----------------------------------------------------
size_t mrp_bfs_workspace_bytes(mrp_map map, int n_goals) {
  std::lock_guard<std::mutex> lk(apiMutex());
  if (!map || ensureInit() != 0) return 0;
  if (map->W == 1 && map->S == 1) return 16 * (size_t)std::max(n_goals, 1) + 256;
  return bfsLargeWorkspaceBytes(map, n_goals);
}

// jobs for the single-tile kernel, built on the device side of the ABI
__global__ void make_small_jobs_kernel(const int32_t* goal_cell, int n, int cells,
                                       int4* jobs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const long long off = (long long)k * cells;
  jobs[k] = make_int4(0, goal_cell[k], (int)(off & 0xffffffffll), (int)(off >> 32));
}

static int bfsFieldsDevLocked(mrp_map map, const int32_t* d_goal_cell, int n_goals,
                              int32_t* d_out, void* d_ws, cudaStream_t st) {
  if (n_goals == 0) return 0;
  if (map->W == 1 && map->S == 1) {
    // workspace: [0,256) dims, then jobs
    int32_t* d_dims = static_cast<int32_t*>(d_ws);
    int4* d_jobs = reinterpret_cast<int4*>(static_cast<char*>(d_ws) + 256);
    const int32_t dims[2] = {map->dimx, map->dimy};
    MRP_CUDA(cudaMemcpyAsync(d_dims, dims, sizeof dims, cudaMemcpyHostToDevice, st));
    make_small_jobs_kernel<<<(n_goals + 255) / 256, 256, 0, st>>>(
        d_goal_cell, n_goals, map->dimx * map->dimy, d_jobs);
    countLaunch();
    return launchBfsSmall(map->d_bits, d_dims, d_jobs, n_goals, d_out, st);
  }
  return launchBfsLarge(map, d_goal_cell, n_goals, d_out, d_ws, st);
}

int mrp_bfs_fields_dev(mrp_map map, const int32_t* d_goal_cell, int n_goals,
                       int32_t* d_out, void* d_workspace, void* stream) {
  std::lock_guard<std::mutex> lk(apiMutex());
  MRP_CHECK(map != nullptr, MRP_ERR_INVALID, "map is NULL");
  MRP_CHECK(n_goals >= 0, MRP_ERR_INVALID, "n_goals < 0");
  if (int rc = ensureInit()) return rc;
  MRP_CHECK(n_goals == 0 || (d_goal_cell && d_out && d_workspace), MRP_ERR_INVALID,
            "NULL device pointer");
  return bfsFieldsDevLocked(map, d_goal_cell, n_goals, d_out, d_workspace,
                            static_cast<cudaStream_t>(stream));
}

// ---- packed device-to-host transfer of distance fields -----------------------
// A field leaves the device as uint16 (0xFFFF = MRP_INF) whenever every finite
// distance of the batch is below 0xFFFF, and is expanded to the ABI's int32 by
// host threads (widen.cpp) while the next batch is on the bus: the host-pointer
// call is PCIe-bound (4 B per cell), this halves the bytes.  A batch that does
// not fit (a maze with paths of >= 65535 steps) is sent again as int32.
__global__ void __launch_bounds__(256)
pack_fields_u16_kernel(const int32_t* __restrict__ in, uint16_t* __restrict__ out,
                       size_t n, int* __restrict__ overflow) {
  const size_t n8 = n >> 3;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  bool ovf = false;
  auto cvt = [&](int v) -> uint32_t {
    if (v == MRP_INF) return 0xFFFFu;
    if (v >= 0xFFFF) ovf = true;
    return (uint32_t)v & 0xFFFFu;
  };
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n8; k += stride) {
    const int4 a = __ldcs(reinterpret_cast<const int4*>(in) + 2 * k);
    const int4 b = __ldcs(reinterpret_cast<const int4*>(in) + 2 * k + 1);
    uint4 o;
    o.x = cvt(a.x) | (cvt(a.y) << 16);
    o.y = cvt(a.z) | (cvt(a.w) << 16);
    o.z = cvt(b.x) | (cvt(b.y) << 16);
    o.w = cvt(b.z) | (cvt(b.w) << 16);
    reinterpret_cast<uint4*>(out)[k] = o;
  }
  for (size_t k = (n8 << 3) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
    out[k] = (uint16_t)cvt(in[k]);
  if (__syncthreads_or(ovf) && threadIdx.x == 0) atomicOr(overflow, 1);
}

// One byte per cell: h = (distance - Manhattan distance to the goal) / 2.  On a
// 4-connected grid the two have the same parity and the BFS distance is never
// the smaller one, so h is a non-negative integer: the detour, in pairs of
// steps, that the obstacles force.  255 = MRP_INF; a batch with h >= 255
// somewhere goes to the uint16 format.  VEC: dimx % 16 == 0, one thread packs
// 16 cells of one row.
extern "C++" {
template <bool VEC>
__global__ void __launch_bounds__(256)
pack_fields_u8_kernel(const int32_t* __restrict__ in, uint8_t* __restrict__ out, size_t n,
                      int dimx, int cells, const int32_t* __restrict__ goalCell,
                      int* __restrict__ overflow) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  bool ovf = false;
  auto cvt = [&](int v, int m) -> uint32_t {
    if (v == MRP_INF) return 255u;
    const int h = (v - m) >> 1;
    if (h >= 255) ovf = true;
    return (uint32_t)h & 255u;
  };
  if (VEC) {
    const size_t n16 = n >> 4;  // cells % 16 == 0
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n16; k += stride) {
      const size_t i = k << 4;
      const int f = (int)(i / (size_t)cells), cell = (int)(i - (size_t)f * cells);
      const int g = __ldg(goalCell + f);
      const int gx = g % dimx, gy = g / dimx, x0 = cell % dimx, y = cell / dimx;
      const int base = abs(y - gy);
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int4 a = __ldcs(reinterpret_cast<const int4*>(in) + 4 * k + q);
        const int x = x0 + 4 * q;
        w[q] = cvt(a.x, base + abs(x - gx)) | (cvt(a.y, base + abs(x + 1 - gx)) << 8) |
               (cvt(a.z, base + abs(x + 2 - gx)) << 16) | (cvt(a.w, base + abs(x + 3 - gx)) << 24);
      }
      reinterpret_cast<uint4*>(out)[k] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const int f = (int)(i / (size_t)cells), cell = (int)(i - (size_t)f * cells);
      const int g = __ldg(goalCell + f);
      const int m = abs(cell % dimx - g % dimx) + abs(cell / dimx - g / dimx);
      out[i] = (uint8_t)cvt(in[i], m);
    }
  }
  if (__syncthreads_or(ovf) && threadIdx.x == 0) atomicOr(overflow, 1);
}
}  // extern "C++"

extern "C++" {
// Compact variant: only the free cells of the map, in cell order.  One warp per (field, word
// of 32 cells): the byte of cell c goes to position prefix[c / 32] + popc(free bits below c).
__global__ void __launch_bounds__(256)
pack_fields_compact_kernel(const int32_t* __restrict__ in, uint8_t* __restrict__ out, int nFields,
                           int dimx, int cells, int nWords, int nFree,
                           const uint32_t* __restrict__ freeBits, const int32_t* __restrict__ prefix,
                           const int32_t* __restrict__ goalCell, int* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nWarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t total = (size_t)nFields * nWords;
  bool ovf = false;
  for (size_t k = warp0; k < total; k += nWarps) {
    const int f = (int)(k / (size_t)nWords), w = (int)(k - (size_t)f * nWords);
    const uint32_t bits = __ldg(freeBits + w);
    if (bits == 0u) continue;
    const int cell = 32 * w + lane;
    if (cell < cells && ((bits >> lane) & 1u)) {
      const int v = __ldcs(in + (size_t)f * cells + cell);
      uint32_t b = 255u;
      if (v != MRP_INF) {
        const int g = __ldg(goalCell + f);
        const int h = (v - abs(cell % dimx - g % dimx) - abs(cell / dimx - g / dimx)) >> 1;
        if (h >= 255) ovf = true;
        b = (uint32_t)h & 255u;
      }
      out[(size_t)f * nFree + __ldg(prefix + w) + __popc(bits & ((1u << lane) - 1u))] = (uint8_t)b;
    }
  }
  if (__syncthreads_or(ovf) && threadIdx.x == 0) atomicOr(overflow, 1);
}
}  // extern "C++"

extern "C++" {
namespace mrp {
// widen.cpp
void widenFieldU16(const uint16_t* src, int32_t* dst, size_t n, int threads);
void widenFieldU8(const uint8_t* src, int32_t* dst, int dimx, int dimy, const int32_t* goalCell,
                  size_t n_fields, int threads);
}
}

// device-to-host bytes of the last mrp_bfs_fields call (mrp_bfs_d2h_bytes)
static std::atomic<long long> g_bfsD2hBytes{0};

static int envInt(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// int32 fields straight into the caller's buffer: batch k+1 computes while batch
// k is copied back on the copy stream (double-buffered device output).  `todo`
// selects the batches (nullptr: all).
static int bfsFieldsHostDirect(mrp_map map, const int32_t* d_goals, int n_goals, size_t batch,
                               int32_t* out, const std::vector<char>* todo) {
  Context& c = ctx();
  const size_t cells = (size_t)map->dimx * map->dimy;
  const size_t fieldBytes = cells * 4;
  int rc = 0;
  int32_t* d_out[2] = {nullptr, nullptr};
  const bool twoBuffers = (size_t)n_goals > batch;
  const size_t wsBytes = (map->W == 1 && map->S == 1) ? 16 * batch + 256
                                                      : bfsLargeWorkspaceBytes(map, (int)batch);
  cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
  do {
    if ((rc = scratch(1, batch * cells, &d_out[0]))) break;
    if (twoBuffers && (rc = scratch(2, batch * cells, &d_out[1]))) break;
    char* wsp = nullptr;
    if ((rc = scratch(3, wsBytes, &wsp))) break;
    for (int b = 0; b < 2; ++b) {
      cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming);
    }
    int b = 0;
    size_t k = 0;
    for (size_t g0 = 0; g0 < (size_t)n_goals && rc == 0; g0 += batch, ++k) {
      if (todo && !(*todo)[k]) continue;
      const size_t n = std::min(batch, (size_t)n_goals - g0);
      int32_t* dst = twoBuffers ? d_out[b] : d_out[0];
      // the buffer must have been drained by its previous D2H copy
      cudaStreamWaitEvent(c.stream, copied[b], 0);
      rc = bfsFieldsDevLocked(map, d_goals + g0, (int)n, dst, wsp, c.stream);
      if (rc) break;
      cudaEventRecord(done[b], c.stream);
      cudaStreamWaitEvent(c.copyStream, done[b], 0);
      cudaError_t e = cudaMemcpyAsync(out + g0 * cells, dst, n * fieldBytes,
                                      cudaMemcpyDeviceToHost, c.copyStream);
      if (e != cudaSuccess) {
        rc = fail(MRP_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
        break;
      }
      g_bfsD2hBytes += (long long)(n * fieldBytes);
      cudaEventRecord(copied[b], c.copyStream);
      b ^= 1;
    }
    cudaError_t e1 = cudaStreamSynchronize(c.stream);
    cudaError_t e2 = cudaStreamSynchronize(c.copyStream);
    if (rc == 0 && (e1 != cudaSuccess || e2 != cudaSuccess))
      rc = fail(MRP_ERR_CUDA, "bfs fields failed: %s",
                cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  } while (0);
  for (int b = 0; b < 2; ++b) {
    if (done[b]) cudaEventDestroy(done[b]);
    if (copied[b]) cudaEventDestroy(copied[b]);
  }
  return rc;
}

// Packed fields (bpc = 1: detour bytes, bpc = 2: uint16) through page-locked
// staging slots, expanded by host threads.  Three slots: one being filled by
// the copy engine, one being expanded, one in between.  `order` lists the
// batches to send; overflowed[k] = 1 for those that need a wider format.
static int bfsFieldsHostPacked(mrp_map map, const int32_t* d_goals, const int32_t* h_goals,
                               int n_goals, size_t batch, int bpc, int32_t* out,
                               const std::vector<size_t>& order, std::vector<char>& overflowed) {
  constexpr int NS = 3;
  Context& c = ctx();
  const size_t cells = (size_t)map->dimx * map->dimy;
  const size_t nBatches = ((size_t)n_goals + batch - 1) / batch;
  const size_t nOrder = order.size();
  overflowed.assign(nBatches, 0);
  const int threads = std::max(1, envInt("MRP_WIDEN_THREADS",
                                         std::min(16, (int)std::thread::hardware_concurrency())));
  const size_t wsBytes = (map->W == 1 && map->S == 1) ? 16 * batch + 256
                                                      : bfsLargeWorkspaceBytes(map, (int)batch);
  // staging slot: batch*cells packed values, then the overflow flag of the batch
  const size_t slotBytes = ((batch * cells * (size_t)bpc + 255) & ~(size_t)255) + 256;
  int rc = 0;
  int32_t* d_out = nullptr;
  char* wsp = nullptr;
  char* d_pack = nullptr;
  void* h_pack = nullptr;
  cudaEvent_t done[NS] = {}, copied[NS] = {};
  if ((rc = scratch(1, batch * cells, &d_out))) return rc;
  if ((rc = scratch(3, wsBytes, &wsp))) return rc;
  if ((rc = scratch(4, NS * slotBytes, &d_pack))) return rc;
  if ((rc = pinnedScratch(3, NS * slotBytes, &h_pack))) return rc;
  for (int s = 0; s < NS; ++s) {
    cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&copied[s], cudaEventBlockingSync | cudaEventDisableTiming);
  }
  std::mutex m;
  std::condition_variable cv;
  size_t enqueued = 0, widened = 0;
  bool stop = false;
  cudaError_t workerErr = cudaSuccess;
  const int device = c.device;
  std::thread worker([&] {
    cudaSetDevice(device);
    for (size_t j = 0;; ++j) {
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return enqueued > j || stop; });
        if (enqueued <= j) return;
      }
      const int s = (int)(j % NS);
      cudaError_t e = cudaEventSynchronize(copied[s]);
      if (e != cudaSuccess) workerErr = e;
      const char* slot = static_cast<const char*>(h_pack) + (size_t)s * slotBytes;
      const size_t k = order[j];
      const size_t g0 = k * batch, n = std::min(batch, (size_t)n_goals - g0);
      if (e == cudaSuccess) {
        if (*reinterpret_cast<const int*>(slot + slotBytes - 256))
          overflowed[k] = 1;
        else if (bpc == 1)
          widenFieldU8(reinterpret_cast<const uint8_t*>(slot), out + g0 * cells, map->dimx,
                       map->dimy, h_goals + g0, n, threads);
        else
          widenFieldU16(reinterpret_cast<const uint16_t*>(slot), out + g0 * cells, n * cells,
                        threads);
      }
      {
        std::lock_guard<std::mutex> lk(m);
        widened = j + 1;
      }
      cv.notify_all();
    }
  });
  for (size_t j = 0; j < nOrder && rc == 0; ++j) {
    const int s = (int)(j % NS);
    const size_t k = order[j];
    const size_t g0 = k * batch, n = std::min(batch, (size_t)n_goals - g0);
    if (j >= NS) {  // the slot's previous batch must have been expanded
      std::unique_lock<std::mutex> lk(m);
      cv.wait(lk, [&] { return widened + NS > j; });
    }
    rc = bfsFieldsDevLocked(map, d_goals + g0, (int)n, d_out, wsp, c.stream);
    if (rc) break;
    char* dslot = d_pack + (size_t)s * slotBytes;
    int* dflag = reinterpret_cast<int*>(dslot + slotBytes - 256);
    cudaMemsetAsync(dflag, 0, sizeof(int), c.stream);
    const size_t total = n * cells;
    const int grid = (int)std::min<size_t>((total / 8 + 255) / 256 + 1, (size_t)c.smCount * 16);
    if (bpc == 2)
      pack_fields_u16_kernel<<<grid, 256, 0, c.stream>>>(
          d_out, reinterpret_cast<uint16_t*>(dslot), total, dflag);
    else if (map->dimx % 16 == 0)
      pack_fields_u8_kernel<true><<<grid, 256, 0, c.stream>>>(
          d_out, reinterpret_cast<uint8_t*>(dslot), total, map->dimx, (int)cells, d_goals + g0,
          dflag);
    else
      pack_fields_u8_kernel<false><<<grid, 256, 0, c.stream>>>(
          d_out, reinterpret_cast<uint8_t*>(dslot), total, map->dimx, (int)cells, d_goals + g0,
          dflag);
    countLaunch();
    cudaEventRecord(done[s], c.stream);
    cudaStreamWaitEvent(c.copyStream, done[s], 0);
    char* hslot = static_cast<char*>(h_pack) + (size_t)s * slotBytes;
    cudaError_t e = cudaMemcpyAsync(hslot, dslot, total * (size_t)bpc, cudaMemcpyDeviceToHost,
                                    c.copyStream);
    g_bfsD2hBytes += (long long)(total * (size_t)bpc);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(hslot + slotBytes - 256, dflag, sizeof(int), cudaMemcpyDeviceToHost,
                          c.copyStream);
    if (e != cudaSuccess) {
      rc = fail(MRP_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
      break;
    }
    cudaEventRecord(copied[s], c.copyStream);
    {
      std::lock_guard<std::mutex> lk(m);
      enqueued = j + 1;
    }
    cv.notify_all();
  }
  {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return widened == enqueued; });
    stop = true;
  }
  cv.notify_all();
  worker.join();
  cudaError_t e1 = cudaStreamSynchronize(c.stream);
  cudaError_t e2 = cudaStreamSynchronize(c.copyStream);
  if (rc == 0 && workerErr != cudaSuccess) e1 = workerErr;
  if (rc == 0 && (e1 != cudaSuccess || e2 != cudaSuccess))
    rc = fail(MRP_ERR_CUDA, "bfs fields failed: %s",
              cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  for (int s = 0; s < NS; ++s) {
    cudaEventDestroy(done[s]);
    cudaEventDestroy(copied[s]);
  }
  return rc;
}

int mrp_bfs_fields(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                   const int32_t* goal_xy, int n_goals, int32_t* out) {
  MRP_CHECK(n_goals >= 0, MRP_ERR_INVALID, "n_goals < 0");
  MRP_CHECK(n_goals == 0 || (goal_xy && out), MRP_ERR_INVALID, "NULL pointer");
  if (int rc = validateGoals(dimx, dimy, goal_xy, n_goals)) return rc;
  mrp_map map = nullptr;
  if (int rc = mrp_map_create(dimx, dimy, obst_xy, n_obst, &map)) return rc;
  int rc = 0;
  {
    std::lock_guard<std::mutex> lk(apiMutex());
    Context& c = ctx();
    const size_t cells = (size_t)dimx * dimy;
    std::vector<int32_t> goalCell(n_goals);
    for (int k = 0; k < n_goals; ++k)
      goalCell[k] = goal_xy[2 * k] + dimx * goal_xy[2 * k + 1];
    int32_t* d_goals = nullptr;
    g_bfsD2hBytes = 0;
    do {
      if ((rc = scratch(0, (size_t)std::max(n_goals, 1), &d_goals))) break;
      cudaError_t e = cudaMemcpyAsync(d_goals, goalCell.data(), (size_t)n_goals * 4,
                                      cudaMemcpyHostToDevice, c.stream);
      if (e != cudaSuccess) {
        rc = fail(MRP_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
        break;
      }
      // Large results travel packed (MRP_BFS_PACK=0 switches it off, =1 forces it).
      const int packMode = envInt("MRP_BFS_PACK", -1);
      const bool packed = packMode == 1 ||
                          (packMode != 0 && (size_t)n_goals * cells >= ((size_t)16 << 20));
      if (packed) {
        // one wave of goals per batch for large maps, ~128 MiB of uint16 otherwise
        size_t batch = std::max<size_t>((size_t)c.smCount, ((size_t)1 << 26) / cells);
        if (envInt("MRP_BFS_BATCH", 0) > 0) batch = (size_t)envInt("MRP_BFS_BATCH", 0);  // tests
        batch = std::min<size_t>(batch, (size_t)n_goals);
        // narrowest format first; the batches that do not fit move on to the next one
        const size_t nBatches = ((size_t)n_goals + batch - 1) / batch;
        std::vector<size_t> order(nBatches);
        for (size_t k = 0; k < nBatches; ++k) order[k] = k;
        std::vector<char> overflowed;
        for (int bpc = envInt("MRP_BFS_FMT", 8) == 16 ? 2 : 1; bpc <= 2 && rc == 0 && !order.empty();
             ++bpc) {
          rc = bfsFieldsHostPacked(map, d_goals, goalCell.data(), n_goals, batch, bpc, out, order,
                                   overflowed);
          order.clear();
          for (size_t k = 0; k < nBatches; ++k)
            if (overflowed[k]) order.push_back(k);
        }
        if (rc == 0 && !order.empty())
          rc = bfsFieldsHostDirect(map, d_goals, n_goals, batch, out, &overflowed);
      } else {
        size_t batch = std::max<size_t>(4 * (size_t)c.smCount, ((size_t)1 << 28) / cells);
        batch = std::min<size_t>(batch, (size_t)std::max(n_goals, 1));
        rc = bfsFieldsHostDirect(map, d_goals, n_goals, batch, out, nullptr);
      }
    } while (0);
  }
  mrp_map_destroy(map);
  return rc;
}

// The packed result mode: one detour byte per cell, copied straight into the
// caller's buffer (no host expansion); see include/mrp_b200.h.
int mrp_free_cell_index(int dimx, int dimy, const int32_t* obst_xy, int n_obst, uint32_t* free_bits,
                        int32_t* prefix) {
  MRP_CHECK(dimx > 0 && dimy > 0 && n_obst >= 0 && (n_obst == 0 || obst_xy) && free_bits && prefix,
            MRP_ERR_INVALID, "bad arguments");
  MRP_CHECK((long long)dimx * dimy < (1ll << 31), MRP_ERR_UNSUPPORTED, "map too large");
  const int cells = dimx * dimy, nWords = (cells + 31) / 32;
  for (int w = 0; w < nWords; ++w) {
    const int n = std::min(32, cells - 32 * w);
    free_bits[w] = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
  }
  for (int k = 0; k < n_obst; ++k) {
    const int x = obst_xy[2 * k], y = obst_xy[2 * k + 1];
    if (x < 0 || y < 0 || x >= dimx || y >= dimy) continue;
    const int c = x + dimx * y;
    free_bits[c >> 5] &= ~(1u << (c & 31));
  }
  int run = 0;
  for (int w = 0; w < nWords; ++w) {
    prefix[w] = run;
    run += __builtin_popcount(free_bits[w]);
  }
  prefix[nWords] = run;
  return run;
}

static int packedFieldsImpl(int dimx, int dimy, const int32_t* obst_xy, int n_obst, const int32_t* goal_xy,
                            int n_goals, uint8_t* out, int32_t* overflowed, bool compact);

int mrp_bfs_fields_packed(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                          const int32_t* goal_xy, int n_goals, uint8_t* out, int32_t* overflowed) {
  return packedFieldsImpl(dimx, dimy, obst_xy, n_obst, goal_xy, n_goals, out, overflowed, false);
}

int mrp_bfs_fields_compact(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                           const int32_t* goal_xy, int n_goals, uint8_t* out, int32_t* overflowed) {
  return packedFieldsImpl(dimx, dimy, obst_xy, n_obst, goal_xy, n_goals, out, overflowed, true);
}

static int packedFieldsImpl(int dimx, int dimy, const int32_t* obst_xy, int n_obst, const int32_t* goal_xy,
                            int n_goals, uint8_t* out, int32_t* overflowed, bool compact) {
  MRP_CHECK(n_goals >= 0, MRP_ERR_INVALID, "n_goals < 0");
  MRP_CHECK(n_goals == 0 || (goal_xy && out), MRP_ERR_INVALID, "NULL pointer");
  if (int rc = validateGoals(dimx, dimy, goal_xy, n_goals)) return rc;
  if (overflowed) std::memset(overflowed, 0, sizeof(int32_t) * (size_t)n_goals);
  if (n_goals == 0) return 0;
  mrp_map map = nullptr;
  if (int rc = mrp_map_create(dimx, dimy, obst_xy, n_obst, &map)) return rc;
  // compact result: the free cells only (prefix sums of the free mask give their positions)
  const int nWords = (dimx * dimy + 31) / 32;
  std::vector<uint32_t> freeBits;
  std::vector<int32_t> prefix;
  size_t perField = (size_t)dimx * dimy;  // bytes of one field in the caller's array
  if (compact) {
    freeBits.resize(nWords);
    prefix.resize(nWords + 1);
    const int nFree = mrp_free_cell_index(dimx, dimy, obst_xy, n_obst, freeBits.data(), prefix.data());
    if (nFree < 0) {
      mrp_map_destroy(map);
      return nFree;
    }
    perField = (size_t)nFree;
  }
  int rc = 0, nOverflowed = 0;
  {
    std::lock_guard<std::mutex> lk(apiMutex());
    Context& c = ctx();
    const size_t cells = (size_t)dimx * dimy;
    std::vector<int32_t> goalCell(n_goals);
    for (int k = 0; k < n_goals; ++k) goalCell[k] = goal_xy[2 * k] + dimx * goal_xy[2 * k + 1];
    size_t batch = std::max<size_t>((size_t)c.smCount, ((size_t)1 << 26) / cells);
    if (envInt("MRP_BFS_BATCH", 0) > 0) batch = (size_t)envInt("MRP_BFS_BATCH", 0);  // tests
    batch = std::min<size_t>(batch, (size_t)n_goals);
    const size_t nBatches = ((size_t)n_goals + batch - 1) / batch;
    const size_t wsBytes = (map->W == 1 && map->S == 1) ? 16 * batch + 256
                                                        : bfsLargeWorkspaceBytes(map, (int)batch);
    const size_t slotBytes = ((batch * cells + 255) & ~(size_t)255) + 256;
    int32_t *d_goals = nullptr, *d_out = nullptr, *d_prefix = nullptr;
    uint32_t* d_freeBits = nullptr;
    char *wsp = nullptr, *d_pack = nullptr;
    void* h_flags = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    g_bfsD2hBytes = 0;
    do {
      if ((rc = scratch(0, (size_t)n_goals, &d_goals))) break;
      if ((rc = scratch(1, batch * cells, &d_out))) break;
      if ((rc = scratch(3, wsBytes, &wsp))) break;
      if ((rc = scratch(4, 2 * slotBytes, &d_pack))) break;
      if ((rc = pinnedScratch(4, nBatches * sizeof(int), &h_flags))) break;
      if (compact) {
        if ((rc = scratch(13, (size_t)nWords, &d_freeBits))) break;
        if ((rc = scratch(14, (size_t)nWords + 1, &d_prefix))) break;
        cudaMemcpyAsync(d_freeBits, freeBits.data(), (size_t)nWords * 4, cudaMemcpyHostToDevice, c.stream);
        cudaMemcpyAsync(d_prefix, prefix.data(), ((size_t)nWords + 1) * 4, cudaMemcpyHostToDevice, c.stream);
      }
      cudaError_t e = cudaMemcpyAsync(d_goals, goalCell.data(), (size_t)n_goals * 4,
                                      cudaMemcpyHostToDevice, c.stream);
      if (e != cudaSuccess) {
        rc = fail(MRP_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
        break;
      }
      for (int b = 0; b < 2; ++b) {
        cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming);
      }
      for (size_t j = 0; j < nBatches && rc == 0; ++j) {
        const int sl = (int)(j & 1);
        const size_t g0 = j * batch, n = std::min(batch, (size_t)n_goals - g0);
        rc = bfsFieldsDevLocked(map, d_goals + g0, (int)n, d_out, wsp, c.stream);
        if (rc) break;
        char* dslot = d_pack + (size_t)sl * slotBytes;
        int* dflag = reinterpret_cast<int*>(dslot + slotBytes - 256);
        cudaStreamWaitEvent(c.stream, copied[sl], 0);  // the slot's previous batch has left
        cudaMemsetAsync(dflag, 0, sizeof(int), c.stream);
        const size_t total = n * cells;
        const int grid = (int)std::min<size_t>((total / 8 + 255) / 256 + 1, (size_t)c.smCount * 16);
        if (compact)
          pack_fields_compact_kernel<<<grid, 256, 0, c.stream>>>(
              d_out, reinterpret_cast<uint8_t*>(dslot), (int)n, map->dimx, (int)cells, nWords, (int)perField,
              d_freeBits, d_prefix, d_goals + g0, dflag);
        else if (map->dimx % 16 == 0)
          pack_fields_u8_kernel<true><<<grid, 256, 0, c.stream>>>(
              d_out, reinterpret_cast<uint8_t*>(dslot), total, map->dimx, (int)cells, d_goals + g0, dflag);
        else
          pack_fields_u8_kernel<false><<<grid, 256, 0, c.stream>>>(
              d_out, reinterpret_cast<uint8_t*>(dslot), total, map->dimx, (int)cells, d_goals + g0, dflag);
        countLaunch();
        cudaEventRecord(done[sl], c.stream);
        cudaStreamWaitEvent(c.copyStream, done[sl], 0);
        const size_t wire = n * perField;
        e = cudaMemcpyAsync(out + g0 * perField, dslot, wire, cudaMemcpyDeviceToHost, c.copyStream);
        if (e == cudaSuccess)
          e = cudaMemcpyAsync(static_cast<int*>(h_flags) + j, dflag, sizeof(int), cudaMemcpyDeviceToHost,
                              c.copyStream);
        if (e != cudaSuccess) {
          rc = fail(MRP_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
          break;
        }
        g_bfsD2hBytes += (long long)wire;
        cudaEventRecord(copied[sl], c.copyStream);
      }
      cudaError_t e1 = cudaStreamSynchronize(c.stream);
      cudaError_t e2 = cudaStreamSynchronize(c.copyStream);
      if (rc == 0 && (e1 != cudaSuccess || e2 != cudaSuccess))
        rc = fail(MRP_ERR_CUDA, "bfs fields failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      if (rc == 0)
        for (size_t j = 0; j < nBatches; ++j)
          if (static_cast<const int*>(h_flags)[j]) {
            const size_t g0 = j * batch, n = std::min(batch, (size_t)n_goals - g0);
            nOverflowed += (int)n;
            if (overflowed)
              for (size_t k = 0; k < n; ++k) overflowed[g0 + k] = 1;
          }
    } while (0);
    for (int b = 0; b < 2; ++b) {
      if (done[b]) cudaEventDestroy(done[b]);
      if (copied[b]) cudaEventDestroy(copied[b]);
    }
  }
  mrp_map_destroy(map);
  return rc ? rc : nOverflowed;
}

long long mrp_bfs_d2h_bytes(void) { return g_bfsD2hBytes.load(); }

int mrp_widen_u8(const uint8_t* src, int32_t* dst, int dimx, int dimy, const int32_t* goal_cell,
                 int n_fields, int threads) {
  MRP_CHECK(dimx > 0 && dimy > 0 && n_fields >= 0, MRP_ERR_INVALID, "bad dimensions");
  MRP_CHECK(n_fields == 0 || (src && dst && goal_cell), MRP_ERR_INVALID, "NULL pointer");
  for (int k = 0; k < n_fields; ++k)
    MRP_CHECK(goal_cell[k] >= 0 && (size_t)goal_cell[k] < (size_t)dimx * dimy, MRP_ERR_INVALID,
              "goal cell outside the map");
  widenFieldU8(src, dst, dimx, dimy, goal_cell, (size_t)n_fields, std::max(1, threads));
  return 0;
}

int mrp_widen_u16(const uint16_t* src, int32_t* dst, size_t n, int threads) {
  MRP_CHECK(n == 0 || (src && dst), MRP_ERR_INVALID, "NULL pointer");
  widenFieldU16(src, dst, n, std::max(1, threads));
  return 0;
}

int mrp_bfs_fields_batch(int n_maps, const int32_t* dims, const int32_t* obst_off,
                         const int32_t* obst_xy, const int32_t* goal_off,
                         const int32_t* goal_xy, int32_t* out) {
  MRP_CHECK(n_maps >= 0, MRP_ERR_INVALID, "n_maps < 0");
  if (n_maps == 0) return 0;
  MRP_CHECK(dims && obst_off && goal_off, MRP_ERR_INVALID, "NULL pointer");
  bool allSmall = true;
  for (int m = 0; m < n_maps; ++m) {
    MRP_CHECK(dims[2 * m] > 0 && dims[2 * m + 1] > 0, MRP_ERR_INVALID,
              "map %d has bad dimensions", m);
    if (dims[2 * m] > 32 || dims[2 * m + 1] > 32) allSmall = false;
    if (int rc = validateGoals(dims[2 * m], dims[2 * m + 1], goal_xy + 2 * goal_off[m],
                               goal_off[m + 1] - goal_off[m]))
      return rc;
  }
  if (!allSmall) {
    // large maps: one map at a time through the tiled kernel
    size_t off = 0;
    for (int m = 0; m < n_maps; ++m) {
      const int ng = goal_off[m + 1] - goal_off[m];
      int rc = mrp_bfs_fields(dims[2 * m], dims[2 * m + 1], obst_xy + 2 * obst_off[m],
                              obst_off[m + 1] - obst_off[m], goal_xy + 2 * goal_off[m],
                              ng, out + off);
      if (rc) return rc;
      off += (size_t)ng * dims[2 * m] * dims[2 * m + 1];
    }
    return 0;
  }
  std::lock_guard<std::mutex> lk(apiMutex());
  if (int rc = ensureInit()) return rc;
  Context& c = ctx();
  const int n_goals = goal_off[n_maps];
  if (n_goals == 0) return 0;
  std::vector<uint32_t> rows((size_t)n_maps * 32);
  std::vector<int4> jobs(n_goals);
  long long off = 0;
  for (int m = 0; m < n_maps; ++m) {
    std::vector<uint32_t> bits;
    int W, S;
    packMapBits(dims[2 * m], dims[2 * m + 1], obst_xy + 2 * obst_off[m],
                obst_off[m + 1] - obst_off[m], bits, &W, &S);
    std::memcpy(&rows[(size_t)m * 32], bits.data(), 32 * 4);
    const int cells = dims[2 * m] * dims[2 * m + 1];
    for (int k = goal_off[m]; k < goal_off[m + 1]; ++k) {
      jobs[k] = make_int4(m, goal_xy[2 * k] + dims[2 * m] * goal_xy[2 * k + 1],
                          (int)(off & 0xffffffffll), (int)(off >> 32));
      off += cells;
    }
  }
  uint32_t* d_rows;
  int32_t* d_dims;
  int4* d_jobs;
  int32_t* d_out;
  if (int rc = scratch(4, rows.size(), &d_rows)) return rc;
  if (int rc = scratch(5, (size_t)2 * n_maps, &d_dims)) return rc;
  if (int rc = scratch(6, jobs.size(), &d_jobs)) return rc;
  if (int rc = scratch(1, (size_t)off, &d_out)) return rc;
  MRP_CUDA(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, c.stream));
  MRP_CUDA(cudaMemcpyAsync(d_dims, dims, (size_t)2 * n_maps * 4, cudaMemcpyHostToDevice, c.stream));
  MRP_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(int4), cudaMemcpyHostToDevice, c.stream));
  if (int rc = launchBfsSmall(d_rows, d_dims, d_jobs, n_goals, d_out, c.stream)) return rc;
  MRP_CUDA(cudaMemcpyAsync(out, d_out, (size_t)off * 4, cudaMemcpyDeviceToHost, c.stream));
  MRP_CUDA(cudaStreamSynchronize(c.stream));
  return 0;
}

// ---- field sets ---------------------------------------------------------------
__global__ void gather_rows_kernel(const uint32_t* const* mapBits, int n_maps,
                                   uint32_t* rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_maps * 32) rows[i] = mapBits[i >> 5][i & 31];
}
__global__ void make_fs_jobs_kernel(const int32_t* goal_map, const int32_t* goal_cell, int n,
                                    int cells, int4* jobs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const long long off = (long long)k * cells;
  jobs[k] = make_int4(goal_map[k], goal_cell[k], (int)(off & 0xffffffffll), (int)(off >> 32));
}

int mrp_fieldset_create(const mrp_map* maps, int n_maps, const int32_t* goal_map,
                        const int32_t* goal_cell, int n_goals, mrp_fieldset* out) {
  std::lock_guard<std::mutex> lk(apiMutex());
  MRP_CHECK(out != nullptr, MRP_ERR_INVALID, "out is NULL");
  *out = nullptr;
  MRP_CHECK(maps && n_maps > 0 && n_goals >= 0, MRP_ERR_INVALID, "bad arguments");
  MRP_CHECK(n_goals == 0 || (goal_map && goal_cell), MRP_ERR_INVALID, "NULL goal arrays");
  if (int rc = ensureInit()) return rc;
  Context& c = ctx();
  const int dimx = maps[0]->dimx, dimy = maps[0]->dimy, cells = dimx * dimy;
  for (int m = 0; m < n_maps; ++m)
    MRP_CHECK(maps[m] && maps[m]->dimx == dimx && maps[m]->dimy == dimy, MRP_ERR_INVALID,
              "all maps of a field set must share their dimensions");
  for (int k = 0; k < n_goals; ++k) {
    MRP_CHECK(goal_map[k] >= 0 && goal_map[k] < n_maps, MRP_ERR_INVALID, "goal %d: bad map", k);
    MRP_CHECK(goal_cell[k] >= 0 && goal_cell[k] < cells, MRP_ERR_INVALID,
              "goal %d: cell outside the map", k);
  }
  mrp_fieldset_s* fs = new mrp_fieldset_s();
  fs->dimx = dimx;
  fs->dimy = dimy;
  fs->n_fields = n_goals;
  fs->d_fields = nullptr;
  if (n_goals == 0) {
    *out = fs;
    return 0;
  }
  cudaError_t e = cudaMalloc(&fs->d_fields, (size_t)n_goals * cells * 4);
  if (e != cudaSuccess) {
    delete fs;
    return fail(MRP_ERR_NOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e));
  }
  int rc = 0;
  int32_t *d_gm = nullptr, *d_gc = nullptr;
  do {
    if ((rc = scratch(5, (size_t)n_goals, &d_gm))) break;
    if ((rc = scratch(0, (size_t)n_goals, &d_gc))) break;
    cudaMemcpyAsync(d_gm, goal_map, (size_t)n_goals * 4, cudaMemcpyHostToDevice, c.stream);
    cudaMemcpyAsync(d_gc, goal_cell, (size_t)n_goals * 4, cudaMemcpyHostToDevice, c.stream);
    if (maps[0]->W == 1 && maps[0]->S == 1) {
      // single-tile maps: one launch over all (map, goal) jobs
      const uint32_t** d_ptrs = nullptr;
      uint32_t* d_rows = nullptr;
      int32_t* d_dims = nullptr;
      int4* d_jobs = nullptr;
      if ((rc = scratch(3, (size_t)n_maps, &d_ptrs))) break;
      if ((rc = scratch(4, (size_t)n_maps * 32, &d_rows))) break;
      if ((rc = scratch(2, (size_t)n_maps * 2, &d_dims))) break;
      if ((rc = scratch(6, (size_t)n_goals, &d_jobs))) break;
      std::vector<const uint32_t*> ptrs(n_maps);
      std::vector<int32_t> dims(2 * (size_t)n_maps);
      for (int m = 0; m < n_maps; ++m) {
        ptrs[m] = maps[m]->d_bits;
        dims[2 * m] = dimx;
        dims[2 * m + 1] = dimy;
      }
      cudaMemcpyAsync(d_ptrs, ptrs.data(), sizeof(void*) * n_maps, cudaMemcpyHostToDevice, c.stream);
      cudaMemcpyAsync(d_dims, dims.data(), dims.size() * 4, cudaMemcpyHostToDevice, c.stream);
      gather_rows_kernel<<<(n_maps * 32 + 255) / 256, 256, 0, c.stream>>>(d_ptrs, n_maps, d_rows);
      make_fs_jobs_kernel<<<(n_goals + 255) / 256, 256, 0, c.stream>>>(d_gm, d_gc, n_goals, cells, d_jobs);
      countLaunch(2);
      rc = launchBfsSmall(d_rows, d_dims, d_jobs, n_goals, fs->d_fields, c.stream);
      if (rc) break;
      cudaStreamSynchronize(c.stream);  // host staging vectors go out of scope
    } else {
      // tiled kernel: goals grouped by map (consecutive runs)
      char* ws = nullptr;
      if ((rc = scratch(3, bfsLargeWorkspaceBytes(maps[0], n_goals), &ws))) break;
      int k = 0;
      while (k < n_goals && rc == 0) {
        int k1 = k;
        while (k1 < n_goals && goal_map[k1] == goal_map[k]) ++k1;
        rc = launchBfsLarge(maps[goal_map[k]], d_gc + k, k1 - k,
                            fs->d_fields + (size_t)k * cells, ws, c.stream);
        k = k1;
      }
      if (rc) break;
    }
    e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) rc = fail(MRP_ERR_CUDA, "field set failed: %s", cudaGetErrorString(e));
  } while (0);
  if (rc) {
    cudaFree(fs->d_fields);
    delete fs;
    return rc;
  }
  *out = fs;
  return 0;
}

int mrp_fieldset_read(mrp_fieldset fs, int first, int count, int32_t* out) {
  std::lock_guard<std::mutex> lk(apiMutex());
  MRP_CHECK(fs && out && first >= 0 && count >= 0 && first + count <= fs->n_fields,
            MRP_ERR_INVALID, "bad field range");
  if (count == 0) return 0;
  if (int rc = ensureInit()) return rc;
  const size_t cells = (size_t)fs->dimx * fs->dimy;
  MRP_CUDA(cudaMemcpy(out, fs->d_fields + (size_t)first * cells, (size_t)count * cells * 4,
                      cudaMemcpyDeviceToHost));
  return 0;
}

int mrp_fieldset_destroy(mrp_fieldset fs) {
  std::lock_guard<std::mutex> lk(apiMutex());
  if (!fs) return 0;
  if (ctx().ready) cudaSetDevice(ctx().device);
  cudaFree(fs->d_fields);
  delete fs;
  return 0;
}

// ---- (2) conflicts -------------------------------------------------------------
int mrp_decode_conflict(unsigned long long key, int dimx, int32_t cell_i_t,
                        int32_t cell_i_t1, mrp_conflict* out) {
  if (key == kNoConflict) return 0;
  out->time = (int32_t)(key >> 41);
  out->type = (int32_t)((key >> 40) & 1);
  out->agent1 = (int32_t)((key >> 20) & 0xfffff);
  out->agent2 = (int32_t)(key & 0xfffff);
  out->x1 = cell_i_t % dimx;
  out->y1 = cell_i_t / dimx;
  if (out->type == 1) {
    out->x2 = cell_i_t1 % dimx;
    out->y2 = cell_i_t1 / dimx;
  } else {
    out->x2 = -1;
    out->y2 = -1;
  }
  return 1;
}

static int checkTable(const int32_t* cell, const int32_t* len, int N, int Tpad) {
  MRP_CHECK(N >= 0 && Tpad >= 0, MRP_ERR_INVALID, "negative table shape");
  MRP_CHECK(N < kMaxAgents, MRP_ERR_UNSUPPORTED, "N=%d exceeds %d agents", N, kMaxAgents);
  MRP_CHECK(Tpad < kMaxTime, MRP_ERR_UNSUPPORTED, "Tpad=%d exceeds %d", Tpad, kMaxTime);
  MRP_CHECK(N == 0 || (cell && len), MRP_ERR_INVALID, "NULL table pointer");
  for (int i = 0; i < N; ++i)
    MRP_CHECK(len[i] >= 0 && len[i] <= Tpad, MRP_ERR_INVALID,
              "len[%d]=%d outside [0,%d]", i, len[i], Tpad);
  return 0;
}

static inline int32_t hostPos(const int32_t* cell, const int32_t* len, int Tpad, int i,
                              int t) {
  const int L = len[i];
  return cell[(size_t)i * Tpad + (t < L ? t : L - 1)];
}

static int conflictsHost(const int32_t* cell, const int32_t* len, int B, int N, int Tpad,
                         int dimx, int mode, bool wantFirst, bool wantCount,
                         int32_t* found, mrp_conflict* conflicts, int32_t* counts) {
  for (int b = 0; b < B; ++b)
    if (int rc = checkTable(cell + (size_t)b * N * Tpad, len + (size_t)b * N, N, Tpad))
      return rc;
  MRP_CHECK(mode == 0 || mode == 1, MRP_ERR_INVALID, "mode must be 0 or 1");
  for (int b = 0; b < B; ++b) {
    if (found) found[b] = 0;
    if (counts) counts[b] = 0;
  }
  if (N < 2 || Tpad == 0 || B == 0) return 0;
  std::lock_guard<std::mutex> lk(apiMutex());
  if (int rc = ensureInit()) return rc;
  Context& c = ctx();
  int32_t *d_cell, *d_len;
  unsigned long long* d_res;
  const size_t nCell = (size_t)B * N * Tpad;
  if (int rc = scratch(7, nCell, &d_cell)) return rc;
  if (int rc = scratch(8, (size_t)B * N, &d_len)) return rc;
  if (int rc = scratch(9, (size_t)4 * B, &d_res)) return rc;
  MRP_CUDA(cudaMemcpyAsync(d_cell, cell, nCell * 4, cudaMemcpyHostToDevice, c.stream));
  MRP_CUDA(cudaMemcpyAsync(d_len, len, (size_t)B * N * 4, cudaMemcpyHostToDevice, c.stream));
  char* d_ws = nullptr;
  const size_t wsBytes = B == 1 ? conflictsWorkspaceBytes(N, Tpad) : 0;
  if (wsBytes)
    if (int rc2 = scratch(12, wsBytes, &d_ws)) return rc2;
  int rc = (B == 1) ? launchConflicts(d_cell, d_len, N, Tpad, mode, wantFirst, wantCount,
                                      d_res, d_ws, wsBytes, c.stream)
                    : launchConflictsBatch(d_cell, d_len, B, N, Tpad, mode, d_res, c.stream);
  if (rc) return rc;
  void* hout = nullptr;
  if (int rc2 = pinnedScratch(1, (size_t)4 * B * 8, &hout)) return rc2;
  const unsigned long long* res = static_cast<const unsigned long long*>(hout);
  MRP_CUDA(cudaMemcpyAsync(hout, d_res, (size_t)4 * B * 8, cudaMemcpyDeviceToHost, c.stream));
  MRP_CUDA(waitStream(c.stream));
  for (int b = 0; b < B; ++b) {
    const int32_t* tc = cell + (size_t)b * N * Tpad;
    const int32_t* tl = len + (size_t)b * N;
    if (counts) counts[b] = (int32_t)res[4 * b + 1];
    if (found && conflicts) {
      const unsigned long long key = res[4 * b];
      if (key != kNoConflict) {
        const int t = (int)(key >> 41), i = (int)((key >> 20) & 0xfffff);
        found[b] = mrp_decode_conflict(key, dimx, hostPos(tc, tl, Tpad, i, t),
                                       hostPos(tc, tl, Tpad, i, t + 1), &conflicts[b]);
      }
    }
  }
  return 0;
}

int mrp_first_conflict(const int32_t* cell, const int32_t* len, int N, int Tpad,
                       int dimx, int mode, mrp_conflict* out) {
  MRP_CHECK(out != nullptr && dimx > 0, MRP_ERR_INVALID, "bad arguments");
  std::memset(out, 0xff, sizeof *out);
  int32_t found = 0;
  int rc = conflictsHost(cell, len, 1, N, Tpad, dimx, mode, true, false, &found, out,
                         nullptr);
  return rc ? rc : found;
}

int mrp_count_conflicts(const int32_t* cell, const int32_t* len, int N, int Tpad,
                        int mode, int32_t* count) {
  MRP_CHECK(count != nullptr, MRP_ERR_INVALID, "count is NULL");
  return conflictsHost(cell, len, 1, N, Tpad, 1, mode, false, true, nullptr, nullptr,
                       count);
}

int mrp_conflicts_batch(const int32_t* cell, const int32_t* len, int B, int N, int Tpad,
                        int dimx, int mode, int32_t* found, mrp_conflict* conflicts,
                        int32_t* counts) {
  MRP_CHECK(B >= 0 && dimx > 0, MRP_ERR_INVALID, "bad arguments");
  MRP_CHECK(B == 0 || (found && conflicts), MRP_ERR_INVALID, "NULL output");
  for (int b = 0; b < B; ++b) std::memset(&conflicts[b], 0xff, sizeof(mrp_conflict));
  return conflictsHost(cell, len, B, N, Tpad, dimx, mode, true, true, found, conflicts,
                       counts);
}

int mrp_conflicts_dev(const int32_t* d_cell, const int32_t* d_len, int N, int Tpad,
                      int mode, int want_first, int want_count,
                      unsigned long long* d_result, void* stream) {
  std::lock_guard<std::mutex> lk(apiMutex());
  MRP_CHECK(d_cell && d_len && d_result, MRP_ERR_INVALID, "NULL device pointer");
  MRP_CHECK(N >= 2 && N < kMaxAgents && Tpad > 0 && Tpad < kMaxTime, MRP_ERR_INVALID,
            "bad table shape N=%d Tpad=%d", N, Tpad);
  MRP_CHECK(want_first || want_count, MRP_ERR_INVALID, "nothing requested");
  if (int rc = ensureInit()) return rc;
  // large N runs the hashed formulation; its transposed table lives in a
  // grow-only scratch buffer of the context (allocated on first use)
  char* d_ws = nullptr;
  const size_t wsBytes = conflictsWorkspaceBytes(N, Tpad);
  if (wsBytes)
    if (int rc = scratch(12, wsBytes, &d_ws)) return rc;
  return launchConflicts(d_cell, d_len, N, Tpad, mode, want_first != 0, want_count != 0,
                         d_result, d_ws, wsBytes, static_cast<cudaStream_t>(stream));
}

int mrp_focal_counts(const int32_t* cell, const int32_t* len, int N, int Tpad, int self,
                     const int32_t* cand_t, const int32_t* cand_from,
                     const int32_t* cand_to, int n_cand, int32_t* state_cnt,
                     int32_t* trans_cnt) {
  if (int rc = checkTable(cell, len, N, Tpad)) return rc;
  MRP_CHECK(n_cand >= 0, MRP_ERR_INVALID, "n_cand < 0");
  if (n_cand == 0) return 0;
  MRP_CHECK(cand_t && cand_from && cand_to && state_cnt && trans_cnt, MRP_ERR_INVALID,
            "NULL pointer");
  for (int k = 0; k < n_cand; ++k)
    MRP_CHECK(cand_t[k] >= 0, MRP_ERR_INVALID, "candidate %d: negative time %d", k, cand_t[k]);
  if (N == 0 || Tpad == 0) {
    std::memset(state_cnt, 0, (size_t)n_cand * 4);
    std::memset(trans_cnt, 0, (size_t)n_cand * 4);
    return 0;
  }
  std::lock_guard<std::mutex> lk(apiMutex());
  if (int rc = ensureInit()) return rc;
  Context& c = ctx();
  int32_t *d_cell, *d_len, *d_cand, *d_out;
  if (int rc = scratch(7, (size_t)N * Tpad, &d_cell)) return rc;
  if (int rc = scratch(8, (size_t)N, &d_len)) return rc;
  if (int rc = scratch(10, (size_t)3 * n_cand, &d_cand)) return rc;
  if (int rc = scratch(11, (size_t)2 * n_cand, &d_out)) return rc;
  MRP_CUDA(cudaMemcpyAsync(d_cell, cell, (size_t)N * Tpad * 4, cudaMemcpyHostToDevice, c.stream));
  MRP_CUDA(cudaMemcpyAsync(d_len, len, (size_t)N * 4, cudaMemcpyHostToDevice, c.stream));
  MRP_CUDA(cudaMemcpyAsync(d_cand, cand_t, (size_t)n_cand * 4, cudaMemcpyHostToDevice, c.stream));
  MRP_CUDA(cudaMemcpyAsync(d_cand + n_cand, cand_from, (size_t)n_cand * 4, cudaMemcpyHostToDevice, c.stream));
  MRP_CUDA(cudaMemcpyAsync(d_cand + 2 * n_cand, cand_to, (size_t)n_cand * 4, cudaMemcpyHostToDevice, c.stream));
  if (int rc = launchFocalCounts(d_cell, d_len, N, Tpad, self, d_cand, d_cand + n_cand,
                                 d_cand + 2 * n_cand, n_cand, d_out, d_out + n_cand,
                                 c.stream))
    return rc;
  MRP_CUDA(cudaMemcpyAsync(state_cnt, d_out, (size_t)n_cand * 4, cudaMemcpyDeviceToHost, c.stream));
  MRP_CUDA(cudaMemcpyAsync(trans_cnt, d_out + n_cand, (size_t)n_cand * 4, cudaMemcpyDeviceToHost, c.stream));
  MRP_CUDA(cudaStreamSynchronize(c.stream));
  return 0;
}

}  // extern "C"
