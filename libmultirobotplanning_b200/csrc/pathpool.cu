// pathpool.cu — device-resident paths of the constraint-tree nodes.
//
// The reference deep-copies a high-level node with all its paths for every
// child (cbs.hpp:144, ecbs.hpp:233) and hands them to getFirstConflict /
// focalHeuristic as host vectors (cbs.cpp:335-386, ecbs.cpp:315-350).  A child
// differs from its parent in ONE path, and that path is produced on the device
// by the replan kernels.  So paths stay where they are made: every path owns a
// row of a pool in HBM (cells only; with cbs / ecbs moves the g-score of a state
// is its time step), a node is a list of row numbers on the host, and
//   * mrp_lowlevel_batch_pool writes the new paths straight into their rows,
//   * mrp_conflicts_batch_pool / the focal tables of the replans gather the
//     rows of a node into the dense [B][N][Tpad] tables the kernels sweep
//     (a device-to-device copy: tens of MB in tens of microseconds),
// and only the row numbers (4 B per agent) cross PCIe per lock-step iteration
// instead of the tables (4 B per agent and time step), plus the results.
#include <algorithm>
#include <cstring>
#include <vector>

#include "lowlevel.cuh"

struct mrp_pathpool_s {
  int rowCap = 0;
  int nSlots = 0;  // capacity, a multiple of kPoolChunk
  std::vector<int32_t*> cells, len;  // device allocations per chunk
  int32_t** d_cells = nullptr;       // device copies of the two pointer lists
  int32_t** d_len = nullptr;
  int maxChunks = 0;
  // state blobs of sliced searches (lowlevel_tile.cu)
  std::vector<unsigned char*> states;
  unsigned char** d_states = nullptr;
  int nStates = 0;
  size_t blobBytes = 0;
  int tileTB = 0, stateMaxNodes = 0;
};

namespace mrp {

// one warp per row: dst[r][0..L) = pool row slots[r], dstLen[r] = L (0 for slot < 0)
__global__ void pool_gather_kernel(int32_t* const* __restrict__ poolCells, int32_t* const* __restrict__ poolLen,
                                   int rowCap, const int32_t* __restrict__ slots, int nRows, int Tpad,
                                   int32_t* __restrict__ dst, int32_t* __restrict__ dstLen) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= nRows) return;
  const int slot = slots[r];
  int L = 0;
  if (slot >= 0) {
    L = min(poolLen[slot >> kPoolChunkBits][slot & (kPoolChunk - 1)], Tpad);
    const int32_t* src = poolCells[slot >> kPoolChunkBits] + (size_t)(slot & (kPoolChunk - 1)) * rowCap;
    int32_t* d = dst + (size_t)r * Tpad;
    for (int t = lane; t < L; t += 32) d[t] = src[t];
  }
  if (lane == 0) dstLen[r] = L;
}

// rows given by the host -> pool rows
__global__ void pool_scatter_kernel(int32_t* const* __restrict__ poolCells, int32_t* const* __restrict__ poolLen,
                                    int rowCap, const int32_t* __restrict__ slots, int nRows,
                                    const int32_t* __restrict__ src, const int32_t* __restrict__ srcLen) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= nRows) return;
  const int slot = slots[r];
  const int L = srcLen[r];
  int32_t* d = poolCells[slot >> kPoolChunkBits] + (size_t)(slot & (kPoolChunk - 1)) * rowCap;
  for (int t = lane; t < L; t += 32) d[t] = src[(size_t)r * rowCap + t];
  if (lane == 0) poolLen[slot >> kPoolChunkBits][slot & (kPoolChunk - 1)] = L;
}

// positions of agent1 at t and t + 1 for the first conflict of every table
// (what mrp_decode_conflict needs; the tables never reach the host)
__global__ void conflict_positions_kernel(const unsigned long long* __restrict__ res, const int32_t* __restrict__ cell,
                                          const int32_t* __restrict__ len, int B, int N, int Tpad,
                                          int32_t* __restrict__ pos) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const unsigned long long key = res[4 * b];
  int p0 = -1, p1 = -1;
  if (key != kNoConflict) {
    const int t = (int)(key >> 41), i = (int)((key >> 20) & 0xfffff);
    const int L = len[(size_t)b * N + i];
    const int32_t* row = cell + ((size_t)b * N + i) * Tpad;
    p0 = row[min(t, L - 1)];
    p1 = row[min(t + 1, L - 1)];
  }
  pos[2 * b] = p0;
  pos[2 * b + 1] = p1;
}

static int checkSlots(mrp_pathpool pool, const int32_t* slots, size_t n, bool allowNone) {
  for (size_t i = 0; i < n; ++i)
    MRP_CHECK((slots[i] >= 0 || (allowNone && slots[i] == -1)) && slots[i] < pool->nSlots, MRP_ERR_INVALID,
              "pool row %d out of range (capacity %d)", slots[i], pool->nSlots);
  return 0;
}

// dense tables of `nRows` pool rows in the lane's scratch: slots 16 (row numbers), 17 (cells), 18 (lengths)
static int gatherRows(mrp_pathpool pool, const int32_t* slots, int nRows, int Tpad, int32_t** d_cell,
                      int32_t** d_len, cudaStream_t st) {
  void *ds = nullptr, *dc = nullptr, *dl = nullptr;
  if (int rc = deviceScratch(16, (size_t)nRows * 4, &ds)) return rc;
  if (int rc = deviceScratch(17, (size_t)nRows * Tpad * 4, &dc)) return rc;
  if (int rc = deviceScratch(18, (size_t)nRows * 4, &dl)) return rc;
  MRP_CUDA(cudaMemcpyAsync(ds, slots, (size_t)nRows * 4, cudaMemcpyHostToDevice, st));
  pool_gather_kernel<<<(nRows + 7) / 8, 256, 0, st>>>(pool->d_cells, pool->d_len, pool->rowCap,
                                                      static_cast<const int32_t*>(ds), nRows, Tpad,
                                                      static_cast<int32_t*>(dc), static_cast<int32_t*>(dl));
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  *d_cell = static_cast<int32_t*>(dc);
  *d_len = static_cast<int32_t*>(dl);
  return 0;
}

}  // namespace mrp

using namespace mrp;

extern "C" {

int mrp_pathpool_create(int row_cap, mrp_pathpool* out) {
  MRP_CHECK(out != nullptr && row_cap > 0 && row_cap <= (1 << 20), MRP_ERR_INVALID, "bad arguments");
  std::lock_guard<std::mutex> lk(apiMutex());
  if (int rc = ensureInit()) return rc;
  mrp_pathpool_s* p = new mrp_pathpool_s();
  p->rowCap = row_cap;
  p->maxChunks = 4096;
  if (cudaMalloc(&p->d_cells, sizeof(void*) * p->maxChunks) != cudaSuccess ||
      cudaMalloc(&p->d_len, sizeof(void*) * p->maxChunks) != cudaSuccess ||
      cudaMalloc(&p->d_states, sizeof(void*) * p->maxChunks) != cudaSuccess) {
    if (p->d_cells) cudaFree(p->d_cells);
    delete p;
    return fail(MRP_ERR_NOMEM, "cudaMalloc failed for the path pool");
  }
  *out = p;
  return 0;
}

int mrp_pathpool_destroy(mrp_pathpool pool) {
  if (!pool) return 0;
  for (int32_t* c : pool->cells) cudaFree(c);
  for (int32_t* l : pool->len) cudaFree(l);
  for (unsigned char* b : pool->states) cudaFree(b);
  cudaFree(pool->d_states);
  cudaFree(pool->d_cells);
  cudaFree(pool->d_len);
  delete pool;
  return 0;
}

int mrp_pathpool_reserve(mrp_pathpool pool, int n_slots) {
  MRP_CHECK(pool != nullptr && n_slots >= 0, MRP_ERR_INVALID, "bad arguments");
  if (n_slots <= pool->nSlots) return 0;
  std::lock_guard<std::mutex> lk(apiMutex());
  if (int rc = ensureInit()) return rc;
  while (pool->nSlots < n_slots) {
    MRP_CHECK((int)pool->cells.size() < pool->maxChunks, MRP_ERR_NOMEM, "path pool is full (%d rows)", pool->nSlots);
    int32_t *c = nullptr, *l = nullptr;
    cudaError_t e = cudaMalloc(&c, (size_t)kPoolChunk * pool->rowCap * 4);
    if (e == cudaSuccess) e = cudaMalloc(&l, (size_t)kPoolChunk * 4);
    if (e != cudaSuccess) {
      if (c) cudaFree(c);
      return fail(MRP_ERR_NOMEM, "cudaMalloc failed for a path pool chunk: %s", cudaGetErrorString(e));
    }
    MRP_CUDA(cudaMemset(l, 0, (size_t)kPoolChunk * 4));
    const size_t k = pool->cells.size();
    pool->cells.push_back(c);
    pool->len.push_back(l);
    // blocking copies: whatever stream uses the pool next sees the new chunk
    MRP_CUDA(cudaMemcpy(pool->d_cells + k, &c, sizeof(void*), cudaMemcpyHostToDevice));
    MRP_CUDA(cudaMemcpy(pool->d_len + k, &l, sizeof(void*), cudaMemcpyHostToDevice));
    pool->nSlots += kPoolChunk;
  }
  return 0;
}

int mrp_pathpool_reserve_states(mrp_pathpool pool, int n_states, int dimx, int dimy, int max_expanded) {
  MRP_CHECK(pool != nullptr && n_states >= 0 && max_expanded > 0 && dimx > 0 && dimy > 0, MRP_ERR_INVALID,
            "bad arguments");
  // rows of the visited bitmap.  192 time steps: with 128 (six warps per SM instead of four) the
  // 100-agent ECBS batch took 1.3-2.1 s instead of 1.0 s: searches that look past t = 128 are
  // handed to the general kernel, which runs them to the end inside the launch
  const int TB = dimx * dimy <= 64 ? 64 : 192;
  const int maxNodes = 5 * max_expanded + 8;
  if (pool->blobBytes == 0) {
    pool->tileTB = TB;
    pool->stateMaxNodes = maxNodes;
    pool->blobBytes = lowlevelTileBlobBytes(TB, maxNodes);
  }
  MRP_CHECK(pool->tileTB == TB && pool->stateMaxNodes >= maxNodes, MRP_ERR_INVALID,
            "the state blobs of this pool were laid out for another map size / expansion cap");
  if (n_states <= pool->nStates) return 0;
  std::lock_guard<std::mutex> lk(apiMutex());
  if (int rc = ensureInit()) return rc;
  while (pool->nStates < n_states) {
    MRP_CHECK((int)pool->states.size() < pool->maxChunks, MRP_ERR_NOMEM, "too many search states (%d)", pool->nStates);
    unsigned char* b = nullptr;
    cudaError_t e = cudaMalloc(&b, (size_t)kStateChunk * pool->blobBytes);
    if (e != cudaSuccess)
      return fail(MRP_ERR_NOMEM, "cudaMalloc failed for a chunk of search states: %s", cudaGetErrorString(e));
    const size_t k = pool->states.size();
    pool->states.push_back(b);
    MRP_CUDA(cudaMemcpy(pool->d_states + k, &b, sizeof(void*), cudaMemcpyHostToDevice));
    pool->nStates += kStateChunk;
  }
  return 0;
}

int mrp_pathpool_write(mrp_pathpool pool, const int32_t* slots, int n, const int32_t* cells, const int32_t* len) {
  MRP_CHECK(pool != nullptr && n >= 0, MRP_ERR_INVALID, "bad arguments");
  if (n == 0) return 0;
  MRP_CHECK(slots && cells && len, MRP_ERR_INVALID, "NULL pointer");
  if (int rc = checkSlots(pool, slots, (size_t)n, false)) return rc;
  for (int i = 0; i < n; ++i)
    MRP_CHECK(len[i] >= 0 && len[i] <= pool->rowCap, MRP_ERR_INVALID, "path %d: bad length %d", i, len[i]);
  std::lock_guard<std::mutex> lk(apiMutex());
  if (int rc = ensureInit()) return rc;
  cudaStream_t st = ctx().stream;
  void *ds = nullptr, *dc = nullptr, *dl = nullptr;
  if (int rc = deviceScratch(16, (size_t)n * 4, &ds)) return rc;
  if (int rc = deviceScratch(17, (size_t)n * pool->rowCap * 4, &dc)) return rc;
  if (int rc = deviceScratch(18, (size_t)n * 4, &dl)) return rc;
  MRP_CUDA(cudaMemcpyAsync(ds, slots, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  MRP_CUDA(cudaMemcpyAsync(dc, cells, (size_t)n * pool->rowCap * 4, cudaMemcpyHostToDevice, st));
  MRP_CUDA(cudaMemcpyAsync(dl, len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  pool_scatter_kernel<<<(n + 7) / 8, 256, 0, st>>>(pool->d_cells, pool->d_len, pool->rowCap,
                                                   static_cast<const int32_t*>(ds), n,
                                                   static_cast<const int32_t*>(dc), static_cast<const int32_t*>(dl));
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  MRP_CUDA(waitStream(st));
  return 0;
}

int mrp_pathpool_read(mrp_pathpool pool, const int32_t* slots, int n, int32_t* cells, int32_t* len) {
  MRP_CHECK(pool != nullptr && n >= 0, MRP_ERR_INVALID, "bad arguments");
  if (n == 0) return 0;
  MRP_CHECK(slots && cells && len, MRP_ERR_INVALID, "NULL pointer");
  if (int rc = checkSlots(pool, slots, (size_t)n, false)) return rc;
  std::lock_guard<std::mutex> lk(apiMutex());
  if (int rc = ensureInit()) return rc;
  cudaStream_t st = ctx().stream;
  int32_t *d_cell = nullptr, *d_len = nullptr;
  if (int rc = gatherRows(pool, slots, n, pool->rowCap, &d_cell, &d_len, st)) return rc;
  MRP_CUDA(cudaMemcpyAsync(cells, d_cell, (size_t)n * pool->rowCap * 4, cudaMemcpyDeviceToHost, st));
  MRP_CUDA(cudaMemcpyAsync(len, d_len, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  MRP_CUDA(waitStream(st));
  return 0;
}

int mrp_conflicts_batch_pool(mrp_pathpool pool, const int32_t* table_slots, int B, int N, int Tpad, int dimx,
                             int mode, int32_t* found, mrp_conflict* conflicts, int32_t* counts) {
  MRP_CHECK(pool != nullptr && B >= 0 && N >= 0 && dimx > 0, MRP_ERR_INVALID, "bad arguments");
  MRP_CHECK(B == 0 || (table_slots && found && conflicts), MRP_ERR_INVALID, "NULL pointer");
  MRP_CHECK(mode == 0 || mode == 1, MRP_ERR_INVALID, "mode must be 0 or 1");
  MRP_CHECK(N < kMaxAgents && Tpad >= 0 && Tpad < kMaxTime, MRP_ERR_UNSUPPORTED, "table shape N=%d Tpad=%d", N, Tpad);
  for (int b = 0; b < B; ++b) {
    std::memset(&conflicts[b], 0xff, sizeof(mrp_conflict));
    found[b] = 0;
    if (counts) counts[b] = 0;
  }
  if (N < 2 || Tpad == 0 || B == 0) return 0;
  if (int rc = checkSlots(pool, table_slots, (size_t)B * N, true)) return rc;
  std::lock_guard<std::mutex> lk(apiMutex());
  if (int rc = ensureInit()) return rc;
  cudaStream_t st = ctx().stream;
  int32_t *d_cell = nullptr, *d_len = nullptr;
  if (int rc = gatherRows(pool, table_slots, B * N, Tpad, &d_cell, &d_len, st)) return rc;
  void *dres = nullptr, *dpos = nullptr;
  if (int rc = deviceScratch(19, (size_t)4 * B * 8, &dres)) return rc;
  if (int rc = deviceScratch(20, (size_t)2 * B * 4, &dpos)) return rc;
  unsigned long long* d_res = static_cast<unsigned long long*>(dres);
  void* d_ws = nullptr;
  const size_t wsBytes = B == 1 ? conflictsWorkspaceBytes(N, Tpad) : 0;
  if (wsBytes)
    if (int rc = deviceScratch(21, wsBytes, &d_ws)) return rc;
  int rc = (B == 1) ? launchConflicts(d_cell, d_len, N, Tpad, mode, true, true, d_res, d_ws, wsBytes, st)
                    : launchConflictsBatch(d_cell, d_len, B, N, Tpad, mode, d_res, st);
  if (rc) return rc;
  conflict_positions_kernel<<<(B + 127) / 128, 128, 0, st>>>(d_res, d_cell, d_len, B, N, Tpad,
                                                             static_cast<int32_t*>(dpos));
  countLaunch();
  MRP_CUDA(cudaGetLastError());
  void* hout = nullptr;
  const size_t resBytes = (size_t)4 * B * 8;
  if (int rc2 = pinnedScratch(6, resBytes + (size_t)2 * B * 4, &hout)) return rc2;
  MRP_CUDA(cudaMemcpyAsync(hout, d_res, resBytes, cudaMemcpyDeviceToHost, st));
  MRP_CUDA(cudaMemcpyAsync(static_cast<char*>(hout) + resBytes, dpos, (size_t)2 * B * 4, cudaMemcpyDeviceToHost, st));
  MRP_CUDA(waitStream(st));
  const unsigned long long* res = static_cast<const unsigned long long*>(hout);
  const int32_t* pos = reinterpret_cast<const int32_t*>(static_cast<char*>(hout) + resBytes);
  for (int b = 0; b < B; ++b) {
    if (counts) counts[b] = (int32_t)res[4 * b + 1];
    if (res[4 * b] != kNoConflict)
      found[b] = mrp_decode_conflict(res[4 * b], dimx, pos[2 * b], pos[2 * b + 1], &conflicts[b]);
  }
  return 0;
}

int mrp_lowlevel_batch_pool(const mrp_map* maps, int n_maps, mrp_fieldset fs, const int32_t* vc, int n_vc,
                            const int32_t* ec, int n_ec, mrp_pathpool pool, const int32_t* table_slots,
                            int n_tables, int N, int Tpad, const mrp_job* jobs, int n_jobs,
                            const mrp_lowlevel_params* params, const int32_t* out_slots, mrp_path_info* info) {
  return mrp_lowlevel_batch_pool_sliced(maps, n_maps, fs, vc, n_vc, ec, n_ec, pool, table_slots, n_tables, N, Tpad,
                                        jobs, n_jobs, params, out_slots, nullptr, nullptr, 0, info);
}

int mrp_lowlevel_batch_pool_sliced(const mrp_map* maps, int n_maps, mrp_fieldset fs, const int32_t* vc, int n_vc,
                                   const int32_t* ec, int n_ec, mrp_pathpool pool, const int32_t* table_slots,
                                   int n_tables, int N, int Tpad, const mrp_job* jobs, int n_jobs,
                                   const mrp_lowlevel_params* params, const int32_t* out_slots,
                                   const int32_t* state_ids, const int32_t* resume, int slice_expanded,
                                   mrp_path_info* info) {
  MRP_CHECK(pool != nullptr && fs != nullptr, MRP_ERR_INVALID, "pool or field set is NULL");
  MRP_CHECK(n_jobs >= 0 && n_tables >= 0 && N >= 0 && Tpad >= 0, MRP_ERR_INVALID, "negative count");
  if (n_jobs == 0) return 0;
  MRP_CHECK(out_slots != nullptr, MRP_ERR_INVALID, "out_slots is NULL");
  MRP_CHECK(n_tables == 0 || table_slots != nullptr, MRP_ERR_INVALID, "table_slots is NULL");
  if (n_maps > 0 && maps && maps[0])
    MRP_CHECK(maps[0]->dimx == fs->dimx && maps[0]->dimy == fs->dimy, MRP_ERR_INVALID,
              "field set and maps differ in dimensions");
  if (int rc = checkSlots(pool, out_slots, (size_t)n_jobs, false)) return rc;
  if (int rc = checkSlots(pool, table_slots, (size_t)n_tables * N, true)) return rc;
  LLPool lp;
  if (state_ids && slice_expanded > 0) {
    MRP_CHECK(resume != nullptr && params != nullptr, MRP_ERR_INVALID, "NULL pointer");
    MRP_CHECK(pool->blobBytes > 0 && 5 * params->max_expanded + 8 <= pool->stateMaxNodes, MRP_ERR_INVALID,
              "mrp_pathpool_reserve_states first (with this expansion cap)");
    for (int j = 0; j < n_jobs; ++j)
      MRP_CHECK(state_ids[j] >= -1 && state_ids[j] < pool->nStates, MRP_ERR_INVALID, "job %d: bad search state %d",
                j, state_ids[j]);
    lp.h_jobState = state_ids;
    lp.h_jobResume = resume;
    lp.d_stateChunks = pool->d_states;
    lp.blobBytes = pool->blobBytes;
    lp.sliceCap = slice_expanded;
    lp.tileTB = pool->tileTB;
  }
  lp.d_poolCells = pool->d_cells;
  lp.d_poolLen = pool->d_len;
  lp.h_outSlots = out_slots;
  lp.rowCap = pool->rowCap;
  const bool haveTables = n_tables > 0 && N > 0 && Tpad > 0;
  if (haveTables) {
    // the gather runs on the lane's stream, in front of the replan kernels
    std::lock_guard<std::mutex> lk(apiMutex());
    if (int rc = ensureInit()) return rc;
    int32_t *d_cell = nullptr, *d_len = nullptr;
    if (int rc = gatherRows(pool, table_slots, n_tables * N, Tpad, &d_cell, &d_len, ctx().stream)) return rc;
    lp.d_tables = d_cell;
    lp.d_tlen = d_len;
  }
  return lowlevelRun(maps, n_maps, nullptr, fs->d_fields, fs->n_fields, vc, n_vc, ec, n_ec, nullptr, nullptr,
                     haveTables ? n_tables : 0, N, Tpad, jobs, n_jobs, params, info, nullptr, nullptr, &lp);
}

}  // extern "C"
