/*
 * mrp_b200.h — C ABI of the B200-native hot path of libMultiRobotPlanning.
 *
 * The reference is a header-only C++14 template library with NO plugin / FFI
 * layer; its hot path is the duck-typed `Environment` concept that the
 * CBS / ECBS / CBS-TA templates call back into.  This header is the boundary a
 * maintainer binds instead: plain pointers and sizes, no C++ or torch types.
 * Every entry point cites the reference code it replaces.  The C++ adapters in
 * libmultirobotplanning_b200/host/ re-create the reference's `Environment`
 * signatures on top of these calls.
 *
 * Conventions
 *   - cell index = x + dimx*y (example/shortest_path_heuristic.hpp:65)
 *   - all costs/distances are int32; MRP_INF = INT_MAX marks unreachable
 *   - return value: 0 (or a non-negative answer) on success, negative
 *     mrp_status on error; mrp_last_error() describes the last error of the
 *     calling thread.  No exception crosses the ABI.
 *   - "host" entry points take host pointers and do their own H2D/D2H copies;
 *     `_dev` entry points take device pointers + a cudaStream_t (as void*) and
 *     are asynchronous with respect to the host.
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     MRP_ERR_NO_DEVICE.
 */
#ifndef MRP_B200_H
#define MRP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRP_INF 2147483647

typedef enum {
  MRP_OK = 0,
  MRP_ERR_NO_DEVICE = -1,
  MRP_ERR_INVALID = -2,
  MRP_ERR_CUDA = -3,
  MRP_ERR_UNSUPPORTED = -4,
  MRP_ERR_NOMEM = -5
} mrp_status;

/* ---- context --------------------------------------------------------- */
/* Binds the calling process to CUDA device `device` (-1: keep the current
 * one) and creates the library's streams.  Idempotent. */
int mrp_init(int device);
int mrp_shutdown(void);
/* Lanes.  The reference is single-threaded and its Environment is not
 * re-entrant (setLowLevelContext, example/cbs.cpp:266-276); this library is
 * thread-compatible per LANE: a lane owns its streams, staging buffers and
 * replan workspace, calls made in one lane are serialised, calls made in
 * different lanes overlap on the device (the batched drivers run sub-batches
 * of instances in separate lanes so that one long replan does not hold up the
 * others).  mrp_set_lane binds the calling host thread to lane 0 <= k <
 * mrp_max_lanes(); threads start in lane 0.  Maps and field sets may be
 * shared between lanes (they are read-only on the device). */
int mrp_set_lane(int lane);
int mrp_max_lanes(void);
int mrp_device_count(void);
const char* mrp_last_error(void);
/* "major.minor name smcount" of the bound device, for logs */
const char* mrp_device_info(void);

/* ---- maps --------------------------------------------------------------
 * A map handle owns the device-resident, bit-packed free-cell mask of one grid
 * (1 bit per cell, 32x32-cell tiles).  Replaces the
 * `std::unordered_set<Location> obstacles` every reference Environment keeps
 * (example/cbs.cpp:561, example/cbs_ta.cpp:498). */
typedef struct mrp_map_s* mrp_map;
int mrp_map_create(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                   mrp_map* out);
int mrp_map_destroy(mrp_map map);

/* ---- (1) distance fields -----------------------------------------------
 * Replaces ShortestPathHeuristic::ShortestPathHeuristic (Floyd–Warshall over
 * all cells, example/shortest_path_heuristic.hpp:12-54) and getValue (:58-62)
 * by the goal rows only, i.e. the layout of the reference's disabled
 * Environment::computeHeuristic (example/cbs.cpp:445-557):
 *   out[g][x + dimx*y] = BFS distance from goal g, MRP_INF if unreachable or
 *   obstacle; a goal that is itself an obstacle yields 0 at the goal and
 *   MRP_INF elsewhere (the Floyd–Warshall row of an isolated vertex).
 *
 * Transfer: results of 16 Mi cells or more leave the device packed and are
 * expanded to int32 by host threads while the next batch is on the bus (the
 * call is PCIe-bound at 4 B per cell).  Formats, narrowest first, chosen per
 * batch of goals: one byte per cell, (distance - Manhattan distance to the
 * goal) / 2, 255 = MRP_INF (the two distances have the same parity on a
 * 4-connected grid); uint16, 0xFFFF = MRP_INF; int32.  A batch moves on to the
 * next format when a value does not fit (detours of 510 steps or more, finite
 * distances >= 65535).  `out` holds the same int32 values either way.
 * Environment: MRP_BFS_PACK=0 never packs, =1 always packs; MRP_BFS_FMT=16
 * starts at uint16; MRP_WIDEN_THREADS = host threads of the expansion
 * (default: all cores, at most 16). */
int mrp_bfs_fields(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                   const int32_t* goal_xy, int n_goals, int32_t* out);
/* The host half of the packed transfer, exported for device-resident
 * pipelines that copy uint16 fields themselves:
 * dst[i] = src[i] == 0xFFFF ? MRP_INF : src[i] on `threads` host threads.
 * Needs no device. */
int mrp_widen_u16(const uint16_t* src, int32_t* dst, size_t n, int threads);
/* Same for the one-byte format: dst[k][y][x] = src == 255 ? MRP_INF :
 * 2*src + |x - gx_k| + |y - gy_k|, goal_cell[k] = gx_k + dimx*gy_k. */
int mrp_widen_u8(const uint8_t* src, int32_t* dst, int dimx, int dimy,
                 const int32_t* goal_cell, int n_fields, int threads);
/* The packed RESULT mode: the same fields, handed to the caller as one byte per
 * cell instead of int32 — out[g][x + dimx*y] = (distance - (|x-gx| + |y-gy|)) / 2,
 * 255 = MRP_INF — copied straight into `out` (use page-locked memory), a quarter
 * of the bytes over PCIe and no expansion on the host.  mrp_packed_value() below
 * is ShortestPathHeuristic::getValue (example/shortest_path_heuristic.hpp:58-62)
 * on such a field.  A field with a detour of 510 steps or more (mazes) does not
 * fit: the call returns the number of such goals (0 = all fine) and marks them
 * in overflowed[n_goals] (may be NULL); recompute those with mrp_bfs_fields. */
int mrp_bfs_fields_packed(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                          const int32_t* goal_xy, int n_goals, uint8_t* out,
                          int32_t* overflowed);
static inline int32_t mrp_packed_value(const uint8_t* field, int dimx, int x, int y,
                                       int gx, int gy) {
  const uint8_t b = field[x + dimx * y];
  if (b == 255) return MRP_INF;
  return 2 * (int32_t)b + (x > gx ? x - gx : gx - x) + (y > gy ? y - gy : gy - y);
}
/* Compact result: the same detour bytes for the FREE cells of the map only, in cell order
 * (an obstacle cell is MRP_INF in every field, and the caller has the map): out is
 * [n_goals][n_free].  mrp_free_cell_index (host only, no device needed) returns n_free and
 * fills free_bits[ceil(cells/32)] (bit c & 31 of word c >> 5 = cell c = x + dimx*y is free)
 * and prefix[ceil(cells/32) + 1] (free cells before each word); mrp_compact_value is getValue
 * on such a field.  On the C5 map 20 % fewer bytes cross PCIe, which is what bounds the call. */
int mrp_free_cell_index(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                        uint32_t* free_bits, int32_t* prefix);
int mrp_bfs_fields_compact(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                           const int32_t* goal_xy, int n_goals, uint8_t* out,
                           int32_t* overflowed);
static inline int32_t mrp_compact_value(const uint8_t* field, const uint32_t* free_bits,
                                        const int32_t* prefix, int dimx, int x, int y, int gx,
                                        int gy) {
  const int c = x + dimx * y;
  const uint32_t w = free_bits[c >> 5], bit = 1u << (c & 31);
  if (x == gx && y == gy) return 0; /* also for a goal on an obstacle (its Floyd-Warshall row) */
  if (!(w & bit)) return MRP_INF;
  const uint8_t b = field[prefix[c >> 5] + __builtin_popcount(w & (bit - 1u))];
  if (b == 255) return MRP_INF;
  return 2 * (int32_t)b + (x > gx ? x - gx : gx - x) + (y > gy ? y - gy : gy - y);
}
/* Bytes the last mrp_bfs_fields call moved from the device to the host
 * (bench.py's `d2h_bytes_per_step`). */
long long mrp_bfs_d2h_bytes(void);
/* Many small maps in one launch (configs C2/C4: every instance has its own
 * obstacle layout).  dims[m] = (dimx, dimy); obstacles and goals are CSR over
 * maps; fields are written back to back in goal order (field of goal k has
 * dimx*dimy entries of its map). */
int mrp_bfs_fields_batch(int n_maps, const int32_t* dims,
                         const int32_t* obst_off, const int32_t* obst_xy,
                         const int32_t* goal_off, const int32_t* goal_xy,
                         int32_t* out);
/* Device-resident variant: goals as cell indices, out = device int32
 * [n_goals][dimx*dimy].  `workspace` must hold
 * mrp_bfs_workspace_bytes(map, n_goals) bytes (may be 0). */
size_t mrp_bfs_workspace_bytes(mrp_map map, int n_goals);
int mrp_bfs_fields_dev(mrp_map map, const int32_t* d_goal_cell, int n_goals,
                       int32_t* d_out, void* d_workspace, void* stream);

/* ---- (2) conflicts over packed path tables -----------------------------
 * Path table: cell[N][Tpad] int32 + len[N] (= PlanResult::states.size());
 * positions past the end clamp to the last cell exactly like
 * Environment::getState (example/cbs.cpp:420-429).  Entries with t >= len are
 * never read.  mode 0: max_t = max(len)-1 (cbs/ecbs, example/cbs.cpp:338-341);
 * mode 1: max_t = max(len) (cbs_ta, example/cbs_ta.cpp:372-375).  An agent with
 * len = 0 (no path yet: the reference's getState asserts, only the focal
 * counters skip such agents, example/ecbs.cpp:287) matches nothing. */
typedef struct {
  int32_t time;
  int32_t agent1;
  int32_t agent2;
  int32_t type; /* 0 Vertex, 1 Edge — Conflict::Type, example/cbs.cpp:81-84 */
  int32_t x1, y1, x2, y2; /* x2,y2 = -1 for Vertex (unset in the reference) */
} mrp_conflict;

/* Environment::getFirstConflict (example/cbs.cpp:335-386,
 * example/cbs_ta.cpp:369-420): returns 1 and fills *out with the first hit in
 * the order (t, Vertex<Edge, agent1, agent2); 0 if conflict-free. */
int mrp_first_conflict(const int32_t* cell, const int32_t* len, int N, int Tpad,
                       int dimx, int mode, mrp_conflict* out);
/* Environment::focalHeuristic (example/ecbs.cpp:315-350): number of vertex +
 * edge conflicts over all agent pairs and timesteps, no early exit. */
int mrp_count_conflicts(const int32_t* cell, const int32_t* len, int N,
                        int Tpad, int mode, int32_t* count);
/* Both answers for B tables of identical shape in one launch (one table per
 * constraint-tree node): cell[B][N][Tpad], len[B][N]; found[B] (0/1),
 * conflicts[B], counts[B] (counts may be NULL). */
int mrp_conflicts_batch(const int32_t* cell, const int32_t* len, int B, int N,
                        int Tpad, int dimx, int mode, int32_t* found,
                        mrp_conflict* conflicts, int32_t* counts);
/* Environment::focalStateHeuristic / focalTransitionHeuristic
 * (example/ecbs.cpp:282-312) for n_cand candidate moves of agent `self`:
 * candidate k leaves cand_from[k] at time cand_t[k] and arrives at cand_to[k]
 * at cand_t[k]+1.  Agents with len == 0 are skipped (example/ecbs.cpp:287). */
int mrp_focal_counts(const int32_t* cell, const int32_t* len, int N, int Tpad,
                     int self, const int32_t* cand_t, const int32_t* cand_from,
                     const int32_t* cand_to, int n_cand, int32_t* state_cnt,
                     int32_t* trans_cnt);
/* Device-resident variant of first-conflict + count.  d_result is 4 x int64:
 * [0] packed first-conflict key (t<<41 | type<<40 | i<<20 | j), ~0 if none;
 * [1] conflict count; [2],[3] scratch.  The call resets d_result itself.
 * The hashed kernels (256 < N <= 4096) keep their transposed table in a scratch
 * buffer of the calling thread's lane: calls of one lane must use one stream (or
 * be ordered by the caller); calls from different lanes are independent.
 * len[i] <= Tpad is the caller's contract for device tables (not checked). */
int mrp_conflicts_dev(const int32_t* d_cell, const int32_t* d_len, int N,
                      int Tpad, int mode, int want_first, int want_count,
                      unsigned long long* d_result, void* stream);
/* Decodes d_result[0] (copied to the host) into a conflict; returns 1/0.
 * cell_i_t / cell_i_t1 are agent1's cells at `time` and `time+1`. */
int mrp_decode_conflict(unsigned long long key, int dimx, int32_t cell_i_t,
                        int32_t cell_i_t1, mrp_conflict* out);

/* ---- (3) batched low-level replans --------------------------------------
 * One job = one call of AStar::search (a_star.hpp:63-161) or
 * AStarEpsilon::search (a_star_epsilon.hpp:86-285) for one agent under one
 * constraint set, i.e. what CBS::search / ECBS::search / CBSTA::search issue
 * at cbs.hpp:155-157, ecbs.hpp:265-268, cbs_ta.hpp:192-195.
 * Constraints follow VertexConstraint(time,x,y) (arrival time) and
 * EdgeConstraint(time,x1,y1,x2,y2) (departure time), example/cbs.cpp:108-163. */
typedef struct {
  int32_t map;        /* index into the maps[] array of the call */
  int32_t start_cell; /* State(0, x, y) */
  int32_t goal_cell;  /* -1: agent without task (cbs_ta, cbs_ta.cpp:283-319) */
  int32_t field;      /* index into the fields of the call: distance field of
                         goal_cell on that map (admissible heuristic) */
  int32_t vc_begin, vc_end; /* range in vc[][2] = (time, cell) */
  int32_t ec_begin, ec_end; /* range in ec[][3] = (time, from, to) */
  int32_t table;  /* ECBS: index of the other agents' path table, -1: none */
  int32_t self;   /* ECBS: this agent's row in that table (skipped) */
} mrp_job;

typedef struct {
  int32_t status;   /* 0 solved, 1 no solution, 2 capped (horizon/expansions) */
  int32_t cost;     /* PlanResult::cost */
  int32_t fmin;     /* PlanResult::fmin */
  int32_t length;   /* number of states (path rows written) */
  int32_t expanded; /* low-level expansions (onExpandLowLevelNode) */
} mrp_path_info;

typedef struct {
  int32_t variant;      /* 0: cbs/ecbs moves (all cost 1); 1: cbs_ta (waiting
                           on the goal is free, example/cbs_ta.cpp:329-339) */
  float w;              /* < 1 (e.g. 0): A*.  >= 1: A*-epsilon focal search; w == 1.0
                           is the focal search among the nodes with f == fmin, i.e.
                           optimal cost with the fewest conflicts, what `ecbs -w 1.0`
                           runs in the reference (a_star_epsilon.hpp:240) */
  int32_t max_expanded; /* per job cap (the reference has none) */
  int32_t path_cap;     /* rows available per job in out_cells / out_g */
} mrp_lowlevel_params;

/* fields: [n_fields][cells] distance fields (e.g. from mrp_bfs_fields); all
 * maps of one call must have the same dimensions.  tables: [n_tables][N][Tpad]
 * + table_len[n_tables][N] (may be NULL when no job has table >= 0).
 * out_cells/out_g: [n_jobs][path_cap] cell and g-score per state. */
int mrp_lowlevel_batch(const mrp_map* maps, int n_maps, const int32_t* fields,
                       int n_fields, const int32_t* vc, int n_vc,
                       const int32_t* ec, int n_ec, const int32_t* tables,
                       const int32_t* table_len, int n_tables, int N, int Tpad,
                       const mrp_job* jobs, int n_jobs,
                       const mrp_lowlevel_params* params, mrp_path_info* info,
                       int32_t* out_cells, int32_t* out_g);

/* Device-resident distance fields ("field set"): computed once per instance
 * batch and kept in HBM so that the replans of every constraint-tree node read
 * them in place — the role ShortestPathHeuristic::m_shortestDistance plays for
 * the lifetime of a reference Environment (example/cbs_ta.cpp:511).  All maps
 * must share their dimensions; goal k lives on map goal_map[k]. */
typedef struct mrp_fieldset_s* mrp_fieldset;
int mrp_fieldset_create(const mrp_map* maps, int n_maps, const int32_t* goal_map,
                        const int32_t* goal_cell, int n_goals, mrp_fieldset* out);
int mrp_fieldset_read(mrp_fieldset fs, int first, int count, int32_t* out);
int mrp_fieldset_destroy(mrp_fieldset fs);
/* mrp_lowlevel_batch with job.field indexing a field set instead of a host
 * array. */
int mrp_lowlevel_batch_fs(const mrp_map* maps, int n_maps, mrp_fieldset fs,
                          const int32_t* vc, int n_vc, const int32_t* ec, int n_ec,
                          const int32_t* tables, const int32_t* table_len,
                          int n_tables, int N, int Tpad, const mrp_job* jobs,
                          int n_jobs, const mrp_lowlevel_params* params,
                          mrp_path_info* info, int32_t* out_cells, int32_t* out_g);

/* ---- device-resident paths of the constraint-tree nodes -------------------
 * The reference copies a high-level node with all its paths for every child
 * (HighLevelNode newNode = P, include/libMultiRobotPlanning/cbs.hpp:144,
 * ecbs.hpp:233) and passes them to getFirstConflict / focalHeuristic as host
 * vectors (example/cbs.cpp:335-386, example/ecbs.cpp:315-350).  A child differs
 * from its parent in one path, and the replan kernels produce that path on the
 * device.  A path pool keeps every path in a row of device memory (cells only:
 * with cbs / ecbs moves the g-score of a state is its time step, so this mode is
 * limited to variant 0); a node is a list of row numbers owned by the caller,
 * who also decides which rows are free.  Per call only row numbers and results
 * cross PCIe; the rows of a node are gathered on the device into the dense
 * tables the kernels sweep.  Row numbers: 0 <= row < capacity (mrp_pathpool_reserve),
 * -1 in a table = that agent has no path (yet). */
typedef struct mrp_pathpool_s* mrp_pathpool;
int mrp_pathpool_create(int row_cap, mrp_pathpool* out);
int mrp_pathpool_destroy(mrp_pathpool pool);
/* makes rows 0 .. n_slots-1 valid (grows in chunks; existing rows keep their content) */
int mrp_pathpool_reserve(mrp_pathpool pool, int n_slots);
/* cells[n][row_cap] + len[n] <-> rows slots[n] (tests, output of the final solution) */
int mrp_pathpool_write(mrp_pathpool pool, const int32_t* slots, int n, const int32_t* cells,
                       const int32_t* len);
int mrp_pathpool_read(mrp_pathpool pool, const int32_t* slots, int n, int32_t* cells,
                      int32_t* len);
/* mrp_conflicts_batch over tables given as pool rows: table_slots[B][N]; Tpad >= the
 * longest path of the call */
int mrp_conflicts_batch_pool(mrp_pathpool pool, const int32_t* table_slots, int B, int N,
                             int Tpad, int dimx, int mode, int32_t* found,
                             mrp_conflict* conflicts, int32_t* counts);
/* mrp_lowlevel_batch_fs with the other agents' paths given as pool rows
 * (table_slots[n_tables][N]) and the path of job j written to row out_slots[j]
 * (rows of failed jobs keep their old content); info[j] as usual, no cells come back */
int mrp_lowlevel_batch_pool(const mrp_map* maps, int n_maps, mrp_fieldset fs,
                            const int32_t* vc, int n_vc, const int32_t* ec, int n_ec,
                            mrp_pathpool pool, const int32_t* table_slots, int n_tables,
                            int N, int Tpad, const mrp_job* jobs, int n_jobs,
                            const mrp_lowlevel_params* params, const int32_t* out_slots,
                            mrp_path_info* info);

/* Sliced searches (single-tile maps): a lock-step batch lasts as long as its longest
 * search, and on 100-agent instances the longest of a few hundred replans is
 * thousands of expansions while the typical one is a few dozen.  With a slice, job j
 * stops after slice_expanded expansions of this call (info[j].status ==
 * MRP_SUSPENDED) and leaves its search state (OPEN, visited set, node pool) in blob
 * state_ids[j] of the pool; a later call with resume[j] != 0 and the same job
 * continues it — the same search, expansion for expansion, cut into launches, while
 * the instances whose replans are done move on.  state_ids[j] == -1: that job runs to
 * its end.  Maps the shared-memory kernel does not take (more than one 32x32 tile)
 * ignore the slice.  Blobs are reserved with mrp_pathpool_reserve_states (one layout
 * per pool: map size and expansion cap of the first call). */
#define MRP_SUSPENDED 4
int mrp_pathpool_reserve_states(mrp_pathpool pool, int n_states, int dimx, int dimy,
                                int max_expanded);
int mrp_lowlevel_batch_pool_sliced(const mrp_map* maps, int n_maps, mrp_fieldset fs,
                                   const int32_t* vc, int n_vc, const int32_t* ec, int n_ec,
                                   mrp_pathpool pool, const int32_t* table_slots,
                                   int n_tables, int N, int Tpad, const mrp_job* jobs,
                                   int n_jobs, const mrp_lowlevel_params* params,
                                   const int32_t* out_slots, const int32_t* state_ids,
                                   const int32_t* resume, int slice_expanded,
                                   mrp_path_info* info);

/* ---- multi-GPU ----------------------------------------------------------
 * One process (or host thread) per GPU, as SURVEY.md §8(e) shards the path:
 * distance fields by goal, conflict checks by agent-pair block, replans by
 * constraint-tree node / instance (no collective: results return to the host
 * that owns the high-level OPEN list).  The reference has no counterpart: it is
 * single-threaded (example/cbs.cpp:571-667).  NCCL is bound at run time
 * (dlopen of libnccl.so.2); without it these calls fail with
 * MRP_ERR_UNSUPPORTED and everything else works.
 *
 * Rank 0 makes an id (mrp_comm_unique_id), the caller hands its 128 bytes to
 * the other ranks by any side channel (file, MPI, torch.distributed, a socket),
 * and every rank — after mrp_init(device) — joins with mrp_comm_init_rank. */
#define MRP_COMM_ID_BYTES 128
int mrp_comm_unique_id(void* id128);
int mrp_comm_init_rank(const void* id128, int n_ranks, int rank);
/* returns 1 if a communicator exists; any out pointer may be NULL */
int mrp_comm_info(int* rank, int* n_ranks, int* nccl_version);
int mrp_comm_destroy(void);
/* Distance fields sharded by goal, resident everywhere afterwards: every rank
 * passes the same goal list; rank r computes goals [r*per, (r+1)*per), per =
 * ceil(n_goals / n_ranks), in chunks of one wave of goals; a finished chunk is
 * all-gathered as ONE BYTE per cell ((distance - Manhattan distance)/2, 255 =
 * MRP_INF) on the library's second stream while the next chunk is computed,
 * and expanded to int32 on the device: d_out[n_goals][dimx*dimy] holds every
 * field on every rank when the stream reaches the end of the call.  If a
 * detour does not fit a byte anywhere (all ranks agree through an all-reduce;
 * this is the one host synchronisation of the call) the gather is repeated with
 * int32.  Needs dimx*dimy % 16 == 0 and a 16-byte aligned d_out.  Without a
 * communicator it is mrp_bfs_fields_dev.  This is what the cost matrix of
 * cbs_ta reads (example/cbs_ta.cpp:272-280: getValue(start_i, goal_j) for all
 * i, j) when the goals' fields are computed on several GPUs. */
size_t mrp_bfs_allgather_workspace_bytes(mrp_map map, int n_goals);
int mrp_bfs_fields_allgather_dev(mrp_map map, const int32_t* d_goal_cell, int n_goals,
                                 int32_t* d_out, void* d_workspace, void* stream);
/* After the stream of the last gather has finished: summed device time of its
 * (first 64) all-gathers, the bytes those and all of them brought into this
 * GPU, and the wire format used (1 = detour bytes, 4 = int32). */
int mrp_comm_last_gather(double* collective_ms, long long* timed_bytes_in,
                         long long* wire_bytes_in, int* bytes_per_cell);
/* Conflict checks by agent-pair block: every rank holds the whole path table
 * and sweeps every n_ranks-th 64x64 block of agent pairs (all-pairs kernel);
 * the first-conflict keys meet in an all-reduce MIN, the counts in a SUM
 * (order of the reference's loops, example/cbs.cpp:343-383).  d_result as in
 * mrp_conflicts_dev, identical on every rank afterwards.  The formulation for
 * tables beyond one GPU's hashed path (more than 4096 agents); on the C5
 * table one GPU's hashed sweep is faster than any sharding. */
int mrp_conflicts_sharded_dev(const int32_t* d_cell, const int32_t* d_len, int N,
                              int Tpad, int mode, int want_first, int want_count,
                              unsigned long long* d_result, void* stream);

/* ---- instrumentation ---------------------------------------------------- */
/* number of kernel launches issued by this library since mrp_init (bench.py's
 * `gpu_launches`) */
long long mrp_launch_count(void);
/* The geometry of the bordered row-major bitmap the queue BFS kernel keeps in shared memory for a map
 * of `dimx` columns: words per row (odd, at least 3) and the multiply-shift constants with which the
 * kernel turns a word index into a row (row = umulhi(word, magic) >> shift).  Host code, needs no
 * device: exported so that the CPU tests can check the constants for every width (a width of at most 30
 * columns once got a stride of one word, for which no 32-bit constant exists). */
int mrp_bitmap_row_division(int dimx, int32_t* row_words, uint32_t* magic, int32_t* shift);

#ifdef __cplusplus
}
#endif
#endif
