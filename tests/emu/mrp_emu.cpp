// mrp_emu.cpp — a HOST-ONLY stand-in for libmrp_b200.so.  TEST INFRASTRUCTURE, never shipped:
// it lives under tests/, is built by tests/test_host_driver_emu.py into tests/emu/_build/ and is
// loaded only by the subprocess that test starts.  The product libraries never link or load it
// (libmrp_b200.so has no CPU fallback; without a device its calls fail with MRP_ERR_NO_DEVICE).
//
// Why it exists: the batched high-level drivers (libmultirobotplanning_b200/host/hl_search.hpp —
// CBS / ECBS / CBS-TA / ECBS-TA loops, flights, sliced replans, pool rows, lanes) are host logic
// that only ever runs against a GPU.  This file implements the 21 C-ABI entry points that
// libmrp_host.so imports (include/mrp_b200.h) on top of the CPU oracle (oracle/mrp_oracle.cpp:
// the restatement of the reference's A*, A*-epsilon and conflict loops), so that `-m "not gpu"`
// tests can drive the real driver code on a machine without a device and compare its answers
// with the reference's.  It checks the ABI contracts the device library relies on (row numbers
// inside the reserved range, paths no longer than Tpad, resumed jobs that were suspended, state
// blobs not shared by two jobs) and fails loudly on a violation.
//
// Sliced searches are emulated at the interface: the first call runs the search to its end,
// keeps the result in the state blob and reports MRP_SUSPENDED until the job has been given as
// many expansions as the search took — the same sequence of statuses the driver sees from the
// device, with results that do not depend on the slice.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mrp_b200.h"
#include "../../oracle/mrp_oracle.h"

namespace {

thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define EMU_CHECK(cond, ...) \
  do {                       \
    if (!(cond)) return fail(MRP_ERR_INVALID, __VA_ARGS__); \
  } while (0)

struct Counters {
  std::mutex mu;
  // conflicts, conflicts_pool, ll_fs, ll_pool, ll_sliced, suspended, resumed, jobs, then the sum of
  // the content hashes of all calls (what the device would have been given, row numbers and state
  // ids left out: two drivers with equal sums made the same calls up to their order)
  long calls[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
} g_cnt;
void bump(int k, long n = 1) {
  std::lock_guard<std::mutex> lk(g_cnt.mu);
  g_cnt.calls[k] += n;
}
struct Fnv {
  unsigned long long h = 1469598103934665603ull;
  void add(const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
  }
  template <class T>
  void val(const T& v) { add(&v, sizeof v); }
};
void bumpHash(const Fnv& f) {
  std::lock_guard<std::mutex> lk(g_cnt.mu);
  g_cnt.calls[8] = (long)((unsigned long long)g_cnt.calls[8] + f.h);
}
// content of dense tables [B][N][Tpad] / len[B][N]: the cells up to each path's end
void hashTables(Fnv& f, const int32_t* cell, const int32_t* len, int B, int N, int Tpad) {
  for (size_t k = 0; k < (size_t)B * N; ++k) {
    f.val(len[k]);
    f.add(cell + k * Tpad, (size_t)len[k] * 4);
  }
}
void hashJob(Fnv& f, const mrp_job& q, const int32_t* vc, const int32_t* ec) {
  f.val(q.map);
  f.val(q.start_cell);
  f.val(q.goal_cell);
  f.val(q.field);
  f.val(q.table);
  f.val(q.self);
  f.add(vc + 2 * (size_t)q.vc_begin, (size_t)(q.vc_end - q.vc_begin) * 8);
  f.add(ec + 3 * (size_t)q.ec_begin, (size_t)(q.ec_end - q.ec_begin) * 12);
}

}  // namespace

struct mrp_map_s {
  int dimx, dimy;
  std::vector<int32_t> obst;  // x, y pairs inside the map
};
struct mrp_fieldset_s {
  int cells = 0;
  std::vector<int32_t> data;  // [n_goals][cells]
  std::vector<int32_t> goalCell;
};
struct mrp_pathpool_s {
  int rowCap = 0;
  std::mutex mu;  // reserve() may run while another lane's pool is in use; one pool = one lane
  std::deque<std::vector<int32_t> > rows;
  std::deque<int32_t> len;
  struct Blob {
    bool live = false;
    mrp_path_info info;
    std::vector<int32_t> cells;
    long given = 0;
    mrp_job job;
  };
  std::deque<Blob> blobs;
  int blobDimx = 0, blobDimy = 0, blobCap = 0;
};

namespace {

int checkJobs(const mrp_job* jobs, int n_jobs, int n_maps, int n_fields, int n_vc, int n_ec, int n_tables,
              int N) {
  for (int j = 0; j < n_jobs; ++j) {
    const mrp_job& q = jobs[j];
    EMU_CHECK(q.map >= 0 && q.map < n_maps, "job %d: map %d of %d", j, q.map, n_maps);
    EMU_CHECK(q.field >= -1 && q.field < n_fields, "job %d: field %d of %d", j, q.field, n_fields);
    EMU_CHECK(0 <= q.vc_begin && q.vc_begin <= q.vc_end && q.vc_end <= n_vc, "job %d: vc range", j);
    EMU_CHECK(0 <= q.ec_begin && q.ec_begin <= q.ec_end && q.ec_end <= n_ec, "job %d: ec range", j);
    EMU_CHECK(q.table >= -1 && q.table < n_tables, "job %d: table %d of %d", j, q.table, n_tables);
    EMU_CHECK(q.table < 0 || (q.self >= 0 && q.self < N), "job %d: self %d of %d", j, q.self, N);
  }
  return 0;
}

// one replan through the oracle; cells / gs receive the path (may be NULL)
int runJob(const mrp_map* maps, const mrp_fieldset fs, const int32_t* vc, const int32_t* ec,
           const int32_t* tables, const int32_t* tlen, int N, int Tpad, const mrp_job& q,
           const mrp_lowlevel_params& prm, mrp_path_info& info, std::vector<int32_t>& cells,
           std::vector<int32_t>& gs) {
  const mrp_map m = maps[q.map];
  if (fs && q.field >= 0 && q.goal_cell >= 0)
    EMU_CHECK(fs->goalCell[q.field] == q.goal_cell, "field %d belongs to goal cell %d, the job's goal is %d",
              q.field, fs->goalCell[q.field], q.goal_cell);
  int32_t cost = 0, fmin = 0, plen = 0;
  int64_t expanded = 0;
  std::vector<int32_t> tcg((size_t)3 * (prm.path_cap + 1));
  const int32_t* oc = q.table >= 0 ? tables + (size_t)q.table * N * Tpad : nullptr;
  const int32_t* ol = q.table >= 0 ? tlen + (size_t)q.table * N : nullptr;
  const int st = orc_lowlevel(m->dimx, m->dimy, m->obst.data(), (int)m->obst.size() / 2, prm.variant,
                              q.start_cell, q.goal_cell, vc + 2 * (size_t)q.vc_begin, q.vc_end - q.vc_begin,
                              ec + 3 * (size_t)q.ec_begin, q.ec_end - q.ec_begin, prm.w >= 1.0f ? prm.w : 0.0f,
                              oc, ol, q.table >= 0 ? N : 0, Tpad, q.self, prm.max_expanded, &cost, &fmin,
                              &expanded, tcg.data(), prm.path_cap + 1, &plen);
  info.status = st;
  info.cost = info.fmin = info.length = 0;
  info.expanded = (int32_t)expanded;
  cells.clear();
  gs.clear();
  if (st != ORC_SOLVED) return 0;
  if (plen > prm.path_cap) {  // does not fit the caller's rows: the device reports it as capped
    info.status = 2;
    return 0;
  }
  info.cost = cost;
  info.fmin = fmin;
  info.length = plen;
  for (int k = 0; k < plen; ++k) {
    cells.push_back(tcg[3 * k + 1]);
    gs.push_back(tcg[3 * k + 2]);
  }
  return 0;
}

// the dense tables of `table_slots[B][N]` (row -1: no path, length 0)
int gather(mrp_pathpool pool, const int32_t* slots, int B, int N, int Tpad, std::vector<int32_t>& cell,
           std::vector<int32_t>& len) {
  cell.assign((size_t)B * N * std::max(Tpad, 1), -7);  // cells past a path's end are never to be read
  len.assign((size_t)B * N, 0);
  for (size_t k = 0; k < (size_t)B * N; ++k) {
    const int r = slots[k];
    if (r < 0) continue;
    EMU_CHECK(r < (int)pool->rows.size(), "row %d beyond the reserved %zu", r, pool->rows.size());
    const int L = pool->len[r];
    EMU_CHECK(L >= 1, "row %d is read but holds no path", r);
    EMU_CHECK(L <= Tpad, "row %d holds %d states, Tpad is %d", r, L, Tpad);
    std::memcpy(&cell[k * Tpad], pool->rows[r].data(), (size_t)L * 4);
    len[k] = L;
  }
  return 0;
}

void fillConflict(const orc_conflict& c, mrp_conflict& o) {
  o.time = c.time;
  o.agent1 = c.agent1;
  o.agent2 = c.agent2;
  o.type = c.type;
  o.x1 = c.x1;
  o.y1 = c.y1;
  o.x2 = c.x2;
  o.y2 = c.y2;
}

}  // namespace

extern "C" {

int mrp_init(int) { return 0; }
int mrp_shutdown(void) { return 0; }
int mrp_set_lane(int lane) {
  EMU_CHECK(lane >= 0 && lane < 64, "lane %d", lane);
  return 0;
}
int mrp_max_lanes(void) { return 64; }
int mrp_device_count(void) { return 1; }
const char* mrp_last_error(void) { return g_err.c_str(); }
const char* mrp_device_info(void) { return "0.0 host emulation of the C ABI (tests only) 0"; }
long long mrp_launch_count(void) { return 0; }

// what the driver did, for the tests: calls of conflicts, conflicts_pool, lowlevel_fs, lowlevel_pool,
// lowlevel_sliced; SUSPENDED answers, resumed jobs, jobs finished
void mrp_emu_counters(long* out9) {
  std::lock_guard<std::mutex> lk(g_cnt.mu);
  for (int k = 0; k < 9; ++k) out9[k] = g_cnt.calls[k];
}

int mrp_map_create(int dimx, int dimy, const int32_t* obst_xy, int n_obst, mrp_map* out) {
  EMU_CHECK(dimx > 0 && dimy > 0 && out, "bad map");
  mrp_map m = new mrp_map_s();
  m->dimx = dimx;
  m->dimy = dimy;
  for (int k = 0; k < n_obst; ++k) {
    const int x = obst_xy[2 * k], y = obst_xy[2 * k + 1];
    if (x < 0 || y < 0 || x >= dimx || y >= dimy) continue;  // ignored like in the reference
    m->obst.push_back(x);
    m->obst.push_back(y);
  }
  *out = m;
  return 0;
}
int mrp_map_destroy(mrp_map m) {
  delete m;
  return 0;
}

int mrp_bfs_fields(int dimx, int dimy, const int32_t* obst_xy, int n_obst, const int32_t* goal_xy,
                   int n_goals, int32_t* out) {
  mrp_map m = nullptr;
  if (int rc = mrp_map_create(dimx, dimy, obst_xy, n_obst, &m)) return rc;
  for (int k = 0; k < n_goals; ++k)
    EMU_CHECK(goal_xy[2 * k] >= 0 && goal_xy[2 * k] < dimx && goal_xy[2 * k + 1] >= 0 && goal_xy[2 * k + 1] < dimy,
              "goal %d outside the map", k);
  orc_bfs_fields(dimx, dimy, m->obst.data(), (int)m->obst.size() / 2, goal_xy, n_goals, out);
  delete m;
  return 0;
}

int mrp_fieldset_create(const mrp_map* maps, int n_maps, const int32_t* goal_map, const int32_t* goal_cell,
                        int n_goals, mrp_fieldset* out) {
  EMU_CHECK(n_maps > 0 && out, "no maps");
  mrp_fieldset fs = new mrp_fieldset_s();
  fs->cells = maps[0]->dimx * maps[0]->dimy;
  fs->data.resize((size_t)n_goals * fs->cells);
  fs->goalCell.assign(goal_cell, goal_cell + n_goals);
  for (int k = 0; k < n_goals; ++k) {
    const mrp_map m = maps[goal_map[k]];
    EMU_CHECK(m->dimx * m->dimy == fs->cells, "maps of one field set must share their dimensions");
    const int32_t g[2] = {goal_cell[k] % m->dimx, goal_cell[k] / m->dimx};
    orc_bfs_fields(m->dimx, m->dimy, m->obst.data(), (int)m->obst.size() / 2, g, 1,
                   fs->data.data() + (size_t)k * fs->cells);
  }
  *out = fs;
  return 0;
}
int mrp_fieldset_read(mrp_fieldset fs, int first, int count, int32_t* out) {
  EMU_CHECK(fs && first >= 0 && count >= 0 && (size_t)(first + count) * fs->cells <= fs->data.size(), "field range");
  std::memcpy(out, fs->data.data() + (size_t)first * fs->cells, (size_t)count * fs->cells * 4);
  return 0;
}
int mrp_fieldset_destroy(mrp_fieldset fs) {
  delete fs;
  return 0;
}

int mrp_first_conflict(const int32_t* cell, const int32_t* len, int N, int Tpad, int dimx, int mode,
                       mrp_conflict* out) {
  orc_conflict c;
  const int found = orc_first_conflict(cell, len, N, Tpad, dimx, mode, &c);
  if (found) fillConflict(c, *out);
  return found;
}
int mrp_count_conflicts(const int32_t* cell, const int32_t* len, int N, int Tpad, int mode, int32_t* count) {
  return orc_count_conflicts(cell, len, N, Tpad, mode, count) < 0 ? MRP_ERR_INVALID : 0;
}
int mrp_conflicts_batch(const int32_t* cell, const int32_t* len, int B, int N, int Tpad, int dimx, int mode,
                        int32_t* found, mrp_conflict* conflicts, int32_t* counts) {
  bump(0);
  {
    Fnv f;
    f.val(B); f.val(N); f.val(dimx); f.val(mode);
    for (size_t k = 0; k < (size_t)B * N; ++k) EMU_CHECK(len[k] >= 0 && len[k] <= Tpad, "len %d, Tpad %d", len[k], Tpad);
    hashTables(f, cell, len, B, N, Tpad);
    bumpHash(f);
  }
  for (int b = 0; b < B; ++b) {
    const int32_t* c = cell + (size_t)b * N * Tpad;
    const int32_t* l = len + (size_t)b * N;
    for (int a = 0; a < N; ++a) EMU_CHECK(l[a] >= 0 && l[a] <= Tpad, "table %d agent %d: len %d, Tpad %d", b, a, l[a], Tpad);
    orc_conflict oc;
    found[b] = orc_first_conflict(c, l, N, Tpad, dimx, mode, &oc);
    if (found[b]) fillConflict(oc, conflicts[b]);
    if (counts) orc_count_conflicts(c, l, N, Tpad, mode, &counts[b]);
  }
  return 0;
}

int mrp_lowlevel_batch_fs(const mrp_map* maps, int n_maps, mrp_fieldset fs, const int32_t* vc, int n_vc,
                          const int32_t* ec, int n_ec, const int32_t* tables, const int32_t* table_len,
                          int n_tables, int N, int Tpad, const mrp_job* jobs, int n_jobs,
                          const mrp_lowlevel_params* prm, mrp_path_info* info, int32_t* out_cells,
                          int32_t* out_g) {
  bump(2);
  bump(7, n_jobs);
  if (int rc = checkJobs(jobs, n_jobs, n_maps, fs ? (int)fs->goalCell.size() : 0, n_vc, n_ec, n_tables, N)) return rc;
  {
    Fnv f;
    f.val(prm->variant); f.val(prm->w); f.val(prm->max_expanded); f.val(n_jobs);
    if (n_tables > 0) hashTables(f, tables, table_len, n_tables, N, Tpad);
    for (int j = 0; j < n_jobs; ++j) hashJob(f, jobs[j], vc, ec);
    bumpHash(f);
  }
  std::vector<int32_t> cells, gs;
  for (int j = 0; j < n_jobs; ++j) {
    if (int rc = runJob(maps, fs, vc, ec, tables, table_len, N, Tpad, jobs[j], *prm, info[j], cells, gs)) return rc;
    if (info[j].status == 0) {
      std::memcpy(out_cells + (size_t)j * prm->path_cap, cells.data(), cells.size() * 4);
      std::memcpy(out_g + (size_t)j * prm->path_cap, gs.data(), gs.size() * 4);
    }
  }
  return 0;
}

int mrp_pathpool_create(int row_cap, mrp_pathpool* out) {
  EMU_CHECK(row_cap > 0 && out, "row_cap");
  *out = new mrp_pathpool_s();
  (*out)->rowCap = row_cap;
  return 0;
}
int mrp_pathpool_destroy(mrp_pathpool p) {
  delete p;
  return 0;
}
int mrp_pathpool_reserve(mrp_pathpool p, int n) {
  std::lock_guard<std::mutex> lk(p->mu);
  while ((int)p->rows.size() < n) {
    p->rows.emplace_back((size_t)p->rowCap, -9);
    p->len.push_back(0);
  }
  return 0;
}
int mrp_pathpool_reserve_states(mrp_pathpool p, int n, int dimx, int dimy, int max_expanded) {
  std::lock_guard<std::mutex> lk(p->mu);
  if (p->blobCap == 0) {
    p->blobDimx = dimx;
    p->blobDimy = dimy;
    p->blobCap = max_expanded;
  }
  // one layout per pool: the same map size class and cap as the first call
  EMU_CHECK((dimx * dimy <= 64) == (p->blobDimx * p->blobDimy <= 64) && max_expanded == p->blobCap,
            "state blobs were laid out for %dx%d / %d expansions", p->blobDimx, p->blobDimy, p->blobCap);
  while ((int)p->blobs.size() < n) p->blobs.emplace_back();
  return 0;
}
int mrp_pathpool_write(mrp_pathpool p, const int32_t* slots, int n, const int32_t* cells, const int32_t* len) {
  for (int k = 0; k < n; ++k) {
    EMU_CHECK(slots[k] >= 0 && slots[k] < (int)p->rows.size() && len[k] >= 0 && len[k] <= p->rowCap, "write row");
    std::memcpy(p->rows[slots[k]].data(), cells + (size_t)k * p->rowCap, (size_t)len[k] * 4);
    p->len[slots[k]] = len[k];
  }
  return 0;
}
int mrp_pathpool_read(mrp_pathpool p, const int32_t* slots, int n, int32_t* cells, int32_t* len) {
  for (int k = 0; k < n; ++k) {
    EMU_CHECK(slots[k] >= 0 && slots[k] < (int)p->rows.size(), "read row %d", slots[k]);
    std::memcpy(cells + (size_t)k * p->rowCap, p->rows[slots[k]].data(), (size_t)p->rowCap * 4);
    len[k] = p->len[slots[k]];
  }
  return 0;
}

int mrp_conflicts_batch_pool(mrp_pathpool pool, const int32_t* table_slots, int B, int N, int Tpad, int dimx,
                             int mode, int32_t* found, mrp_conflict* conflicts, int32_t* counts) {
  bump(1);
  std::vector<int32_t> cell, len;
  if (int rc = gather(pool, table_slots, B, N, Tpad, cell, len)) return rc;
  {
    Fnv f;
    f.val(B); f.val(N); f.val(dimx); f.val(mode);
    hashTables(f, cell.data(), len.data(), B, N, Tpad);
    bumpHash(f);
  }
  for (int b = 0; b < B; ++b) {
    const int32_t* c = cell.data() + (size_t)b * N * Tpad;
    const int32_t* l = len.data() + (size_t)b * N;
    orc_conflict oc;
    found[b] = orc_first_conflict(c, l, N, Tpad, dimx, mode, &oc);
    if (found[b]) fillConflict(oc, conflicts[b]);
    if (counts) orc_count_conflicts(c, l, N, Tpad, mode, &counts[b]);
  }
  return 0;
}

int mrp_lowlevel_batch_pool_sliced(const mrp_map* maps, int n_maps, mrp_fieldset fs, const int32_t* vc, int n_vc,
                                   const int32_t* ec, int n_ec, mrp_pathpool pool, const int32_t* table_slots,
                                   int n_tables, int N, int Tpad, const mrp_job* jobs, int n_jobs,
                                   const mrp_lowlevel_params* prm, const int32_t* out_slots,
                                   const int32_t* state_ids, const int32_t* resume, int slice_expanded,
                                   mrp_path_info* info) {
  bump(state_ids ? 4 : 3);
  EMU_CHECK(prm->variant == 0, "pool rows hold cells only: variant 0");
  EMU_CHECK(prm->path_cap <= pool->rowCap, "path_cap %d exceeds the rows (%d)", prm->path_cap, pool->rowCap);
  if (int rc = checkJobs(jobs, n_jobs, n_maps, fs ? (int)fs->goalCell.size() : 0, n_vc, n_ec, n_tables, N)) return rc;
  std::vector<int32_t> cell, len;
  if (int rc = gather(pool, table_slots, n_tables, N, Tpad, cell, len)) return rc;
  {
    Fnv f;
    f.val(prm->variant); f.val(prm->w); f.val(prm->max_expanded); f.val(n_jobs); f.val(slice_expanded);
    hashTables(f, cell.data(), len.data(), n_tables, N, Tpad);
    for (int j = 0; j < n_jobs; ++j) {
      hashJob(f, jobs[j], vc, ec);
      const int r = resume ? resume[j] : 0;
      f.val(r);
    }
    bumpHash(f);
  }
  std::vector<char> blobSeen(pool->blobs.size(), 0);
  std::vector<int32_t> cells, gs;
  for (int j = 0; j < n_jobs; ++j) {
    EMU_CHECK(out_slots[j] >= 0 && out_slots[j] < (int)pool->rows.size(), "job %d: output row %d beyond the reserved %zu",
              j, out_slots[j], pool->rows.size());
    const int sid = (state_ids && slice_expanded > 0) ? state_ids[j] : -1;
    mrp_path_info res;
    if (sid >= 0) {
      EMU_CHECK(sid < (int)pool->blobs.size(), "job %d: state %d beyond the reserved %zu", j, sid, pool->blobs.size());
      EMU_CHECK(!blobSeen[sid], "state %d is used by two jobs of one call", sid);
      blobSeen[sid] = 1;
      mrp_pathpool_s::Blob& b = pool->blobs[sid];
      if (resume && resume[j]) {
        EMU_CHECK(b.live, "job %d resumes state %d, which holds no suspended search", j, sid);
        EMU_CHECK(std::memcmp(&b.job.start_cell, &jobs[j].start_cell, 8) == 0 && b.job.self == jobs[j].self,
                  "job %d resumes state %d with another job", j, sid);
        bump(6);
      } else {
        if (int rc = runJob(maps, fs, vc, ec, cell.data(), len.data(), N, Tpad, jobs[j], *prm, b.info, b.cells, gs))
          return rc;
        b.live = true;
        b.given = 0;
        b.job = jobs[j];
      }
      b.given += slice_expanded;
      if (b.given < b.info.expanded) {
        info[j] = b.info;
        info[j].status = MRP_SUSPENDED;
        info[j].cost = info[j].fmin = info[j].length = 0;
        info[j].expanded = 0;  // reported once, with the final answer
        bump(5);
        continue;
      }
      res = b.info;
      cells = b.cells;
      b.live = false;
    } else {
      if (int rc = runJob(maps, fs, vc, ec, cell.data(), len.data(), N, Tpad, jobs[j], *prm, res, cells, gs)) return rc;
    }
    bump(7);
    info[j] = res;
    if (res.status == 0) {  // rows of failed jobs keep their old content
      std::memcpy(pool->rows[out_slots[j]].data(), cells.data(), cells.size() * 4);
      pool->len[out_slots[j]] = (int32_t)cells.size();
    }
  }
  return 0;
}

int mrp_lowlevel_batch_pool(const mrp_map* maps, int n_maps, mrp_fieldset fs, const int32_t* vc, int n_vc,
                            const int32_t* ec, int n_ec, mrp_pathpool pool, const int32_t* table_slots,
                            int n_tables, int N, int Tpad, const mrp_job* jobs, int n_jobs,
                            const mrp_lowlevel_params* prm, const int32_t* out_slots, mrp_path_info* info) {
  return mrp_lowlevel_batch_pool_sliced(maps, n_maps, fs, vc, n_vc, ec, n_ec, pool, table_slots, n_tables, N, Tpad,
                                        jobs, n_jobs, prm, out_slots, nullptr, nullptr, 0, info);
}

}  // extern "C"
