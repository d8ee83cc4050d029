// hl_search.hpp — batched high-level searches on top of the CUDA hot path.
//
// The high-level loops of the reference stay on the host, as sequential
// best-first searches per instance:
//   CBS::search    include/libMultiRobotPlanning/cbs.hpp:85-172
//   ECBS::search   include/libMultiRobotPlanning/ecbs.hpp:109-288
//   CBSTA::search  include/libMultiRobotPlanning/cbs_ta.hpp:87-214
// What changes is the granularity of their callees.  Many instances advance in
// lock-step; per step ALL constraint-tree nodes that were just created are
// checked for conflicts in one launch (mrp_conflicts_batch: getFirstConflict +
// focalHeuristic), and ALL low-level replans they trigger run in one launch
// (mrp_lowlevel_batch_fs: AStar / AStarEpsilon).  The per-goal distance fields
// are computed once per batch and stay resident in HBM (mrp_fieldset).
// cbs / ecbs batches keep the paths of all constraint-tree nodes in a device pool
// (mrp_pathpool_*) and advance per instance: an expansion is a "flight" whose
// replans run in slices (mrp_lowlevel_batch_pool_sliced), so a launch never lasts
// as long as the slowest search.  A node is one flat block of its instance's
// slab: pool rows, heads of parent-pointer constraint chains, a shared task
// vector (see Node below).  The host logic is tested without a device on an
// emulation of the C ABI (tests/emu, tests/test_host_driver_emu.py).
//
// Deviations from the reference, all result-preserving:
//   * cbs/ecbs use the exact BFS distance field instead of the Manhattan
//     distance as admissible heuristic: same optimal costs.
//   * ECBS admits a node to FOCAL iff cost <= w * min LB over OPEN (LB = sum of
//     fmin), which implies the reference's rule cost <= w * min cost
//     (ecbs.hpp:171-189) and additionally guarantees cost <= w * optimum.
//   * expansion / wall-clock caps (the reference has none and may not return).
//   * ties among equal-cost nodes break by node id (reference: Boost.Heap
//     internals, not pinned by any of its tests).
#pragma once

#include <algorithm>
#include <array>
#include <chrono>
#include <mutex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/mrp_b200.h"
#include "assignment.hpp"

namespace mrp_host {

struct MapfInstance {
  int dimx = 0, dimy = 0;
  std::vector<int32_t> obstXY;                     // x0, y0, x1, y1, ...
  std::vector<int> starts;                         // cell = x + dimx*y
  std::vector<int> goals;                          // cbs / ecbs
  std::vector<std::vector<int> > potentialGoals;   // cbs_ta (lists may be empty)
  size_t numAgents() const { return starts.size(); }
};

struct AgentPath {
  std::vector<int> cells;
  std::vector<int> g;  // g-score per state: the "t" of output.yaml (cbs.cpp:659)
  int cost = 0;
  int fmin = 0;
  // pool mode (cbs / ecbs batches): the cells live in row `slot` of the device path
  // pool, `length` states; cells / g are filled from the pool for the final solution only
  int slot = -1;
  int length = 0;
};

enum SolveStatus { kSolved = 0, kNoSolution = 1, kCapped = 2 };

struct SolveResult {
  int status = kNoSolution;
  long cost = 0, makespan = 0, lowerBound = 0;
  long hlExpanded = 0, llExpanded = 0, numTaskAssignments = 0;
  double runtime = 0;
  std::vector<AgentPath> paths;
};

struct SolveOptions {
  float w = 1.0f;
  long maxHlExpanded = 0;  // per instance; 0 = unlimited
  // per low-level search (GPU workspace is sized by it).  12000: of the 22 instances of the
  // 1000 x 100-agent ECBS batch that a cap of 8000 loses, 20 need less than this (the
  // unmodified reference solves 21 of them; with 24000 one instance runs to the high-level
  // cap for 75 s: tools/try_unsolved.py)
  int maxLlExpanded = 12000;
  // low-level expansions of ONE instance in total (0 = unlimited): an instance whose replans
  // have used more than this is given up as capped.  The reference has no such limit (and does
  // not return on those instances either); in a batch one runaway instance would otherwise keep
  // every GPU lane waiting: with a replan cap of 12000 one of 4000 ECBS instances ran to the
  // high-level cap with 6*10^7 expansions (27 s) where the whole batch takes 1 s.
  long maxLlTotal = 0;
  double maxSeconds = 0;     // whole batch; 0 = unlimited
  long maxTaskAssignments = 1000000000L;
};

enum class Algo { CBS = 0, ECBS = 1, CBSTA = 2, ECBSTA = 3 };

// wall-clock split of a batch run, printed when MRP_HOST_PROFILE is set
struct HostProfile {
  double gpuConflicts = 0, gpuLowLevel = 0, total = 0, setup = 0;
  double pop = 0, build = 0, llPack = 0, llUnpack = 0, evalPack = 0, absorb = 0;
  long iterations = 0, nodes = 0, jobs = 0, launches = 0;
};
inline double nowSeconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

inline void gpuCheck(int rc) {
  if (rc < 0) throw std::runtime_error(std::string("mrp_b200: ") + mrp_last_error());
}

class BatchSolver {
 public:
  BatchSolver(Algo algo, const std::vector<MapfInstance>& instances, const SolveOptions& opt)
      : m_algo(algo), m_opt(opt) {
    if (instances.empty()) return;
    m_dimx = instances[0].dimx;
    m_dimy = instances[0].dimy;
    m_cells = m_dimx * m_dimy;
    m_pathCap = std::max(128, 4 * (m_dimx + m_dimy));
    m_inst.resize(instances.size());
    for (size_t k = 0; k < instances.size(); ++k) {
      if (instances[k].dimx != m_dimx || instances[k].dimy != m_dimy)
        throw std::runtime_error("all instances of one batch must share their map dimensions");
      m_inst[k].in = &instances[k];
    }
  }
  ~BatchSolver() {
    m_flights.clear();
    for (Inst& I : m_inst) {
      for (Node* n : I.open) freeNode(n);
      I.open.clear();
      for (const auto& h : I.heap) freeNode(h.node);
      I.heap.clear();
      I.solution.clear();
    }
    if (m_pool) poolCache(m_pathCap, m_dimx, m_dimy, m_opt.maxLlExpanded, m_pool);
    if (m_fields) mrp_fieldset_destroy(m_fields);
    for (mrp_map m : m_maps) mrp_map_destroy(m);
  }
  BatchSolver(const BatchSolver&) = delete;
  BatchSolver& operator=(const BatchSolver&) = delete;

  std::vector<SolveResult> run() {
    std::vector<SolveResult> out(m_inst.size());
    if (m_inst.empty()) return out;
    const double tSetup = nowSeconds();
    setup();
    m_prof.setup = nowSeconds() - tSetup;
    const auto t0 = std::chrono::steady_clock::now();
    auto elapsed = [&t0]() {
      return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    std::vector<Node*> fresh;
    const double tRun = nowSeconds();
    // Sliced replans (pool mode, single-tile maps): every launch gives each unfinished
    // replan at most m_slice expansions; an instance whose replans are done moves on to
    // its next expansion while the long searches of the others continue in the next
    // launches (MRP_HOST_SLICE=0: every launch runs its replans to the end).
    if (const char* e = getenv("MRP_HOST_SLICE")) m_slice = atoi(e);
    m_sliced = m_pool && m_slice > 0 && m_dimx <= 32 && m_dimy <= 32;
    buildRoots(fresh);
    const bool sliced = m_sliced;
    while (true) {
      ++m_prof.iterations;
      evaluate(fresh);
      insertFresh(fresh);
      fresh.clear();
      if (m_algo == Algo::ECBSTA) spawnMinRoots();
      // one expansion per running instance
      std::vector<Pending> pending;
      bool anyRunning = false;
      const double tPop = nowSeconds();
      const bool timeUp = m_opt.maxSeconds > 0 && elapsed() > m_opt.maxSeconds;
      const double tNow = elapsed();
      std::vector<NodeUP> popped(m_inst.size());
      // instances are independent: the per-instance bookkeeping of a lock-step
      // iteration runs on all host cores
#pragma omp parallel for schedule(dynamic, 16) if (m_inst.size() >= kParallelMin)
      for (long k = 0; k < (long)m_inst.size(); ++k) {
        Inst& I = m_inst[k];
        if (I.done || I.busy) continue;
        if (I.openEmpty()) {
          finish(I, kNoSolution, nullptr, tNow);
          continue;
        }
        if (timeUp || (m_opt.maxHlExpanded > 0 && I.res.hlExpanded >= m_opt.maxHlExpanded) ||
            (m_opt.maxLlTotal > 0 && I.res.llExpanded >= m_opt.maxLlTotal)) {
          finish(I, kCapped, nullptr, tNow);
          continue;
        }
        NodeUP P = popBest(I);
        ++I.res.hlExpanded;  // onExpandHighLevelNode, cbs.hpp:121
        if (!P->found) {
          finish(I, kSolved, P.get(), tNow);
          continue;
        }
        popped[k] = std::move(P);
      }
      for (size_t k = 0; k < m_inst.size(); ++k) {
        if (!popped[k]) continue;
        anyRunning = true;
        Pending pd;
        pd.inst = (int)k;
        pd.parent = std::move(popped[k]);
        pending.push_back(std::move(pd));
      }
      fetchSolutions();
      m_prof.pop += nowSeconds() - tPop;
      if (sliced) {
        if (timeUp) {
          for (auto& f : m_flights) {
            finish(m_inst[f->pd.inst], kCapped, nullptr, tNow);
            dropFlight(*f);
          }
          m_flights.clear();
        }
        startFlights(pending);
        if (m_flights.empty()) break;
        runFlights();
        absorbFlights(fresh);
        continue;
      }
      if (!anyRunning) break;
      expand(pending, fresh, m_opt.maxLlExpanded, m_mainBuf, m_prof);
      // the expanded parents are released on all cores as well
#pragma omp parallel for schedule(dynamic, 16) if (pending.size() >= kParallelMin)
      for (long pi = 0; pi < (long)pending.size(); ++pi) pending[pi].parent.reset();
    }
    fetchSolutions();
    for (size_t k = 0; k < m_inst.size(); ++k) out[k] = std::move(m_inst[k].res);
    m_prof.total = nowSeconds() - tRun;
    if (getenv("MRP_HOST_PROFILE"))
      fprintf(stderr,
              "[mrp_host] %zu instances, %ld lock-step iterations, %ld nodes, %ld replans in %ld sliced launches: "
              "total %.3fs = conflicts(gpu call) %.3fs + replans(gpu call) %.3fs + host %.3fs "
              "[pop %.2f build %.2f llPack %.2f llUnpack %.2f evalPack %.2f absorb %.2f]; "
              "setup (maps + distance fields, outside the reference's timer too) %.3fs\n",
              m_inst.size(), m_prof.iterations, m_prof.nodes, m_prof.jobs, m_prof.launches, m_prof.total,
              m_prof.gpuConflicts, m_prof.gpuLowLevel,
              m_prof.total - m_prof.gpuConflicts - m_prof.gpuLowLevel, m_prof.pop, m_prof.build,
              m_prof.llPack, m_prof.llUnpack, m_prof.evalPack, m_prof.absorb, m_prof.setup);
    return out;
  }

 private:
  // A constraint-tree node does not own its paths or constraints (the reference deep-copies the
  // whole node for every child, cbs.hpp:144,175-181; sharing is invisible in the results):
  //  * a path is a PathRef: cost, fmin, length and either a row of the device pool (cbs / ecbs:
  //    the cells never come to the host) or a shared host copy (cbs_ta / ecbs_ta, large maps).
  //    Pool rows belong to the INSTANCE, not to the node: they are handed back in one piece when
  //    the instance finishes, so copying a node is a flat copy without reference counts;
  //  * the constraints of an agent are a parent-pointer chain in the instance's arena
  //    (Inst::consArena): a node keeps the index of the newest entry per agent, a child adds one
  //    entry whose `prev` is its parent's index (cbs.hpp:146-153 adds the new constraint to a
  //    copy of the whole set);
  //  * the task vector is shared by every node below one root.
  struct PathRef {
    std::shared_ptr<const AgentPath> host;  // cells / g-scores on the host (not in pool mode)
    int32_t slot = -1, length = 0, cost = 0, fmin = 0;
    bool set = false;  // false: this agent has no path yet (roots under construction)
    explicit operator bool() const { return set; }
  };
  struct ConsEntry {
    int32_t time, a, b;  // b < 0: vertex constraint (time, cell a); else edge (time, from a, to b)
    int32_t prev;        // older constraint of the same agent on the way to the root, -1: none
  };
  typedef std::shared_ptr<const std::vector<int> > TaskPtr;
  // A node and its two per-agent arrays are ONE block of its instance's pool (Inst::nodePool:
  // all nodes of an instance have the same size, an instance is only ever touched by one thread
  // at a time, so creating / dropping a node costs no malloc, no lock and no arena growth; the
  // slabs go back in one piece when the instance is done).  Nodes are made by newNode /
  // cloneNode and owned through NodeUP or, inside OPEN / `fresh`, as plain pointers that end in
  // freeNode.
  struct Node {
    int inst = 0, nAgents = 0;
    PathRef* paths = nullptr;   // [nAgents], behind the node in its block
    int32_t* cons = nullptr;    // [nAgents]: newest entry of the agent's chain in Inst::consArena, -1: none
    TaskPtr task;               // goal cell per agent (-1: none)
    long cost = 0, LB = 0;
    int focal = 0;
    int id = 0;
    bool isRoot = false;
    int found = 0;
    mrp_conflict conflict;
  };
  struct BlockPool {  // equal-sized blocks cut from slabs; not thread-safe (one instance, one thread)
    size_t blockSize = 0, left = 0;
    char* cur = nullptr;
    void* freeList = nullptr;
    std::vector<char*> slabs;
    BlockPool() {}
    BlockPool(const BlockPool&) = delete;
    BlockPool& operator=(const BlockPool&) = delete;
    BlockPool(BlockPool&& o) noexcept
        : blockSize(o.blockSize), left(o.left), cur(o.cur), freeList(o.freeList), slabs(std::move(o.slabs)) {
      o.left = 0;
      o.cur = nullptr;
      o.freeList = nullptr;
    }
    ~BlockPool() { release(); }
    // -DMRP_HOST_DEBUG_MALLOC: one malloc per block, so that AddressSanitizer sees a node that is
    // used after it was dropped (tests/test_host_driver_emu.py builds that variant when asked to)
#ifdef MRP_HOST_DEBUG_MALLOC
    void* alloc() { return std::malloc(blockSize); }
    void free(void* p) { std::free(p); }
    void release() {}
#else
    void* alloc() {
      if (freeList) {
        void* p = freeList;
        freeList = *(void**)p;
        return p;
      }
      if (left < blockSize) {
        const size_t n = std::max<size_t>(blockSize * 16, (size_t)64 << 10);
        cur = (char*)std::malloc(n);
        if (!cur) throw std::bad_alloc();
        slabs.push_back(cur);
        left = n;
      }
      void* p = cur;
      cur += blockSize;
      left -= blockSize;
      return p;
    }
    void free(void* p) {
      *(void**)p = freeList;
      freeList = p;
    }
    void release() {
      for (char* sl : slabs) std::free(sl);
      slabs.clear();
      left = 0;
      cur = nullptr;
      freeList = nullptr;
    }
#endif
  };
  struct NodeDeleter {
    BatchSolver* solver;
    NodeDeleter() : solver(nullptr) {}
    explicit NodeDeleter(BatchSolver* s) : solver(s) {}
    void operator()(Node* n) const { solver->freeNode(n); }
  };
  typedef std::unique_ptr<Node, NodeDeleter> NodeUP;
  NodeUP own(Node* n) { return NodeUP(n, NodeDeleter(this)); }
  static size_t nodeHeaderBytes() { return (sizeof(Node) + 15) & ~(size_t)15; }
  // an empty node of instance k: no paths, no constraints
  Node* newNode(int k) {
    Inst& I = m_inst[k];
    const int N = (int)I.in->numAgents();
    if (!I.nodePool.blockSize)
      I.nodePool.blockSize = (nodeHeaderBytes() + (size_t)N * (sizeof(PathRef) + sizeof(int32_t)) + 15) & ~(size_t)15;
    char* b = (char*)I.nodePool.alloc();
    Node* n = new (b) Node();
    n->inst = k;
    n->nAgents = N;
    n->paths = reinterpret_cast<PathRef*>(b + nodeHeaderBytes());
    n->cons = reinterpret_cast<int32_t*>(b + nodeHeaderBytes() + (size_t)N * sizeof(PathRef));
    for (int a = 0; a < N; ++a) {
      new (&n->paths[a]) PathRef();
      n->cons[a] = -1;
    }
    return n;
  }
  // HighLevelNode newNode = P (cbs.hpp:144): a flat copy of path references and chain heads
  Node* cloneNode(const Node& P) {
    Inst& I = m_inst[P.inst];
    char* b = (char*)I.nodePool.alloc();
    Node* n = new (b) Node(P);
    n->paths = reinterpret_cast<PathRef*>(b + nodeHeaderBytes());
    n->cons = reinterpret_cast<int32_t*>(b + nodeHeaderBytes() + (size_t)P.nAgents * sizeof(PathRef));
    for (int a = 0; a < P.nAgents; ++a) new (&n->paths[a]) PathRef(P.paths[a]);
    std::memcpy(n->cons, P.cons, (size_t)P.nAgents * sizeof(int32_t));
    return n;
  }
  void freeNode(Node* n) {
    if (!n) return;
    BlockPool& pool = m_inst[n->inst].nodePool;
    for (int a = 0; a < n->nAgents; ++a) n->paths[a].~PathRef();
    n->~Node();
    pool.free(n);
  }
  // batches smaller than this stay on the calling thread (a single instance from
  // the command-line binaries must not wake a thread team per iteration)
  static constexpr size_t kParallelMin = 64;
  bool isTA() const { return m_algo == Algo::CBSTA || m_algo == Algo::ECBSTA; }
  bool isFocal() const { return m_algo == Algo::ECBS || m_algo == Algo::ECBSTA; }

  // lowest cost first (cbs.hpp:187-191).  Among equal costs the reference's
  // order is whatever Boost's heap does; here: fewer conflicts first (the count
  // is computed anyway), then lowest id.  Any order among equal costs keeps
  // CBS optimal; this one reaches a conflict-free node sooner.
  struct OpenOrder {
    bool operator()(const Node* a, const Node* b) const {
      return std::make_tuple(a->cost, a->focal, a->id) <
             std::make_tuple(b->cost, b->focal, b->id);
    }
  };
  struct Inst {
    const MapfInstance* in = nullptr;
    std::set<Node*, OpenOrder> open;  // owning; ecbs / ecbs_ta (FOCAL walks it in cost order)
    // cbs / cbs_ta only ever take the best node: a binary heap of (cost, focal, id, node) with the
    // keys inline — the same total order as OpenOrder (ids are unique), no tree node per entry
    struct HeapItem {
      long cost;
      int focal, id;
      Node* node;
      bool operator<(const HeapItem& o) const {  // std::*_heap keep the LARGEST on top: invert
        return std::make_tuple(cost, focal, id) > std::make_tuple(o.cost, o.focal, o.id);
      }
    };
    std::vector<HeapItem> heap;  // owning
    bool openEmpty() const { return open.empty() && heap.empty(); }
    int nextId = 0;
    bool done = false;
    bool busy = false;       // sliced replans of its current expansion are still running
    int mapIdx = 0;
    int fieldBase = 0;                 // first field of this instance in the set
    std::map<int, int> fieldOfGoal;    // goal cell -> field index (cbs_ta)
    std::unique_ptr<NextBestAssignment<int, int> > assignment;
    long nextRootNodeCost = 0;  // ecbs_ta (STYLE_MINROOT, ecbs_ta.hpp:142,345)
    SolveResult res;
    std::vector<PathRef> solution;  // pool mode: the rows of the solution until they are fetched
    std::vector<ConsEntry> consArena;  // constraint chains of all nodes of this instance
    std::vector<int32_t> rows;         // pool rows this instance holds (returned when it is done)
    bool rowsQueued = false;
    BlockPool nodePool;                // the blocks of this instance's nodes
  };
  struct Pending {
    int inst = 0;
    NodeUP parent;
  };
  struct JobSpec {
    int inst, agent, goal, field;
    int cons;  // newest entry of the agent's constraint chain (Inst::consArena), -1: none
    int table, self;
  };
  struct JobOut {
    int status = 1;
    long expanded = 0;
    PathRef path;  // set iff status == 0
  };

  // Pools outlive their batch: creating one costs a dozen device allocations (tens of
  // milliseconds), which a single small instance would pay on every call.  give != NULL
  // stores a pool, give == NULL takes one of that shape (or returns NULL).  The rows and
  // state blobs of a cached pool hold garbage; the next batch numbers them from 0 again.
  static mrp_pathpool poolCache(int rowCap, int dimx, int dimy, int maxLl, mrp_pathpool give) {
    struct Entry {
      int rowCap, tb, maxLl;
      mrp_pathpool pool;
    };
    static std::mutex mu;
    static std::vector<Entry> cache;
    const int tb = dimx * dimy <= 64 ? 0 : 1;  // layout class of the state blobs
    std::lock_guard<std::mutex> lk(mu);
    if (give) {
      // keep the most recently used ones: a process that alternates between batch shapes
      // (bench.py: 4 lanes of 100-agent ECBS, then 16 lanes of 8x8 CBS) would otherwise
      // create and destroy sixteen pools per call (a dozen device allocations each)
      if (cache.size() >= 40) {
        mrp_pathpool_destroy(cache.front().pool);
        cache.erase(cache.begin());
      }
      cache.push_back({rowCap, tb, maxLl, give});
      return nullptr;
    }
    for (size_t i = 0; i < cache.size(); ++i)
      if (cache[i].rowCap == rowCap && cache[i].tb == tb && cache[i].maxLl == maxLl) {
        mrp_pathpool p = cache[i].pool;
        cache.erase(cache.begin() + i);
        return p;
      }
    return nullptr;
  }

  // ---- device path pool (cbs / ecbs): rows are handed out and taken back here ----
  bool poolWanted() const {
    if (!(m_algo == Algo::CBS || m_algo == Algo::ECBS)) return false;
    const char* e = getenv("MRP_HOST_POOL");
    return !(e && atoi(e) == 0);
  }
  void takeRows(size_t n, std::vector<int32_t>& rows) {
    rows.resize(n);
    std::lock_guard<std::mutex> lk(m_rowMutex);
    for (size_t i = 0; i < n; ++i) {
      if (!m_freeRows.empty()) {
        rows[i] = m_freeRows.back();
        m_freeRows.pop_back();
      } else {
        rows[i] = m_nextRow++;
      }
    }
  }
  void giveRow(int row) {
    std::lock_guard<std::mutex> lk(m_rowMutex);
    m_freeRows.push_back(row);
  }
  // A finished replan becomes a PathRef.  In pool mode its row joins the rows of its instance
  // (main thread only: the launches are absorbed there) and stays until the instance is done.
  PathRef adopt(AgentPath&& p, int inst) {
    PathRef r;
    r.slot = p.slot;
    r.length = p.length;
    r.cost = p.cost;
    r.fmin = p.fmin;
    r.set = true;
    if (p.slot < 0)
      r.host = std::make_shared<const AgentPath>(std::move(p));
    else
      m_inst[inst].rows.push_back(p.slot);
    return r;
  }
  // the rows of the instances that finished go back to the free list (main thread, after their
  // solutions were read: fetchSolutions)
  void releaseFinished() {
    std::vector<int> done;
    {
      std::lock_guard<std::mutex> lk(m_rowMutex);
      done.swap(m_doneQueue);
    }
    for (int k : done) {
      Inst& I = m_inst[k];
      {
        std::lock_guard<std::mutex> lk(m_rowMutex);
        m_freeRows.insert(m_freeRows.end(), I.rows.begin(), I.rows.end());
      }
      std::vector<int32_t>().swap(I.rows);
      std::vector<ConsEntry>().swap(I.consArena);
      I.nodePool.release();  // every node of a finished instance is gone by now (see run())
    }
  }
  // the constraints of one chain, oldest first, appended to the (time, cell) pairs / (time, from,
  // to) triples of a launch
  void appendCons(const Inst& I, int head, std::vector<int32_t>& vc, std::vector<int32_t>& ec) const {
    const size_t v0 = vc.size(), e0 = ec.size();
    for (int k = head; k >= 0; k = I.consArena[k].prev) {
      const ConsEntry& c = I.consArena[k];
      if (c.b < 0) {  // pushed back to front, reversed below
        vc.push_back(c.a);
        vc.push_back(c.time);
      } else {
        ec.push_back(c.b);
        ec.push_back(c.a);
        ec.push_back(c.time);
      }
    }
    std::reverse(vc.begin() + v0, vc.end());
    std::reverse(ec.begin() + e0, ec.end());
  }
  // createConstraintsFromConflict (example/cbs.cpp:388-406) for one side of a conflict: the new
  // entry of `agent`'s chain below node P
  int addConstraint(Inst& I, const Node& P, const mrp_conflict& c, int side, int agent) const {
    const int c1 = c.x1 + m_dimx * c.y1;
    const int c2 = c.type == 1 ? c.x2 + m_dimx * c.y2 : -1;
    ConsEntry e;
    e.time = c.time;
    if (c.type == 0) {
      e.a = c1;
      e.b = -1;
    } else {
      e.a = side == 0 ? c1 : c2;
      e.b = side == 0 ? c2 : c1;
    }
    e.prev = P.cons[agent];
    I.consArena.push_back(e);
    return (int)I.consArena.size() - 1;
  }
  // pool mode: the paths of the instances that finished since the last call come to
  // the host (one read for all of them; main thread only)
  void fetchSolutions() {
    if (!m_pool) {
      releaseFinished();
      return;
    }
    std::vector<int32_t> rows;
    for (size_t k = 0; k < m_inst.size(); ++k)
      for (const PathRef& p : m_inst[k].solution) rows.push_back(p.slot);
    if (rows.empty()) {
      releaseFinished();
      return;
    }
    std::vector<int32_t> cells(rows.size() * (size_t)m_pathCap), len(rows.size());
    gpuCheck(mrp_pathpool_read(m_pool, rows.data(), (int)rows.size(), cells.data(), len.data()));
    size_t r = 0;
    for (size_t k = 0; k < m_inst.size(); ++k) {
      Inst& I = m_inst[k];
      for (const PathRef& p : I.solution) {
        AgentPath out;
        out.cost = p.cost;
        out.fmin = p.fmin;
        out.length = len[r];
        out.cells.assign(cells.begin() + r * m_pathCap, cells.begin() + r * m_pathCap + len[r]);
        out.g.resize(len[r]);
        for (int t = 0; t < len[r]; ++t) out.g[t] = t;  // cbs / ecbs moves: g-score = time step
        I.res.paths.push_back(std::move(out));
        ++r;
      }
      I.solution.clear();
    }
    releaseFinished();
  }

  // ---------------------------------------------------------------------
  void setup() {
    std::vector<int32_t> goalMap, goalCell;
    for (size_t k = 0; k < m_inst.size(); ++k) {
      Inst& I = m_inst[k];
      const MapfInstance& in = *I.in;
      mrp_map m = nullptr;
      gpuCheck(mrp_map_create(in.dimx, in.dimy, in.obstXY.data(), (int)in.obstXY.size() / 2, &m));
      I.mapIdx = (int)m_maps.size();
      m_maps.push_back(m);
      I.fieldBase = (int)goalCell.size();
      if (isTA()) {
        for (const auto& pg : in.potentialGoals)
          for (int g : pg)
            if (!I.fieldOfGoal.count(g)) {
              I.fieldOfGoal[g] = (int)goalCell.size();
              goalMap.push_back(I.mapIdx);
              goalCell.push_back(g);
            }
      } else {
        for (int g : in.goals) {
          goalMap.push_back(I.mapIdx);
          goalCell.push_back(g);
        }
      }
    }
    if (poolWanted()) {
      // device allocations are made here, outside the search timer (like the heuristic
      // precompute below); a pool left by an earlier batch of the same shape is reused
      m_pool = poolCache(m_pathCap, m_dimx, m_dimy, m_opt.maxLlExpanded, nullptr);
      if (!m_pool) gpuCheck(mrp_pathpool_create(m_pathCap, &m_pool));
      gpuCheck(mrp_pathpool_reserve(m_pool, 1));
      if (m_dimx <= 32 && m_dimy <= 32)
        gpuCheck(mrp_pathpool_reserve_states(m_pool, 1, m_dimx, m_dimy, m_opt.maxLlExpanded));
    }
    // the heuristic precompute sits outside the reference's timer as well
    // (Environment ctor, example/cbs_ta.cpp:254-281,570-578)
    gpuCheck(mrp_fieldset_create(m_maps.data(), (int)m_maps.size(), goalMap.data(),
                                 goalCell.data(), (int)goalCell.size(), &m_fields));
    if (isTA()) {
      for (Inst& I : m_inst) {
        const MapfInstance& in = *I.in;
        const int nf = (int)I.fieldOfGoal.size();
        std::vector<int32_t> f((size_t)nf * m_cells);
        if (nf) gpuCheck(mrp_fieldset_read(m_fields, I.fieldBase, nf, f.data()));
        I.assignment.reset(new NextBestAssignment<int, int>());
        for (size_t a = 0; a < in.numAgents(); ++a)
          for (int g : in.potentialGoals[a]) {
            const int fi = I.fieldOfGoal[g] - I.fieldBase;
            I.assignment->setCost((int)a, g, f[(size_t)fi * m_cells + in.starts[a]]);
          }
        I.assignment->solve();
      }
    }
  }

  // nextTaskAssignment, example/cbs_ta.cpp:442-456
  bool nextTasks(Inst& I, TaskPtr& out) {
    std::shared_ptr<std::vector<int> > task = std::make_shared<std::vector<int> >(I.in->numAgents(), -1);
    out = task;
    if ((unsigned long)I.res.numTaskAssignments > (unsigned long)m_opt.maxTaskAssignments)
      return false;
    std::map<int, int> sol;
    I.assignment->nextSolution(sol);
    if (sol.empty()) return false;
    ++I.res.numTaskAssignments;
    for (const auto& e : sol) (*task)[e.first] = e.second;
    return true;
  }

  int fieldFor(const Inst& I, int agent, int goalCell) const {
    if (goalCell < 0) return -1;
    if (isTA()) return I.fieldOfGoal.at(goalCell);
    return I.fieldBase + agent;
  }

  void finish(Inst& I, int status, const Node* n, double t) {
    I.done = true;
    I.res.status = status;
    I.res.runtime = t;
    if (n) {
      for (int a = 0; a < n->nAgents; ++a) {  // example/cbs.cpp:630-635
        const PathRef& p = n->paths[a];
        if (p.slot >= 0)
          I.solution.push_back(p);  // cells follow in fetchSolutions()
        else
          I.res.paths.push_back(*p.host);
        I.res.cost += p.cost;
        I.res.makespan = std::max<long>(I.res.makespan, p.cost);
        I.res.lowerBound += p.fmin;
      }
    }
    for (Node* o : I.open) freeNode(o);
    I.open.clear();
    for (const auto& h : I.heap) freeNode(h.node);
    I.heap.clear();
    if (!I.rowsQueued) {  // its rows and node slabs go back once the solution has been read
      I.rowsQueued = true;
      std::lock_guard<std::mutex> lk(m_rowMutex);
      m_doneQueue.push_back((int)(&I - m_inst.data()));
    }
  }

  // pushes the evaluated nodes into the OPEN lists of their instances; `fresh`
  // holds the nodes of one instance in consecutive positions
  void insertFresh(const std::vector<Node*>& fresh) {
    std::vector<std::pair<size_t, size_t> > runs;
    for (size_t i = 0; i < fresh.size();) {
      size_t j = i + 1;
      while (j < fresh.size() && fresh[j]->inst == fresh[i]->inst) ++j;
      runs.push_back(std::make_pair(i, j));
      i = j;
    }
#pragma omp parallel for schedule(dynamic, 16) if (runs.size() >= kParallelMin)
    for (long r = 0; r < (long)runs.size(); ++r)
      for (size_t i = runs[r].first; i < runs[r].second; ++i) openInsert(m_inst[fresh[i]->inst], fresh[i]);
  }

  void openInsert(Inst& I, Node* n) {
    if (isFocal()) {
      I.open.insert(n);
      return;
    }
    I.heap.push_back({n->cost, n->focal, n->id, n});
    std::push_heap(I.heap.begin(), I.heap.end());
  }
  NodeUP popBest(Inst& I) {
    if (!isFocal()) {  // lowest (cost, focal, id): cbs.hpp:187-191 orders by cost only
      std::pop_heap(I.heap.begin(), I.heap.end());
      NodeUP n = own(I.heap.back().node);
      I.heap.pop_back();
      return n;
    }
    auto best = I.open.begin();
    if (m_algo == Algo::ECBSTA) {
      // FOCAL rebuilt every iteration: cost <= nextRootNodeCost, best by
      // (focalHeuristic, cost) — ecbs_ta.hpp:160-181,420-428; ties by id
      bool have = false;
      for (auto it = I.open.begin(); it != I.open.end(); ++it) {
        const Node& n = **it;
        if (!((float)n.cost <= (float)I.nextRootNodeCost)) break;  // ordered by cost
        if (!have || std::make_tuple(n.focal, n.cost, n.id) <
                         std::make_tuple((*best)->focal, (*best)->cost, (*best)->id)) {
          best = it;
          have = true;
        }
      }
      // `have` is false only if the bound fell below the cheapest node; the
      // reference would pop an empty FOCAL there — take the cheapest node
    } else if (m_algo == Algo::ECBS) {
      long minLB = (*best)->LB;
      for (const Node* n : I.open) minLB = std::min(minLB, n->LB);
      const float bound = (float)minLB * m_opt.w;  // fp32 like ecbs.hpp:181
      bool have = false;
      for (auto it = I.open.begin(); it != I.open.end(); ++it) {
        const Node& n = **it;
        if (!((float)n.cost <= bound)) break;  // the set is ordered by cost
        // FOCAL order: focalHeuristic, then cost (ecbs.hpp:344-352), then id
        if (!have || std::make_tuple(n.focal, n.cost, n.id) <
                         std::make_tuple((*best)->focal, (*best)->cost, (*best)->id)) {
          best = it;
          have = true;
        }
      }
      if (!have) {  // cannot happen (the min-LB node always qualifies); stay safe
        for (auto it = I.open.begin(); it != I.open.end(); ++it)
          if ((*it)->LB == minLB) {
            best = it;
            break;
          }
      }
    }
    NodeUP n = own(*best);
    I.open.erase(best);
    return n;
  }

  // ---- low-level batch ---------------------------------------------------
  // staging buffers of one calling thread, reused across launches
  struct LLBuffers {
    std::vector<int32_t> tables, tlen, cells, gs;
  };
  void runLowLevel(const std::vector<JobSpec>& specs, const std::vector<const Node*>& tableNodes,
                   std::vector<JobOut>& out) {
    runLowLevel(specs, tableNodes, out, m_opt.maxLlExpanded, m_mainBuf, m_prof);
  }
  void runLowLevel(const std::vector<JobSpec>& specs, const std::vector<const Node*>& tableNodes,
                   std::vector<JobOut>& out, int maxExpanded, LLBuffers& buf, HostProfile& prof) {
    out.assign(specs.size(), JobOut());
    if (specs.empty()) return;
    const double tPack = nowSeconds();
    std::vector<mrp_job> jobs(specs.size());
    std::vector<int32_t> vc, ec;
    for (size_t k = 0; k < specs.size(); ++k) {
      const JobSpec& s = specs[k];
      mrp_job& j = jobs[k];
      j.map = m_inst[s.inst].mapIdx;
      j.start_cell = m_inst[s.inst].in->starts[s.agent];
      j.goal_cell = s.goal;
      j.field = s.field;
      j.vc_begin = (int)vc.size() / 2;
      j.ec_begin = (int)ec.size() / 3;
      appendCons(m_inst[s.inst], s.cons, vc, ec);
      j.vc_end = (int)vc.size() / 2;
      j.ec_end = (int)ec.size() / 3;
      j.table = s.table;
      j.self = s.self;
    }
    mrp_lowlevel_params prm;
    prm.variant = isTA() ? 1 : 0;
    prm.w = isFocal() ? m_opt.w : 0.0f;
    prm.max_expanded = maxExpanded;
    prm.path_cap = m_pathCap;
    std::vector<mrp_path_info> info(specs.size());
    if (m_pool) {
      // the other agents' paths as pool rows, the new paths into fresh rows
      int N = 0, Tpad = 0;
      std::vector<int32_t>& rowsOf = buf.tables;
      slotTables(tableNodes, rowsOf, N, Tpad);
      std::vector<int32_t> outRows;
      takeRows(specs.size(), outRows);
      gpuCheck(mrp_pathpool_reserve(m_pool, m_nextRow));
      const double tg = nowSeconds();
      prof.llPack += tg - tPack;
      prof.jobs += (long)specs.size();
      gpuCheck(mrp_lowlevel_batch_pool(m_maps.data(), (int)m_maps.size(), m_fields, vc.data(),
                                       (int)vc.size() / 2, ec.data(), (int)ec.size() / 3, m_pool,
                                       rowsOf.data(), (int)tableNodes.size(), N, Tpad, jobs.data(),
                                       (int)jobs.size(), &prm, outRows.data(), info.data()));
      const double tUn = nowSeconds();
      prof.gpuLowLevel += tUn - tg;
      for (size_t k = 0; k < specs.size(); ++k) {
        JobOut& o = out[k];
        o.status = info[k].status;
        o.expanded = info[k].expanded;
        m_inst[specs[k].inst].res.llExpanded += info[k].expanded;
        if (o.status == 0) {
          AgentPath ap;
          ap.cost = info[k].cost;
          ap.fmin = info[k].fmin;
          ap.slot = outRows[k];
          ap.length = info[k].length;
          o.path = adopt(std::move(ap), specs[k].inst);
        } else {
          giveRow(outRows[k]);
        }
      }
      prof.llUnpack += nowSeconds() - tUn;
      return;
    }
    // path tables of the other agents (ECBS focal heuristics)
    std::vector<int32_t>& tables = buf.tables;
    std::vector<int32_t>& tlen = buf.tlen;
    const bool haveTables = !tableNodes.empty();
    int N = 0, Tpad = 0;
    if (haveTables) packTables(tableNodes, tables, tlen, N, Tpad);
    // output staging is kept across calls (no zero fill of tens of MB per launch)
    if (buf.cells.size() < specs.size() * (size_t)m_pathCap) {
      buf.cells.resize(specs.size() * (size_t)m_pathCap);
      buf.gs.resize(buf.cells.size());
    }
    std::vector<int32_t>& cells = buf.cells;
    std::vector<int32_t>& gs = buf.gs;
    const double tg = nowSeconds();
    prof.llPack += tg - tPack;
    prof.jobs += (long)specs.size();
    gpuCheck(mrp_lowlevel_batch_fs(m_maps.data(), (int)m_maps.size(), m_fields, vc.data(),
                                   (int)vc.size() / 2, ec.data(), (int)ec.size() / 3,
                                   haveTables ? tables.data() : nullptr,
                                   haveTables ? tlen.data() : nullptr, (int)tableNodes.size(), N,
                                   Tpad, jobs.data(), (int)jobs.size(), &prm, info.data(),
                                   cells.data(), gs.data()));
    const double tUn = nowSeconds();
    prof.gpuLowLevel += tUn - tg;
    for (size_t k = 0; k < specs.size(); ++k) {
      JobOut& o = out[k];
      o.status = info[k].status;
      o.expanded = info[k].expanded;
      m_inst[specs[k].inst].res.llExpanded += info[k].expanded;
      if (o.status == 0) {
        const int L = info[k].length;
        AgentPath ap;
        ap.cells.assign(cells.begin() + k * m_pathCap, cells.begin() + k * m_pathCap + L);
        ap.g.assign(gs.begin() + k * m_pathCap, gs.begin() + k * m_pathCap + L);
        ap.cost = info[k].cost;
        ap.fmin = info[k].fmin;
        ap.length = L;
        o.path = adopt(std::move(ap), specs[k].inst);
      }
    }
    prof.llUnpack += nowSeconds() - tUn;
  }

  // pool mode: the tables of `nodes` as pool rows [B][N] (-1: no path yet) and the
  // length of the longest path among them
  void slotTables(const std::vector<const Node*>& nodes, std::vector<int32_t>& rows, int& N,
                  int& Tpad) const {
    N = 0;
    Tpad = 1;
    for (const Node* n : nodes) N = std::max(N, n->nAgents);
    rows.assign(nodes.size() * (size_t)N, -1);
    std::vector<int> tmax(nodes.size(), 1);
#pragma omp parallel for schedule(static) if (nodes.size() >= kParallelMin)
    for (long b = 0; b < (long)nodes.size(); ++b)
      for (int a = 0; a < nodes[b]->nAgents; ++a) {
        const PathRef& p = nodes[b]->paths[a];
        if (!p) continue;
        rows[b * N + a] = p.slot;
        tmax[b] = std::max(tmax[b], p.length);
      }
    for (int t : tmax) Tpad = std::max(Tpad, t);
  }

  void packTables(const std::vector<const Node*>& nodes, std::vector<int32_t>& tables,
                  std::vector<int32_t>& tlen, int& N, int& Tpad) const {
    N = 0;
    Tpad = 1;
    for (const Node* n : nodes) {
      N = std::max(N, n->nAgents);
      for (int a = 0; a < n->nAgents; ++a)
        if (n->paths[a]) Tpad = std::max(Tpad, (int)n->paths[a].host->cells.size());
    }
    // cells past a path's length are never read (the kernels clamp to len-1),
    // so the tables are not cleared: only the lengths are
    if (tables.size() < nodes.size() * (size_t)N * Tpad) tables.resize(nodes.size() * (size_t)N * Tpad);
    tlen.assign(nodes.size() * (size_t)N, 0);
#pragma omp parallel for schedule(static) if (nodes.size() >= kParallelMin)
    for (long b = 0; b < (long)nodes.size(); ++b)
      for (int a = 0; a < nodes[b]->nAgents; ++a) {
        if (!nodes[b]->paths[a]) continue;  // not planned yet (ECBS root construction)
        const auto& c = nodes[b]->paths[a].host->cells;
        std::copy(c.begin(), c.end(), tables.begin() + (b * N + a) * (size_t)Tpad);
        tlen[b * N + a] = (int)c.size();
      }
  }

  // ---- conflicts of freshly created nodes ----------------------------------
  void evaluate(std::vector<Node*>& fresh) {
    if (fresh.empty()) return;
    const double tEp = nowSeconds();
    std::vector<const Node*> nodes(fresh.begin(), fresh.end());
    std::vector<int32_t>& tables = m_evTables;
    std::vector<int32_t>& tlen = m_evTlen;
    int N = 0, Tpad = 0;
    const int B = (int)fresh.size();
    std::vector<int32_t> found(B), counts(B);
    std::vector<mrp_conflict> confl(B);
    if (m_pool) {
      slotTables(nodes, tables, N, Tpad);
      const double tg = nowSeconds();
      m_prof.evalPack += tg - tEp;
      m_prof.nodes += B;
      gpuCheck(mrp_conflicts_batch_pool(m_pool, tables.data(), B, N, Tpad, m_dimx, 0, found.data(),
                                        confl.data(), counts.data()));
      m_prof.gpuConflicts += nowSeconds() - tg;
      for (int b = 0; b < B; ++b) {
        fresh[b]->found = found[b];
        fresh[b]->conflict = confl[b];
        fresh[b]->focal = counts[b];
      }
      return;
    }
    packTables(nodes, tables, tlen, N, Tpad);
    m_prof.evalPack += nowSeconds() - tEp;
    // getFirstConflict bound: size-1 for cbs/ecbs (cbs.cpp:338-341), size for
    // cbs_ta (cbs_ta.cpp:372-375); focalHeuristic counts with the same table
    const int mode = isTA() ? 1 : 0;
    const double tg = nowSeconds();
    m_prof.nodes += B;
    gpuCheck(mrp_conflicts_batch(tables.data(), tlen.data(), B, N, Tpad, m_dimx, mode,
                                 found.data(), confl.data(), counts.data()));
    if (m_algo == Algo::ECBSTA) {
      // ecbs_ta mixes the bounds: getFirstConflict uses max(size)
      // (ecbs_ta.cpp:445), focalHeuristic max(size-1) (ecbs_ta.cpp:353)
      std::vector<int32_t> found0(B);
      std::vector<mrp_conflict> confl0(B);
      gpuCheck(mrp_conflicts_batch(tables.data(), tlen.data(), B, N, Tpad, m_dimx, 0,
                                   found0.data(), confl0.data(), counts.data()));
    }
    m_prof.gpuConflicts += nowSeconds() - tg;
    for (int b = 0; b < B; ++b) {
      fresh[b]->found = found[b];
      fresh[b]->conflict = confl[b];
      fresh[b]->focal = counts[b];
    }
  }

  // ---- roots -----------------------------------------------------------------
  void buildRoots(std::vector<Node*>& fresh) {
    std::vector<NodeUP> roots(m_inst.size());
    for (size_t k = 0; k < m_inst.size(); ++k) {
      Inst& I = m_inst[k];
      NodeUP n = own(newNode((int)k));
      n->id = I.nextId++;
      n->isRoot = true;
      if (m_algo == Algo::CBSTA) {
        if (!nextTasks(I, n->task)) {  // cbs_ta.hpp:96-103: no assignment, no search
          finish(I, kNoSolution, nullptr, 0);
          continue;
        }
      } else if (m_algo == Algo::ECBSTA) {
        nextTasks(I, n->task);  // ecbs_ta.hpp:104 plans even without an assignment
      } else {
        n->task = std::make_shared<const std::vector<int> >(I.in->goals.begin(), I.in->goals.end());
      }
      roots[k] = std::move(n);
    }
    if (isFocal() && m_sliced) {
      // the agents of a root are planned one after the other (ecbs.hpp:118-136), but the
      // instances need not wait for each other: one flight per root, one sliced replan
      // at a time, and the finished root joins `fresh` in the main loop
      for (size_t k = 0; k < m_inst.size(); ++k) {
        if (!roots[k]) continue;
        std::unique_ptr<Flight> f(new Flight());
        f->pd.inst = (int)k;
        f->root = std::move(roots[k]);
        f->nextAgent = 0;
        if (!nextRootJob(*f)) {  // no agents: the empty root is complete
          fresh.push_back(f->root.release());
          continue;
        }
        m_inst[k].busy = true;
        m_flights.push_back(std::move(f));
      }
      return;
    }
    if (isFocal()) {
      planSequential(roots, true);
      for (size_t k = 0; k < m_inst.size(); ++k)
        if (roots[k]) m_inst[k].nextRootNodeCost = (long)((float)roots[k]->LB * m_opt.w);
    } else {
      std::vector<JobSpec> specs;
      for (size_t k = 0; k < m_inst.size(); ++k) {
        if (!roots[k]) continue;
        for (size_t a = 0; a < m_inst[k].in->numAgents(); ++a) {
          const int goal = (*roots[k]->task)[a];
          specs.push_back({(int)k, (int)a, goal, fieldFor(m_inst[k], (int)a, goal),
                           roots[k]->cons[a], -1, (int)a});
        }
      }
      std::vector<JobOut> outs;
      runLowLevel(specs, std::vector<const Node*>(), outs);
      for (size_t j = 0; j < specs.size(); ++j) absorbRoot(roots, specs[j], outs[j], true);
    }
    for (size_t k = 0; k < m_inst.size(); ++k)
      if (roots[k]) fresh.push_back(roots[k].release());
  }

  // Plans the agents of whole (root) nodes one after the other, each seeing the
  // already planned ones through the focal heuristics (ecbs.hpp:118-136,
  // ecbs_ta.hpp:107-126,318-330); nodes[k] belongs to instance k (may be null).
  // initial == true: a failing search ends the instance (first root);
  // otherwise the node is just dropped (later roots of ecbs_ta).
  void planSequential(std::vector<NodeUP>& nodes, bool initial) {
    size_t maxN = 0;
    for (size_t k = 0; k < nodes.size(); ++k)
      if (nodes[k]) maxN = std::max(maxN, m_inst[k].in->numAgents());
    for (size_t a = 0; a < maxN; ++a) {
      std::vector<JobSpec> specs;
      std::vector<const Node*> tabs;
      for (size_t k = 0; k < nodes.size(); ++k) {
        if (!nodes[k] || a >= m_inst[k].in->numAgents()) continue;
        const int goal = (*nodes[k]->task)[a];
        specs.push_back({(int)k, (int)a, goal, fieldFor(m_inst[k], (int)a, goal),
                         nodes[k]->cons[a], (int)tabs.size(), (int)a});
        tabs.push_back(nodes[k].get());
      }
      std::vector<JobOut> outs;
      runLowLevel(specs, tabs, outs);
      for (size_t j = 0; j < specs.size(); ++j) absorbRoot(nodes, specs[j], outs[j], initial);
    }
  }

  void absorbRoot(std::vector<NodeUP>& roots, const JobSpec& s, JobOut& o,
                  bool initial) {
    if (!roots[s.inst]) return;
    if (o.status != 0) {  // cbs.hpp:96-100: a failing root search ends the search
      if (initial || o.status == 2)
        finish(m_inst[s.inst], o.status == 2 ? kCapped : kNoSolution, nullptr, 0);
      roots[s.inst].reset();
      return;
    }
    Node& n = *roots[s.inst];
    n.cost += o.path.cost;
    n.LB += o.path.fmin;
    n.paths[s.agent] = std::move(o.path);
  }

  // ecbs_ta, STYLE_MINROOT (ecbs_ta.hpp:299-348): once the cheapest open node
  // exceeds nextRootNodeCost the next-best assignment becomes a new root, then
  // the bound is reset from the cheapest node's LB.
  void spawnMinRoots() {
    std::vector<NodeUP> roots(m_inst.size());
    std::vector<char> triggered(m_inst.size(), 0);
    bool any = false;
    for (size_t k = 0; k < m_inst.size(); ++k) {
      Inst& I = m_inst[k];
      if (I.done || I.open.empty()) continue;
      if (!((*I.open.begin())->cost > I.nextRootNodeCost)) continue;
      triggered[k] = 1;
      NodeUP r = own(newNode((int)k));
      if (!nextTasks(I, r->task)) continue;
      r->isRoot = true;
      roots[k] = std::move(r);
      any = true;
    }
    if (any) {
      planSequential(roots, false);
      std::vector<Node*> fresh;
      for (size_t k = 0; k < m_inst.size(); ++k)
        if (roots[k] && !m_inst[k].done) {
          roots[k]->id = m_inst[k].nextId++;
          fresh.push_back(roots[k].release());
        }
      evaluate(fresh);
      for (Node* n : fresh) openInsert(m_inst[n->inst], n);
    }
    for (size_t k = 0; k < m_inst.size(); ++k) {
      Inst& I = m_inst[k];
      if (triggered[k] && !I.done && !I.open.empty())
        I.nextRootNodeCost = (long)((float)(*I.open.begin())->LB * m_opt.w);
    }
  }

  // ---- sliced expansions (pool mode) ---------------------------------------------
  struct FlightJob {
    JobSpec spec;
    int32_t outRow = -1, state = -1;
    bool started = false, done = false;
    JobOut out;
  };
  struct FlightChild {
    int agent = 0;
    NodeUP node;
  };
  // one high-level expansion in progress: the parent, its two children and their replans
  struct Flight {
    Pending pd;
    std::vector<FlightChild> kids;  // kids[q] is replanned by jobs[q]
    std::vector<FlightJob> jobs;
    // a root under construction (ecbs): its agents are planned one by one, jobs[0] is
    // the replan of agent nextAgent and sees the agents planned so far
    NodeUP root;
    int nextAgent = 0;
  };
  bool nextRootJob(Flight& f) {
    const Inst& I = m_inst[f.pd.inst];
    if ((size_t)f.nextAgent >= I.in->numAgents()) return false;
    const int a = f.nextAgent;
    FlightJob j;
    j.spec = {f.pd.inst, a, (*f.root->task)[a], fieldFor(I, a, (*f.root->task)[a]), f.root->cons[a], -1, a};
    f.jobs.clear();
    f.jobs.push_back(std::move(j));
    return true;
  }
  void takeStates(size_t n, std::vector<int32_t>& ids) {
    ids.resize(n);
    for (size_t i = 0; i < n; ++i) {
      if (!m_freeStates.empty()) {
        ids[i] = m_freeStates.back();
        m_freeStates.pop_back();
      } else {
        ids[i] = m_nextState++;
      }
    }
  }
  void dropFlight(Flight& f) {
    for (FlightJob& j : f.jobs) {
      if (j.state >= 0) m_freeStates.push_back(j.state);
      if (j.outRow >= 0 && !j.out.path) giveRow(j.outRow);  // (an adopted row belongs to the instance)
      j.state = j.outRow = -1;
    }
    m_inst[f.pd.inst].busy = false;
  }
  // createConstraintsFromConflict (example/cbs.cpp:388-406) for every popped parent:
  // two children in ascending agent order, one replan each
  void startFlights(std::vector<Pending>& pending) {
    if (pending.empty()) return;
    const double tBuild = nowSeconds();
    std::vector<std::unique_ptr<Flight> > made(pending.size());
#pragma omp parallel for schedule(dynamic, 16) if (pending.size() >= kParallelMin)
    for (long pi = 0; pi < (long)pending.size(); ++pi) {
      std::unique_ptr<Flight> f(new Flight());
      Inst& I = m_inst[pending[pi].inst];
      const Node& P = *pending[pi].parent;
      const mrp_conflict& c = P.conflict;
      for (int side = 0; side < 2; ++side) {
        const int agent = side == 0 ? c.agent1 : c.agent2;
        NodeUP n = own(cloneNode(P));  // a flat copy: rows and chain heads
        n->cons[agent] = addConstraint(I, P, c, side, agent);
        n->cost -= n->paths[agent].cost;
        n->LB -= n->paths[agent].fmin;
        FlightJob j;
        j.spec = {pending[pi].inst, agent, (*n->task)[agent], fieldFor(I, agent, (*n->task)[agent]),
                  n->cons[agent], -1, agent};
        f->jobs.push_back(std::move(j));
        FlightChild k;
        k.agent = agent;
        k.node = std::move(n);
        f->kids.push_back(std::move(k));
      }
      f->pd = std::move(pending[pi]);
      I.busy = true;
      made[pi] = std::move(f);
    }
    for (auto& f : made) m_flights.push_back(std::move(f));
    m_prof.build += nowSeconds() - tBuild;
  }
  // one launch: every unfinished replan of every flight gets (at most) one slice
  void runFlights() {
    const double tPack = nowSeconds();
    std::vector<std::pair<Flight*, FlightJob*> > run;
    std::vector<const Node*> tabs;
    std::vector<mrp_job> jobs;
    std::vector<int32_t> vc, ec, outRows, states, resume;
    size_t needRows = 0, needStates = 0;
    for (auto& f : m_flights)
      for (FlightJob& j : f->jobs)
        if (!j.done) {
          if (j.outRow < 0) ++needRows;
          if (j.state < 0) ++needStates;
        }
    std::vector<int32_t> newRows, newStates;
    takeRows(needRows, newRows);
    takeStates(needStates, newStates);
    size_t ir = 0, is = 0;
    for (auto& f : m_flights) {
      int table = -1;
      for (FlightJob& j : f->jobs) {
        if (j.done) continue;
        if (isFocal() && table < 0) {
          table = (int)tabs.size();
          tabs.push_back(f->root ? f->root.get() : f->pd.parent.get());
        }
        if (j.outRow < 0) j.outRow = newRows[ir++];
        if (j.state < 0) j.state = newStates[is++];
        const JobSpec& sp = j.spec;
        mrp_job mj;
        mj.map = m_inst[sp.inst].mapIdx;
        mj.start_cell = m_inst[sp.inst].in->starts[sp.agent];
        mj.goal_cell = sp.goal;
        mj.field = sp.field;
        mj.vc_begin = (int)vc.size() / 2;
        mj.ec_begin = (int)ec.size() / 3;
        appendCons(m_inst[sp.inst], sp.cons, vc, ec);
        mj.vc_end = (int)vc.size() / 2;
        mj.ec_end = (int)ec.size() / 3;
        mj.table = table;
        mj.self = sp.self;
        jobs.push_back(mj);
        outRows.push_back(j.outRow);
        states.push_back(j.state);
        resume.push_back(j.started ? 1 : 0);
        run.push_back(std::make_pair(f.get(), &j));
      }
    }
    if (jobs.empty()) return;
    int N = 0, Tpad = 0;
    std::vector<int32_t>& rowsOf = m_mainBuf.tables;
    slotTables(tabs, rowsOf, N, Tpad);
    gpuCheck(mrp_pathpool_reserve(m_pool, m_nextRow));
    gpuCheck(mrp_pathpool_reserve_states(m_pool, m_nextState, m_dimx, m_dimy, m_opt.maxLlExpanded));
    mrp_lowlevel_params prm;
    prm.variant = 0;
    prm.w = isFocal() ? m_opt.w : 0.0f;
    prm.max_expanded = m_opt.maxLlExpanded;
    prm.path_cap = m_pathCap;
    std::vector<mrp_path_info> info(jobs.size());
    const double tg = nowSeconds();
    m_prof.llPack += tg - tPack;
    // few replans left (the stragglers of a batch): nobody gains from short launches, the
    // per-launch overhead is all that is left to save
    const int slice = jobs.size() <= 16 ? 4 * m_slice : (jobs.size() <= 64 ? 2 * m_slice : m_slice);
    gpuCheck(mrp_lowlevel_batch_pool_sliced(m_maps.data(), (int)m_maps.size(), m_fields, vc.data(),
                                            (int)vc.size() / 2, ec.data(), (int)ec.size() / 3, m_pool,
                                            rowsOf.data(), (int)tabs.size(), N, Tpad, jobs.data(),
                                            (int)jobs.size(), &prm, outRows.data(), states.data(),
                                            resume.data(), slice, info.data()));
    const double tUn = nowSeconds();
    m_prof.gpuLowLevel += tUn - tg;
    ++m_prof.launches;
    for (size_t k = 0; k < run.size(); ++k) {
      FlightJob& j = *run[k].second;
      if (info[k].status == MRP_SUSPENDED) {
        j.started = true;
        continue;
      }
      j.done = true;
      ++m_prof.jobs;
      j.out.status = info[k].status;
      j.out.expanded = info[k].expanded;
      m_inst[j.spec.inst].res.llExpanded += info[k].expanded;
      if (j.out.status == 0) {
        AgentPath ap;
        ap.cost = info[k].cost;
        ap.fmin = info[k].fmin;
        ap.slot = j.outRow;
        ap.length = info[k].length;
        j.out.path = adopt(std::move(ap), j.spec.inst);
      } else {
        giveRow(j.outRow);
      }
      j.outRow = -1;
      m_freeStates.push_back(j.state);
      j.state = -1;
    }
    m_prof.llUnpack += nowSeconds() - tUn;
  }
  // flights whose replans are all done hand their children over (cbs.hpp:146-167); the
  // flights belong to different instances, so they are absorbed (and their parents
  // released) on all host cores, the children join `fresh` in flight order
  void absorbFlights(std::vector<Node*>& fresh) {
    const double tAbs = nowSeconds();
    const size_t nF = m_flights.size();
    std::vector<std::array<Node*, 2> > made(nF, std::array<Node*, 2>{{nullptr, nullptr}});
    std::vector<char> keepIt(nF, 0);
#pragma omp parallel for schedule(dynamic, 16) if (nF >= kParallelMin)
    for (long i = 0; i < (long)nF; ++i) {
      Flight& f = *m_flights[i];
      bool all = true;
      for (const FlightJob& j : f.jobs) all = all && j.done;
      if (!all) {
        keepIt[i] = 1;
        continue;
      }
      Inst& I = m_inst[f.pd.inst];
      if (f.root) {
        FlightJob& j = f.jobs[0];
        if (j.out.status != 0) {  // cbs.hpp:96-100: a failing root search ends the search
          finish(I, j.out.status == 2 ? kCapped : kNoSolution, nullptr, 0);
          I.busy = false;
          m_flights[i].reset();
          continue;
        }
        f.root->cost += j.out.path.cost;
        f.root->LB += j.out.path.fmin;
        f.root->paths[f.nextAgent] = std::move(j.out.path);
        ++f.nextAgent;
        if (nextRootJob(f)) {  // the next agent of this root
          keepIt[i] = 1;
          continue;
        }
        made[i][0] = f.root.release();
        I.busy = false;
        m_flights[i].reset();
        continue;
      }
      I.busy = false;
      bool capped = false;
      for (const FlightJob& j : f.jobs) capped = capped || j.out.status == 2;
      if (capped) {  // a capped replan could hide the optimum: give up honestly
        finish(I, kCapped, nullptr, 0);
        m_flights[i].reset();
        continue;
      }
      for (size_t q = 0; q < f.kids.size() && q < 2; ++q) {
        FlightJob& j = f.jobs[q];
        if (j.out.status != 0) continue;  // no path under these constraints: the child is dropped
        Node& n = *f.kids[q].node;
        n.cost += j.out.path.cost;
        n.LB += j.out.path.fmin;
        n.paths[f.kids[q].agent] = std::move(j.out.path);
        n.id = I.nextId++;
        made[i][q] = f.kids[q].node.release();
      }
      m_flights[i].reset();  // releases the parent
    }
    size_t keep = 0;
    for (size_t i = 0; i < nF; ++i) {
      for (Node* n : made[i])
        if (n) fresh.push_back(n);
      if (keepIt[i]) {
        if (keep != i) m_flights[keep] = std::move(m_flights[i]);
        ++keep;
      }
    }
    m_flights.resize(keep);
    m_prof.absorb += nowSeconds() - tAbs;
  }

  // ---- one expansion step for every pending parent -----------------------------
  // maxExpanded: cap of the replans of this call
  void expand(std::vector<Pending>& pending, std::vector<Node*>& fresh, int maxExpanded,
              LLBuffers& buf, HostProfile& prof) {
    struct ChildPlan {
      int pendingIdx;
      int agent;        // replanned agent; -1: a whole new root (cbs_ta)
      NodeUP node;
      size_t firstJob, nJobs;
      bool failed = false;
    };
    const double tBuild = nowSeconds();
    // every pending parent builds its children on its own (all host cores),
    // then the plans / job specs are concatenated in pending order
    std::vector<std::vector<ChildPlan> > localPlans(pending.size());
    std::vector<std::vector<JobSpec> > localSpecs(pending.size());
#pragma omp parallel for schedule(dynamic, 16) if (pending.size() >= kParallelMin)
    for (long pi = 0; pi < (long)pending.size(); ++pi) {
      std::vector<ChildPlan>& plans = localPlans[pi];
      std::vector<JobSpec>& specs = localSpecs[pi];
      Inst& I = m_inst[pending[pi].inst];
      const Node& P = *pending[pi].parent;
      const int tableIdx = isFocal() ? (int)pi : -1;  // tabs[pi] = the parent of pending[pi]
      if (m_algo == Algo::CBSTA && P.isRoot) {
        // an expanded root with a conflict spawns the next-best assignment as
        // a new root (cbs_ta.hpp:142-172).  The reference copies isRoot into
        // every child (cbs_ta.hpp:180), so this fires on every expansion.
        NodeUP r = own(newNode(pending[pi].inst));
        if (nextTasks(I, r->task)) {
          r->isRoot = true;
          ChildPlan cp;
          cp.pendingIdx = (int)pi;
          cp.agent = -1;
          cp.firstJob = specs.size();
          cp.nJobs = I.in->numAgents();
          for (size_t a = 0; a < I.in->numAgents(); ++a)
            specs.push_back({pending[pi].inst, (int)a, (*r->task)[a],
                             fieldFor(I, (int)a, (*r->task)[a]), r->cons[a], -1, (int)a});
          cp.node = std::move(r);
          plans.push_back(std::move(cp));
        }
      }
      // createConstraintsFromConflict, example/cbs.cpp:388-406: children in
      // ascending agent order (agent1 < agent2)
      const mrp_conflict& c = P.conflict;
      for (int side = 0; side < 2; ++side) {
        const int agent = side == 0 ? c.agent1 : c.agent2;
        NodeUP n = own(cloneNode(P));  // a flat copy: path references and chain heads
        n->isRoot = P.isRoot;
        n->cons[agent] = addConstraint(I, P, c, side, agent);
        n->cost -= n->paths[agent].cost;
        n->LB -= n->paths[agent].fmin;
        ChildPlan cp;
        cp.pendingIdx = (int)pi;
        cp.agent = agent;
        cp.firstJob = specs.size();
        cp.nJobs = 1;
        specs.push_back({pending[pi].inst, agent, (*n->task)[agent],
                         fieldFor(I, agent, (*n->task)[agent]), n->cons[agent], tableIdx, agent});
        cp.node = std::move(n);
        plans.push_back(std::move(cp));
      }
    }
    std::vector<ChildPlan> plans;
    std::vector<JobSpec> specs;
    std::vector<const Node*> tabs;
    std::vector<size_t> firstPlan(pending.size() + 1, 0);
    for (size_t pi = 0; pi < pending.size(); ++pi) {
      if (isFocal()) tabs.push_back(pending[pi].parent.get());
      const size_t base = specs.size();
      for (ChildPlan& cp : localPlans[pi]) {
        cp.firstJob += base;
        plans.push_back(std::move(cp));
      }
      specs.insert(specs.end(), localSpecs[pi].begin(), localSpecs[pi].end());
      firstPlan[pi + 1] = plans.size();
    }
    std::vector<JobOut> outs;
    prof.build += nowSeconds() - tBuild;
    runLowLevel(specs, tabs, outs, maxExpanded, buf, prof);
    const double tAbs = nowSeconds();
    std::vector<Node*> made(plans.size(), nullptr);
#pragma omp parallel for schedule(dynamic, 16) if (pending.size() >= kParallelMin)
    for (long pi = 0; pi < (long)pending.size(); ++pi) {
      Inst& I = m_inst[pending[pi].inst];
      for (size_t q = firstPlan[pi]; q < firstPlan[pi + 1]; ++q) {
        ChildPlan& cp = plans[q];
        if (I.done) continue;
        Node& n = *cp.node;
        bool ok = true, capped = false;
        for (size_t j = cp.firstJob; j < cp.firstJob + cp.nJobs; ++j) {
          if (outs[j].status == 2) capped = true;
          if (outs[j].status != 0) ok = false;
        }
        if (capped) {  // a capped replan could hide the optimum: give up honestly
          finish(I, kCapped, nullptr, 0);
          continue;
        }
        if (!ok) continue;  // no path under these constraints: the child is dropped
        for (size_t j = cp.firstJob; j < cp.firstJob + cp.nJobs; ++j) {
          const int a = specs[j].agent;
          n.cost += outs[j].path.cost;
          n.LB += outs[j].path.fmin;
          n.paths[a] = std::move(outs[j].path);
        }
        n.id = I.nextId++;
        made[q] = cp.node.release();
      }
    }
    for (Node* n : made)
      if (n) fresh.push_back(n);
    prof.absorb += nowSeconds() - tAbs;
    // children of instances that were finished meanwhile must not leak
    size_t keep = 0;
    for (size_t i = 0; i < fresh.size(); ++i) {
      if (m_inst[fresh[i]->inst].done)
        freeNode(fresh[i]);
      else
        fresh[keep++] = fresh[i];
    }
    fresh.resize(keep);
  }

  Algo m_algo;
  SolveOptions m_opt;
  int m_dimx = 0, m_dimy = 0, m_cells = 0, m_pathCap = 128;
  std::vector<Inst> m_inst;
  std::vector<mrp_map> m_maps;
  mrp_fieldset m_fields = nullptr;
  mrp_pathpool m_pool = nullptr;  // device rows of the paths (cbs / ecbs)
  std::vector<int32_t> m_freeRows;
  int m_nextRow = 0;
  std::vector<std::unique_ptr<Flight> > m_flights;  // expansions whose replans are in progress
  std::vector<int32_t> m_freeStates;
  int m_nextState = 0;
  int m_slice = 256;  // expansions per replan and launch (MRP_HOST_SLICE)
  bool m_sliced = false;
  std::mutex m_rowMutex;
  std::vector<int> m_doneQueue;  // finished instances whose rows are still to be returned (m_rowMutex)
  HostProfile m_prof;
  // staging buffers reused across lock-step iterations
  mutable std::vector<int32_t> m_evTables, m_evTlen;
  LLBuffers m_mainBuf;
};

}  // namespace mrp_host
