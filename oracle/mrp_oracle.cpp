// mrp_oracle.cpp — CPU ORACLE (test infrastructure, never the product path).
//
// A from-scratch restatement of the semantics of the libMultiRobotPlanning hot
// path.  Each block cites the reference file:line whose behaviour it follows.
// Parity status and what is (un)pinned: see mrp_oracle.h.
//
// Build: see oracle/Makefile (g++ -O3 -shared -fPIC).

#include "mrp_oracle.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <queue>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------
// Grid
// ---------------------------------------------------------------------------
struct Grid {
  int dimx = 0, dimy = 0;
  std::vector<uint8_t> blocked;  // [dimx*dimy], index x + dimx*y
  int cells() const { return dimx * dimy; }
  bool inside(int x, int y) const {
    return x >= 0 && x < dimx && y >= 0 && y < dimy;
  }
  bool freeCell(int x, int y) const {
    return inside(x, y) && !blocked[x + dimx * y];
  }
};

static Grid makeGrid(int dimx, int dimy, const int32_t* obst_xy, int n_obst) {
  Grid g;
  g.dimx = dimx;
  g.dimy = dimy;
  g.blocked.assign((size_t)dimx * dimy, 0);
  for (int k = 0; k < n_obst; ++k) {
    int x = obst_xy[2 * k], y = obst_xy[2 * k + 1];
    // obstacles outside the map are simply never looked up by the reference
    // (unordered_set<Location> membership, example/cbs.cpp:431-436)
    if (g.inside(x, y)) g.blocked[x + dimx * y] = 1;
  }
  return g;
}

// ---------------------------------------------------------------------------
// Distance fields
// ---------------------------------------------------------------------------
// Floyd–Warshall restatement of example/shortest_path_heuristic.hpp:12-54.
// One vertex per cell (obstacles included, :18-22); unit edges between two
// free cells to the right and below (:25-43); Boost defaults: d[v][v] = 0,
// inf = INT_MAX, combine = closed_plus (saturating).
static void floydWarshall(const Grid& g, int32_t* d) {
  const int V = g.cells();
  for (size_t i = 0; i < (size_t)V * V; ++i) d[i] = ORC_INF;
  for (int v = 0; v < V; ++v) d[(size_t)v * V + v] = 0;
  for (int y = 0; y < g.dimy; ++y)
    for (int x = 0; x < g.dimx; ++x) {
      if (!g.freeCell(x, y)) continue;
      int v = x + g.dimx * y;
      if (g.freeCell(x + 1, y)) {
        int u = v + 1;
        d[(size_t)v * V + u] = d[(size_t)u * V + v] = 1;
      }
      if (g.freeCell(x, y + 1)) {
        int u = v + g.dimx;
        d[(size_t)v * V + u] = d[(size_t)u * V + v] = 1;
      }
    }
  for (int k = 0; k < V; ++k) {
    const int32_t* dk = d + (size_t)k * V;
    for (int i = 0; i < V; ++i) {
      int32_t dik = d[(size_t)i * V + k];
      if (dik == ORC_INF) continue;
      int32_t* di = d + (size_t)i * V;
      for (int j = 0; j < V; ++j) {
        if (dk[j] == ORC_INF) continue;
        int32_t c = dik + dk[j];
        if (c < di[j]) di[j] = c;
      }
    }
  }
}

// Per-goal restatement following the disabled computeHeuristic()
// (example/cbs.cpp:445-557): field initialised to INT_MAX (:549), goal = 0
// (:554), neighbours Left/Right/Up/Down over free in-bounds cells (:478-508),
// g-score recorded on discovery (:517-523).  Unit costs => FIFO queue.
static void bfsField(const Grid& g, int gx, int gy, int32_t* out) {
  const int V = g.cells();
  for (int i = 0; i < V; ++i) out[i] = ORC_INF;
  if (!g.inside(gx, gy)) return;
  const int goal = gx + g.dimx * gy;
  out[goal] = 0;
  if (g.blocked[goal]) return;  // FW row of an obstacle vertex: only d[v][v]=0
  std::vector<int> q;
  q.reserve(V);
  q.push_back(goal);
  for (size_t head = 0; head < q.size(); ++head) {
    int c = q[head];
    int x = c % g.dimx, y = c / g.dimx;
    int d = out[c] + 1;
    const int nx[4] = {x - 1, x + 1, x, x};
    const int ny[4] = {y, y, y + 1, y - 1};
    for (int k = 0; k < 4; ++k) {
      if (!g.freeCell(nx[k], ny[k])) continue;
      int n = nx[k] + g.dimx * ny[k];
      if (out[n] != ORC_INF) continue;
      out[n] = d;
      q.push_back(n);
    }
  }
}

// ---------------------------------------------------------------------------
// Paths and conflicts
// ---------------------------------------------------------------------------
struct Step {
  int time;
  int cell;
  int g;
};
struct Plan {  // libMultiRobotPlanning/planresult.hpp:18-28
  std::vector<Step> states;
  int cost = 0;
  int fmin = 0;
};

// getState() clamp, example/cbs.cpp:420-429
static inline int posAt(const std::vector<Plan>& sol, size_t i, int t) {
  const auto& s = sol[i].states;
  if ((size_t)t < s.size()) return s[t].cell;
  return s.back().cell;
}
static inline int posAtTable(const int32_t* cell, const int32_t* len, int Tpad,
                             int i, int t) {
  int L = len[i];
  // An empty path is outside the reference's contract for these loops
  // (getState asserts !states.empty(), example/cbs.cpp:420-429); the C ABI
  // defines it as "matches nothing" (include/mrp_b200.h): a private cell.
  if (L <= 0) return -2 - i;
  return cell[(size_t)i * Tpad + (t < L ? t : L - 1)];
}

template <class PosFn>
static bool firstConflictT(int N, int max_t, int dimx, PosFn pos,
                           orc_conflict* out) {
  // example/cbs.cpp:343-383: for every t, all vertex pairs first, then all
  // edge pairs; i ascending, j > i ascending.
  for (int t = 0; t < max_t; ++t) {
    for (int i = 0; i < N; ++i) {
      int a = pos(i, t);
      for (int j = i + 1; j < N; ++j) {
        if (a == pos(j, t)) {
          out->time = t;
          out->agent1 = i;
          out->agent2 = j;
          out->type = 0;
          out->x1 = a % dimx;
          out->y1 = a / dimx;
          out->x2 = -1;
          out->y2 = -1;
          return true;
        }
      }
    }
    for (int i = 0; i < N; ++i) {
      int a0 = pos(i, t), a1 = pos(i, t + 1);
      for (int j = i + 1; j < N; ++j) {
        int b0 = pos(j, t), b1 = pos(j, t + 1);
        if (a0 == b1 && a1 == b0) {
          out->time = t;
          out->agent1 = i;
          out->agent2 = j;
          out->type = 1;
          out->x1 = a0 % dimx;
          out->y1 = a0 / dimx;
          out->x2 = a1 % dimx;
          out->y2 = a1 / dimx;
          return true;
        }
      }
    }
  }
  return false;
}

template <class PosFn>
static int countConflictsT(int N, int max_t, PosFn pos) {
  // example/ecbs.cpp:315-350
  int n = 0;
  for (int t = 0; t < max_t; ++t) {
    for (int i = 0; i < N; ++i) {
      int a = pos(i, t);
      for (int j = i + 1; j < N; ++j)
        if (a == pos(j, t)) ++n;
    }
    for (int i = 0; i < N; ++i) {
      int a0 = pos(i, t), a1 = pos(i, t + 1);
      for (int j = i + 1; j < N; ++j)
        if (a0 == pos(j, t + 1) && a1 == pos(j, t)) ++n;
    }
  }
  return n;
}

static int maxT(const std::vector<Plan>& sol, int mode) {
  int m = 0;
  for (const auto& p : sol)
    m = std::max<int>(m, (int)p.states.size() - (mode == 0 ? 1 : 0));
  return m;
}

// ---------------------------------------------------------------------------
// Mutable binary heap with stable handles (stands in for
// boost::heap::d_ary_heap<T, arity<2>, mutable_<true>>; the operations used are
// push/top/pop/erase/increase/ordered iteration — a_star.hpp:205-208,
// a_star_epsilon.hpp:296-298,378-381, cbs.hpp:110-112, ecbs.hpp:299-301).
// Tie behaviour among equal keys is implementation-defined and unpinned.
// Better(a,b) == "a leaves the heap strictly before b".
// ---------------------------------------------------------------------------
template <class T, class Better>
class MutableHeap {
 public:
  explicit MutableHeap(Better b = Better()) : better_(b) {}
  bool empty() const { return heap_.empty(); }
  size_t size() const { return heap_.size(); }
  int push(const T& v) {
    int id;
    if (!freeIds_.empty()) {
      id = freeIds_.back();
      freeIds_.pop_back();
      pool_[id] = v;
    } else {
      id = (int)pool_.size();
      pool_.push_back(v);
      pos_.push_back(-1);
    }
    pos_[id] = (int)heap_.size();
    heap_.push_back(id);
    siftUp(pos_[id]);
    return id;
  }
  int topId() const { return heap_[0]; }
  T& get(int id) { return pool_[id]; }
  const T& get(int id) const { return pool_[id]; }
  void pop() { removeAt(0); }
  void erase(int id) { removeAt(pos_[id]); }
  void increase(int id) { siftUp(pos_[id]); }  // priority got better
  // ids in leaving order without modifying the heap (ordered_begin/end)
  template <class Visit>
  void orderedWalk(Visit visit) const {
    if (heap_.empty()) return;
    auto cmp = [&](int a, int b) {
      return better_(pool_[heap_[b]], pool_[heap_[a]]);
    };
    std::priority_queue<int, std::vector<int>, decltype(cmp)> pq(cmp);
    pq.push(0);
    while (!pq.empty()) {
      int h = pq.top();
      pq.pop();
      if (!visit(heap_[h])) return;
      int l = 2 * h + 1, r = l + 1;
      if (l < (int)heap_.size()) pq.push(l);
      if (r < (int)heap_.size()) pq.push(r);
    }
  }

 private:
  void swapAt(int a, int b) {
    std::swap(heap_[a], heap_[b]);
    pos_[heap_[a]] = a;
    pos_[heap_[b]] = b;
  }
  void siftUp(int h) {
    while (h > 0) {
      int p = (h - 1) / 2;
      if (!better_(pool_[heap_[h]], pool_[heap_[p]])) break;
      swapAt(h, p);
      h = p;
    }
  }
  void siftDown(int h) {
    const int n = (int)heap_.size();
    for (;;) {
      int l = 2 * h + 1, r = l + 1, c = l;
      if (l >= n) break;
      if (r < n && better_(pool_[heap_[r]], pool_[heap_[l]])) c = r;
      if (better_(pool_[heap_[h]], pool_[heap_[c]])) break;
      swapAt(h, c);
      h = c;
    }
  }
  void removeAt(int h) {
    int id = heap_[h];
    int last = (int)heap_.size() - 1;
    if (h != last) swapAt(h, last);
    heap_.pop_back();
    pos_[id] = -1;
    freeIds_.push_back(id);
    if (h < (int)heap_.size()) {
      siftUp(h);
      siftDown(h);
    }
  }
  Better better_;
  std::vector<T> pool_;
  std::vector<int> pos_;
  std::vector<int> heap_;
  std::vector<int> freeIds_;
};

// ---------------------------------------------------------------------------
// Low-level environment (example/cbs.cpp:247-444, example/ecbs.cpp:282-312,
// example/cbs_ta.cpp:283-367)
// ---------------------------------------------------------------------------
struct Constraints {
  // vertex key: time*cells + cell; edge key: (time*cells + from)*cells + to
  std::unordered_set<uint64_t> vc;
  std::unordered_set<uint64_t> ec;
  std::vector<std::pair<int, int>> vcList;  // (time, cell) for the goal scan
};

struct LowLevelEnv {
  const Grid* grid = nullptr;
  int variant = 0;    // 0 cbs/ecbs, 1 cbs_ta
  int goal = -1;      // cell, -1 = no task (cbs_ta only)
  const int32_t* field = nullptr;  // variant 1: distance-to-goal field
  const Constraints* cons = nullptr;
  int lastGoalConstraint = -1;
  // focal context (example/ecbs.cpp:282-312)
  const std::vector<Plan>* solution = nullptr;
  int agentIdx = 0;
  int64_t expanded = 0;

  void setContext(int agent, const Constraints* c, int goalCell) {
    // example/cbs.cpp:266-276 ; example/cbs_ta.cpp:283-304
    agentIdx = agent;
    cons = c;
    goal = goalCell;
    lastGoalConstraint = -1;
    for (const auto& v : c->vcList) {
      if (goal < 0 || v.second == goal)
        lastGoalConstraint = std::max(lastGoalConstraint, v.first);
    }
  }
  int h(int cell) const {
    if (variant == 0) {  // example/cbs.cpp:278-284 (Manhattan)
      int dx = cell % grid->dimx - goal % grid->dimx;
      int dy = cell / grid->dimx - goal / grid->dimx;
      return std::abs(dx) + std::abs(dy);
    }
    if (goal < 0) return 0;  // example/cbs_ta.cpp:305-311
    return field[cell];
  }
  bool isSolution(int time, int cell) const {
    bool atGoal = goal < 0 ? true : cell == goal;
    return atGoal && time > lastGoalConstraint;
  }
  bool stateValid(int time, int x, int y) const {
    if (!grid->freeCell(x, y)) return false;
    uint64_t key = (uint64_t)time * grid->cells() + (x + grid->dimx * y);
    return cons->vc.find(key) == cons->vc.end();
  }
  bool transitionValid(int time, int from, int to) const {
    uint64_t key =
        ((uint64_t)time * grid->cells() + from) * grid->cells() + to;
    return cons->ec.find(key) == cons->ec.end();
  }
  struct Nb {
    int cell;
    int cost;
  };
  // Successor order Wait, Left, Right, Up(y+1), Down(y-1):
  // example/cbs.cpp:299-332; free wait on the goal: example/cbs_ta.cpp:329-339
  int neighbors(int time, int cell, Nb* out) const {
    int n = 0;
    const int x = cell % grid->dimx, y = cell / grid->dimx;
    const int dx[5] = {0, -1, 1, 0, 0};
    const int dy[5] = {0, 0, 0, 1, -1};
    for (int k = 0; k < 5; ++k) {
      int nx = x + dx[k], ny = y + dy[k];
      if (!stateValid(time + 1, nx, ny)) continue;
      int nc = nx + grid->dimx * ny;
      if (!transitionValid(time, cell, nc)) continue;
      int cost = 1;
      if (k == 0 && variant == 1) {
        bool atGoal = goal < 0 ? true : cell == goal;
        cost = atGoal ? 0 : 1;
      }
      out[n++] = {nc, cost};
    }
    return n;
  }
  int focalState(int time, int cell) const {  // example/ecbs.cpp:282-295
    if (!solution) return 0;
    int n = 0;
    for (size_t i = 0; i < solution->size(); ++i) {
      if ((int)i == agentIdx || (*solution)[i].states.empty()) continue;
      if (posAt(*solution, i, time) == cell) ++n;
    }
    return n;
  }
  int focalTransition(int t1, int c1, int t2, int c2) const {
    // example/ecbs.cpp:298-312
    if (!solution) return 0;
    int n = 0;
    for (size_t i = 0; i < solution->size(); ++i) {
      if ((int)i == agentIdx || (*solution)[i].states.empty()) continue;
      int s2a = posAt(*solution, i, t1), s2b = posAt(*solution, i, t2);
      if (c1 == s2b && c2 == s2a) ++n;
    }
    return n;
  }
};

struct CameFrom {
  uint64_t parent;
  int g;
};

static inline uint64_t stateKey(const Grid& g, int time, int cell) {
  return (uint64_t)time * g.cells() + cell;
}

static void reconstruct(const Grid& grid,
                        const std::unordered_map<uint64_t, CameFrom>& cameFrom,
                        uint64_t goalKey, int startCell, int initialCost,
                        Plan& plan) {
  // a_star.hpp:89-104
  plan.states.clear();
  uint64_t k = goalKey;
  auto it = cameFrom.find(k);
  while (it != cameFrom.end()) {
    plan.states.push_back(
        {(int)(k / grid.cells()), (int)(k % grid.cells()), it->second.g});
    k = it->second.parent;
    it = cameFrom.find(k);
  }
  plan.states.push_back({0, startCell, initialCost});
  std::reverse(plan.states.begin(), plan.states.end());
}

struct ANode {
  uint64_t key;
  int f, g, focal;
};
struct OpenBetter {  // a_star.hpp:168-179: lowest f, then highest g
  bool operator()(const ANode& a, const ANode& b) const {
    if (a.f != b.f) return a.f < b.f;
    return a.g > b.g;
  }
};

// a_star.hpp:63-161.  status: ORC_SOLVED / ORC_NO_SOLUTION / ORC_CAPPED
static int aStar(LowLevelEnv& env, int startCell, Plan& plan,
                 int64_t maxExpanded) {
  const Grid& grid = *env.grid;
  plan.states.assign(1, {0, startCell, 0});
  plan.cost = 0;
  MutableHeap<ANode, OpenBetter> open;
  std::unordered_map<uint64_t, int> stateToHeap;
  std::unordered_set<uint64_t> closed;
  std::unordered_map<uint64_t, CameFrom> cameFrom;
  const uint64_t startKey = stateKey(grid, 0, startCell);
  stateToHeap[startKey] = open.push({startKey, env.h(startCell), 0, 0});
  LowLevelEnv::Nb nb[5];
  int64_t expandedHere = 0;
  while (!open.empty()) {
    ANode cur = open.get(open.topId());
    ++env.expanded;  // onExpandNode, a_star.hpp:87
    const int time = (int)(cur.key / grid.cells());
    const int cell = (int)(cur.key % grid.cells());
    if (env.isSolution(time, cell)) {
      reconstruct(grid, cameFrom, cur.key, startCell, 0, plan);
      plan.cost = cur.g;
      plan.fmin = cur.f;
      return ORC_SOLVED;
    }
    if (maxExpanded > 0 && ++expandedHere > maxExpanded) return ORC_CAPPED;
    open.pop();
    stateToHeap.erase(cur.key);
    closed.insert(cur.key);
    int n = env.neighbors(time, cell, nb);
    for (int k = 0; k < n; ++k) {
      uint64_t nk = stateKey(grid, time + 1, nb[k].cell);
      if (closed.count(nk)) continue;
      int tg = cur.g + nb[k].cost;
      auto it = stateToHeap.find(nk);
      if (it == stateToHeap.end()) {
        int h = env.h(nb[k].cell);
        int f = (h == ORC_INF) ? ORC_INF : tg + h;  // saturate (ref overflows)
        stateToHeap[nk] = open.push({nk, f, tg, 0});
      } else {
        ANode& o = open.get(it->second);
        if (tg >= o.g) continue;
        int delta = o.g - tg;
        o.g = tg;
        if (o.f != ORC_INF) o.f -= delta;
        open.increase(it->second);
      }
      cameFrom[nk] = {cur.key, tg};
    }
  }
  return ORC_NO_SOLUTION;
}

// a_star_epsilon.hpp:86-285 (the non-REBUILT_FOCAL_LIST branch, :134-154)
static int aStarEpsilon(LowLevelEnv& env, float w, int startCell, Plan& plan,
                        int64_t maxExpanded) {
  const Grid& grid = *env.grid;
  plan.states.assign(1, {0, startCell, 0});
  plan.cost = 0;
  MutableHeap<ANode, OpenBetter> open;
  // FOCAL holds OPEN handles ordered by (focal, f, -g):
  // a_star_epsilon.hpp:346-366
  auto focalBetter = [&open](int a, int b) {
    const ANode& x = open.get(a);
    const ANode& y = open.get(b);
    if (x.focal != y.focal) return x.focal < y.focal;
    if (x.f != y.f) return x.f < y.f;
    return x.g > y.g;
  };
  MutableHeap<int, decltype(focalBetter)> focal(focalBetter);
  std::unordered_map<uint64_t, int> stateToHeap;
  std::unordered_map<int, int> focalHandleOf;  // open id -> focal id
  std::unordered_set<uint64_t> closed;
  std::unordered_map<uint64_t, CameFrom> cameFrom;
  const uint64_t startKey = stateKey(grid, 0, startCell);
  int h0 = open.push({startKey, env.h(startCell), 0, 0});
  stateToHeap[startKey] = h0;
  focalHandleOf[h0] = focal.push(h0);
  int bestF = open.get(h0).f;
  LowLevelEnv::Nb nb[5];
  int64_t expandedHere = 0;
  while (!open.empty()) {
    {  // incremental focal refill when min f rises, :134-154 (fp32 compare)
      int oldBest = bestF;
      bestF = open.get(open.topId()).f;
      if (bestF > oldBest) {
        open.orderedWalk([&](int id) {
          int val = open.get(id).f;
          if (val > oldBest * w && val <= bestF * w)
            focalHandleOf[id] = focal.push(id);
          return !(val > bestF * w);
        });
      }
    }
    const int curId = focal.get(focal.topId());
    ANode cur = open.get(curId);
    ++env.expanded;
    const int time = (int)(cur.key / grid.cells());
    const int cell = (int)(cur.key % grid.cells());
    if (env.isSolution(time, cell)) {
      reconstruct(grid, cameFrom, cur.key, startCell, 0, plan);
      plan.cost = cur.g;
      plan.fmin = open.get(open.topId()).f;  // :210
      return ORC_SOLVED;
    }
    if (maxExpanded > 0 && ++expandedHere > maxExpanded) return ORC_CAPPED;
    focal.pop();
    focalHandleOf.erase(curId);
    open.erase(curId);
    stateToHeap.erase(cur.key);
    closed.insert(cur.key);
    int n = env.neighbors(time, cell, nb);
    for (int k = 0; k < n; ++k) {
      uint64_t nk = stateKey(grid, time + 1, nb[k].cell);
      if (closed.count(nk)) continue;
      int tg = cur.g + nb[k].cost;
      auto it = stateToHeap.find(nk);
      if (it == stateToHeap.end()) {
        int f = tg + env.h(nb[k].cell);
        int fh = cur.focal + env.focalState(time + 1, nb[k].cell) +
                 env.focalTransition(time, cell, time + 1, nb[k].cell);
        int id = open.push({nk, f, tg, fh});
        if (f <= bestF * w) focalHandleOf[id] = focal.push(id);
        stateToHeap[nk] = id;
      } else {
        ANode& o = open.get(it->second);
        if (tg >= o.g) continue;
        int lastF = o.f;
        int delta = o.g - tg;
        o.g = tg;
        o.f -= delta;
        open.increase(it->second);
        // FOCAL is deliberately not re-ordered here (:248-270)
        if (o.f <= bestF * w && lastF > bestF * w)
          focalHandleOf[it->second] = focal.push(it->second);
      }
      cameFrom[nk] = {cur.key, tg};
    }
  }
  return ORC_NO_SOLUTION;
}

// ---------------------------------------------------------------------------
// High level: shared pieces
// ---------------------------------------------------------------------------
struct Caps {
  int64_t maxHL = 0, maxLL = 0;
  double maxSeconds = 0;
  std::chrono::steady_clock::time_point t0;
  void start() { t0 = std::chrono::steady_clock::now(); }
  double elapsed() const {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0)
        .count();
  }
  bool timeUp() const { return maxSeconds > 0 && elapsed() > maxSeconds; }
};
static Caps toCaps(const orc_caps* c) {
  Caps k;
  if (c) {
    k.maxHL = c->max_hl_expanded;
    k.maxLL = c->max_ll_expanded;
    k.maxSeconds = c->max_seconds;
  }
  return k;
}

static void addVertexConstraint(const Grid& g, Constraints& c, int t, int cell) {
  c.vc.insert((uint64_t)t * g.cells() + cell);
  c.vcList.push_back({t, cell});
}
static void addEdgeConstraint(const Grid& g, Constraints& c, int t, int from,
                              int to) {
  c.ec.insert(((uint64_t)t * g.cells() + from) * g.cells() + to);
}

// createConstraintsFromConflict, example/cbs.cpp:388-406.  Returned in
// ascending agent order (std::map iteration, cbs.hpp:140).
struct NewConstraint {
  int agent;
  bool edge;
  int t, a, b;
};
static int constraintsFromConflict(const Grid& g, const orc_conflict& c,
                                   NewConstraint out[2]) {
  int c1 = c.x1 + g.dimx * c.y1;
  if (c.type == 0) {
    out[0] = {c.agent1, false, c.time, c1, -1};
    out[1] = {c.agent2, false, c.time, c1, -1};
  } else {
    int c2 = c.x2 + g.dimx * c.y2;
    out[0] = {c.agent1, true, c.time, c1, c2};
    out[1] = {c.agent2, true, c.time, c2, c1};
  }
  return 2;  // agent1 < agent2 always
}

static bool firstConflict(const Grid& g, const std::vector<Plan>& sol, int mode,
                          orc_conflict* out) {
  return firstConflictT(
      (int)sol.size(), maxT(sol, mode), g.dimx,
      [&](int i, int t) { return posAt(sol, i, t); }, out);
}
static int countConflicts(const std::vector<Plan>& sol) {
  return countConflictsT((int)sol.size(), maxT(sol, 0),
                         [&](int i, int t) { return posAt(sol, i, t); });
}

static void exportPaths(const Grid& g, const std::vector<Plan>& sol,
                        int32_t* path_off, int32_t* path_xyg, int path_cap) {
  if (!path_off) return;
  int off = 0;
  for (size_t i = 0; i < sol.size(); ++i) {
    path_off[i] = off;
    for (const auto& s : sol[i].states) {
      if (path_xyg && off < path_cap) {
        path_xyg[3 * off + 0] = s.cell % g.dimx;
        path_xyg[3 * off + 1] = s.cell / g.dimx;
        path_xyg[3 * off + 2] = s.g;  // "t" of output.yaml, example/cbs.cpp:659
      }
      ++off;
    }
  }
  path_off[sol.size()] = off;
}

static void finishResult(const std::vector<Plan>& sol, orc_result* res) {
  res->cost = 0;
  res->makespan = 0;
  res->lower_bound = 0;
  for (const auto& p : sol) {  // example/cbs.cpp:630-635
    res->cost += p.cost;
    res->makespan = std::max<int64_t>(res->makespan, p.cost);
    res->lower_bound += p.fmin;
  }
}

// ---------------------------------------------------------------------------
// CBS — cbs.hpp:85-172
// ---------------------------------------------------------------------------
struct HLNode {
  std::vector<Plan> solution;
  std::vector<Constraints> constraints;
  std::map<int, int> tasks;  // cbs_ta: agent -> goal cell
  int cost = 0;
  int LB = 0;
  int focal = 0;
  int id = 0;
  bool isRoot = false;
};
struct HLBetter {  // cbs.hpp:187-191: lowest cost
  bool operator()(const HLNode& a, const HLNode& b) const {
    return a.cost < b.cost;
  }
};

static int cbsSearch(const Grid& grid, const std::vector<int>& starts,
                     const std::vector<int>& goals, Caps caps,
                     std::vector<Plan>& solution, orc_result* res) {
  const size_t N = starts.size();
  LowLevelEnv env;
  env.grid = &grid;
  env.variant = 0;
  HLNode start;
  start.solution.resize(N);
  start.constraints.resize(N);
  int64_t hl = 0;
  auto fail = [&](int status) {
    res->hl_expanded = hl;
    res->ll_expanded = env.expanded;
    return status;
  };
  for (size_t i = 0; i < N; ++i) {
    env.setContext((int)i, &start.constraints[i], goals[i]);
    int st = aStar(env, starts[i], start.solution[i], caps.maxLL);
    if (st != ORC_SOLVED) return fail(st);
    start.cost += start.solution[i].cost;
  }
  MutableHeap<HLNode, HLBetter> open;
  open.push(start);
  int id = 1;
  while (!open.empty()) {
    if ((caps.maxHL > 0 && hl >= caps.maxHL) || caps.timeUp())
      return fail(ORC_CAPPED);
    HLNode P = open.get(open.topId());
    ++hl;  // onExpandHighLevelNode, cbs.hpp:121
    open.pop();
    orc_conflict conflict;
    if (!firstConflict(grid, P.solution, 0, &conflict)) {
      solution = P.solution;
      return fail(ORC_SOLVED);
    }
    NewConstraint nc[2];
    int n = constraintsFromConflict(grid, conflict, nc);
    for (int k = 0; k < n; ++k) {
      int i = nc[k].agent;
      HLNode child = P;
      child.id = id;
      if (nc[k].edge)
        addEdgeConstraint(grid, child.constraints[i], nc[k].t, nc[k].a, nc[k].b);
      else
        addVertexConstraint(grid, child.constraints[i], nc[k].t, nc[k].a);
      child.cost -= child.solution[i].cost;
      env.setContext(i, &child.constraints[i], goals[i]);
      int st = aStar(env, starts[i], child.solution[i], caps.maxLL);
      if (st == ORC_CAPPED) return fail(ORC_CAPPED);
      child.cost += child.solution[i].cost;
      if (st == ORC_SOLVED) open.push(child);
      ++id;
    }
  }
  return fail(ORC_NO_SOLUTION);
}

// ---------------------------------------------------------------------------
// ECBS — ecbs.hpp:109-288 (non-REBUILT_FOCAL_LIST branch)
// ---------------------------------------------------------------------------
static int ecbsSearch(const Grid& grid, const std::vector<int>& starts,
                      const std::vector<int>& goals, float w, Caps caps,
                      std::vector<Plan>& solution, orc_result* res) {
  const size_t N = starts.size();
  LowLevelEnv env;
  env.grid = &grid;
  env.variant = 0;
  HLNode start;
  start.solution.resize(N);
  start.constraints.resize(N);
  int64_t hl = 0;
  auto fail = [&](int status) {
    res->hl_expanded = hl;
    res->ll_expanded = env.expanded;
    return status;
  };
  for (size_t i = 0; i < N; ++i) {  // ecbs.hpp:118-136
    env.setContext((int)i, &start.constraints[i], goals[i]);
    env.solution = &start.solution;
    int st = aStarEpsilon(env, w, starts[i], start.solution[i], caps.maxLL);
    if (st != ORC_SOLVED) return fail(st);
    start.cost += start.solution[i].cost;
    start.LB += start.solution[i].fmin;
  }
  start.focal = countConflicts(start.solution);  // ecbs.hpp:137

  MutableHeap<HLNode, HLBetter> open;
  auto focalBetter = [&open](int a, int b) {  // ecbs.hpp:344-352
    const HLNode& x = open.get(a);
    const HLNode& y = open.get(b);
    if (x.focal != y.focal) return x.focal < y.focal;
    return x.cost < y.cost;
  };
  MutableHeap<int, decltype(focalBetter)> focal(focalBetter);
  int h0 = open.push(start);
  focal.push(h0);
  int bestCost = open.get(h0).cost;
  int id = 1;
  while (!open.empty()) {
    if ((caps.maxHL > 0 && hl >= caps.maxHL) || caps.timeUp())
      return fail(ORC_CAPPED);
    {  // ecbs.hpp:170-190
      int oldBest = bestCost;
      bestCost = open.get(open.topId()).cost;
      if (bestCost > oldBest) {
        open.orderedWalk([&](int nid) {
          int val = open.get(nid).cost;
          if (val > oldBest * w && val <= bestCost * w) focal.push(nid);
          return !(val > bestCost * w);
        });
      }
    }
    int hid = focal.get(focal.topId());
    HLNode P = open.get(hid);
    ++hl;
    focal.pop();
    open.erase(hid);
    orc_conflict conflict;
    if (!firstConflict(grid, P.solution, 0, &conflict)) {
      solution = P.solution;
      return fail(ORC_SOLVED);
    }
    NewConstraint nc[2];
    int n = constraintsFromConflict(grid, conflict, nc);
    for (int k = 0; k < n; ++k) {
      int i = nc[k].agent;
      HLNode child = P;
      child.id = id;
      if (nc[k].edge)
        addEdgeConstraint(grid, child.constraints[i], nc[k].t, nc[k].a, nc[k].b);
      else
        addVertexConstraint(grid, child.constraints[i], nc[k].t, nc[k].a);
      child.cost -= child.solution[i].cost;
      child.LB -= child.solution[i].fmin;
      env.setContext(i, &child.constraints[i], goals[i]);
      env.solution = &child.solution;
      int st = aStarEpsilon(env, w, starts[i], child.solution[i], caps.maxLL);
      if (st == ORC_CAPPED) return fail(ORC_CAPPED);
      child.cost += child.solution[i].cost;
      child.LB += child.solution[i].fmin;
      child.focal = countConflicts(child.solution);
      if (st == ORC_SOLVED) {
        int cid = open.push(child);
        if (child.cost <= bestCost * w) focal.push(cid);
      }
      ++id;
    }
  }
  return fail(ORC_NO_SOLUTION);
}

// ---------------------------------------------------------------------------
// Assignment — assignment.hpp:34-118 (min-cost max-flow by successive shortest
// paths; Boost's successive_shortest_path_nonnegative_weights restated as
// Dijkstra with potentials) and next_best_assignment.hpp:37-201.
// The optimal COST is unique; which optimum is returned among ties is unpinned.
// ---------------------------------------------------------------------------
class Assignment {
 public:
  void clear() { edges_.clear(); }  // keeps agents/tasks, assignment.hpp:46-61
  void setCost(int agent, int task, long cost) {
    if (!agentIdx_.count(agent)) {
      agentIdx_[agent] = (int)agents_.size();
      agents_.push_back(agent);
    }
    if (!taskIdx_.count(task)) {
      taskIdx_[task] = (int)tasks_.size();
      tasks_.push_back(task);
    }
    edges_[{agentIdx_[agent], taskIdx_[task]}] = cost;
  }
  long solve(std::map<int, int>& solution) {
    solution.clear();
    const int A = (int)agents_.size(), T = (int)tasks_.size();
    // vertices: 0 = source, 1 = sink, 2.. agents, then tasks
    const int V = 2 + A + T;
    struct E {
      int to;
      long cap, cost;
      int rev;
    };
    std::vector<std::vector<E>> adj(V);
    auto addEdge = [&](int u, int v, long cost) {
      adj[u].push_back({v, 1, cost, (int)adj[v].size()});
      adj[v].push_back({u, 0, -cost, (int)adj[u].size() - 1});
    };
    for (int a = 0; a < A; ++a) addEdge(0, 2 + a, 0);
    for (int t = 0; t < T; ++t) addEdge(2 + A + t, 1, 0);
    for (const auto& e : edges_)
      addEdge(2 + e.first.first, 2 + A + e.first.second, e.second);
    std::vector<long> pot(V, 0), dist(V);
    std::vector<int> prevV(V), prevE(V);
    const long INF = LONG_MAX / 4;
    for (;;) {
      std::fill(dist.begin(), dist.end(), INF);
      dist[0] = 0;
      typedef std::pair<long, int> QE;
      std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
      pq.push({0, 0});
      while (!pq.empty()) {
        auto [d, u] = pq.top();
        pq.pop();
        if (d > dist[u]) continue;
        for (int k = 0; k < (int)adj[u].size(); ++k) {
          const E& e = adj[u][k];
          if (e.cap <= 0) continue;
          long nd = d + e.cost + pot[u] - pot[e.to];
          if (nd < dist[e.to]) {
            dist[e.to] = nd;
            prevV[e.to] = u;
            prevE[e.to] = k;
            pq.push({nd, e.to});
          }
        }
      }
      if (dist[1] >= INF) break;
      for (int v = 0; v < V; ++v)
        if (dist[v] < INF) pot[v] += dist[v];
      for (int v = 1; v != 0; v = prevV[v]) {
        E& e = adj[prevV[v]][prevE[v]];
        e.cap -= 1;
        adj[v][e.rev].cap += 1;
      }
    }
    long cost = 0;
    for (int a = 0; a < A; ++a) {  // assignment.hpp:94-115
      for (const E& e : adj[2 + a]) {
        if (e.to >= 2 + A && e.cap == 0) {
          // forward agent->task edge that is saturated (residual == 0)
          solution[agents_[a]] = tasks_[e.to - 2 - A];
          cost += e.cost;
          break;
        }
      }
    }
    return cost;
  }

 private:
  std::vector<int> agents_, tasks_;
  std::map<int, int> agentIdx_, taskIdx_;
  std::map<std::pair<int, int>, long> edges_;
};

class NextBestAssignment {
 public:
  void setCost(int agent, int task, long cost) {  // next_best_assignment.hpp:41-47
    cost_[{agent, task}] = cost;
    if (!agentsSet_.count(agent)) {
      agentsSet_.insert(agent);
      agentsVec_.push_back(agent);
    }
  }
  void solve() {  // :49-56
    Node n;
    n.cost = constrainedMatching(n.I, n.O, n.Iagents, n.Oagents, n.solution);
    open_.push(n);
    numMatching_ = n.solution.size();
  }
  long nextSolution(std::map<int, int>& solution) {  // :59-122
    solution.clear();
    if (open_.empty()) return LONG_MAX;
    const Node next = open_.top();
    open_.pop();
    solution = next.solution;
    long result = next.cost;
    std::set<int> fixedAgents;
    for (const auto& c : next.I) fixedAgents.insert(c.first);
    for (size_t i = 0; i < agentsVec_.size(); ++i) {
      if (fixedAgents.count(agentsVec_[i])) continue;
      Node n;
      n.I = next.I;
      n.O = next.O;
      n.Iagents = next.Iagents;
      n.Oagents = next.Oagents;
      for (size_t j = 0; j < i; ++j) {
        int agent = agentsVec_[j];
        auto it = solution.find(agent);
        if (it != solution.end())
          n.I.insert({agent, it->second});
        else
          n.Oagents.insert(agent);
      }
      auto it = solution.find(agentsVec_[i]);
      if (it != solution.end())
        n.O.insert({agentsVec_[i], it->second});
      else
        n.Iagents.insert(agentsVec_[i]);
      n.cost = constrainedMatching(n.I, n.O, n.Iagents, n.Oagents, n.solution);
      if (!n.solution.empty()) open_.push(n);
    }
    return result;
  }

 private:
  typedef std::set<std::pair<int, int>> PairSet;
  long constrainedMatching(const PairSet& I, const PairSet& O,
                           const std::set<int>& Iagents,
                           const std::set<int>& Oagents,
                           std::map<int, int>& solution) {  // :129-189
    assignment_.clear();
    for (const auto& c : I)
      if (!Oagents.count(c.first)) assignment_.setCost(c.first, c.second, 0);
    for (const auto& c : cost_) {
      if (!O.count(c.first) && !I.count(c.first) &&
          !Oagents.count(c.first.first)) {
        long costOffset = 1000000000L;  // :148
        if (Iagents.count(c.first.first)) costOffset = 0;
        assignment_.setCost(c.first.first, c.first.second,
                            c.second + costOffset);
      }
    }
    assignment_.solve(solution);
    size_t matching = solution.size();
    bool valid = true;
    for (int agent : Iagents)
      if (!solution.count(agent)) {
        valid = false;
        break;
      }
    for (const auto& c : I) {
      auto it = solution.find(c.first);
      if (it == solution.end() || it->second != c.second) {
        valid = false;
        break;
      }
    }
    if (!valid || matching < numMatching_) {
      solution.clear();
      return LONG_MAX;
    }
    long result = 0;
    for (const auto& e : solution) result += cost_.at(e);
    return result;
  }
  struct Node {
    PairSet I, O;
    std::set<int> Iagents, Oagents;
    std::map<int, int> solution;
    long cost = 0;
    bool operator<(const Node& n) const { return cost > n.cost; }
  };
  Assignment assignment_;
  std::map<std::pair<int, int>, long> cost_;
  std::vector<int> agentsVec_;
  std::set<int> agentsSet_;
  std::priority_queue<Node> open_;
  size_t numMatching_ = 0;
};

// ---------------------------------------------------------------------------
// CBS-TA — cbs_ta.hpp:87-214 with the Environment of example/cbs_ta.cpp:252-514
// ---------------------------------------------------------------------------
static int cbsTaSearch(const Grid& grid, const std::vector<int>& starts,
                       const std::vector<std::vector<int>>& potentialGoals,
                       int64_t maxTaskAssignments, Caps caps,
                       std::vector<Plan>& solution, orc_result* res) {
  const size_t N = starts.size();
  const int V = grid.cells();
  // Environment ctor (example/cbs_ta.cpp:254-281): all-pairs heuristic, cost
  // matrix from start->goal distances, first assignment.  Outside the timer.
  std::vector<int32_t> apsp((size_t)V * V);
  floydWarshall(grid, apsp.data());
  NextBestAssignment nba;
  for (size_t i = 0; i < N; ++i)
    for (int goal : potentialGoals[i])
      nba.setCost((int)i, goal, apsp[(size_t)starts[i] * V + goal]);
  nba.solve();
  int64_t numTA = 0;
  auto nextTaskAssignment = [&](std::map<int, int>& tasks) {
    // example/cbs_ta.cpp:442-456
    if ((uint64_t)numTA > (uint64_t)maxTaskAssignments) return;
    nba.nextSolution(tasks);
    if (!tasks.empty()) ++numTA;
  };

  caps.start();  // Timer starts after the Environment ctor, cbs_ta.cpp:576
  LowLevelEnv env;
  env.grid = &grid;
  env.variant = 1;
  int64_t hl = 0;
  auto fail = [&](int status) {
    res->hl_expanded = hl;
    res->ll_expanded = env.expanded;
    res->n_task_assignments = numTA;
    res->runtime_s = caps.elapsed();
    return status;
  };
  auto taskOf = [](const HLNode& n, int i) {
    auto it = n.tasks.find(i);
    return it == n.tasks.end() ? -1 : it->second;
  };
  auto setCtx = [&](const HLNode& n, int i) {
    int goal = taskOf(n, i);
    env.setContext(i, &n.constraints[i], goal);
    env.field = goal >= 0 ? apsp.data() + (size_t)goal * V : nullptr;
  };

  HLNode start;
  start.solution.resize(N);
  start.constraints.resize(N);
  start.isRoot = true;
  nextTaskAssignment(start.tasks);
  for (size_t i = 0; i < N; ++i) {  // cbs_ta.hpp:98-114
    if (start.tasks.empty()) return fail(ORC_NO_SOLUTION);
    setCtx(start, (int)i);
    int st = aStar(env, starts[i], start.solution[i], caps.maxLL);
    if (st != ORC_SOLVED) return fail(st);
    start.cost += start.solution[i].cost;
  }
  MutableHeap<HLNode, HLBetter> open;
  open.push(start);
  int id = 1;
  while (!open.empty()) {
    if ((caps.maxHL > 0 && hl >= caps.maxHL) || caps.timeUp())
      return fail(ORC_CAPPED);
    HLNode P = open.get(open.topId());
    ++hl;
    open.pop();
    orc_conflict conflict;
    if (!firstConflict(grid, P.solution, 1, &conflict)) {
      solution = P.solution;
      return fail(ORC_SOLVED);
    }
    if (P.isRoot) {  // cbs_ta.hpp:142-172
      HLNode n;
      nextTaskAssignment(n.tasks);
      if (!n.tasks.empty()) {
        n.solution.resize(N);
        n.constraints.resize(N);
        n.id = id;
        n.isRoot = true;
        bool all = true;
        for (size_t i = 0; i < N; ++i) {
          setCtx(n, (int)i);
          int st = aStar(env, starts[i], n.solution[i], caps.maxLL);
          if (st == ORC_CAPPED) return fail(ORC_CAPPED);
          if (st != ORC_SOLVED) {
            all = false;
            break;
          }
          n.cost += n.solution[i].cost;
        }
        if (all) {
          open.push(n);
          ++id;
        }
      }
    }
    NewConstraint nc[2];
    int n = constraintsFromConflict(grid, conflict, nc);
    for (int k = 0; k < n; ++k) {
      int i = nc[k].agent;
      HLNode child = P;
      child.id = id;
      // `newNode = P` (cbs_ta.hpp:180) also copies isRoot and nothing resets
      // it, so in the reference every descendant of a root counts as a root
      // and spawns the next-best assignment when expanded.  Kept as is.
      if (nc[k].edge)
        addEdgeConstraint(grid, child.constraints[i], nc[k].t, nc[k].a, nc[k].b);
      else
        addVertexConstraint(grid, child.constraints[i], nc[k].t, nc[k].a);
      child.cost -= child.solution[i].cost;
      setCtx(child, i);
      int st = aStar(env, starts[i], child.solution[i], caps.maxLL);
      if (st == ORC_CAPPED) return fail(ORC_CAPPED);
      child.cost += child.solution[i].cost;
      if (st == ORC_SOLVED) open.push(child);
      ++id;
    }
  }
  return fail(ORC_NO_SOLUTION);
}


// ---------------------------------------------------------------------------
// ECBS-TA — ecbs_ta.hpp:94-352 with REBUILT_FOCAL_LIST and STYLE_MINROOT (both
// defined on, ecbs_ta.hpp:8,11) and the Environment of example/ecbs_ta.cpp
// (cbs_ta moves + the ecbs focal heuristics; getFirstConflict with the
// cbs_ta bound, focalHeuristic with the cbs bound — ecbs_ta.cpp:353,445).
// ---------------------------------------------------------------------------
static int ecbsTaSearch(const Grid& grid, const std::vector<int>& starts,
                        const std::vector<std::vector<int>>& potentialGoals, float w,
                        int64_t maxTaskAssignments, Caps caps, std::vector<Plan>& solution,
                        orc_result* res) {
  const size_t N = starts.size();
  const int V = grid.cells();
  std::vector<int32_t> apsp((size_t)V * V);
  floydWarshall(grid, apsp.data());
  NextBestAssignment nba;
  for (size_t i = 0; i < N; ++i)
    for (int goal : potentialGoals[i])
      nba.setCost((int)i, goal, apsp[(size_t)starts[i] * V + goal]);
  nba.solve();
  int64_t numTA = 0;
  auto nextTaskAssignment = [&](std::map<int, int>& tasks) {
    if ((uint64_t)numTA > (uint64_t)maxTaskAssignments) return;
    nba.nextSolution(tasks);
    if (!tasks.empty()) ++numTA;
  };
  caps.start();
  LowLevelEnv env;
  env.grid = &grid;
  env.variant = 1;
  int64_t hl = 0;
  auto fail = [&](int status) {
    res->hl_expanded = hl;
    res->ll_expanded = env.expanded;
    res->n_task_assignments = numTA;
    res->runtime_s = caps.elapsed();
    return status;
  };
  auto planAll = [&](HLNode& n) -> int {  // agents one after the other (ecbs_ta.hpp:107-126)
    for (size_t i = 0; i < N; ++i) {
      auto it = n.tasks.find((int)i);
      const int goal = it == n.tasks.end() ? -1 : it->second;
      env.setContext((int)i, &n.constraints[i], goal);
      env.field = goal >= 0 ? apsp.data() + (size_t)goal * V : nullptr;
      env.solution = &n.solution;
      int st = aStarEpsilon(env, w, starts[i], n.solution[i], caps.maxLL);
      if (st != ORC_SOLVED) return st;
      n.cost += n.solution[i].cost;
      n.LB += n.solution[i].fmin;
    }
    n.focal = countConflicts(n.solution);
    return ORC_SOLVED;
  };
  HLNode start;
  start.solution.resize(N);
  start.constraints.resize(N);
  start.isRoot = true;
  nextTaskAssignment(start.tasks);
  {
    int st = planAll(start);
    if (st != ORC_SOLVED) return fail(st);
  }
  int nextRootNodeCost = (int)(start.LB * w);  // Cost = float product, ecbs_ta.hpp:142
  std::vector<HLNode> open;  // small: linear scans stand in for the two heaps
  open.push_back(start);
  int id = 1;
  auto topOf = [&]() {
    size_t b = 0;
    for (size_t k = 1; k < open.size(); ++k)
      if (open[k].cost < open[b].cost) b = k;
    return b;
  };
  while (!open.empty()) {
    if ((caps.maxHL > 0 && hl >= caps.maxHL) || caps.timeUp()) return fail(ORC_CAPPED);
    // FOCAL rebuilt every iteration: nodes with cost <= nextRootNodeCost,
    // best by (focalHeuristic, cost) — ecbs_ta.hpp:160-181,420-428
    int best = -1;
    for (size_t k = 0; k < open.size(); ++k) {
      if (!((float)open[k].cost <= (float)nextRootNodeCost)) continue;
      if (best < 0 || std::make_pair(open[k].focal, open[k].cost) <
                          std::make_pair(open[best].focal, open[best].cost))
        best = (int)k;
    }
    if (best < 0) return fail(ORC_NO_SOLUTION);  // the reference would dereference an empty heap
    HLNode P = open[best];
    open.erase(open.begin() + best);
    ++hl;
    orc_conflict conflict;
    if (!firstConflict(grid, P.solution, 1, &conflict)) {
      solution = P.solution;
      return fail(ORC_SOLVED);
    }
    NewConstraint nc[2];
    int n = constraintsFromConflict(grid, conflict, nc);
    for (int k = 0; k < n; ++k) {
      int i = nc[k].agent;
      HLNode child = P;
      child.id = id;
      if (nc[k].edge)
        addEdgeConstraint(grid, child.constraints[i], nc[k].t, nc[k].a, nc[k].b);
      else
        addVertexConstraint(grid, child.constraints[i], nc[k].t, nc[k].a);
      child.cost -= child.solution[i].cost;
      child.LB -= child.solution[i].fmin;
      auto it = child.tasks.find(i);
      const int goal = it == child.tasks.end() ? -1 : it->second;
      env.setContext(i, &child.constraints[i], goal);
      env.field = goal >= 0 ? apsp.data() + (size_t)goal * V : nullptr;
      env.solution = &child.solution;
      int st = aStarEpsilon(env, w, starts[i], child.solution[i], caps.maxLL);
      if (st == ORC_CAPPED) return fail(ORC_CAPPED);
      child.cost += child.solution[i].cost;
      child.LB += child.solution[i].fmin;
      child.focal = countConflicts(child.solution);
      if (st == ORC_SOLVED) open.push_back(child);
      ++id;
    }
    // STYLE_MINROOT: a new root once the cheapest open node exceeds the bound
    // (ecbs_ta.hpp:299-348)
    if (open.empty()) return fail(ORC_NO_SOLUTION);
    if (open[topOf()].cost > nextRootNodeCost) {
      HLNode r;
      nextTaskAssignment(r.tasks);
      if (!r.tasks.empty()) {
        r.solution.resize(N);
        r.constraints.resize(N);
        r.id = id;
        r.isRoot = true;
        int st = planAll(r);
        if (st == ORC_CAPPED) return fail(ORC_CAPPED);
        if (st == ORC_SOLVED) {
          open.push_back(r);
          ++id;
        }
      }
      nextRootNodeCost = (int)(open[topOf()].LB * w);
    }
  }
  return fail(ORC_NO_SOLUTION);
}

}  // namespace orc

// ===========================================================================
// C interface
// ===========================================================================
using namespace orc;

extern "C" {

int orc_floyd_warshall(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                       int32_t* out) {
  Grid g = makeGrid(dimx, dimy, obst_xy, n_obst);
  floydWarshall(g, out);
  return 0;
}

int orc_bfs_fields(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                   const int32_t* goal_xy, int n_goals, int32_t* out) {
  Grid g = makeGrid(dimx, dimy, obst_xy, n_obst);
  for (int k = 0; k < n_goals; ++k)
    bfsField(g, goal_xy[2 * k], goal_xy[2 * k + 1],
             out + (size_t)k * g.cells());
  return 0;
}

int orc_first_conflict(const int32_t* cell, const int32_t* len, int N, int Tpad,
                       int dimx, int mode, orc_conflict* out) {
  int max_t = 0;
  for (int i = 0; i < N; ++i)
    max_t = std::max(max_t, len[i] - (mode == 0 ? 1 : 0));
  orc_conflict c;
  std::memset(&c, 0xff, sizeof c);
  bool found = firstConflictT(
      N, max_t, dimx,
      [&](int i, int t) { return posAtTable(cell, len, Tpad, i, t); }, &c);
  *out = c;
  return found ? 1 : 0;
}

int orc_count_conflicts(const int32_t* cell, const int32_t* len, int N,
                        int Tpad, int mode, int32_t* count) {
  int max_t = 0;
  for (int i = 0; i < N; ++i)
    max_t = std::max(max_t, len[i] - (mode == 0 ? 1 : 0));
  *count = countConflictsT(N, max_t, [&](int i, int t) {
    return posAtTable(cell, len, Tpad, i, t);
  });
  return 0;
}

int orc_focal_counts(const int32_t* cell, const int32_t* len, int N, int Tpad,
                     int self, const int32_t* cand_t, const int32_t* cand_from,
                     const int32_t* cand_to, int n_cand, int32_t* state_cnt,
                     int32_t* trans_cnt) {
  for (int k = 0; k < n_cand; ++k) {
    int t = cand_t[k], s = 0, tr = 0;
    for (int i = 0; i < N; ++i) {
      if (i == self || len[i] <= 0) continue;  // example/ecbs.cpp:287
      int pa = posAtTable(cell, len, Tpad, i, t);
      int pb = posAtTable(cell, len, Tpad, i, t + 1);
      if (pb == cand_to[k]) ++s;                       // ecbs.cpp:288-291
      if (cand_from[k] == pb && cand_to[k] == pa) ++tr;  // ecbs.cpp:304-308
    }
    state_cnt[k] = s;
    trans_cnt[k] = tr;
  }
  return 0;
}

int orc_lowlevel(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                 int variant, int start_cell, int goal_cell, const int32_t* vc,
                 int n_vc, const int32_t* ec, int n_ec, float w,
                 const int32_t* oth_cell, const int32_t* oth_len, int oth_n,
                 int oth_tpad, int self, int64_t max_expanded, int32_t* cost,
                 int32_t* fmin, int64_t* expanded, int32_t* path_tcg,
                 int path_cap, int32_t* path_len) {
  Grid g = makeGrid(dimx, dimy, obst_xy, n_obst);
  Constraints c;
  for (int k = 0; k < n_vc; ++k)
    addVertexConstraint(g, c, vc[2 * k], vc[2 * k + 1]);
  for (int k = 0; k < n_ec; ++k)
    addEdgeConstraint(g, c, ec[3 * k], ec[3 * k + 1], ec[3 * k + 2]);
  std::vector<int32_t> field;
  LowLevelEnv env;
  env.grid = &g;
  env.variant = variant;
  if (variant == 1 && goal_cell >= 0) {
    field.resize(g.cells());
    bfsField(g, goal_cell % dimx, goal_cell / dimx, field.data());
    env.field = field.data();
  }
  std::vector<Plan> others;
  if (oth_cell && oth_n > 0) {
    others.resize(oth_n);
    for (int i = 0; i < oth_n; ++i)
      for (int t = 0; t < oth_len[i]; ++t)
        others[i].states.push_back({t, oth_cell[(size_t)i * oth_tpad + t], t});
    env.solution = &others;
  }
  env.setContext(self, &c, goal_cell);
  Plan plan;
  int st = (w > 0) ? aStarEpsilon(env, w, start_cell, plan, max_expanded)
                   : aStar(env, start_cell, plan, max_expanded);
  if (expanded) *expanded = env.expanded;
  if (st != ORC_SOLVED) return st;
  *cost = plan.cost;
  *fmin = plan.fmin;
  if (path_len) *path_len = (int)plan.states.size();
  if (path_tcg)
    for (size_t k = 0; k < plan.states.size() && (int)k < path_cap; ++k) {
      path_tcg[3 * k] = plan.states[k].time;
      path_tcg[3 * k + 1] = plan.states[k].cell;
      path_tcg[3 * k + 2] = plan.states[k].g;
    }
  return ORC_SOLVED;
}

static void unpackInstance(const orc_instance* inst, Grid& g,
                           std::vector<int>& starts, std::vector<int>& goals) {
  g = makeGrid(inst->dimx, inst->dimy, inst->obst_xy, inst->n_obst);
  for (int i = 0; i < inst->n_agents; ++i) {
    starts.push_back(inst->start_xy[2 * i] + g.dimx * inst->start_xy[2 * i + 1]);
    if (inst->goal_xy)
      goals.push_back(inst->goal_xy[2 * i] + g.dimx * inst->goal_xy[2 * i + 1]);
  }
}

int orc_cbs(const orc_instance* inst, const orc_caps* caps, orc_result* res,
            int32_t* path_off, int32_t* path_xyg, int path_cap) {
  Grid g;
  std::vector<int> starts, goals;
  unpackInstance(inst, g, starts, goals);
  std::memset(res, 0, sizeof *res);
  Caps k = toCaps(caps);
  k.start();
  std::vector<Plan> sol;
  int st = cbsSearch(g, starts, goals, k, sol, res);
  res->runtime_s = k.elapsed();
  res->status = st;
  if (st == ORC_SOLVED) {
    finishResult(sol, res);
    exportPaths(g, sol, path_off, path_xyg, path_cap);
  }
  return st;
}

int orc_ecbs(const orc_instance* inst, float w, const orc_caps* caps,
             orc_result* res, int32_t* path_off, int32_t* path_xyg,
             int path_cap) {
  Grid g;
  std::vector<int> starts, goals;
  unpackInstance(inst, g, starts, goals);
  std::memset(res, 0, sizeof *res);
  Caps k = toCaps(caps);
  k.start();
  std::vector<Plan> sol;
  int st = ecbsSearch(g, starts, goals, w, k, sol, res);
  res->runtime_s = k.elapsed();
  res->status = st;
  if (st == ORC_SOLVED) {
    finishResult(sol, res);
    exportPaths(g, sol, path_off, path_xyg, path_cap);
  }
  return st;
}

int orc_cbs_ta(const orc_instance* inst, int64_t max_task_assignments,
               const orc_caps* caps, orc_result* res, int32_t* path_off,
               int32_t* path_xyg, int path_cap) {
  Grid g;
  std::vector<int> starts, goals;
  unpackInstance(inst, g, starts, goals);
  std::vector<std::vector<int>> pg(inst->n_agents);
  for (int i = 0; i < inst->n_agents; ++i) {
    // goals[i] is an unordered_set<Location> in the reference
    // (example/cbs_ta.cpp:549,565-567): duplicates collapse
    std::set<int> seen;
    for (int k = inst->pg_off[i]; k < inst->pg_off[i + 1]; ++k) {
      int cell = inst->pg_xy[2 * k] + g.dimx * inst->pg_xy[2 * k + 1];
      if (seen.insert(cell).second) pg[i].push_back(cell);
    }
  }
  std::memset(res, 0, sizeof *res);
  Caps k = toCaps(caps);
  std::vector<Plan> sol;
  int st = cbsTaSearch(g, starts, pg, max_task_assignments, k, sol, res);
  res->status = st;
  if (st == ORC_SOLVED) {
    finishResult(sol, res);
    exportPaths(g, sol, path_off, path_xyg, path_cap);
  }
  return st;
}

int orc_ecbs_ta(const orc_instance* inst, float w, int64_t max_task_assignments,
                const orc_caps* caps, orc_result* res, int32_t* path_off, int32_t* path_xyg,
                int path_cap) {
  Grid g;
  std::vector<int> starts, goals;
  unpackInstance(inst, g, starts, goals);
  std::vector<std::vector<int>> pg(inst->n_agents);
  for (int i = 0; i < inst->n_agents; ++i) {
    std::set<int> seen;
    for (int k = inst->pg_off[i]; k < inst->pg_off[i + 1]; ++k) {
      int cell = inst->pg_xy[2 * k] + g.dimx * inst->pg_xy[2 * k + 1];
      if (seen.insert(cell).second) pg[i].push_back(cell);
    }
  }
  std::memset(res, 0, sizeof *res);
  Caps k = toCaps(caps);
  std::vector<Plan> sol;
  int st = ecbsTaSearch(g, starts, pg, w, max_task_assignments, k, sol, res);
  res->status = st;
  if (st == ORC_SOLVED) {
    finishResult(sol, res);
    exportPaths(g, sol, path_off, path_xyg, path_cap);
  }
  return st;
}

int64_t orc_assignment(const int64_t* edges, int n_edges, int n_agents,
                       int n_tasks, int32_t* sol_task) {
  (void)n_tasks;
  Assignment a;
  for (int k = 0; k < n_edges; ++k)
    a.setCost((int)edges[3 * k], (int)edges[3 * k + 1], (long)edges[3 * k + 2]);
  std::map<int, int> sol;
  long c = a.solve(sol);
  for (int i = 0; i < n_agents; ++i) sol_task[i] = -1;
  for (const auto& e : sol) sol_task[e.first] = e.second;
  return c;
}

int orc_next_best_assignments(const int64_t* edges, int n_edges, int n_agents,
                              int n_tasks, int max_solutions, int64_t* costs,
                              int32_t* sol_task) {
  (void)n_tasks;
  NextBestAssignment a;
  for (int k = 0; k < n_edges; ++k)
    a.setCost((int)edges[3 * k], (int)edges[3 * k + 1], (long)edges[3 * k + 2]);
  a.solve();
  int n = 0;
  std::map<int, int> sol;
  while (n < max_solutions) {
    long c = a.nextSolution(sol);
    if (sol.empty()) break;
    costs[n] = c;
    for (int i = 0; i < n_agents; ++i) sol_task[(size_t)n * n_agents + i] = -1;
    for (const auto& e : sol) sol_task[(size_t)n * n_agents + e.first] = e.second;
    ++n;
  }
  return n;
}

}  // extern "C"
