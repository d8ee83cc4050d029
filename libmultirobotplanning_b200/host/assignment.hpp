// assignment.hpp — host-side task assignment for the cbs_ta / ecbs_ta path.
//
// Re-creates the interface of include/libMultiRobotPlanning/assignment.hpp:34-118
// (Assignment: setCost / solve / clear) and next_best_assignment.hpp:37-201
// (NextBestAssignment: setCost / solve / nextSolution) without Boost.Graph.
//
// The reference solves every (sub)problem as a min-cost max-flow on a BGL
// adjacency_list (successive_shortest_path_nonnegative_weights,
// assignment.hpp:84-89) and enumerates the next-best assignments with Murty's
// partitioning, one full matching per agent per enumerated solution
// (next_best_assignment.hpp:59-122).  At the size of config C4 (100 agents x 100
// potential goals, one enumeration step per high-level expansion because the
// children of a root inherit isRoot, cbs_ta.hpp:180) this was 59 of 60 seconds of
// a cbs_ta run.  Same semantics here — maximum cardinality first, then minimum
// cost; the 1e9 offset of next_best_assignment.hpp:148; identical validity rules —
// on a dense cost matrix: successive shortest augmenting paths from ALL unmatched
// agents at once (dense Dijkstra over the tasks with potentials, O(A*T) per
// augmentation), and the sub-problems of one enumeration step solved in parallel
// (OpenMP) and queued in agent order.  The optimal costs and their order are
// unique; which optimum is returned among ties is not pinned by the reference
// (Boost internals).  Sequential by nature across steps and tiny (N <= 200): stays
// on the host.
#pragma once

#include <algorithm>
#include <climits>
#include <map>
#include <queue>
#include <set>
#include <utility>
#include <vector>

namespace mrp_host {

namespace detail {

constexpr long kAbsent = LONG_MAX;  // no edge between this agent and this task

// Minimum-cost matching among the matchings of maximum cardinality.
// cost: A x T row-major (kAbsent = no edge, all others >= 0).  taskOf[a] = task
// of agent a or -1.  Equivalent to the min-cost max-flow the reference builds
// (unit capacities, source -> agents -> tasks -> sink): every iteration augments
// along a globally shortest residual path, started from every unmatched agent.
inline void minCostMaxMatching(int A, int T, const std::vector<long>& cost, std::vector<int>& taskOf) {
  const long INF = LONG_MAX / 4;
  taskOf.assign(A, -1);
  std::vector<int> agentOf(T, -1), prevAgent(T);
  std::vector<long> potA(A, 0), potT(T, 0), dist(T);
  std::vector<char> done(T);
  for (;;) {
    // distances of the tasks from the set of unmatched agents
    for (int t = 0; t < T; ++t) {
      dist[t] = INF;
      done[t] = 0;
      prevAgent[t] = -1;
    }
    bool anySource = false;
    for (int a = 0; a < A; ++a) {
      if (taskOf[a] >= 0) continue;
      const long* row = &cost[(size_t)a * T];
      for (int t = 0; t < T; ++t) {
        if (row[t] == kAbsent) continue;
        anySource = true;
        const long nd = row[t] + potA[a] - potT[t];
        if (nd < dist[t]) {
          dist[t] = nd;
          prevAgent[t] = a;
        }
      }
    }
    if (!anySource) break;
    int target = -1;
    for (;;) {
      int best = -1;
      for (int t = 0; t < T; ++t)
        if (!done[t] && dist[t] < INF && (best < 0 || dist[t] < dist[best])) best = t;
      if (best < 0) break;
      done[best] = 1;
      if (agentOf[best] < 0) {
        target = best;
        break;
      }
      // through the matched edge (reduced cost 0) to its agent, on to the other tasks
      const int a = agentOf[best];
      const long* row = &cost[(size_t)a * T];
      const long base = dist[best] + potA[a];
      for (int t = 0; t < T; ++t) {
        if (done[t] || row[t] == kAbsent) continue;
        const long nd = base + row[t] - potT[t];
        if (nd < dist[t]) {
          dist[t] = nd;
          prevAgent[t] = a;
        }
      }
    }
    if (target < 0) break;  // no augmenting path: the matching has maximum cardinality
    // potentials += min(distance, D): reduced costs stay non-negative, the
    // edges of the shortest path become tight (an unmatched agent is at distance
    // 0, a matched one at the distance of its task)
    const long D = dist[target];
    for (int a = 0; a < A; ++a)
      if (taskOf[a] >= 0) potA[a] += std::min(dist[taskOf[a]], D);
    for (int t = 0; t < T; ++t) potT[t] += std::min(dist[t], D);
    for (int t = target; t >= 0;) {
      const int a = prevAgent[t];
      const int next = taskOf[a];
      taskOf[a] = t;
      agentOf[t] = a;
      t = next;
    }
  }
}

}  // namespace detail

template <typename Agent, typename Task>
class Assignment {
 public:
  void clear() { m_edges.clear(); }
  void setCost(const Agent& agent, const Task& task, long cost) {
    auto ai = m_agentIdx.find(agent);
    if (ai == m_agentIdx.end()) {
      ai = m_agentIdx.emplace(agent, (int)m_agents.size()).first;
      m_agents.push_back(agent);
    }
    auto ti = m_taskIdx.find(task);
    if (ti == m_taskIdx.end()) {
      ti = m_taskIdx.emplace(task, (int)m_tasks.size()).first;
      m_tasks.push_back(task);
    }
    m_edges[std::make_pair(ai->second, ti->second)] = cost;
  }

  long solve(std::map<Agent, Task>& solution) {
    solution.clear();
    const int A = (int)m_agents.size(), T = (int)m_tasks.size();
    std::vector<long> cost((size_t)A * T, detail::kAbsent);
    for (const auto& e : m_edges) cost[(size_t)e.first.first * T + e.first.second] = e.second;
    std::vector<int> taskOf;
    detail::minCostMaxMatching(A, T, cost, taskOf);
    long total = 0;
    for (int a = 0; a < A; ++a)
      if (taskOf[a] >= 0) {
        solution[m_agents[a]] = m_tasks[taskOf[a]];
        total += cost[(size_t)a * T + taskOf[a]];
      }
    return total;
  }

 private:
  std::vector<Agent> m_agents;
  std::vector<Task> m_tasks;
  std::map<Agent, int> m_agentIdx;
  std::map<Task, int> m_taskIdx;
  std::map<std::pair<int, int>, long> m_edges;
};

template <typename Agent, typename Task>
class NextBestAssignment {
 public:
  void setCost(const Agent& agent, const Task& task, long cost) {
    m_cost[std::make_pair(agent, task)] = cost;
    if (!m_agentIdx.count(agent)) {
      m_agentIdx.emplace(agent, (int)m_agentsVec.size());
      m_agentsVec.push_back(agent);
    }
    if (!m_taskIdx.count(task)) {
      m_taskIdx.emplace(task, (int)m_tasksVec.size());
      m_tasksVec.push_back(task);
    }
  }

  void solve() {
    // dense base matrix (agents in the order they were first seen, as the
    // reference's m_agentsVec, next_best_assignment.hpp:41-46)
    m_A = (int)m_agentsVec.size();
    m_T = (int)m_tasksVec.size();
    m_base.assign((size_t)m_A * m_T, detail::kAbsent);
    for (const auto& c : m_cost)
      m_base[(size_t)m_agentIdx[c.first.first] * m_T + m_taskIdx[c.first.second]] = c.second;
    Node n;
    n.fixed.assign(m_A, -1);
    n.mustHave.assign(m_A, 0);
    n.mustNot.assign(m_A, 0);
    n.cost = constrainedMatching(n);
    m_numMatching = 0;
    for (int t : n.solution) m_numMatching += t >= 0;
    m_open.push(n);
  }

  // next best solution; `solution` stays empty when the enumeration is over
  long nextSolution(std::map<Agent, Task>& solution) {
    solution.clear();
    if (m_open.empty()) return LONG_MAX;
    const Node next = m_open.top();
    m_open.pop();
    for (int a = 0; a < m_A; ++a)
      if (next.solution[a] >= 0) solution[m_agentsVec[a]] = m_tasksVec[next.solution[a]];
    const long result = next.cost;
    // Murty's partition (next_best_assignment.hpp:77-119): for every agent i
    // without a fixed pair, agents before i keep their assignment (or their lack
    // of one) and agent i must change.  The sub-problems are independent.
    std::vector<int> free;
    for (int i = 0; i < m_A; ++i)
      if (next.fixed[i] < 0) free.push_back(i);
    std::vector<Node> made(free.size());
#pragma omp parallel for schedule(dynamic, 1) if (free.size() >= 8 && m_A >= 24)
    for (long q = 0; q < (long)free.size(); ++q) {
      const int i = free[q];
      Node n;
      n.fixed = next.fixed;
      n.forbidden = next.forbidden;
      n.mustHave = next.mustHave;
      n.mustNot = next.mustNot;
      for (int j = 0; j < i; ++j) {
        if (next.solution[j] >= 0)
          n.fixed[j] = next.solution[j];
        else
          n.mustNot[j] = 1;
      }
      if (next.solution[i] >= 0)
        n.forbidden.push_back(std::make_pair(i, next.solution[i]));
      else
        n.mustHave[i] = 1;
      n.cost = constrainedMatching(n);
      made[q] = std::move(n);
    }
    for (Node& n : made)
      if (n.valid) m_open.push(std::move(n));
    return result;
  }

 private:
  // I, O, Iagents, Oagents of the reference's Node (next_best_assignment.hpp:204-231)
  struct Node {
    std::vector<int> fixed;                      // I: task an agent must keep, or -1
    std::vector<std::pair<int, int> > forbidden;  // O: pairs that must not be used
    std::vector<char> mustHave;                  // Iagents: must get a task
    std::vector<char> mustNot;                   // Oagents: must stay without a task
    std::vector<int> solution;                   // task per agent or -1
    long cost = 0;
    bool valid = false;
    bool operator<(const Node& n) const { return cost > n.cost; }
  };

  // next_best_assignment.hpp:129-189
  long constrainedMatching(Node& n) const {
    const int A = m_A, T = m_T;
    std::vector<long> cost(m_base);
    for (int a = 0; a < A; ++a) {
      long* row = &cost[(size_t)a * T];
      if (n.mustNot[a]) {
        for (int t = 0; t < T; ++t) row[t] = detail::kAbsent;
        continue;
      }
      // the offset makes every free agent cheaper to leave unassigned than an
      // enforced one (next_best_assignment.hpp:148)
      const long offset = n.mustHave[a] ? 0 : 1000000000L;
      for (int t = 0; t < T; ++t)
        if (row[t] != detail::kAbsent) row[t] += offset;
    }
    for (const auto& o : n.forbidden) cost[(size_t)o.first * T + o.second] = detail::kAbsent;
    for (int a = 0; a < A; ++a)
      if (n.fixed[a] >= 0 && !n.mustNot[a]) cost[(size_t)a * T + n.fixed[a]] = 0;
    detail::minCostMaxMatching(A, T, cost, n.solution);
    size_t matched = 0;
    for (int t : n.solution) matched += t >= 0;
    bool valid = matched >= m_numMatching;
    for (int a = 0; a < A; ++a) {
      if (n.mustHave[a] && n.solution[a] < 0) valid = false;
      if (n.fixed[a] >= 0 && n.solution[a] != n.fixed[a]) valid = false;
    }
    n.valid = valid && matched > 0;
    if (!valid) {
      n.solution.assign(A, -1);
      return LONG_MAX;
    }
    long result = 0;
    for (int a = 0; a < A; ++a)
      if (n.solution[a] >= 0) result += m_base[(size_t)a * T + n.solution[a]];
    return result;
  }

  std::map<std::pair<Agent, Task>, long> m_cost;
  std::vector<Agent> m_agentsVec;
  std::vector<Task> m_tasksVec;
  std::map<Agent, int> m_agentIdx;
  std::map<Task, int> m_taskIdx;
  std::vector<long> m_base;
  int m_A = 0, m_T = 0;
  std::priority_queue<Node> m_open;
  size_t m_numMatching = 0;
};

}  // namespace mrp_host
