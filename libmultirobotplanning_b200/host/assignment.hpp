// assignment.hpp — host-side task assignment for the cbs_ta path.
//
// Re-creates the interface of include/libMultiRobotPlanning/assignment.hpp:34-118
// (Assignment: setCost / solve / clear) and next_best_assignment.hpp:37-201
// (NextBestAssignment: setCost / solve / nextSolution) without Boost.Graph:
// the min-cost maximum matching is found by successive shortest augmenting
// paths (Dijkstra with potentials).  The optimal cost is unique; which optimum
// is returned among ties is not pinned by the reference (Boost internals).
// Sequential and tiny (N <= 200): stays on the host.
#pragma once

#include <climits>
#include <map>
#include <queue>
#include <set>
#include <utility>
#include <vector>

namespace mrp_host {

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
    const int V = 2 + A + T;  // 0 source, 1 sink
    struct Arc {
      int to;
      long cap, cost;
      int rev;
    };
    std::vector<std::vector<Arc> > adj(V);
    auto addArc = [&adj](int u, int v, long cost) {
      adj[u].push_back({v, 1, cost, (int)adj[v].size()});
      adj[v].push_back({u, 0, -cost, (int)adj[u].size() - 1});
    };
    for (int a = 0; a < A; ++a) addArc(0, 2 + a, 0);
    for (int t = 0; t < T; ++t) addArc(2 + A + t, 1, 0);
    for (const auto& e : m_edges) addArc(2 + e.first.first, 2 + A + e.first.second, e.second);
    const long INF = LONG_MAX / 4;
    std::vector<long> pot(V, 0), dist(V);
    std::vector<int> pv(V), pe(V);
    for (;;) {
      std::fill(dist.begin(), dist.end(), INF);
      dist[0] = 0;
      typedef std::pair<long, int> QE;
      std::priority_queue<QE, std::vector<QE>, std::greater<QE> > pq;
      pq.push(QE(0, 0));
      while (!pq.empty()) {
        QE top = pq.top();
        pq.pop();
        const int u = top.second;
        if (top.first > dist[u]) continue;
        for (int k = 0; k < (int)adj[u].size(); ++k) {
          const Arc& e = adj[u][k];
          if (e.cap <= 0) continue;
          const long nd = top.first + e.cost + pot[u] - pot[e.to];
          if (nd < dist[e.to]) {
            dist[e.to] = nd;
            pv[e.to] = u;
            pe[e.to] = k;
            pq.push(QE(nd, e.to));
          }
        }
      }
      if (dist[1] >= INF) break;
      for (int v = 0; v < V; ++v)
        if (dist[v] < INF) pot[v] += dist[v];
      for (int v = 1; v != 0; v = pv[v]) {
        Arc& e = adj[pv[v]][pe[v]];
        e.cap -= 1;
        adj[v][e.rev].cap += 1;
      }
    }
    long cost = 0;
    for (int a = 0; a < A; ++a)
      for (const Arc& e : adj[2 + a])
        if (e.to >= 2 + A && e.cap == 0) {
          solution[m_agents[a]] = m_tasks[e.to - 2 - A];
          cost += e.cost;
          break;
        }
    return cost;
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
    if (!m_agentsSet.count(agent)) {
      m_agentsSet.insert(agent);
      m_agentsVec.push_back(agent);
    }
  }

  void solve() {
    Node n;
    n.cost = constrainedMatching(n.I, n.O, n.Iagents, n.Oagents, n.solution);
    m_open.push(n);
    m_numMatching = n.solution.size();
  }

  // next best solution; `solution` stays empty when the enumeration is over
  long nextSolution(std::map<Agent, Task>& solution) {
    solution.clear();
    if (m_open.empty()) return LONG_MAX;
    const Node next = m_open.top();
    m_open.pop();
    solution = next.solution;
    const long result = next.cost;
    std::set<Agent> fixedAgents;
    for (const auto& c : next.I) fixedAgents.insert(c.first);
    for (size_t i = 0; i < m_agentsVec.size(); ++i) {
      if (fixedAgents.count(m_agentsVec[i])) continue;
      Node n;
      n.I = next.I;
      n.O = next.O;
      n.Iagents = next.Iagents;
      n.Oagents = next.Oagents;
      // agents before i keep their assignment (or their lack of one) ...
      for (size_t j = 0; j < i; ++j) {
        const Agent& agent = m_agentsVec[j];
        auto it = solution.find(agent);
        if (it != solution.end())
          n.I.insert(std::make_pair(agent, it->second));
        else
          n.Oagents.insert(agent);
      }
      // ... agent i must change
      auto it = solution.find(m_agentsVec[i]);
      if (it != solution.end())
        n.O.insert(std::make_pair(m_agentsVec[i], it->second));
      else
        n.Iagents.insert(m_agentsVec[i]);
      n.cost = constrainedMatching(n.I, n.O, n.Iagents, n.Oagents, n.solution);
      if (!n.solution.empty()) m_open.push(n);
    }
    return result;
  }

 private:
  typedef std::set<std::pair<Agent, Task> > PairSet;
  long constrainedMatching(const PairSet& I, const PairSet& O, const std::set<Agent>& Iagents,
                           const std::set<Agent>& Oagents, std::map<Agent, Task>& solution) {
    m_assignment.clear();
    for (const auto& c : I)
      if (!Oagents.count(c.first)) m_assignment.setCost(c.first, c.second, 0);
    for (const auto& c : m_cost) {
      if (O.count(c.first) || I.count(c.first) || Oagents.count(c.first.first)) continue;
      // the offset makes every free agent cheaper to leave unassigned than an
      // enforced one (next_best_assignment.hpp:148)
      const long offset = Iagents.count(c.first.first) ? 0 : 1000000000L;
      m_assignment.setCost(c.first.first, c.first.second, c.second + offset);
    }
    m_assignment.solve(solution);
    bool valid = solution.size() >= m_numMatching;
    for (const auto& agent : Iagents)
      if (!solution.count(agent)) valid = false;
    for (const auto& c : I) {
      auto it = solution.find(c.first);
      if (it == solution.end() || !(it->second == c.second)) valid = false;
    }
    if (!valid) {
      solution.clear();
      return LONG_MAX;
    }
    long result = 0;
    for (const auto& e : solution) result += m_cost.at(e);
    return result;
  }

  struct Node {
    PairSet I, O;
    std::set<Agent> Iagents, Oagents;
    std::map<Agent, Task> solution;
    long cost = 0;
    bool operator<(const Node& n) const { return cost > n.cost; }
  };

  Assignment<Agent, Task> m_assignment;
  std::map<std::pair<Agent, Task>, long> m_cost;
  std::vector<Agent> m_agentsVec;
  std::set<Agent> m_agentsSet;
  std::priority_queue<Node> m_open;
  size_t m_numMatching = 0;
};

}  // namespace mrp_host
