// gpu_environment.hpp — the reference's duck-typed `Environment` concept
// (the callback surface CBS / ECBS / CBSTA and AStar / AStarEpsilon consume,
// cbs.hpp:209-244, ecbs.hpp:365-416, cbs_ta.hpp:262-297) backed by the CUDA
// library.  Same method names, argument meaning and return conventions as the
// Environment classes of example/cbs.cpp:247-569, example/ecbs.cpp:247-522 and
// example/cbs_ta.cpp:252-514, so the reference's unmodified templates can be
// instantiated with these classes:
//
// (the plan / neighbour containers are template parameters of the methods, so
// both libMultiRobotPlanning::PlanResult / Neighbor and the look-alikes of
// mapf_types.hpp work; the parity tests instantiate the reference's own
// cbs.hpp / ecbs.hpp / cbs_ta.hpp with these classes):
//
//   mrp_host::Environment mapf(dimx, dimy, obstacles, goals);
//   libMultiRobotPlanning::CBS<State, Action, int, Conflict, Constraints,
//                              mrp_host::Environment> cbs(mapf);
//
// What runs where: getFirstConflict and focalHeuristic pack the solution into
// a path table and run on the GPU (mrp_first_conflict / mrp_count_conflicts);
// admissibleHeuristic reads the BFS distance field computed on the GPU at
// construction (mrp_bfs_fields) — the job of ShortestPathHeuristic in the
// reference; the per-state callbacks (getNeighbors, isSolution, the two focal
// per-node counters) are O(1)/O(N) host code, exactly as in the reference:
// called one state at a time there is nothing to batch.  The batched path that
// bypasses this non-reentrant per-call context is hl_search.hpp.
#pragma once

#include <climits>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/mrp_b200.h"
#include "assignment.hpp"
#include "mapf_types.hpp"

namespace mrp_host {

namespace detail {
inline void check(int rc) {
  if (rc < 0) throw std::runtime_error(std::string("mrp_b200: ") + mrp_last_error());
}

template <typename Plan>
void packSolution(const std::vector<Plan>& solution, int dimx, std::vector<int32_t>& cell,
                  std::vector<int32_t>& len, int& Tpad) {
  Tpad = 1;
  for (const auto& s : solution) Tpad = std::max<int>(Tpad, (int)s.states.size());
  cell.assign(solution.size() * (size_t)Tpad, 0);
  len.assign(solution.size(), 0);
  for (size_t i = 0; i < solution.size(); ++i) {
    len[i] = (int)solution[i].states.size();
    for (size_t t = 0; t < solution[i].states.size(); ++t)
      cell[i * (size_t)Tpad + t] =
          solution[i].states[t].first.x + dimx * solution[i].states[t].first.y;
  }
}

inline std::vector<int32_t> flatten(const std::unordered_set<Location>& obstacles) {
  std::vector<int32_t> xy;
  for (const auto& o : obstacles) {
    xy.push_back(o.x);
    xy.push_back(o.y);
  }
  return xy;
}
}  // namespace detail

// ---------------------------------------------------------------------------
// cbs / ecbs
// ---------------------------------------------------------------------------
class Environment {
 public:
  typedef PlanResult<State, Action, int> Plan;

  Environment(size_t dimx, size_t dimy, std::unordered_set<Location> obstacles,
              std::vector<Location> goals)
      : m_dimx((int)dimx),
        m_dimy((int)dimy),
        m_obstacles(std::move(obstacles)),
        m_goals(std::move(goals)),
        m_agentIdx(0),
        m_constraints(nullptr),
        m_lastGoalConstraint(-1),
        m_highLevelExpanded(0),
        m_lowLevelExpanded(0) {
    // one BFS distance field per goal, on the GPU (the reference's disabled
    // computeHeuristic, example/cbs.cpp:445-557)
    std::vector<int32_t> obst = detail::flatten(m_obstacles), goalXY;
    for (const auto& g : m_goals) {
      goalXY.push_back(g.x);
      goalXY.push_back(g.y);
    }
    m_heuristic.resize(m_goals.size() * (size_t)m_dimx * m_dimy);
    detail::check(mrp_bfs_fields(m_dimx, m_dimy, obst.data(), (int)obst.size() / 2,
                                 goalXY.data(), (int)m_goals.size(), m_heuristic.data()));
  }
  Environment(const Environment&) = delete;
  Environment& operator=(const Environment&) = delete;

  void setLowLevelContext(size_t agentIdx, const Constraints* constraints) {
    m_agentIdx = agentIdx;
    m_constraints = constraints;
    m_lastGoalConstraint = -1;
    for (const auto& vc : constraints->vertexConstraints)
      if (vc.x == m_goals[m_agentIdx].x && vc.y == m_goals[m_agentIdx].y)
        m_lastGoalConstraint = std::max(m_lastGoalConstraint, vc.time);
  }

  int admissibleHeuristic(const State& s) {
    return m_heuristic[m_agentIdx * (size_t)m_dimx * m_dimy + s.x + m_dimx * s.y];
  }

  template <class PlanT>
  int focalStateHeuristic(const State& s, int /*gScore*/, const std::vector<PlanT>& solution) {
    int n = 0;
    for (size_t i = 0; i < solution.size(); ++i)
      if (i != m_agentIdx && !solution[i].states.empty() &&
          s.equalExceptTime(getState(i, solution, s.time)))
        ++n;
    return n;
  }

  template <class PlanT>
  int focalTransitionHeuristic(const State& s1a, const State& s1b, int /*g1a*/, int /*g1b*/,
                               const std::vector<PlanT>& solution) {
    int n = 0;
    for (size_t i = 0; i < solution.size(); ++i)
      if (i != m_agentIdx && !solution[i].states.empty()) {
        State s2a = getState(i, solution, s1a.time), s2b = getState(i, solution, s1b.time);
        if (s1a.equalExceptTime(s2b) && s1b.equalExceptTime(s2a)) ++n;
      }
    return n;
  }

  // number of vertex + edge conflicts of a joint plan — on the GPU
  template <class PlanT>
  int focalHeuristic(const std::vector<PlanT>& solution) {
    std::vector<int32_t> cell, len;
    int Tpad;
    detail::packSolution(solution, m_dimx, cell, len, Tpad);
    int32_t count = 0;
    detail::check(mrp_count_conflicts(cell.data(), len.data(), (int)solution.size(), Tpad, 0,
                                      &count));
    return count;
  }

  bool isSolution(const State& s) {
    return s.x == m_goals[m_agentIdx].x && s.y == m_goals[m_agentIdx].y &&
           s.time > m_lastGoalConstraint;
  }

  template <class NeighborT>
  void getNeighbors(const State& s, std::vector<NeighborT>& neighbors) {
    neighbors.clear();
    static const int dx[5] = {0, -1, 1, 0, 0}, dy[5] = {0, 0, 0, 1, -1};
    static const Action act[5] = {Action::Wait, Action::Left, Action::Right, Action::Up,
                                  Action::Down};
    for (int k = 0; k < 5; ++k) {
      State n(s.time + 1, s.x + dx[k], s.y + dy[k]);
      if (stateValid(n) && transitionValid(s, n))
        neighbors.emplace_back(NeighborT(n, act[k], 1));
    }
  }

  // first conflict in the order (time, Vertex < Edge, agent1, agent2) — on the GPU
  template <class PlanT>
  bool getFirstConflict(const std::vector<PlanT>& solution, Conflict& result) {
    std::vector<int32_t> cell, len;
    int Tpad;
    detail::packSolution(solution, m_dimx, cell, len, Tpad);
    mrp_conflict c;
    int rc = mrp_first_conflict(cell.data(), len.data(), (int)solution.size(), Tpad, m_dimx, 0,
                                &c);
    detail::check(rc);
    if (rc == 0) return false;
    result.time = c.time;
    result.agent1 = (size_t)c.agent1;
    result.agent2 = (size_t)c.agent2;
    result.type = c.type == 0 ? Conflict::Vertex : Conflict::Edge;
    result.x1 = c.x1;
    result.y1 = c.y1;
    if (c.type == 1) {
      result.x2 = c.x2;
      result.y2 = c.y2;
    }
    return true;
  }

  void createConstraintsFromConflict(const Conflict& conflict,
                                     std::map<size_t, Constraints>& constraints) {
    if (conflict.type == Conflict::Vertex) {
      Constraints c1;
      c1.vertexConstraints.emplace(VertexConstraint(conflict.time, conflict.x1, conflict.y1));
      constraints[conflict.agent1] = c1;
      constraints[conflict.agent2] = c1;
    } else {
      Constraints c1, c2;
      c1.edgeConstraints.emplace(
          EdgeConstraint(conflict.time, conflict.x1, conflict.y1, conflict.x2, conflict.y2));
      c2.edgeConstraints.emplace(
          EdgeConstraint(conflict.time, conflict.x2, conflict.y2, conflict.x1, conflict.y1));
      constraints[conflict.agent1] = c1;
      constraints[conflict.agent2] = c2;
    }
  }

  void onExpandHighLevelNode(int /*cost*/) { m_highLevelExpanded++; }
  void onExpandLowLevelNode(const State& /*s*/, int /*fScore*/, int /*gScore*/) {
    m_lowLevelExpanded++;
  }
  int highLevelExpanded() { return m_highLevelExpanded; }
  int lowLevelExpanded() const { return m_lowLevelExpanded; }

 private:
  template <class PlanT>
  State getState(size_t agentIdx, const std::vector<PlanT>& solution, size_t t) {
    if (t < solution[agentIdx].states.size()) return solution[agentIdx].states[t].first;
    return solution[agentIdx].states.back().first;
  }
  bool stateValid(const State& s) {
    const auto& con = m_constraints->vertexConstraints;
    return s.x >= 0 && s.x < m_dimx && s.y >= 0 && s.y < m_dimy &&
           m_obstacles.find(Location(s.x, s.y)) == m_obstacles.end() &&
           con.find(VertexConstraint(s.time, s.x, s.y)) == con.end();
  }
  bool transitionValid(const State& s1, const State& s2) {
    const auto& con = m_constraints->edgeConstraints;
    return con.find(EdgeConstraint(s1.time, s1.x, s1.y, s2.x, s2.y)) == con.end();
  }

  int m_dimx, m_dimy;
  std::unordered_set<Location> m_obstacles;
  std::vector<Location> m_goals;
  std::vector<int32_t> m_heuristic;  // [goal][cell], from the GPU
  size_t m_agentIdx;
  const Constraints* m_constraints;
  int m_lastGoalConstraint;
  int m_highLevelExpanded, m_lowLevelExpanded;
};

// ---------------------------------------------------------------------------
// cbs_ta
// ---------------------------------------------------------------------------
class EnvironmentTA {
 public:
  typedef PlanResult<State, Action, int> Plan;

  EnvironmentTA(size_t dimx, size_t dimy, const std::unordered_set<Location>& obstacles,
                const std::vector<State>& startStates,
                const std::vector<std::unordered_set<Location> >& goals,
                size_t maxTaskAssignments)
      : m_dimx((int)dimx),
        m_dimy((int)dimy),
        m_obstacles(obstacles),
        m_agentIdx(0),
        m_goal(nullptr),
        m_constraints(nullptr),
        m_lastGoalConstraint(-1),
        m_maxTaskAssignments(maxTaskAssignments),
        m_numTaskAssignments(0),
        m_highLevelExpanded(0),
        m_lowLevelExpanded(0) {
    // distance fields of every potential goal, on the GPU — replaces the
    // Floyd–Warshall of ShortestPathHeuristic (shortest_path_heuristic.hpp:47-53)
    std::vector<int32_t> obst = detail::flatten(m_obstacles), goalXY;
    for (const auto& gs : goals)
      for (const auto& g : gs)
        if (!m_fieldOf.count(g)) {
          const size_t idx = m_fieldOf.size();
          m_fieldOf[g] = idx;
          goalXY.push_back(g.x);
          goalXY.push_back(g.y);
        }
    m_heuristic.resize(m_fieldOf.size() * (size_t)m_dimx * m_dimy);
    detail::check(mrp_bfs_fields(m_dimx, m_dimy, obst.data(), (int)obst.size() / 2,
                                 goalXY.data(), (int)m_fieldOf.size(), m_heuristic.data()));
    for (size_t i = 0; i < startStates.size(); ++i)
      for (const auto& goal : goals[i])
        m_assignment.setCost(i, goal,
                             getValue(Location(startStates[i].x, startStates[i].y), goal));
    m_assignment.solve();
  }

  void setLowLevelContext(size_t agentIdx, const Constraints* constraints,
                          const Location* task) {
    m_agentIdx = agentIdx;
    m_goal = task;
    m_constraints = constraints;
    m_lastGoalConstraint = -1;
    for (const auto& vc : constraints->vertexConstraints)
      if (m_goal == nullptr || (vc.x == m_goal->x && vc.y == m_goal->y))
        m_lastGoalConstraint = std::max(m_lastGoalConstraint, vc.time);
  }

  int admissibleHeuristic(const State& s) {
    return m_goal != nullptr ? getValue(Location(s.x, s.y), *m_goal) : 0;
  }

  bool isSolution(const State& s) {
    const bool atGoal = m_goal == nullptr || (s.x == m_goal->x && s.y == m_goal->y);
    return atGoal && s.time > m_lastGoalConstraint;
  }

  template <class NeighborT>
  void getNeighbors(const State& s, std::vector<NeighborT>& neighbors) {
    neighbors.clear();
    static const int dx[5] = {0, -1, 1, 0, 0}, dy[5] = {0, 0, 0, 1, -1};
    static const Action act[5] = {Action::Wait, Action::Left, Action::Right, Action::Up,
                                  Action::Down};
    for (int k = 0; k < 5; ++k) {
      State n(s.time + 1, s.x + dx[k], s.y + dy[k]);
      if (!(stateValid(n) && transitionValid(s, n))) continue;
      int cost = 1;
      if (k == 0) {  // waiting on the goal is free (example/cbs_ta.cpp:329-339)
        const bool atGoal = m_goal == nullptr || (s.x == m_goal->x && s.y == m_goal->y);
        cost = atGoal ? 0 : 1;
      }
      neighbors.emplace_back(NeighborT(n, act[k], cost));
    }
  }

  template <class PlanT>
  bool getFirstConflict(const std::vector<PlanT>& solution, Conflict& result) {
    std::vector<int32_t> cell, len;
    int Tpad;
    detail::packSolution(solution, m_dimx, cell, len, Tpad);
    mrp_conflict c;
    // mode 1: the cbs_ta loop bound max(states.size()) (example/cbs_ta.cpp:372-375)
    int rc = mrp_first_conflict(cell.data(), len.data(), (int)solution.size(), Tpad, m_dimx, 1,
                                &c);
    detail::check(rc);
    if (rc == 0) return false;
    result.time = c.time;
    result.agent1 = (size_t)c.agent1;
    result.agent2 = (size_t)c.agent2;
    result.type = c.type == 0 ? Conflict::Vertex : Conflict::Edge;
    result.x1 = c.x1;
    result.y1 = c.y1;
    if (c.type == 1) {
      result.x2 = c.x2;
      result.y2 = c.y2;
    }
    return true;
  }

  void createConstraintsFromConflict(const Conflict& conflict,
                                     std::map<size_t, Constraints>& constraints) {
    if (conflict.type == Conflict::Vertex) {
      Constraints c1;
      c1.vertexConstraints.emplace(VertexConstraint(conflict.time, conflict.x1, conflict.y1));
      constraints[conflict.agent1] = c1;
      constraints[conflict.agent2] = c1;
    } else {
      Constraints c1, c2;
      c1.edgeConstraints.emplace(
          EdgeConstraint(conflict.time, conflict.x1, conflict.y1, conflict.x2, conflict.y2));
      c2.edgeConstraints.emplace(
          EdgeConstraint(conflict.time, conflict.x2, conflict.y2, conflict.x1, conflict.y1));
      constraints[conflict.agent1] = c1;
      constraints[conflict.agent2] = c2;
    }
  }

  void nextTaskAssignment(std::map<size_t, Location>& tasks) {
    if (m_numTaskAssignments > m_maxTaskAssignments) return;
    m_assignment.nextSolution(tasks);
    if (!tasks.empty()) ++m_numTaskAssignments;
  }

  void onExpandHighLevelNode(int /*cost*/) { m_highLevelExpanded++; }
  void onExpandLowLevelNode(const State& /*s*/, int /*fScore*/, int /*gScore*/) {
    m_lowLevelExpanded++;
  }
  int highLevelExpanded() { return m_highLevelExpanded; }
  int lowLevelExpanded() const { return m_lowLevelExpanded; }
  size_t numTaskAssignments() const { return m_numTaskAssignments; }

  // ShortestPathHeuristic::getValue (shortest_path_heuristic.hpp:58-62);
  // b must be one of the potential goals
  int getValue(const Location& a, const Location& b) {
    return m_heuristic[m_fieldOf.at(b) * (size_t)m_dimx * m_dimy + a.x + m_dimx * a.y];
  }

 private:
  template <class PlanT>
  State getState(size_t agentIdx, const std::vector<PlanT>& solution, size_t t) {
    if (t < solution[agentIdx].states.size()) return solution[agentIdx].states[t].first;
    return solution[agentIdx].states.back().first;
  }
  bool stateValid(const State& s) {
    const auto& con = m_constraints->vertexConstraints;
    return s.x >= 0 && s.x < m_dimx && s.y >= 0 && s.y < m_dimy &&
           m_obstacles.find(Location(s.x, s.y)) == m_obstacles.end() &&
           con.find(VertexConstraint(s.time, s.x, s.y)) == con.end();
  }
  bool transitionValid(const State& s1, const State& s2) {
    const auto& con = m_constraints->edgeConstraints;
    return con.find(EdgeConstraint(s1.time, s1.x, s1.y, s2.x, s2.y)) == con.end();
  }

  int m_dimx, m_dimy;
  std::unordered_set<Location> m_obstacles;
  size_t m_agentIdx;
  const Location* m_goal;
  const Constraints* m_constraints;
  int m_lastGoalConstraint;
  NextBestAssignment<size_t, Location> m_assignment;
  size_t m_maxTaskAssignments, m_numTaskAssignments;
  int m_highLevelExpanded, m_lowLevelExpanded;
  std::map<Location, size_t> m_fieldOf;
  std::vector<int32_t> m_heuristic;
};

}  // namespace mrp_host
