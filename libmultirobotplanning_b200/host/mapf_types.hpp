// mapf_types.hpp — the value types that cross the boundary between the
// high-level searches and the Environment, with the names and members the
// reference's examples use (example/cbs.cpp:16-232 for State / Action /
// Conflict / VertexConstraint / EdgeConstraint / Constraints / Location,
// include/libMultiRobotPlanning/neighbor.hpp:14-25 for Neighbor and
// planresult.hpp:18-28 for PlanResult), so that code written against the
// reference's Environment concept reads the same against this one.
#pragma once

#include <cstddef>
#include <functional>
#include <ostream>
#include <tuple>
#include <unordered_set>
#include <utility>
#include <vector>

namespace mrp_host {

inline void hashCombine(std::size_t& seed, std::size_t v) {
  seed ^= v + 0x9e3779b97f4a7c15ull + (seed << 6) + (seed >> 2);
}

struct State {
  State(int time, int x, int y) : time(time), x(x), y(y) {}
  bool operator==(const State& s) const { return time == s.time && x == s.x && y == s.y; }
  bool equalExceptTime(const State& s) const { return x == s.x && y == s.y; }
  friend std::ostream& operator<<(std::ostream& os, const State& s) {
    return os << s.time << ": (" << s.x << "," << s.y << ")";
  }
  int time, x, y;
};

enum class Action { Up, Down, Left, Right, Wait };
inline std::ostream& operator<<(std::ostream& os, const Action& a) {
  static const char* names[] = {"Up", "Down", "Left", "Right", "Wait"};
  return os << names[static_cast<int>(a)];
}

struct Conflict {
  enum Type { Vertex, Edge };
  int time;
  std::size_t agent1, agent2;
  Type type;
  int x1, y1, x2, y2;
  friend std::ostream& operator<<(std::ostream& os, const Conflict& c) {
    if (c.type == Vertex) return os << c.time << ": Vertex(" << c.x1 << "," << c.y1 << ")";
    return os << c.time << ": Edge(" << c.x1 << "," << c.y1 << "," << c.x2 << "," << c.y2 << ")";
  }
};

struct VertexConstraint {
  VertexConstraint(int time, int x, int y) : time(time), x(x), y(y) {}
  int time, x, y;
  bool operator<(const VertexConstraint& o) const {
    return std::tie(time, x, y) < std::tie(o.time, o.x, o.y);
  }
  bool operator==(const VertexConstraint& o) const {
    return std::tie(time, x, y) == std::tie(o.time, o.x, o.y);
  }
  friend std::ostream& operator<<(std::ostream& os, const VertexConstraint& c) {
    return os << "VC(" << c.time << "," << c.x << "," << c.y << ")";
  }
};

struct EdgeConstraint {
  EdgeConstraint(int time, int x1, int y1, int x2, int y2)
      : time(time), x1(x1), y1(y1), x2(x2), y2(y2) {}
  int time, x1, y1, x2, y2;
  bool operator<(const EdgeConstraint& o) const {
    return std::tie(time, x1, y1, x2, y2) < std::tie(o.time, o.x1, o.y1, o.x2, o.y2);
  }
  bool operator==(const EdgeConstraint& o) const {
    return std::tie(time, x1, y1, x2, y2) == std::tie(o.time, o.x1, o.y1, o.x2, o.y2);
  }
  friend std::ostream& operator<<(std::ostream& os, const EdgeConstraint& c) {
    return os << "EC(" << c.time << "," << c.x1 << "," << c.y1 << "," << c.x2 << "," << c.y2
              << ")";
  }
};

struct Location {
  Location() = default;
  Location(int x, int y) : x(x), y(y) {}
  int x = 0, y = 0;
  bool operator<(const Location& o) const { return std::tie(x, y) < std::tie(o.x, o.y); }
  bool operator==(const Location& o) const { return x == o.x && y == o.y; }
  friend std::ostream& operator<<(std::ostream& os, const Location& c) {
    return os << "(" << c.x << "," << c.y << ")";
  }
};

template <typename StateT, typename ActionT, typename Cost>
struct Neighbor {
  Neighbor(const StateT& state, const ActionT& action, Cost cost)
      : state(state), action(action), cost(cost) {}
  StateT state;
  ActionT action;
  Cost cost;
};

template <typename StateT, typename ActionT, typename Cost>
struct PlanResult {
  std::vector<std::pair<StateT, Cost> > states;
  std::vector<std::pair<ActionT, Cost> > actions;
  Cost cost;
  Cost fmin;
};

}  // namespace mrp_host

namespace std {
template <>
struct hash<mrp_host::State> {
  size_t operator()(const mrp_host::State& s) const {
    size_t seed = 0;
    mrp_host::hashCombine(seed, s.time);
    mrp_host::hashCombine(seed, s.x);
    mrp_host::hashCombine(seed, s.y);
    return seed;
  }
};
template <>
struct hash<mrp_host::VertexConstraint> {
  size_t operator()(const mrp_host::VertexConstraint& s) const {
    size_t seed = 0;
    mrp_host::hashCombine(seed, s.time);
    mrp_host::hashCombine(seed, s.x);
    mrp_host::hashCombine(seed, s.y);
    return seed;
  }
};
template <>
struct hash<mrp_host::EdgeConstraint> {
  size_t operator()(const mrp_host::EdgeConstraint& s) const {
    size_t seed = 0;
    mrp_host::hashCombine(seed, s.time);
    mrp_host::hashCombine(seed, s.x1);
    mrp_host::hashCombine(seed, s.y1);
    mrp_host::hashCombine(seed, s.x2);
    mrp_host::hashCombine(seed, s.y2);
    return seed;
  }
};
template <>
struct hash<mrp_host::Location> {
  size_t operator()(const mrp_host::Location& s) const {
    size_t seed = 0;
    mrp_host::hashCombine(seed, s.x);
    mrp_host::hashCombine(seed, s.y);
    return seed;
  }
};
}  // namespace std

namespace mrp_host {

struct Constraints {
  std::unordered_set<VertexConstraint> vertexConstraints;
  std::unordered_set<EdgeConstraint> edgeConstraints;
  void add(const Constraints& other) {
    vertexConstraints.insert(other.vertexConstraints.begin(), other.vertexConstraints.end());
    edgeConstraints.insert(other.edgeConstraints.begin(), other.edgeConstraints.end());
  }
  // true if the two sets share a constraint (what the reference's assert means
  // to check, cbs.hpp:149; its own implementation intersects unordered ranges)
  bool overlap(const Constraints& other) const {
    for (const auto& v : other.vertexConstraints)
      if (vertexConstraints.count(v)) return true;
    for (const auto& e : other.edgeConstraints)
      if (edgeConstraints.count(e)) return true;
    return false;
  }
  friend std::ostream& operator<<(std::ostream& os, const Constraints& c) {
    for (const auto& vc : c.vertexConstraints) os << vc << std::endl;
    for (const auto& ec : c.edgeConstraints) os << ec << std::endl;
    return os;
  }
};

}  // namespace mrp_host
