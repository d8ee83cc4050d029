// astar_probe — the reference's OWN low-level search behind a pipe (TEST INFRASTRUCTURE): AStar::search
// (include/libMultiRobotPlanning/a_star.hpp:63-161) driven through the reference's Environment of
//   -DPROBE_CBS    example/cbs.cpp     (every move costs 1, Manhattan heuristic, goal test after the
//                                       last constraint on the goal: cbs.cpp:266-333)
//   -DPROBE_CBSTA  example/cbs_ta.cpp  (field heuristic, waiting on the goal is free, agents without a
//                                       task: cbs_ta.cpp:283-367)
// included UNMODIFIED (its main() renamed by the preprocessor).  The adapter below forwards to the
// Environment exactly as CBS's private LowLevelEnvironment does (cbs.hpp:209-244).
// stdin:  dimx dimy n_obst, the obstacles "x y"; n_jobs; per job:
//         sx sy gx gy (gx < 0: no task, cbs_ta only)  n_vc, then "t x y" each;  n_ec, then "t x1 y1 x2 y2" each
// stdout: per job "J cost length expanded" (the search is unbounded, as in the reference: only give it
//         jobs that have a solution)
#define main reference_example_main
#if defined(PROBE_CBS)
#include "cbs.cpp"
#elif defined(PROBE_CBSTA)
#include "cbs_ta.cpp"
#else
#error "define PROBE_CBS or PROBE_CBSTA"
#endif
#undef main

#include <cstdio>

struct ProbeLowLevel {
  explicit ProbeLowLevel(Environment& env) : m_env(env) {}
  int admissibleHeuristic(const State& s) { return m_env.admissibleHeuristic(s); }
  bool isSolution(const State& s) { return m_env.isSolution(s); }
  void getNeighbors(const State& s, std::vector<Neighbor<State, Action, int> >& n) { m_env.getNeighbors(s, n); }
  void onExpandNode(const State& s, int f, int g) { m_env.onExpandLowLevelNode(s, f, g); }
  void onDiscover(const State&, int, int) {}
  Environment& m_env;
};

int main() {
  int dimx, dimy, nObst;
  if (scanf("%d %d %d", &dimx, &dimy, &nObst) != 3) return 2;
  std::unordered_set<Location> obstacles;
  for (int i = 0; i < nObst; ++i) {
    int x, y;
    if (scanf("%d %d", &x, &y) != 2) return 2;
    obstacles.insert(Location(x, y));
  }
  int nJobs;
  if (scanf("%d", &nJobs) != 1) return 2;
#if defined(PROBE_CBSTA)
  // one Environment for all jobs (its constructor runs Floyd-Warshall over the map); no agents
  Environment env(dimx, dimy, obstacles, std::vector<State>(), std::vector<std::unordered_set<Location> >(), 1);
#endif
  for (int j = 0; j < nJobs; ++j) {
    int sx, sy, gx, gy, nvc, nec;
    if (scanf("%d %d %d %d %d", &sx, &sy, &gx, &gy, &nvc) != 5) return 2;
    Constraints c;
    for (int k = 0; k < nvc; ++k) {
      int t, x, y;
      if (scanf("%d %d %d", &t, &x, &y) != 3) return 2;
      c.vertexConstraints.emplace(VertexConstraint(t, x, y));
    }
    if (scanf("%d", &nec) != 1) return 2;
    for (int k = 0; k < nec; ++k) {
      int t, x1, y1, x2, y2;
      if (scanf("%d %d %d %d %d", &t, &x1, &y1, &x2, &y2) != 5) return 2;
      c.edgeConstraints.emplace(EdgeConstraint(t, x1, y1, x2, y2));
    }
#if defined(PROBE_CBS)
    Environment env(dimx, dimy, obstacles, std::vector<Location>(1, Location(gx, gy)));
    env.setLowLevelContext(0, &c);
#else
    const Location goal(gx, gy);
    env.setLowLevelContext(0, &c, gx >= 0 ? &goal : nullptr);
#endif
    const int before = env.lowLevelExpanded();
    ProbeLowLevel ll(env);
    libMultiRobotPlanning::AStar<State, Action, int, ProbeLowLevel> astar(ll);
    PlanResult<State, Action, int> sol;
    const bool ok = astar.search(State(0, sx, sy), sol);
    if (!ok) {
      printf("J -1 0 %d\n", env.lowLevelExpanded() - before);
      continue;
    }
    printf("J %d %d %d\n", sol.cost, (int)sol.states.size(), env.lowLevelExpanded() - before);
  }
  return 0;
}
