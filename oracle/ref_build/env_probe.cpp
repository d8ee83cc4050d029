// env_probe — the reference's OWN Environment methods behind a pipe (TEST INFRASTRUCTURE).
// The example file is included UNMODIFIED from /root/reference/example (its main() renamed by
// the preprocessor), so the methods called below are the reference's code, not a restatement:
//   -DPROBE_CBS    example/cbs.cpp     Environment::getFirstConflict            (:335-386)
//   -DPROBE_ECBS   example/ecbs.cpp    getFirstConflict (:401-452), focalHeuristic (:315-350),
//                                      focalStateHeuristic / focalTransitionHeuristic (:282-312)
//   -DPROBE_CBSTA  example/cbs_ta.cpp  Environment::getFirstConflict, the other loop bound (:369-420)
// stdin:  n_tables, then per table: N; per agent: len x0 y0 ... ; n_queries; per query (ecbs only):
//         self t fx fy tx ty  = agent `self` moves from (fx, fy) at time t to (tx, ty) at t + 1
// stdout: per table "F found time agent1 agent2 type x1 y1 x2 y2" (x2 = y2 = -1 unless Edge),
//         ecbs: "C count", and per query "Q stateCount transitionCount"
#define main reference_example_main
#if defined(PROBE_CBS)
#include "cbs.cpp"
#elif defined(PROBE_ECBS)
#include "ecbs.cpp"
#elif defined(PROBE_CBSTA)
#include "cbs_ta.cpp"
#else
#error "define PROBE_CBS, PROBE_ECBS or PROBE_CBSTA"
#endif
#undef main

#include <cstdio>

int main() {
  int nTables;
  if (scanf("%d", &nTables) != 1) return 2;
  std::unordered_set<Location> obstacles;
#if defined(PROBE_CBSTA)
  // 2 x 2 map, no agents: the constructor's Floyd-Warshall and assignment have nothing to do
  Environment env(2, 2, obstacles, std::vector<State>(), std::vector<std::unordered_set<Location> >(), 1);
#else
  Environment env(1 << 20, 1 << 20, obstacles, std::vector<Location>());
#endif
  for (int tb = 0; tb < nTables; ++tb) {
    int N;
    if (scanf("%d", &N) != 1) return 2;
    std::vector<PlanResult<State, Action, int> > solution(N);
    for (int a = 0; a < N; ++a) {
      int len;
      if (scanf("%d", &len) != 1) return 2;
      for (int t = 0; t < len; ++t) {
        int x, y;
        if (scanf("%d %d", &x, &y) != 2) return 2;
        solution[a].states.push_back(std::make_pair(State(t, x, y), t));
      }
      solution[a].cost = len ? len - 1 : 0;
    }
    Conflict c;
    c.time = -1;
    c.agent1 = c.agent2 = 0;
    c.type = Conflict::Vertex;
    c.x1 = c.y1 = c.x2 = c.y2 = -1;
    const bool found = env.getFirstConflict(solution, c);
    if (!found || c.type == Conflict::Vertex) c.x2 = c.y2 = -1;
    printf("F %d %d %d %d %d %d %d %d %d\n", (int)found, c.time, (int)c.agent1, (int)c.agent2, (int)c.type,
           c.x1, c.y1, c.x2, c.y2);
#if defined(PROBE_ECBS)
    printf("C %d\n", env.focalHeuristic(solution));
#endif
    int nq;
    if (scanf("%d", &nq) != 1) return 2;
    for (int q = 0; q < nq; ++q) {
      int self, t, fx, fy, tx, ty;
      if (scanf("%d %d %d %d %d %d", &self, &t, &fx, &fy, &tx, &ty) != 6) return 2;
#if defined(PROBE_ECBS)
      Constraints none;
      env.setLowLevelContext(self, &none);  // needs a goal per agent only for its goal test
      const State a(t, fx, fy), b(t + 1, tx, ty);
      printf("Q %d %d\n", env.focalStateHeuristic(b, t + 1, solution),
             env.focalTransitionHeuristic(a, b, t, t + 1, solution));
#else
      printf("Q 0 0\n");
#endif
    }
  }
  return 0;
}
