/*
 * mrp_oracle.h — C interface of the CPU ORACLE.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It restates, on the CPU, the
 * algorithms of the libMultiRobotPlanning hot path so that the CUDA path can
 * be checked against them.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.
 *
 * Parity status: the reference cannot be compiled as shipped in this image
 * (Boost and yaml-cpp are absent).  What can be compiled is its own source,
 * UNMODIFIED, against small stand-in headers for those libraries
 * (oracle/ref_build -> oracle/_ref/), and the oracle is pinned against that:
 *  (1) whole searches: cbs / ecbs on 582 + 36 benchmark instances with identical
 *      cost, makespan and expansion counts (tests/golden/ref_binary_golden.json);
 *      cbs_ta / ecbs_ta (with a stand-in for the slice of Boost.Graph and
 *      boost::bimap that assignment.hpp and shortest_path_heuristic.hpp use) on
 *      594 task-assignment instances: every optimal sum of costs and every "no
 *      solution" (tests/golden/ref_binary_golden_ta.json);
 *  (2) functions: distance fields against getValue() of the reference's own
 *      ShortestPathHeuristic class on 70 maps (oracle/_ref/sph_fields,
 *      tests/golden/sph_fields_golden.json); getFirstConflict under both loop
 *      bounds, focalHeuristic, focalStateHeuristic and focalTransitionHeuristic
 *      against the reference's own Environment classes on 40 path tables
 *      (oracle/_ref/env_probe_*, tests/golden/env_probe_golden.json); the costs of 140
 *      constrained low-level replans against the reference's AStar::search through
 *      its cbs / cbs_ta Environments (oracle/_ref/astar_probe_*,
 *      tests/golden/astar_probe_golden.json);
 *  (3) the known answers held by the reference's own tests (test/test_cbs.py:24-34,
 *      test/test_ecbs.py:25-35, test/test_cbs_ta.py:24-38, test/test_assignment.py:19-63,
 *      test/test_next_best_assignment.py:19-110) — the compiled binaries pass all
 *      21 of those tests, and tests/test_oracle_pinned.py asserts the same values
 *      for the oracle.
 * Tie-breaking among equal-cost optimal paths / equal-cost tree nodes / equal-cost
 * assignments depends on Boost.Heap and Boost.Graph internals and is "parity
 * unpinned": only costs (and, for cbs / ecbs, the expansion counts the stand-in
 * heap happens to reproduce) are compared.
 */
#ifndef MRP_ORACLE_H
#define MRP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_INF 2147483647 /* std::numeric_limits<int>::max() */

/* status codes of the search drivers */
#define ORC_SOLVED 0
#define ORC_NO_SOLUTION 1 /* OPEN ran empty (reference returns false) */
#define ORC_CAPPED 2      /* an expansion cap was hit (the reference has none) */

typedef struct {
  int32_t time;
  int32_t agent1;
  int32_t agent2;
  int32_t type; /* 0 = Vertex, 1 = Edge (example/cbs.cpp:81-84) */
  int32_t x1, y1, x2, y2; /* x2,y2 = -1 for Vertex */
} orc_conflict;

typedef struct {
  int32_t dimx, dimy;
  int32_t n_obst;
  const int32_t* obst_xy; /* [n_obst][2] */
  int32_t n_agents;
  const int32_t* start_xy; /* [n_agents][2] */
  const int32_t* goal_xy;  /* [n_agents][2]  (cbs / ecbs) */
  const int32_t* pg_off;   /* [n_agents+1]   (cbs_ta: CSR of potentialGoals) */
  const int32_t* pg_xy;    /* [pg_off[n_agents]][2] */
} orc_instance;

typedef struct {
  int32_t status;
  int64_t cost;
  int64_t makespan;
  int64_t lower_bound; /* ECBS: sum of fmin of the returned node */
  int64_t hl_expanded;
  int64_t ll_expanded;
  int64_t n_task_assignments;
  double runtime_s; /* wall time of search() only (example/cbs.cpp:624-626) */
} orc_result;

typedef struct {
  int64_t max_hl_expanded; /* <=0: unlimited */
  int64_t max_ll_expanded; /* per low-level search; <=0: unlimited */
  double max_seconds;      /* <=0: unlimited */
} orc_caps;

/* ---- distance fields (example/shortest_path_heuristic.hpp:12-65) ---- */
/* Floyd–Warshall all pairs, out[V*V], V = dimx*dimy, vertex id x + dimx*y. */
int orc_floyd_warshall(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                       int32_t* out);
/* queue BFS per goal, out[n_goals][dimy*dimx]; ORC_INF = unreachable/obstacle
 * (example/cbs.cpp:445-557).  A goal that is itself an obstacle yields the
 * Floyd–Warshall row of that vertex: 0 at the goal, ORC_INF elsewhere. */
int orc_bfs_fields(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                   const int32_t* goal_xy, int n_goals, int32_t* out);

/* ---- conflicts over packed path tables cell[N][Tpad], len[N] ---- */
/* mode 0: max_t = max(len)-1 (example/cbs.cpp:338-341); mode 1: max(len)
 * (example/cbs_ta.cpp:372-375). returns 1 if a conflict was found. */
int orc_first_conflict(const int32_t* cell, const int32_t* len, int N, int Tpad,
                       int dimx, int mode, orc_conflict* out);
/* example/ecbs.cpp:315-350 (bound as mode 0; mode 1 offered for symmetry). */
int orc_count_conflicts(const int32_t* cell, const int32_t* len, int N,
                        int Tpad, int mode, int32_t* count);
/* example/ecbs.cpp:282-312 for a list of candidate moves of agent `self`:
 * candidate k goes from cell cand_from[k] at time cand_t[k] to cell cand_to[k]
 * at time cand_t[k]+1. */
int orc_focal_counts(const int32_t* cell, const int32_t* len, int N, int Tpad,
                     int self, const int32_t* cand_t, const int32_t* cand_from,
                     const int32_t* cand_to, int n_cand, int32_t* state_cnt,
                     int32_t* trans_cnt);

/* ---- low-level search (a_star.hpp:63-161, a_star_epsilon.hpp:86-285) ---- */
/* variant 0: cbs/ecbs Environment (Manhattan h, every move costs 1);
 * variant 1: cbs_ta Environment (field h, waiting on the goal is free,
 *            goal_cell < 0 = agent without task).
 * vc: [n_vc][2] = (time, cell); ec: [n_ec][3] = (time, from, to).
 * w <= 0: A*; w >= 1: A*-epsilon with the focal heuristics evaluated against
 * the path table (oth_cell/oth_len may be NULL => all counts 0).
 * path_tcg: [path_cap][3] = (time, cell, g). */
int orc_lowlevel(int dimx, int dimy, const int32_t* obst_xy, int n_obst,
                 int variant, int start_cell, int goal_cell,
                 const int32_t* vc, int n_vc, const int32_t* ec, int n_ec,
                 float w, const int32_t* oth_cell, const int32_t* oth_len,
                 int oth_n, int oth_tpad, int self, int64_t max_expanded,
                 int32_t* cost, int32_t* fmin, int64_t* expanded,
                 int32_t* path_tcg, int path_cap, int32_t* path_len);

/* ---- high-level drivers ---- */
/* paths: path_off[n_agents+1] offsets into path_xyg[cap][3] = (x, y, g). */
int orc_cbs(const orc_instance* inst, const orc_caps* caps, orc_result* res,
            int32_t* path_off, int32_t* path_xyg, int path_cap);
int orc_ecbs(const orc_instance* inst, float w, const orc_caps* caps,
             orc_result* res, int32_t* path_off, int32_t* path_xyg,
             int path_cap);
int orc_cbs_ta(const orc_instance* inst, int64_t max_task_assignments,
               const orc_caps* caps, orc_result* res, int32_t* path_off,
               int32_t* path_xyg, int path_cap);

/* ecbs_ta.hpp:94-352 (REBUILT_FOCAL_LIST + STYLE_MINROOT), example/ecbs_ta.cpp */
int orc_ecbs_ta(const orc_instance* inst, float w, int64_t max_task_assignments,
                const orc_caps* caps, orc_result* res, int32_t* path_off,
                int32_t* path_xyg, int path_cap);

/* ---- assignment (assignment.hpp:34-118, next_best_assignment.hpp:37-201) ---- */
/* edges: [n_edges][3] = (agent, task, cost).  sol_task[a] = task or -1.
 * returns total cost. n_agents_out agents are those that appear in edges. */
int64_t orc_assignment(const int64_t* edges, int n_edges, int n_agents,
                       int n_tasks, int32_t* sol_task);
/* enumerates up to max_solutions next-best assignments;
 * sol_task[k][n_agents], costs[k]; returns number of solutions produced. */
int orc_next_best_assignments(const int64_t* edges, int n_edges, int n_agents,
                              int n_tasks, int max_solutions, int64_t* costs,
                              int32_t* sol_task);

#ifdef __cplusplus
}
#endif
#endif
