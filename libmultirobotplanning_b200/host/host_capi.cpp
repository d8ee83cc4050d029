// host_capi.cpp — C entry point of the batched high-level drivers
// (libmrp_host.so), used by the Python tests and bench.py.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <atomic>
#include <map>
#include <string>
#include <thread>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "cli.hpp"
#include "gpu_environment.hpp"
#include "hl_search.hpp"

using namespace mrp_host;

static thread_local std::string g_err;

extern "C" {

const char* mrph_last_error(void) { return g_err.c_str(); }

// algo: 0 cbs, 1 ecbs, 2 cbs_ta, 3 ecbs_ta.  Instances are CSR-flattened:
// dims[n][2]; obst_off[n+1] into obst_xy[][2]; agent_off[n+1] into
// start_cell[] / goal_cell[]; cbs_ta: pg_off[total_agents+1] into pg_cell[].
// Outputs per instance; paths: path_off[total_agents+1] into path_cell/path_g.
// per-instance budget of low-level expansions for the batches that follow (SolveOptions::maxLlTotal)
static std::atomic<long long> g_llTotal{0};
void mrph_set_ll_total(int64_t n) { g_llTotal.store(n); }

int mrph_solve_batch(int algo, int n_inst, const int32_t* dims, const int32_t* obst_off,
                     const int32_t* obst_xy, const int32_t* agent_off,
                     const int32_t* start_cell, const int32_t* goal_cell,
                     const int32_t* pg_off, const int32_t* pg_cell, float w, int64_t max_hl,
                     int32_t max_ll, double max_seconds, int64_t max_ta, int32_t* status,
                     int64_t* cost, int64_t* makespan, int64_t* lower_bound, int64_t* hl,
                     int64_t* ll, int64_t* nta, double* runtime, int32_t* path_off,
                     int32_t* path_cell, int32_t* path_g, int64_t path_cap) {
  try {
    std::vector<MapfInstance> insts(n_inst);
    for (int k = 0; k < n_inst; ++k) {
      MapfInstance& in = insts[k];
      in.dimx = dims[2 * k];
      in.dimy = dims[2 * k + 1];
      in.obstXY.assign(obst_xy + 2 * obst_off[k], obst_xy + 2 * obst_off[k + 1]);
      for (int a = agent_off[k]; a < agent_off[k + 1]; ++a) {
        in.starts.push_back(start_cell[a]);
        if (algo == 2 || algo == 3)
          in.potentialGoals.emplace_back(pg_cell + pg_off[a], pg_cell + pg_off[a + 1]);
        else
          in.goals.push_back(goal_cell[a]);
      }
    }
    SolveOptions opt;
    opt.w = w;
    opt.maxHlExpanded = max_hl;
    opt.maxLlExpanded = max_ll;
    opt.maxLlTotal = (long)g_llTotal.load();
    opt.maxSeconds = max_seconds;
    opt.maxTaskAssignments = max_ta;
    // One lock-step batch per map size, cut into sub-batches that run in their
    // own host thread and their own lane of the C ABI (own streams and replan
    // workspace, mrp_set_lane): a lock-step iteration lasts as long as its
    // slowest replan, so with one batch every instance waits for the slowest
    // replan of all of them; sub-batches only wait for their own, and their
    // launches overlap on the device.  Instances do not interact, so the
    // results do not depend on the cut.  MRP_HOST_LANES=1 restores one batch.
    std::map<std::pair<int, int>, std::vector<int> > groups;
    for (int k = 0; k < n_inst; ++k) groups[{insts[k].dimx, insts[k].dimy}].push_back(k);
    // Small instances are bound by the host bookkeeping of their trees (8x8 set:
    // 2.4 s with 4 lanes, 1.0 s with 16), large ones by the longest replan of an
    // iteration, and more than 8 lanes of long kernels need
    // CUDA_DEVICE_MAX_CONNECTIONS > 8 to overlap.
    long agents = 0;
    for (int k = 0; k < n_inst; ++k) agents += (long)insts[k].starts.size();
    int maxLanes = std::min(n_inst > 0 && agents / n_inst < 32 ? 16 : 4, mrp_max_lanes());
    if (const char* e = getenv("MRP_HOST_LANES")) maxLanes = std::max(1, std::min(atoi(e), mrp_max_lanes()));
    int perLane = 48;  // instances per sub-batch the cut aims at
    if (const char* e = getenv("MRP_HOST_LANE_SIZE")) perLane = std::max(1, atoi(e));
    std::vector<std::vector<int> > parts;
    for (const auto& g : groups) {
      const int n = (int)g.second.size();
      const int L = std::max(1, std::min(maxLanes, n / perLane));
      const size_t base = parts.size();
      parts.resize(base + L);
      for (int j = 0; j < n; ++j) parts[base + j % L].push_back(g.second[j]);  // round-robin: mixes difficulty
    }
    std::vector<SolveResult> results(n_inst);
    const int nThreads = std::min<int>(maxLanes, (int)parts.size());
    std::vector<std::string> errors(nThreads);
#ifdef _OPENMP
    const int ompBudget = omp_get_max_threads();  // OMP_NUM_THREADS: ranks that share a host divide the cores
#endif
    auto work = [&](int t) {
      try {
        SolveOptions myOpt = opt;
        if (nThreads > 1) {
          mrp_set_lane(t);
#ifdef _OPENMP
          omp_set_num_threads(std::max(1, ompBudget / nThreads));
#endif
        }
        for (size_t q = t; q < parts.size(); q += nThreads) {
          std::vector<MapfInstance> sub;
          for (int k : parts[q]) sub.push_back(insts[k]);
          BatchSolver solver(static_cast<Algo>(algo), sub, myOpt);
          std::vector<SolveResult> r = solver.run();
          for (size_t j = 0; j < parts[q].size(); ++j) results[parts[q][j]] = std::move(r[j]);
        }
      } catch (const std::exception& e) {
        errors[t] = e.what();
        if (errors[t].empty()) errors[t] = "unknown error";
      }
    };
    if (nThreads <= 1) {
      work(0);
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < nThreads; ++t) pool.emplace_back(work, t);
      for (auto& th : pool) th.join();
    }
    for (const std::string& e : errors)
      if (!e.empty()) throw std::runtime_error(e);
    int64_t off = 0;
    for (int k = 0; k < n_inst; ++k) {
      const SolveResult& r = results[k];
      status[k] = r.status;
      cost[k] = r.cost;
      makespan[k] = r.makespan;
      lower_bound[k] = r.lowerBound;
      hl[k] = r.hlExpanded;
      ll[k] = r.llExpanded;
      nta[k] = r.numTaskAssignments;
      runtime[k] = r.runtime;
      for (int a = agent_off[k]; a < agent_off[k + 1]; ++a) {
        path_off[a] = (int32_t)off;
        if (r.status != kSolved) continue;
        const AgentPath& p = r.paths[a - agent_off[k]];
        for (size_t t = 0; t < p.cells.size(); ++t) {
          if (off < path_cap) {
            path_cell[off] = p.cells[t];
            path_g[off] = p.g[t];
          }
          ++off;
        }
      }
    }
    path_off[agent_off[n_inst]] = (int32_t)off;
    return off > path_cap ? -2 : 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// Smoke test of the concept adapter (gpu_environment.hpp): builds an
// Environment, exercises every callback once and returns the first-conflict
// time of two straight-line plans (-1: none).  Keeps the adapter compiled and
// linked even though the CLIs use the batched path.
int mrph_environment_selftest(void) {
  try {
    std::unordered_set<Location> obstacles;
    std::vector<Location> goals = {Location(2, 0), Location(0, 0)};
    Environment env(3, 1, obstacles, goals);
    Constraints none;
    env.setLowLevelContext(0, &none);
    if (env.admissibleHeuristic(State(0, 0, 0)) != 2) return -10;
    if (env.isSolution(State(0, 0, 0)) || !env.isSolution(State(2, 2, 0))) return -11;
    std::vector<Neighbor<State, Action, int> > nb;
    env.getNeighbors(State(0, 1, 0), nb);
    if (nb.size() != 3) return -12;
    std::vector<Environment::Plan> sol(2);
    for (int t = 0; t < 3; ++t) {
      sol[0].states.push_back({State(t, t, 0), t});
      sol[1].states.push_back({State(t, 2 - t, 0), t});
    }
    Conflict c;
    if (!env.getFirstConflict(sol, c)) return -13;
    std::map<size_t, Constraints> cons;
    env.createConstraintsFromConflict(c, cons);
    if (cons.size() != 2) return -14;
    if (env.focalHeuristic(sol) < 1) return -15;
    if (env.focalStateHeuristic(State(1, 1, 0), 1, sol) != 1) return -16;
    env.onExpandHighLevelNode(0);
    env.onExpandLowLevelNode(State(0, 0, 0), 0, 0);
    return c.time * 10 + (c.type == Conflict::Vertex ? 0 : 1);
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"

// The host assignment module on its own (CPU only; tests and timing):
// edges[n][3] = (agent, task, cost); enumerates up to max_solutions assignments
// in non-decreasing cost (next_best_assignment.hpp:37-122 of the reference).
// sol[k][a] = task of agent a in solution k or -1.  Returns the number of
// solutions written.
extern "C" int mrph_next_best_assignments(const int64_t* edges, int n_edges, int n_agents,
                                          int n_tasks, int max_solutions, int64_t* costs,
                                          int32_t* sol) {
  try {
    (void)n_tasks;
    NextBestAssignment<int, int> nba;
    for (int e = 0; e < n_edges; ++e)
      nba.setCost((int)edges[3 * e], (int)edges[3 * e + 1], (long)edges[3 * e + 2]);
    nba.solve();
    int n = 0;
    for (; n < max_solutions; ++n) {
      std::map<int, int> s;
      const long c = nba.nextSolution(s);
      if (s.empty()) break;
      costs[n] = c;
      for (int a = 0; a < n_agents; ++a) sol[(size_t)n * n_agents + a] = -1;
      for (const auto& kv : s) sol[(size_t)n * n_agents + kv.first] = kv.second;
    }
    return n;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// Parses an input YAML with the CLI's reader (CPU only; used by the tests).
// Returns the number of agents, or -1.  Arrays must hold `cap` entries.
extern "C" int mrph_load_instance(const char* path, int ta, int32_t* dims, int32_t* n_obst,
                                  int32_t* obst_xy, int32_t* start_cell, int32_t* goal_cell,
                                  int32_t* pg_off, int32_t* pg_cell, int cap) {
  try {
    const MapfInstance in = loadInstance(path, ta != 0);
    if ((int)in.obstXY.size() > cap || (int)in.numAgents() >= cap) {
      g_err = "instance exceeds the buffers";
      return -1;
    }
    dims[0] = in.dimx;
    dims[1] = in.dimy;
    *n_obst = (int)in.obstXY.size() / 2;
    for (size_t k = 0; k < in.obstXY.size(); ++k) obst_xy[k] = in.obstXY[k];
    int off = 0;
    for (size_t a = 0; a < in.numAgents(); ++a) {
      start_cell[a] = in.starts[a];
      if (ta) {
        pg_off[a] = off;
        for (int g : in.potentialGoals[a]) {
          if (off >= cap) return -1;
          pg_cell[off++] = g;
        }
      } else {
        goal_cell[a] = in.goals[a];
      }
    }
    if (ta) pg_off[in.numAgents()] = off;
    return (int)in.numAgents();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
