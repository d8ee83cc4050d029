// templates_gpuenv.cpp — drop-in proof at the concept level (TEST INFRASTRUCTURE).
//
// Instantiates the reference's UNMODIFIED search templates
//   libMultiRobotPlanning::CBS / ECBS / CBSTA   (include/libMultiRobotPlanning/)
// with this repository's GPU-backed Environment classes
// (libmultirobotplanning_b200/host/gpu_environment.hpp) and runs them.  The
// reference's own high-level and low-level loops then drive our callbacks:
// getFirstConflict / focalHeuristic / the heuristic precompute execute on the
// GPU through the C ABI.  Built only where /root/reference exists; the binary
// lands in oracle/_ref/ and travels to the GPU box.
//
//   templates_gpuenv cbs|ecbs|cbs_ta <input.yaml> [w]   ->  prints "cost <n>"
#include <cstdlib>
#include <iostream>
#include <string>
#include <unordered_set>
#include <vector>

#include <libMultiRobotPlanning/cbs.hpp>
#include <libMultiRobotPlanning/cbs_ta.hpp>
#include <libMultiRobotPlanning/ecbs.hpp>

#include "../../libmultirobotplanning_b200/host/gpu_environment.hpp"
#include "../../libmultirobotplanning_b200/host/yaml_lite.hpp"

using namespace mrp_host;
using libMultiRobotPlanning::CBS;
using libMultiRobotPlanning::CBSTA;
using libMultiRobotPlanning::ECBS;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string algo = argv[1];
  const float w = argc > 3 ? (float)atof(argv[3]) : 1.0f;
  const yaml::Node cfg = yaml::loadFile(argv[2]);
  const int dimx = cfg["map"]["dimensions"][0].asInt(), dimy = cfg["map"]["dimensions"][1].asInt();
  std::unordered_set<Location> obstacles;
  if (cfg["map"].has("obstacles"))
    for (const auto& o : cfg["map"]["obstacles"].seq)
      obstacles.insert(Location(o[0].asInt(), o[1].asInt()));
  std::vector<State> starts;
  std::vector<Location> goals;
  std::vector<std::unordered_set<Location> > potentialGoals;
  for (const auto& a : cfg["agents"].seq) {
    starts.emplace_back(State(0, a["start"][0].asInt(), a["start"][1].asInt()));
    if (a.has("goal")) goals.emplace_back(Location(a["goal"][0].asInt(), a["goal"][1].asInt()));
    if (a.has("potentialGoals")) {
      potentialGoals.emplace_back();
      for (const auto& g : a["potentialGoals"].seq)
        potentialGoals.back().insert(Location(g[0].asInt(), g[1].asInt()));
    }
  }
  typedef libMultiRobotPlanning::PlanResult<State, Action, int> RefPlan;
  std::vector<RefPlan> solution;
  bool ok = false;
  try {
    if (algo == "cbs") {
      Environment env(dimx, dimy, obstacles, goals);
      CBS<State, Action, int, Conflict, Constraints, Environment> search(env);
      ok = search.search(starts, solution);
    } else if (algo == "ecbs") {
      Environment env(dimx, dimy, obstacles, goals);
      ECBS<State, Action, int, Conflict, Constraints, Environment> search(env, w);
      ok = search.search(starts, solution);
    } else if (algo == "cbs_ta") {
      EnvironmentTA env(dimx, dimy, obstacles, starts, potentialGoals, 1000000000);
      CBSTA<State, Action, int, Conflict, Constraints, Location, EnvironmentTA> search(env);
      ok = search.search(starts, solution);
    } else {
      return 2;
    }
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 3;
  }
  if (!ok) {
    std::cout << "no solution" << std::endl;
    return 1;
  }
  long cost = 0;
  for (const auto& s : solution) cost += s.cost;
  std::cout << "cost " << cost << std::endl;
  return 0;
}
