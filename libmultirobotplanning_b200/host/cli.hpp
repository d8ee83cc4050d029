// cli.hpp — command line, input and output of the cbs / ecbs / cbs_ta drop-in
// binaries.  Flags, messages, exit codes and the output.yaml layout follow
// example/cbs.cpp:571-667, example/ecbs.cpp:524-623, example/cbs_ta.cpp:516-620
// (boost::program_options and yaml-cpp are not available in this image).
#pragma once

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "hl_search.hpp"
#include "yaml_lite.hpp"

namespace mrp_host {

struct CliArgs {
  std::string input, output;
  float w = 1.0f;
  size_t maxTaskAssignments = 1000000000;
  bool help = false;
  // extensions (not in the reference): caps, because the reference's searches
  // are unbounded
  long maxHlExpanded = 0;
  double maxSeconds = 0;
};

inline void printUsage(std::ostream& os, bool hasW, bool hasTA) {
  os << "Allowed options:\n"
     << "  --help                       produce help message\n"
     << "  -i [ --input ] arg           input file (YAML)\n"
     << "  -o [ --output ] arg          output file (YAML)\n";
  if (hasW) os << "  -w [ --suboptimality ] arg (=1) suboptimality bound\n";
  if (hasTA)
    os << "  --maxTaskAssignments arg (=1000000000)\n"
       << "                               maximum number of task assignments to try\n";
  os << "  --maxHighLevelExpansions arg (=0, unlimited)\n"
     << "  --maxSeconds arg (=0, unlimited)\n";
}

// returns false (and fills err) on an option error, like po::error
inline bool parseArgs(int argc, char** argv, bool hasW, bool hasTA, CliArgs& a,
                      std::string& err) {
  bool haveIn = false, haveOut = false;
  for (int i = 1; i < argc; ++i) {
    std::string arg = argv[i], val;
    bool hasVal = false;
    if (arg.rfind("--", 0) == 0) {
      size_t eq = arg.find('=');
      if (eq != std::string::npos) {
        val = arg.substr(eq + 1);
        arg = arg.substr(0, eq);
        hasVal = true;
      }
    } else if (arg.size() > 2 && arg[0] == '-') {  // -ifile
      val = arg.substr(2);
      arg = arg.substr(0, 2);
      hasVal = true;
    }
    auto need = [&](std::string& out) {
      if (hasVal) {
        out = val;
        return true;
      }
      if (i + 1 >= argc) {
        err = "the required argument for option '" + arg + "' is missing";
        return false;
      }
      out = argv[++i];
      return true;
    };
    std::string v;
    if (arg == "--help") {
      a.help = true;
    } else if (arg == "-i" || arg == "--input") {
      if (!need(a.input)) return false;
      haveIn = true;
    } else if (arg == "-o" || arg == "--output") {
      if (!need(a.output)) return false;
      haveOut = true;
    } else if (hasW && (arg == "-w" || arg == "--suboptimality")) {
      if (!need(v)) return false;
      char* end = nullptr;
      a.w = std::strtof(v.c_str(), &end);
      if (end == v.c_str() || *end) {
        err = "the argument ('" + v + "') for option '--suboptimality' is invalid";
        return false;
      }
    } else if (hasTA && arg == "--maxTaskAssignments") {
      if (!need(v)) return false;
      char* end = nullptr;
      a.maxTaskAssignments = (size_t)std::strtod(v.c_str(), &end);
      if (end == v.c_str() || *end) {
        err = "the argument ('" + v + "') for option '--maxTaskAssignments' is invalid";
        return false;
      }
    } else if (arg == "--maxHighLevelExpansions") {
      if (!need(v)) return false;
      a.maxHlExpanded = std::atol(v.c_str());
    } else if (arg == "--maxSeconds") {
      if (!need(v)) return false;
      a.maxSeconds = std::atof(v.c_str());
    } else {
      err = "unrecognised option '" + std::string(argv[i]) + "'";
      return false;
    }
  }
  if (a.help) return true;
  if (!haveIn) {
    err = "the option '--input' is required but missing";
    return false;
  }
  if (!haveOut) {
    err = "the option '--output' is required but missing";
    return false;
  }
  return true;
}

inline MapfInstance loadInstance(const std::string& path, bool ta) {
  const yaml::Node cfg = yaml::loadFile(path);
  MapfInstance in;
  const yaml::Node& dim = cfg["map"]["dimensions"];
  in.dimx = dim[0].asInt();
  in.dimy = dim[1].asInt();
  if (cfg["map"].has("obstacles")) {
    const yaml::Node& obst = cfg["map"]["obstacles"];
    for (size_t k = 0; k < obst.seq.size(); ++k) {
      in.obstXY.push_back(obst[k][0].asInt());
      in.obstXY.push_back(obst[k][1].asInt());
    }
  }
  auto cellOf = [&in](const yaml::Node& xy) {
    const int x = xy[0].asInt(), y = xy[1].asInt();
    if (x < 0 || y < 0 || x >= in.dimx || y >= in.dimy)
      throw std::runtime_error("location outside the map");
    return x + in.dimx * y;
  };
  const yaml::Node& agents = cfg["agents"];
  for (size_t k = 0; k < agents.seq.size(); ++k) {
    const yaml::Node& ag = agents[k];
    in.starts.push_back(cellOf(ag["start"]));
    if (ta) {
      std::vector<int> pg;
      if (ag.has("potentialGoals")) {
        const yaml::Node& g = ag["potentialGoals"];
        for (size_t j = 0; j < g.seq.size(); ++j) {
          const int c = cellOf(g[j]);
          // an unordered_set<Location> in the reference: duplicates collapse
          if (std::find(pg.begin(), pg.end(), c) == pg.end()) pg.push_back(c);
        }
      }
      in.potentialGoals.push_back(pg);
    } else {
      in.goals.push_back(cellOf(ag["goal"]));
    }
  }
  return in;
}

// statistics + schedule, example/cbs.cpp:637-661 / example/cbs_ta.cpp:589-614
inline void writeOutput(std::ostream& out, const SolveResult& r, int dimx, bool ta) {
  out << "statistics:" << std::endl;
  out << "  cost: " << r.cost << std::endl;
  out << "  makespan: " << r.makespan << std::endl;
  out << "  runtime: " << r.runtime << std::endl;
  out << "  highLevelExpanded: " << r.hlExpanded << std::endl;
  out << "  lowLevelExpanded: " << r.llExpanded << std::endl;
  if (ta) out << "  numTaskAssignments: " << r.numTaskAssignments << std::endl;
  out << "schedule:" << std::endl;
  for (size_t a = 0; a < r.paths.size(); ++a) {
    out << "  agent" << a << ":" << std::endl;
    for (size_t t = 0; t < r.paths[a].cells.size(); ++t) {
      out << "    - x: " << r.paths[a].cells[t] % dimx << std::endl
          << "      y: " << r.paths[a].cells[t] / dimx << std::endl
          << "      t: " << r.paths[a].g[t] << std::endl;
    }
  }
}

// shared main(): returns the process exit code
inline int runCli(int argc, char** argv, Algo algo) {
  const bool hasW = algo == Algo::ECBS || algo == Algo::ECBSTA;
  const bool hasTA = algo == Algo::CBSTA || algo == Algo::ECBSTA;
  CliArgs args;
  std::string err;
  if (!parseArgs(argc, argv, hasW, hasTA, args, err)) {
    std::cerr << err << std::endl << std::endl;
    printUsage(std::cerr, hasW, hasTA);
    return 1;
  }
  if (args.help) {
    printUsage(std::cout, hasW, hasTA);
    std::cout << "\n";
    return 0;
  }
  std::vector<MapfInstance> insts(1, loadInstance(args.input, hasTA));
  SolveOptions opt;
  opt.w = args.w;
  opt.maxTaskAssignments = (long)args.maxTaskAssignments;
  opt.maxHlExpanded = args.maxHlExpanded;
  opt.maxSeconds = args.maxSeconds;
  BatchSolver solver(algo, insts, opt);
  const SolveResult r = solver.run()[0];
  if (r.status == kSolved) {
    std::cout << "done; cost: " << r.cost << std::endl;  // cbs.hpp:128
    std::cout << "Planning successful! " << std::endl;
    std::ofstream out(args.output);
    writeOutput(out, r, insts[0].dimx, hasTA);
  } else {
    std::cout << "Planning NOT successful!" << std::endl;
  }
  return 0;
}

}  // namespace mrp_host
