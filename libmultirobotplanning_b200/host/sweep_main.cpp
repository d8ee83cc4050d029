// mapf_sweep — the reference's benchmark sweep (one `cbs` / `ecbs` / `cbs_ta`
// process per YAML file of benchmark/8x8_obst12 and benchmark/32x32_obst204) as
// ONE streaming run: a reader thread parses the input files (the reference's
// input format, example/cbs.cpp:604-618 / example/cbs_ta.cpp:561-567) while the
// batched driver (hl_search.hpp) solves the previous batch on the GPU, so the
// 3000-file sweep needs neither 3000 process starts nor all files in memory.
//
//   mapf_sweep --algo cbs|ecbs|cbs_ta|ecbs_ta [-w W] [--batch N] [--outputDir DIR]
//              [--maxHighLevelExpansions N] [--maxSeconds S] [--csv FILE] FILE.yaml ... | --list FILE
//
// Per input file one CSV row `file,status,cost,makespan,lowerBound,highLevelExpanded,
// lowLevelExpanded,runtime` (status 0 solved, 1 no solution, 2 capped); with
// --outputDir also DIR/<name>.output.yaml in the reference's output format
// (example/cbs.cpp:637-661) for the solved ones.  Files of different map sizes
// may be mixed: a batch is cut where the size changes.
#include <condition_variable>
#include <deque>
#include <fstream>
#include <iostream>
#include <mutex>
#include <thread>

#include "cli.hpp"

using namespace mrp_host;

namespace {
struct Item {
  std::string path;
  MapfInstance inst;
  std::string error;
};
std::string baseName(const std::string& p) {
  const size_t s = p.find_last_of('/');
  std::string b = s == std::string::npos ? p : p.substr(s + 1);
  const size_t d = b.rfind(".yaml");
  if (d != std::string::npos && d + 5 == b.size()) b = b.substr(0, d);
  return b;
}
}  // namespace

int main(int argc, char* argv[]) {
  std::string algoName = "cbs", outputDir, csvPath, listPath;
  SolveOptions opt;
  size_t batch = 256;
  std::vector<std::string> files;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto val = [&](const char* name) -> std::string {
      if (i + 1 >= argc) {
        std::cerr << "the required argument for option '" << name << "' is missing\n";
        exit(1);
      }
      return argv[++i];
    };
    if (a == "--algo") algoName = val("--algo");
    else if (a == "-w" || a == "--suboptimality") opt.w = std::strtof(val("-w").c_str(), nullptr);
    else if (a == "--batch") batch = (size_t)std::max(1L, std::atol(val("--batch").c_str()));
    else if (a == "--outputDir") outputDir = val("--outputDir");
    else if (a == "--csv") csvPath = val("--csv");
    else if (a == "--list") listPath = val("--list");
    else if (a == "--maxHighLevelExpansions") opt.maxHlExpanded = std::atol(val("--maxHighLevelExpansions").c_str());
    else if (a == "--maxSeconds") opt.maxSeconds = std::atof(val("--maxSeconds").c_str());
    else if (a == "--maxTaskAssignments") opt.maxTaskAssignments = (long)std::strtod(val("--maxTaskAssignments").c_str(), nullptr);
    else if (a == "--help") {
      std::cout << "usage: mapf_sweep --algo cbs|ecbs|cbs_ta|ecbs_ta [-w W] [--batch N] [--outputDir DIR] "
                   "[--maxHighLevelExpansions N] [--maxSeconds S] [--csv FILE] FILE.yaml ... | --list FILE\n";
      return 0;
    } else if (a.rfind("--", 0) == 0) {
      std::cerr << "unrecognised option '" << a << "'\n";
      return 1;
    } else files.push_back(a);
  }
  if (!listPath.empty()) {
    std::ifstream in(listPath);
    for (std::string line; std::getline(in, line);)
      if (!line.empty()) files.push_back(line);
  }
  Algo algo;
  if (algoName == "cbs") algo = Algo::CBS;
  else if (algoName == "ecbs") algo = Algo::ECBS;
  else if (algoName == "cbs_ta") algo = Algo::CBSTA;
  else if (algoName == "ecbs_ta") algo = Algo::ECBSTA;
  else {
    std::cerr << "unknown --algo '" << algoName << "'\n";
    return 1;
  }
  const bool ta = algo == Algo::CBSTA || algo == Algo::ECBSTA;

  // reader thread: at most two batches ahead of the solver
  std::mutex m;
  std::condition_variable cv;
  std::deque<Item> queue;
  bool readerDone = false;
  std::thread reader([&] {
    for (const std::string& f : files) {
      Item it;
      it.path = f;
      try {
        it.inst = loadInstance(f, ta);
      } catch (const std::exception& e) {
        it.error = e.what();
        if (it.error.empty()) it.error = "cannot read";
      }
      std::unique_lock<std::mutex> lk(m);
      cv.wait(lk, [&] { return queue.size() < 2 * batch; });
      queue.push_back(std::move(it));
      cv.notify_all();
    }
    std::lock_guard<std::mutex> lk(m);
    readerDone = true;
    cv.notify_all();
  });

  std::ofstream csvFile;
  if (!csvPath.empty()) csvFile.open(csvPath);
  std::ostream& csv = csvPath.empty() ? std::cout : csvFile;
  csv << "file,status,cost,makespan,lowerBound,highLevelExpanded,lowLevelExpanded,runtime\n";
  long nSolved = 0, nTotal = 0;
  int rc = 0;
  try {
    while (true) {
      std::vector<Item> items;
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return !queue.empty() || readerDone; });
        if (queue.empty()) break;
        // one batch: up to `batch` consecutive files of one map size
        const int dx = queue.front().inst.dimx, dy = queue.front().inst.dimy;
        const bool bad = !queue.front().error.empty();
        while (!queue.empty() && items.size() < batch && !queue.front().error.empty() == bad &&
               (bad || (queue.front().inst.dimx == dx && queue.front().inst.dimy == dy))) {
          items.push_back(std::move(queue.front()));
          queue.pop_front();
        }
        cv.notify_all();
      }
      if (!items[0].error.empty()) {
        for (const Item& it : items) {
          csv << it.path << ",-1,0,0,0,0,0,0\n";
          std::cerr << it.path << ": " << it.error << "\n";
          ++nTotal;
          rc = 2;
        }
        continue;
      }
      std::vector<MapfInstance> insts;
      for (Item& it : items) insts.push_back(std::move(it.inst));
      BatchSolver solver(algo, insts, opt);
      const std::vector<SolveResult> res = solver.run();
      for (size_t k = 0; k < res.size(); ++k) {
        const SolveResult& r = res[k];
        csv << items[k].path << ',' << r.status << ',' << r.cost << ',' << r.makespan << ',' << r.lowerBound << ','
            << r.hlExpanded << ',' << r.llExpanded << ',' << r.runtime << '\n';
        ++nTotal;
        if (r.status == kSolved) {
          ++nSolved;
          if (!outputDir.empty()) {
            std::ofstream out(outputDir + "/" + baseName(items[k].path) + ".output.yaml");
            writeOutput(out, r, insts[k].dimx, ta);
          }
        }
      }
      csv.flush();
    }
  } catch (const std::exception& e) {
    std::cerr << "mapf_sweep: " << e.what() << std::endl;
    rc = 134;
    {
      std::lock_guard<std::mutex> lk(m);
      queue.clear();
      files.clear();
    }
    cv.notify_all();
  }
  {
    // let the reader run out (its waits see an empty queue)
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] {
      queue.clear();
      cv.notify_all();
      return readerDone;
    });
  }
  reader.join();
  std::cerr << "mapf_sweep: " << nSolved << "/" << nTotal << " solved" << std::endl;
  return rc;
}
