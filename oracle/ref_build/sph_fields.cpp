// sph_fields — the reference's OWN ShortestPathHeuristic (example/shortest_path_heuristic.hpp,
// included unmodified from /root/reference; Boost.Graph calls resolved by the stand-in headers)
// behind a pipe, so that the oracle's per-goal BFS fields — and through them the CUDA fields —
// can be compared with getValue() of the reference class itself.  TEST INFRASTRUCTURE.
//
// stdin  (text):  dimx dimy n_obst n_goals, then n_obst pairs "x y", then n_goals pairs "x y"
// stdout (binary): int32 out[n_goals][dimy][dimx], out[g][y][x] = getValue(Location(x, y), goal g)
//                  (shortest_path_heuristic.hpp:58-62; unreachable = INT_MAX as Floyd-Warshall leaves it)
// The class writes "searchGraph.dot" into the working directory (shortest_path_heuristic.hpp:44):
// run it in a scratch directory.
#include <cstdint>
#include <cstdio>
#include <functional>
#include <iostream>
#include <tuple>
#include <unordered_set>
#include <vector>

struct Location {
  Location(int x, int y) : x(x), y(y) {}
  int x, y;
  bool operator==(const Location& o) const { return x == o.x && y == o.y; }
};
namespace std {
template <>
struct hash<Location> {
  size_t operator()(const Location& l) const { return std::hash<long long>()(((long long)l.x << 32) ^ (unsigned)l.y); }
};
}  // namespace std

#include "shortest_path_heuristic.hpp"

int main() {
  int dimx, dimy, nObst, nGoals;
  if (scanf("%d %d %d %d", &dimx, &dimy, &nObst, &nGoals) != 4) return 2;
  std::unordered_set<Location> obstacles;
  for (int i = 0; i < nObst; ++i) {
    int x, y;
    if (scanf("%d %d", &x, &y) != 2) return 2;
    obstacles.insert(Location(x, y));
  }
  std::vector<Location> goals;
  for (int i = 0; i < nGoals; ++i) {
    int x, y;
    if (scanf("%d %d", &x, &y) != 2) return 2;
    goals.push_back(Location(x, y));
  }
  ShortestPathHeuristic h(dimx, dimy, obstacles);
  std::vector<int32_t> row((size_t)dimx * dimy);
  for (const Location& g : goals) {
    for (int y = 0; y < dimy; ++y)
      for (int x = 0; x < dimx; ++x) row[x + (size_t)dimx * y] = h.getValue(Location(x, y), g);
    fwrite(row.data(), sizeof(int32_t), row.size(), stdout);
  }
  return 0;
}
