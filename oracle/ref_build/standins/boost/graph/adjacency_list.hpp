// Stand-in for the part of Boost.Graph the reference's task-assignment path uses
// (TEST INFRASTRUCTURE; Boost is absent from this image and there is no network):
//   include/libMultiRobotPlanning/assignment.hpp:30-165 — adjacency_list<vecS, vecS,
//     bidirectionalS, Vertex, Edge> with bundled properties, add_vertex / add_edge /
//     edge / remove_edge / out_edges / target, get(&Edge::member, g), named parameters,
//     successive_shortest_path_nonnegative_weights
//   example/shortest_path_heuristic.hpp:12-54 — adjacency_list<vecS, vecS, undirectedS,
//     Vertex, Edge>, exterior_vertex_property, floyd_warshall_all_pairs_shortest_paths,
//     write_graphviz
// Written from the published interface of those calls; the two algorithms are the
// textbook ones Boost documents (successive shortest augmenting paths with vertex
// potentials; the Floyd-Warshall triple loop with an "infinity" of
// numeric_limits<T>::max() that is never added to).  With real Boost the SAME optimal
// costs come out; WHICH of several equal-cost matchings the flow returns may differ
// (Dijkstra's tie-breaking), which changes the order of equal-cost assignments in
// cbs_ta / ecbs_ta but not an optimal sum of costs.
#pragma once
#include <cstddef>
#include <functional>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <memory>
#include <ostream>
#include <queue>
#include <utility>
#include <vector>

namespace boost {

struct vecS {};
struct directedS {};
struct undirectedS {};
struct bidirectionalS {};

namespace standin_detail {
struct EdgeDesc {
  std::size_t src = 0, dst = 0, id = (std::size_t)-1;
  bool operator<(const EdgeDesc& o) const { return id < o.id; }
  bool operator==(const EdgeDesc& o) const { return id == o.id; }
  bool operator!=(const EdgeDesc& o) const { return id != o.id; }
};
}  // namespace standin_detail

template <class OutEdgeList, class VertexList, class Directed>
struct adjacency_list_traits {
  typedef std::size_t vertex_descriptor;
  typedef standin_detail::EdgeDesc edge_descriptor;
};

struct no_property {};

template <class OutEdgeList, class VertexList, class Directed, class VertexProp = no_property,
          class EdgeProp = no_property>
class adjacency_list {
 public:
  typedef std::size_t vertex_descriptor;
  typedef standin_detail::EdgeDesc edge_descriptor;
  typedef Directed directed_selector;
  typedef EdgeProp edge_bundled;
  typedef VertexProp vertex_bundled;

  struct Stored {
    std::size_t dst, id;
  };
  class out_edge_iterator {
   public:
    out_edge_iterator() {}
    out_edge_iterator(std::size_t v, const Stored* p) : v_(v), p_(p) {}
    edge_descriptor operator*() const {
      edge_descriptor e;
      e.src = v_;
      e.dst = p_->dst;
      e.id = p_->id;
      return e;
    }
    out_edge_iterator& operator++() {
      ++p_;
      return *this;
    }
    bool operator!=(const out_edge_iterator& o) const { return p_ != o.p_; }
    bool operator==(const out_edge_iterator& o) const { return p_ == o.p_; }

   private:
    std::size_t v_ = 0;
    const Stored* p_ = nullptr;
  };

  EdgeProp& operator[](const edge_descriptor& e) { return edgeProps_[e.id]; }
  const EdgeProp& operator[](const edge_descriptor& e) const { return edgeProps_[e.id]; }
  VertexProp& operator[](vertex_descriptor v) { return vertexProps_[v]; }
  const VertexProp& operator[](vertex_descriptor v) const { return vertexProps_[v]; }

  // --- used by the free functions below ---
  std::vector<std::vector<Stored> > out_;
  std::vector<VertexProp> vertexProps_;
  std::vector<EdgeProp> edgeProps_;      // by edge id; ids of removed edges are reused
  std::vector<std::size_t> freeIds_;
  std::size_t numEdges_ = 0;
};

#define MRP_STANDIN_GRAPH_T adjacency_list<O, V, D, VP, EP>
#define MRP_STANDIN_GRAPH_TPL template <class O, class V, class D, class VP, class EP>

MRP_STANDIN_GRAPH_TPL
std::size_t add_vertex(MRP_STANDIN_GRAPH_T& g) {
  g.out_.emplace_back();
  g.vertexProps_.emplace_back();
  return g.out_.size() - 1;
}
MRP_STANDIN_GRAPH_TPL
std::size_t num_vertices(const MRP_STANDIN_GRAPH_T& g) { return g.out_.size(); }

MRP_STANDIN_GRAPH_TPL
std::pair<standin_detail::EdgeDesc, bool> add_edge(std::size_t u, std::size_t v,
                                                   MRP_STANDIN_GRAPH_T& g) {
  std::size_t id;
  if (!g.freeIds_.empty()) {
    id = g.freeIds_.back();
    g.freeIds_.pop_back();
    g.edgeProps_[id] = EP();
  } else {
    id = g.edgeProps_.size();
    g.edgeProps_.emplace_back();
  }
  g.out_[u].push_back({v, id});
  if (std::is_same<D, undirectedS>::value && u != v) g.out_[v].push_back({u, id});
  ++g.numEdges_;
  standin_detail::EdgeDesc e;
  e.src = u;
  e.dst = v;
  e.id = id;
  return std::make_pair(e, true);
}

MRP_STANDIN_GRAPH_TPL
std::pair<standin_detail::EdgeDesc, bool> edge(std::size_t u, std::size_t v,
                                               const MRP_STANDIN_GRAPH_T& g) {
  standin_detail::EdgeDesc e;
  for (const auto& s : g.out_[u])
    if (s.dst == v) {
      e.src = u;
      e.dst = v;
      e.id = s.id;
      return std::make_pair(e, true);
    }
  return std::make_pair(e, false);
}

MRP_STANDIN_GRAPH_TPL
void remove_edge(const standin_detail::EdgeDesc& e, MRP_STANDIN_GRAPH_T& g) {
  auto drop = [&](std::size_t from) {
    auto& lst = g.out_[from];
    for (std::size_t i = 0; i < lst.size(); ++i)
      if (lst[i].id == e.id) {
        lst.erase(lst.begin() + i);
        return true;
      }
    return false;
  };
  bool was = drop(e.src);
  if (std::is_same<D, undirectedS>::value) was = drop(e.dst) || was;
  if (was) {
    g.freeIds_.push_back(e.id);
    --g.numEdges_;
  }
}

MRP_STANDIN_GRAPH_TPL
std::pair<typename MRP_STANDIN_GRAPH_T::out_edge_iterator,
          typename MRP_STANDIN_GRAPH_T::out_edge_iterator>
out_edges(std::size_t v, const MRP_STANDIN_GRAPH_T& g) {
  typedef typename MRP_STANDIN_GRAPH_T::out_edge_iterator It;
  const auto& lst = g.out_[v];
  return std::make_pair(It(v, lst.data()), It(v, lst.data() + lst.size()));
}

MRP_STANDIN_GRAPH_TPL
std::size_t target(const standin_detail::EdgeDesc& e, const MRP_STANDIN_GRAPH_T&) { return e.dst; }
MRP_STANDIN_GRAPH_TPL
std::size_t source(const standin_detail::EdgeDesc& e, const MRP_STANDIN_GRAPH_T&) { return e.src; }

// get(&Edge::member, g): a property map over the bundled edge properties
template <class Graph, class T>
struct bundle_member_map {
  Graph* g;
  T Graph::edge_bundled::*member;
  T& operator[](const standin_detail::EdgeDesc& e) const { return ((*g)[e]).*member; }
};
template <class O, class V, class D, class VP, class EP, class T>
bundle_member_map<MRP_STANDIN_GRAPH_T, T> get(T EP::*member, MRP_STANDIN_GRAPH_T& g) {
  return bundle_member_map<MRP_STANDIN_GRAPH_T, T>{&g, member};
}

// named parameters: capacity_map(a).residual_capacity_map(b).weight_map(c).reverse_edge_map(d)
struct standin_none {};
template <class Cap = standin_none, class Res = standin_none, class Wgt = standin_none,
          class Rev = standin_none>
struct bgl_named_params {
  Cap cap;
  Res res;
  Wgt wgt;
  Rev rev;
  template <class X>
  bgl_named_params<X, Res, Wgt, Rev> capacity_map(const X& x) const {
    return bgl_named_params<X, Res, Wgt, Rev>{x, res, wgt, rev};
  }
  template <class X>
  bgl_named_params<Cap, X, Wgt, Rev> residual_capacity_map(const X& x) const {
    return bgl_named_params<Cap, X, Wgt, Rev>{cap, x, wgt, rev};
  }
  template <class X>
  bgl_named_params<Cap, Res, X, Rev> weight_map(const X& x) const {
    return bgl_named_params<Cap, Res, X, Rev>{cap, res, x, rev};
  }
  template <class X>
  bgl_named_params<Cap, Res, Wgt, X> reverse_edge_map(const X& x) const {
    return bgl_named_params<Cap, Res, Wgt, X>{cap, res, wgt, x};
  }
};
template <class X>
bgl_named_params<X> capacity_map(const X& x) {
  return bgl_named_params<X>{x, standin_none(), standin_none(), standin_none()};
}
template <class X>
bgl_named_params<standin_none, standin_none, X> weight_map(const X& x) {
  return bgl_named_params<standin_none, standin_none, X>{standin_none(), standin_none(), x,
                                                         standin_none()};
}

// Minimum-cost maximum flow by successive shortest augmenting paths (Dijkstra on the
// reduced costs of the residual network; all forward weights are non-negative, the
// reverse edges start without residual capacity).  Residual capacities are reset from the
// capacities first, as the Boost function documents.
template <class O, class V, class D, class VP, class EP, class P>
void successive_shortest_path_nonnegative_weights(MRP_STANDIN_GRAPH_T& g, std::size_t s,
                                                  std::size_t t, const P& p) {
  typedef standin_detail::EdgeDesc E;
  typedef long W;
  const std::size_t n = num_vertices(g);
  for (std::size_t v = 0; v < n; ++v) {
    auto es = out_edges(v, g);
    for (auto it = es.first; it != es.second; ++it) p.res[*it] = p.cap[*it];
  }
  const W inf = std::numeric_limits<W>::max();
  std::vector<W> pot(n, 0), dist(n);
  std::vector<E> pred(n);
  std::vector<char> havePred(n);
  typedef std::pair<W, std::size_t> QE;
  while (true) {
    std::fill(dist.begin(), dist.end(), inf);
    std::fill(havePred.begin(), havePred.end(), 0);
    std::priority_queue<QE, std::vector<QE>, std::greater<QE> > q;
    dist[s] = 0;
    q.push(QE(0, s));
    while (!q.empty()) {
      const QE top = q.top();
      q.pop();
      const std::size_t u = top.second;
      if (top.first != dist[u]) continue;
      auto es = out_edges(u, g);
      for (auto it = es.first; it != es.second; ++it) {
        const E e = *it;
        if (p.res[e] <= 0) continue;
        const W nd = dist[u] + (W)p.wgt[e] + pot[u] - pot[e.dst];
        if (nd < dist[e.dst]) {
          dist[e.dst] = nd;
          pred[e.dst] = e;
          havePred[e.dst] = 1;
          q.push(QE(nd, e.dst));
        }
      }
    }
    if (dist[t] == inf) break;
    for (std::size_t v = 0; v < n; ++v)
      if (dist[v] != inf) pot[v] += dist[v];
    W bottleneck = inf;
    for (std::size_t v = t; v != s; v = pred[v].src) bottleneck = std::min<W>(bottleneck, p.res[pred[v]]);
    for (std::size_t v = t; v != s; v = pred[v].src) {
      const E e = pred[v];
      p.res[e] -= bottleneck;
      p.res[p.rev[e]] += bottleneck;
    }
  }
}

// exterior_vertex_property<Graph, T>: a |V| x |V| matrix and its property-map view
template <class T>
class standin_matrix {
 public:
  explicit standin_matrix(std::size_t n) : rows_(n, std::vector<T>(n)) {}
  std::vector<T>& operator[](std::size_t i) { return rows_[i]; }
  const std::vector<T>& operator[](std::size_t i) const { return rows_[i]; }
  std::size_t size() const { return rows_.size(); }

 private:
  std::vector<std::vector<T> > rows_;
};
template <class Graph, class T>
struct standin_matrix_map {
  standin_matrix_map(standin_matrix<T>& m, const Graph&) : m(&m) {}
  standin_matrix<T>* m;
};
template <class Graph, class T>
struct exterior_vertex_property {
  typedef standin_matrix<T> matrix_type;
  typedef standin_matrix_map<Graph, T> matrix_map_type;
};

// d[i][i] = 0, d[u][v] = min over the edges u-v of their weight (both directions for an
// undirected graph), everything else numeric_limits<T>::max(); then the triple loop, in
// which an "infinite" entry is never extended.  Returns false on a negative cycle.
template <class O, class V, class D, class VP, class EP, class T, class P>
bool floyd_warshall_all_pairs_shortest_paths(const MRP_STANDIN_GRAPH_T& g,
                                             standin_matrix_map<MRP_STANDIN_GRAPH_T, T>& d,
                                             const P& p) {
  standin_matrix<T>& m = *d.m;
  const std::size_t n = num_vertices(g);
  const T inf = std::numeric_limits<T>::max();
  for (std::size_t i = 0; i < n; ++i)
    for (std::size_t j = 0; j < n; ++j) m[i][j] = inf;
  for (std::size_t i = 0; i < n; ++i) m[i][i] = T();
  for (std::size_t u = 0; u < n; ++u) {
    auto es = out_edges(u, g);
    for (auto it = es.first; it != es.second; ++it) {
      const standin_detail::EdgeDesc e = *it;
      const T w = p.wgt[e];
      if (w < m[u][e.dst]) m[u][e.dst] = w;
    }
  }
  for (std::size_t k = 0; k < n; ++k)
    for (std::size_t i = 0; i < n; ++i) {
      const T dik = m[i][k];
      if (dik == inf) continue;
      std::vector<T>& ri = m[i];
      const std::vector<T>& rk = m[k];
      for (std::size_t j = 0; j < n; ++j) {
        if (rk[j] == inf) continue;
        const T c = dik + rk[j];
        if (c < ri[j]) ri[j] = c;
      }
    }
  for (std::size_t i = 0; i < n; ++i)
    if (m[i][i] < T()) return false;
  return true;
}
// write_graphviz(out, g, vertexWriter, edgeWriter): the same dot structure Boost emits
template <class O, class V, class D, class VP, class EP, class VW, class EW>
void write_graphviz(std::ostream& out, const MRP_STANDIN_GRAPH_T& g, VW vw, EW ew) {
  const bool undirected = std::is_same<D, undirectedS>::value;
  out << (undirected ? "graph" : "digraph") << " G {\n";
  const std::size_t n = num_vertices(g);
  for (std::size_t v = 0; v < n; ++v) {
    out << v;
    vw(out, v);
    out << ";\n";
  }
  for (std::size_t u = 0; u < n; ++u) {
    auto es = out_edges(u, g);
    for (auto it = es.first; it != es.second; ++it) {
      const standin_detail::EdgeDesc e = *it;
      if (undirected && e.dst < u) continue;
      out << u << (undirected ? "--" : "->") << e.dst << " ";
      ew(out, e);
      out << ";\n";
    }
  }
  out << "}\n";
}

#undef MRP_STANDIN_GRAPH_T
#undef MRP_STANDIN_GRAPH_TPL

}  // namespace boost
