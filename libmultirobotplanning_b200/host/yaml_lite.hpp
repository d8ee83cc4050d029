// yaml_lite.hpp — reader for the YAML subset of the reference's input files
// (example/cbs.cpp:598-618, example/cbs_ta.cpp:547-568): block mappings and
// sequences (2- or 4-space, "-   key:" items, sequences at the indentation of
// their key), flow sequences such as [3, 2] or [[0,1],[1,1]] or [], integer and
// string scalars, comments.  yaml-cpp is not available in this image.
#pragma once

#include <cctype>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mrp_host {
namespace yaml {

struct Node {
  enum Kind { Null, Scalar, Seq, Map } kind = Null;
  std::string scalar;
  std::vector<Node> seq;
  std::vector<std::pair<std::string, Node> > map;

  bool has(const std::string& key) const {
    for (const auto& kv : map)
      if (kv.first == key) return true;
    return false;
  }
  const Node& operator[](const std::string& key) const {
    for (const auto& kv : map)
      if (kv.first == key) return kv.second;
    throw std::runtime_error("yaml: missing key '" + key + "'");
  }
  const Node& operator[](size_t i) const {
    if (kind != Seq || i >= seq.size()) throw std::runtime_error("yaml: bad sequence index");
    return seq[i];
  }
  size_t size() const { return kind == Seq ? seq.size() : map.size(); }
  int asInt() const {
    if (kind != Scalar) throw std::runtime_error("yaml: expected an integer");
    size_t pos = 0;
    int v = std::stoi(scalar, &pos);
    if (pos != scalar.size()) throw std::runtime_error("yaml: bad integer '" + scalar + "'");
    return v;
  }
};

namespace detail {

struct Line {
  int indent;
  std::string text;
};

inline std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) ++a;
  while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

inline std::string stripComment(const std::string& s) {
  bool inS = false, inD = false;
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[i];
    if (c == '\'' && !inD) inS = !inS;
    if (c == '"' && !inS) inD = !inD;
    if (c == '#' && !inS && !inD && (i == 0 || std::isspace((unsigned char)s[i - 1])))
      return s.substr(0, i);
  }
  return s;
}

inline Node parseFlow(const std::string& s, size_t& i) {
  while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
  Node n;
  if (i < s.size() && s[i] == '[') {
    n.kind = Node::Seq;
    ++i;
    while (true) {
      while (i < s.size() && (std::isspace((unsigned char)s[i]) || s[i] == ',')) ++i;
      if (i >= s.size()) throw std::runtime_error("yaml: unterminated '['");
      if (s[i] == ']') {
        ++i;
        break;
      }
      n.seq.push_back(parseFlow(s, i));
    }
    return n;
  }
  size_t j = i;
  if (i < s.size() && (s[i] == '"' || s[i] == '\'')) {
    char q = s[i];
    j = s.find(q, i + 1);
    if (j == std::string::npos) throw std::runtime_error("yaml: unterminated string");
    n.kind = Node::Scalar;
    n.scalar = s.substr(i + 1, j - i - 1);
    i = j + 1;
    return n;
  }
  while (j < s.size() && s[j] != ',' && s[j] != ']') ++j;
  n.kind = Node::Scalar;
  n.scalar = trim(s.substr(i, j - i));
  i = j;
  if (n.scalar.empty() || n.scalar == "~" || n.scalar == "null") n.kind = Node::Null;
  return n;
}

inline Node parseInline(const std::string& text) {
  size_t i = 0;
  Node n = parseFlow(text, i);
  return n;
}

inline bool splitKey(const std::string& t, std::string& key, std::string& rest) {
  if (t.empty() || t[0] == '[' || t[0] == '-') return false;
  size_t i = 0;
  if (t[0] == '"' || t[0] == '\'') {
    i = t.find(t[0], 1);
    if (i == std::string::npos) return false;
    ++i;
  } else {
    while (i < t.size() && t[i] != ':' && !std::isspace((unsigned char)t[i])) ++i;
  }
  size_t j = i;
  while (j < t.size() && std::isspace((unsigned char)t[j])) ++j;
  if (j >= t.size() || t[j] != ':') return false;
  if (j + 1 < t.size() && !std::isspace((unsigned char)t[j + 1])) return false;
  key = trim(t.substr(0, i));
  if (!key.empty() && (key[0] == '"' || key[0] == '\'')) key = key.substr(1, key.size() - 2);
  rest = trim(t.substr(j + 1));
  return true;
}

inline Node parseNode(const std::vector<Line>& L, size_t& i, int indent) {
  Node n;
  if (i >= L.size()) return n;
  std::string key, rest;
  if (L[i].text == "-") {
    n.kind = Node::Seq;
    while (i < L.size() && L[i].indent == indent && L[i].text == "-") {
      ++i;
      if (i < L.size() && L[i].indent > indent)
        n.seq.push_back(parseNode(L, i, L[i].indent));
      else
        n.seq.push_back(Node());
    }
    return n;
  }
  if (splitKey(L[i].text, key, rest)) {
    n.kind = Node::Map;
    while (i < L.size() && L[i].indent == indent && splitKey(L[i].text, key, rest)) {
      ++i;
      Node v;
      if (!rest.empty())
        v = parseInline(rest);
      else if (i < L.size() && L[i].indent > indent)
        v = parseNode(L, i, L[i].indent);
      else if (i < L.size() && L[i].indent == indent && L[i].text == "-")
        v = parseNode(L, i, indent);
      n.map.emplace_back(key, std::move(v));
    }
    return n;
  }
  n = parseInline(L[i].text);
  ++i;
  return n;
}

}  // namespace detail

inline Node parse(std::istream& in) {
  std::vector<detail::Line> lines;
  std::string raw;
  while (std::getline(in, raw)) {
    if (!raw.empty() && raw.back() == '\r') raw.pop_back();
    std::string s = detail::stripComment(raw);
    if (detail::trim(s).empty()) continue;
    if (detail::trim(s) == "---") continue;
    int indent = 0;
    while (indent < (int)s.size() && s[indent] == ' ') ++indent;
    if (indent < (int)s.size() && s[indent] == '\t')
      throw std::runtime_error("yaml: tabs are not allowed for indentation");
    std::string t = detail::trim(s);
    // "- item" lines become a "-" marker plus the item at its own column;
    // nested dashes ("- - 1") repeat
    while (t == "-" || (t.size() > 1 && t[0] == '-' && t[1] == ' ')) {
      lines.push_back({indent, "-"});
      if (t == "-") {
        t.clear();
        break;
      }
      size_t k = 1;
      while (k < t.size() && t[k] == ' ') ++k;
      indent += (int)k;
      t = t.substr(k);
    }
    if (!t.empty()) lines.push_back({indent, t});
  }
  size_t i = 0;
  if (lines.empty()) return Node();
  Node root = detail::parseNode(lines, i, lines[0].indent);
  if (i != lines.size())
    throw std::runtime_error("yaml: could not parse line '" + lines[i].text + "'");
  return root;
}

inline Node loadFile(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("bad file: " + path);  // YAML::BadFile
  return parse(f);
}

}  // namespace yaml
}  // namespace mrp_host
