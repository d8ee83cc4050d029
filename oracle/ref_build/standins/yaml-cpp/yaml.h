// Stand-in for <yaml-cpp/yaml.h> (TEST INFRASTRUCTURE): LoadFile, operator[],
// as<T>() and sequence iteration — what example/cbs.cpp:598-618 uses.  Parsing
// is delegated to the small reader of this repository.
#pragma once
#include <string>
#include <vector>

#include "../../../../libmultirobotplanning_b200/host/yaml_lite.hpp"

namespace YAML {

class Node {
 public:
  Node() {}
  explicit Node(const mrp_host::yaml::Node& n) : n_(n) {}
  Node operator[](const char* key) const { return Node(n_[std::string(key)]); }
  Node operator[](const std::string& key) const { return Node(n_[key]); }
  Node operator[](int i) const { return Node(n_[(size_t)i]); }
  Node operator[](size_t i) const { return Node(n_[i]); }
  template <class T>
  T as() const;
  size_t size() const { return n_.size(); }

  class const_iterator {
   public:
    const_iterator(const mrp_host::yaml::Node* n, size_t i) : n_(n), i_(i) {}
    Node operator*() const { return Node(n_->seq[i_]); }
    const_iterator& operator++() {
      ++i_;
      return *this;
    }
    bool operator!=(const const_iterator& o) const { return i_ != o.i_; }

   private:
    const mrp_host::yaml::Node* n_;
    size_t i_;
  };
  const_iterator begin() const { return const_iterator(&n_, 0); }
  const_iterator end() const { return const_iterator(&n_, n_.seq.size()); }

 private:
  mrp_host::yaml::Node n_;
};

template <>
inline int Node::as<int>() const {
  return n_.asInt();
}
template <>
inline std::string Node::as<std::string>() const {
  return n_.scalar;
}

inline Node LoadFile(const std::string& path) { return Node(mrp_host::yaml::loadFile(path)); }

}  // namespace YAML
