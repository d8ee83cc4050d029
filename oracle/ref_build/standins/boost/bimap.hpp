// Stand-in for <boost/bimap.hpp> (TEST INFRASTRUCTURE): the one-to-one map between agents /
// tasks and their graph vertices of include/libMultiRobotPlanning/assignment.hpp:36-116,171-176
// — insert(value_type(l, r)), left.find / left.end / ->second, right.at, and iteration over
// the relations (ordered by the left key, as a bimap of two set_of views iterates) with
// .left / .right members.
#pragma once
#include <map>
#include <stdexcept>

namespace boost {

template <class L, class R>
class bimap {
 public:
  struct value_type {
    value_type(const L& l, const R& r) : left(l), right(r) {}
    L left;
    R right;
  };
  class left_view {
   public:
    typedef typename std::map<L, R>::const_iterator const_iterator;
    const_iterator find(const L& l) const { return m_.find(l); }
    const_iterator end() const { return m_.end(); }
    const R& at(const L& l) const { return m_.at(l); }
    std::map<L, R> m_;
  };
  class right_view {
   public:
    typedef typename std::map<R, L>::const_iterator const_iterator;
    const_iterator find(const R& r) const { return m_.find(r); }
    const_iterator end() const { return m_.end(); }
    const L& at(const R& r) const { return m_.at(r); }
    std::map<R, L> m_;
  };
  class const_iterator {
   public:
    explicit const_iterator(typename std::map<L, R>::const_iterator it) : it_(it) {}
    value_type operator*() const { return value_type(it_->first, it_->second); }
    const_iterator& operator++() {
      ++it_;
      return *this;
    }
    bool operator!=(const const_iterator& o) const { return it_ != o.it_; }

   private:
    typename std::map<L, R>::const_iterator it_;
  };

  // a relation is only accepted if neither side is taken (both views are sets)
  bool insert(const value_type& v) {
    if (left.m_.count(v.left) || right.m_.count(v.right)) return false;
    left.m_[v.left] = v.right;
    right.m_[v.right] = v.left;
    return true;
  }
  const_iterator begin() const { return const_iterator(left.m_.begin()); }
  const_iterator end() const { return const_iterator(left.m_.end()); }
  size_t size() const { return left.m_.size(); }

  left_view left;
  right_view right;
};

}  // namespace boost
