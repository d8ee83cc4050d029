// Stand-in for <boost/heap/d_ary_heap.hpp> (TEST INFRASTRUCTURE).
//
// Boost is not installed in this image.  This header offers just the part of
// boost::heap::d_ary_heap<T, arity<2>, mutable_<true>[, compare<C>]> that the
// UNMODIFIED reference headers use (a_star.hpp:205-208, a_star_epsilon.hpp:
// 296-298,378-381, cbs.hpp:110-112, ecbs.hpp:299-301,359-362,
// cbs_ta.hpp:119-121): push / top / pop / empty / size / erase / increase /
// clear / ordered_begin / ordered_end and handles.  It lets the reference's own
// template code run as a second oracle (oracle/_ref/).  The order among equal
// keys is this implementation's, not Boost's — like everything tie-related it
// is unpinned by the reference's tests.
#pragma once

#include <algorithm>
#include <cassert>
#include <cstddef>
#include <functional>
#include <memory>
#include <vector>

namespace boost {
namespace heap {

template <int N>
struct arity {};
template <bool B>
struct mutable_ {};
template <class C>
struct compare {
  typedef C type;
};

namespace standin_detail {
template <class T, class... Options>
struct pick_compare {
  typedef std::less<T> type;
};
template <class T, class C, class... Rest>
struct pick_compare<T, compare<C>, Rest...> {
  typedef C type;
};
template <class T, class First, class... Rest>
struct pick_compare<T, First, Rest...> {
  typedef typename pick_compare<T, Rest...>::type type;
};
}  // namespace standin_detail

// max-heap with respect to Compare (Compare(a, b): a has lower priority than b)
template <class T, class... Options>
class d_ary_heap {
  struct node;

 public:
  typedef typename standin_detail::pick_compare<T, Options...>::type value_compare;

  class handle_type {
   public:
    handle_type() : n_(nullptr) {}
    T& operator*() const { return n_->value(); }
    bool operator==(const handle_type& o) const { return n_ == o.n_; }
    bool operator!=(const handle_type& o) const { return n_ != o.n_; }

   private:
    friend class d_ary_heap;
    explicit handle_type(node* n) : n_(n) {}
    node* n_;
  };

  class ordered_iterator {
   public:
    const T& operator*() const { return (*order_)[i_]->value(); }
    const T* operator->() const { return &(*order_)[i_]->value(); }
    ordered_iterator& operator++() {
      ++i_;
      return *this;
    }
    bool operator!=(const ordered_iterator& o) const { return remaining() != o.remaining(); }
    bool operator==(const ordered_iterator& o) const { return remaining() == o.remaining(); }

   private:
    friend class d_ary_heap;
    size_t remaining() const { return order_ ? order_->size() - i_ : 0; }
    std::shared_ptr<std::vector<node*> > order_;
    size_t i_ = 0;
  };

  d_ary_heap() {}
  d_ary_heap(const d_ary_heap&) = delete;
  d_ary_heap& operator=(const d_ary_heap&) = delete;
  ~d_ary_heap() { clear(); }

  bool empty() const { return heap_.empty(); }
  size_t size() const { return heap_.size(); }
  const T& top() const { return heap_[0]->value(); }

  handle_type push(const T& v) {
    node* n = new node(v);
    n->pos = heap_.size();
    heap_.push_back(n);
    siftUp(n->pos);
    return handle_type(n);
  }
  void pop() { removeAt(0); }
  void erase(handle_type h) { removeAt(h.n_->pos); }
  void increase(handle_type h) { siftUp(h.n_->pos); }  // priority went up
  void update(handle_type h) {
    siftUp(h.n_->pos);
    siftDown(h.n_->pos);
  }
  void clear() {
    for (node* n : heap_) delete n;
    heap_.clear();
  }

  ordered_iterator ordered_begin() const {
    ordered_iterator it;
    it.order_ = std::make_shared<std::vector<node*> >(heap_);
    const value_compare& c = cmp_;
    std::stable_sort(it.order_->begin(), it.order_->end(),
                     [&c](node* a, node* b) { return c(b->value(), a->value()); });
    return it;
  }
  ordered_iterator ordered_end() const { return ordered_iterator(); }

 private:
  struct node {
    explicit node(const T& v) : storage(new T(v)), pos(0) {}
    T& value() { return *storage; }
    std::unique_ptr<T> storage;
    size_t pos;
  };
  bool lower(size_t a, size_t b) const { return cmp_(heap_[a]->value(), heap_[b]->value()); }
  void swapAt(size_t a, size_t b) {
    std::swap(heap_[a], heap_[b]);
    heap_[a]->pos = a;
    heap_[b]->pos = b;
  }
  void siftUp(size_t i) {
    while (i > 0) {
      size_t p = (i - 1) / 2;
      if (!lower(p, i)) break;
      swapAt(p, i);
      i = p;
    }
  }
  void siftDown(size_t i) {
    for (;;) {
      size_t l = 2 * i + 1, r = l + 1, c = l;
      if (l >= heap_.size()) break;
      if (r < heap_.size() && lower(l, r)) c = r;
      if (lower(c, i)) break;
      swapAt(c, i);
      i = c;
    }
  }
  void removeAt(size_t i) {
    node* n = heap_[i];
    size_t last = heap_.size() - 1;
    if (i != last) swapAt(i, last);
    heap_.pop_back();
    delete n;
    if (i < heap_.size()) {
      siftUp(i);
      siftDown(i);
    }
  }
  std::vector<node*> heap_;
  value_compare cmp_;
};

}  // namespace heap
}  // namespace boost
