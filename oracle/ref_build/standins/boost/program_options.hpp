// Stand-in for <boost/program_options.hpp> (TEST INFRASTRUCTURE): the handful of
// calls example/cbs.cpp:572-596 / ecbs.cpp:524-552 make.
#pragma once
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace program_options {

class error : public std::logic_error {
 public:
  explicit error(const std::string& w) : std::logic_error(w) {}
};

class value_semantic {
 public:
  virtual ~value_semantic() {}
  virtual void parse(const std::string& s) = 0;
  virtual bool applyDefault() = 0;
  bool isRequired = false;
};

template <class T>
class typed_value : public value_semantic {
 public:
  explicit typed_value(T* store) : store_(store) {}
  typed_value* required() {
    isRequired = true;
    return this;
  }
  typed_value* default_value(const T& v) {
    def_ = v;
    hasDefault_ = true;
    return this;
  }
  void parse(const std::string& s) override {
    std::istringstream in(s);
    T v;
    in >> v;
    if (in.fail()) throw error("the argument ('" + s + "') is invalid");
    *store_ = v;
  }
  bool applyDefault() override {
    if (hasDefault_) *store_ = def_;
    return hasDefault_;
  }

 private:
  T* store_;
  T def_ = T();
  bool hasDefault_ = false;
};
template <>
inline void typed_value<std::string>::parse(const std::string& s) {
  *store_ = s;
}

template <class T>
typed_value<T>* value(T* v) {
  return new typed_value<T>(v);
}

struct option_entry {
  std::string longName, shortName, description;
  std::shared_ptr<value_semantic> semantic;
};

class options_description {
 public:
  explicit options_description(const std::string& caption) : caption_(caption) {}
  class easy_init {
   public:
    explicit easy_init(options_description* o) : o_(o) {}
    easy_init& operator()(const char* name, const char* description) {
      o_->add(name, nullptr, description);
      return *this;
    }
    easy_init& operator()(const char* name, value_semantic* s, const char* description) {
      o_->add(name, s, description);
      return *this;
    }

   private:
    options_description* o_;
  };
  easy_init add_options() { return easy_init(this); }
  const std::vector<option_entry>& options() const { return options_; }
  friend std::ostream& operator<<(std::ostream& os, const options_description& d) {
    os << d.caption_ << ":\n";
    for (const auto& o : d.options_) {
      os << "  ";
      if (!o.shortName.empty()) os << "-" << o.shortName << " [ --" << o.longName << " ]";
      else os << "--" << o.longName;
      if (o.semantic) os << " arg";
      os << "   " << o.description << "\n";
    }
    return os;
  }

 private:
  void add(const char* name, value_semantic* s, const char* description) {
    option_entry e;
    std::string n = name;
    size_t comma = n.find(',');
    e.longName = n.substr(0, comma);
    if (comma != std::string::npos) e.shortName = n.substr(comma + 1);
    e.description = description;
    e.semantic.reset(s);
    options_.push_back(e);
  }
  std::string caption_;
  std::vector<option_entry> options_;
};

struct parsed_options {
  const options_description* desc;
  std::map<std::string, std::string> values;
};

class variables_map {
 public:
  size_t count(const std::string& name) const { return seen.count(name); }
  std::map<std::string, std::string> seen;
  const options_description* desc = nullptr;
};

inline parsed_options parse_command_line(int argc, char** argv, const options_description& d) {
  parsed_options p;
  p.desc = &d;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i], val;
    bool hasVal = false;
    const option_entry* opt = nullptr;
    if (a.rfind("--", 0) == 0) {
      size_t eq = a.find('=');
      std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
      if (eq != std::string::npos) {
        val = a.substr(eq + 1);
        hasVal = true;
      }
      for (const auto& o : d.options())
        if (o.longName == name) opt = &o;
    } else if (a.size() >= 2 && a[0] == '-') {
      for (const auto& o : d.options())
        if (!o.shortName.empty() && o.shortName == a.substr(1, 1)) opt = &o;
      if (a.size() > 2) {
        val = a.substr(2);
        hasVal = true;
      }
    }
    if (!opt) throw error("unrecognised option '" + a + "'");
    if (opt->semantic && !hasVal) {
      if (i + 1 >= argc)
        throw error("the required argument for option '--" + opt->longName + "' is missing");
      val = argv[++i];
    }
    p.values[opt->longName] = val;
  }
  return p;
}

inline void store(const parsed_options& p, variables_map& vm) {
  vm.desc = p.desc;
  for (const auto& kv : p.values) vm.seen[kv.first] = kv.second;
}

inline void notify(variables_map& vm) {
  if (!vm.desc) return;
  for (const auto& o : vm.desc->options()) {
    auto it = vm.seen.find(o.longName);
    if (it != vm.seen.end()) {
      if (o.semantic) o.semantic->parse(it->second);
    } else if (o.semantic) {
      if (!o.semantic->applyDefault() && o.semantic->isRequired && !vm.seen.count("help"))
        throw error("the option '--" + o.longName + "' is required but missing");
    }
  }
}

}  // namespace program_options
}  // namespace boost
