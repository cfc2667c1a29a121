// TEST INFRASTRUCTURE ONLY (oracle build).  binding.cpp:11-33 names these types inside two helper
// templates that are never instantiated; they only have to parse.
#ifndef ORACLE_SHIM_BOOST_PYTHON_HPP_
#define ORACLE_SHIM_BOOST_PYTHON_HPP_
#include <iterator>
namespace boost { namespace python {
struct object {};
struct list { template <class T> void append(const T &) {} };
template <class T> struct stl_input_iterator {
  stl_input_iterator() {}
  explicit stl_input_iterator(const object &) {}
  typedef T value_type; typedef T *pointer; typedef T &reference; typedef long difference_type;
  typedef std::input_iterator_tag iterator_category;
  T operator*() const { return T(); }
  stl_input_iterator &operator++() { return *this; }
  bool operator==(const stl_input_iterator &) const { return true; }
  bool operator!=(const stl_input_iterator &) const { return false; }
};
} }
#endif
