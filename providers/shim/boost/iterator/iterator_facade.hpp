// TEST INFRASTRUCTURE ONLY (oracle build).  KenLM util/tokenize_piece.hh:8 (pulled in by the
// reference's scorer.cpp:10) derives TokenIter from boost::iterator_facade; the reference never
// instantiates it, so the names only have to parse.
#ifndef ORACLE_SHIM_BOOST_ITERATOR_FACADE_HPP_
#define ORACLE_SHIM_BOOST_ITERATOR_FACADE_HPP_
namespace boost {
struct forward_traversal_tag {};
class iterator_core_access {};
template <class Derived, class Value, class Traversal> class iterator_facade {};
}
#endif
