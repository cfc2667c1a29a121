// TEST INFRASTRUCTURE ONLY (oracle build).  See ../python.hpp.
#include <iterator>
#include "boost/python.hpp"
