// TEST INFRASTRUCTURE ONLY (oracle build).  binding.cpp:10 includes this header and uses nothing from it.
