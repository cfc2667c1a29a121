// TEST INFRASTRUCTURE ONLY (oracle build).  Stand-in for OpenFST's fst/log.h: the reference only
// uses `LOG(FATAL) << ...` (decoder_utils.h:17-23), which prints and aborts the process.
#ifndef ORACLE_SHIM_FST_LOG_H_
#define ORACLE_SHIM_FST_LOG_H_
#include <cstdlib>
#include <iostream>
namespace oracle_shim {
struct FatalLogger {
  ~FatalLogger() {
    std::cerr << std::endl;
    std::abort();
  }
  template <class T>
  FatalLogger &operator<<(const T &v) {
    std::cerr << v;
    return *this;
  }
};
}  // namespace oracle_shim
#define LOG(severity) ::oracle_shim::FatalLogger()
#endif
