// C-ABI glue: version, thread-local error string and launch counter (include/dva_b200.h).
#include "dva_common.cuh"

namespace dva {
char* tls_error_buf() {
  static thread_local char buf[256] = {0};
  return buf;
}
// process-wide: autograd runs backward kernels from its own worker thread
std::atomic<int64_t>& launch_counter() {
  static std::atomic<int64_t> n{0};
  return n;
}
}  // namespace dva

extern "C" int dva_abi_version(void) { return DVA_ABI_VERSION; }
extern "C" const char* dva_last_error(void) { return dva::tls_error_buf(); }
extern "C" int64_t dva_launch_count(void) { return dva::launch_counter().load(); }
