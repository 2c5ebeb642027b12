// Library-wide state of the C-ABI: error text, ABI version, launch counter.
#include "common.cuh"
namespace egnn {
std::atomic<uint64_t> g_launches{0};
char* err_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
}  // namespace egnn
extern "C" int egnn_abi_version(void) { return EGNN_ABI_VERSION; }
extern "C" const char* egnn_last_error(void) { return egnn::err_buf(); }
extern "C" uint64_t egnn_launch_count(void) { return egnn::g_launches.load(); }
