// api.cu -- library-wide state of liblob_b200.
#include "common.cuh"

namespace lob {
thread_local std::string g_last_error;
std::atomic<int64_t> g_launch_count{0};
}  // namespace lob

extern "C" int lob_version(void) { return 100; }
extern "C" const char* lob_last_error(void) { return lob::g_last_error.c_str(); }
extern "C" int64_t lob_launch_count(void) { return lob::g_launch_count.load(); }
