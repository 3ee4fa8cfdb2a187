// Library-wide pieces of the C ABI: version, per-thread error string, launch accounting.
#include "common.cuh"

namespace bevf {
std::string &last_error() {
    static thread_local std::string e;
    return e;
}
std::atomic<int64_t> g_launches{0};
}  // namespace bevf

extern "C" int bevf_version(void) { return BEVF_ABI_VERSION; }
extern "C" const char *bevf_last_error(void) { return bevf::last_error().c_str(); }
extern "C" int64_t bevf_launch_count(void) { return bevf::g_launches.load(); }
