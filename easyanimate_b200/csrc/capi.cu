// Library-level entry points of the C ABI: error string, version, launch counter.
#include <atomic>

#include "host.h"
#include "../../include/ea_b200.h"

namespace ea {
const char* last_error_cstr();
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace ea

extern "C" const char* ea_last_error(void) { return ea::last_error_cstr(); }
extern "C" int ea_abi_version(void) { return 2; }
extern "C" uint64_t ea_launch_count(void) { return ea::g_launches.load(std::memory_order_relaxed); }
