// Library-level entry points of the C ABI: error string, version, launch counter.
#include <atomic>
#include <string>

#include "host.h"
#include "../../include/ea_b200.h"

namespace ea {
const char* last_error_cstr();
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace ea

extern "C" const char* ea_last_error(void) { return ea::last_error_cstr(); }
extern "C" int ea_abi_version(void) { return 3; }
extern "C" uint64_t ea_launch_count(void) { return ea::g_launches.load(std::memory_order_relaxed); }

extern "C" int ea_enable_peer_access(int32_t peer_device) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return ea::fail(EA_ERR_CUDA, "ea_enable_peer_access: no current device");
  if (peer_device == dev) return EA_OK;
  int can = 0;
  cudaError_t e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
  if (e != cudaSuccess || !can) {
    (void)cudaGetLastError();
    return ea::fail(EA_ERR_CUDA, "ea_enable_peer_access: device " + std::to_string(dev) + " cannot access device " + std::to_string(peer_device));
  }
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
    return ea::fail(EA_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
  (void)cudaGetLastError();
  return EA_OK;
}
