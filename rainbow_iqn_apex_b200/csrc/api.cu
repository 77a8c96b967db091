// Library identification entry points of the C-ABI (include/riqn_b200.h).
#include "common.cuh"
#include "../../include/riqn_b200.h"

RIQN_API int riqn_version(void) { return RIQN_B200_ABI_VERSION; }

RIQN_API int riqn_device_ok(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return -(int)e;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return -(int)e;
  return major == 10 ? 1 : 0;
}

#include <atomic>
namespace riqn {
static std::atomic<long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace riqn

RIQN_API long long riqn_launch_count(void) { return riqn::g_launches.load(std::memory_order_relaxed); }

RIQN_API int riqn_zero_f32(float* p, long n, void* stream) {
  return (int)cudaMemsetAsync(p, 0, sizeof(float) * (size_t)n, (cudaStream_t)stream);
}
