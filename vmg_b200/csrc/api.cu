// api.cu -- library-level entry points of libvmgb.so (version, arch gate, error reporting).
#include "common.cuh"

namespace vmgb {
int g_last_cuda_error = 0;
long long g_launch_count = 0;
int g_pack_pending = 0;
}  // namespace vmgb

extern "C" int vmgb_version(void) { return 100; }
extern "C" const char* vmgb_arch(void) { return "sm_100a"; }
extern "C" int vmgb_last_cuda_error(void) { return vmgb::g_last_cuda_error; }
extern "C" const char* vmgb_last_cuda_error_string(void) { return cudaGetErrorString((cudaError_t)vmgb::g_last_cuda_error); }
extern "C" long long vmgb_launch_count(void) { return vmgb::g_launch_count; }

extern "C" int vmgb_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { vmgb::g_last_cuda_error = (int)e; return VMGB_ECUDA; }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) { vmgb::g_last_cuda_error = (int)e; return VMGB_ECUDA; }
  return major == 10 ? VMGB_OK : VMGB_EARCH;
}
