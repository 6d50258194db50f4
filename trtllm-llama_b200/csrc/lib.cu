#include "common.cuh"
#include "kernels.h"
extern "C" const char* tb_version(void) { return "trtllm_llama_b200 0.1.0 (sm_100a)"; }
extern "C" int tb_check_device(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  return major == 10 ? 0 : -2;
}
