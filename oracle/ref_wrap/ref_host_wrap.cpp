// oracle/_ref host-side wrapper: exposes the REFERENCE's own weight-only quantiser and layout
// pre-processor (compiled from /root/reference where it lies, never copied) through a C ABI so
// the numpy oracle can be pinned against it.  TEST INFRASTRUCTURE ONLY.
//
// Reference entry points wrapped:
//   tensorrt_llm/kernels/cutlass_kernels/cutlass_preprocessors.cpp:615-721  symmetric_quantize
//   tensorrt_llm/kernels/cutlass_kernels/cutlass_preprocessors.cpp:537-578  preprocess_weights_for_mixed_gemm
//
// The pre-processor asks the CUDA runtime for the SM version (cutlass_preprocessors.cpp:130-150,
// common/cudaUtils.h:230-239) and rejects anything outside [70, 90].  This wrapper answers that
// query itself with SM 8.0 (the authors' A10 is SM 8.6 -> the same "Sm80" layout), so the library
// needs neither a GPU nor libcudart; the stubs below are hidden by the version script.
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "tensorrt_llm/kernels/cutlass_kernels/cutlass_preprocessors.h"

extern "C" {
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr attr, int) {
  *v = (attr == cudaDevAttrComputeCapabilityMajor) ? 8 : 0;
  return cudaSuccess;
}
const char* cudaGetErrorString(cudaError_t) { return "stub"; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
}

using namespace tensorrt_llm::kernels::cutlass_kernels;

extern "C" {

// w: [K, N] fp16 bits; bits = 8 or 4. processed/unprocessed: K * N * bits / 8 bytes; scales: N fp16 bits.
int ref_symmetric_quantize(const uint16_t* w, int64_t K, int64_t N, int bits, int8_t* processed,
                           int8_t* unprocessed, uint16_t* scales) {
  try {
    QuantType qt = bits == 8 ? QuantType::INT8_WEIGHT_ONLY : QuantType::PACKED_INT4_WEIGHT_ONLY;
    std::vector<size_t> shape{(size_t)K, (size_t)N};
    symmetric_quantize<half, half>(processed, unprocessed, reinterpret_cast<half*>(scales),
                                   reinterpret_cast<const half*>(w), shape, qt);
    return 0;
  } catch (...) { return 1; }
}

int ref_preprocess_weights(const int8_t* row_major, int64_t K, int64_t N, int bits, int8_t* processed) {
  try {
    QuantType qt = bits == 8 ? QuantType::INT8_WEIGHT_ONLY : QuantType::PACKED_INT4_WEIGHT_ONLY;
    std::vector<size_t> shape{(size_t)K, (size_t)N};
    preprocess_weights_for_mixed_gemm(processed, row_major, shape, qt);
    return 0;
  } catch (...) { return 1; }
}
}
