// Measurement support ("builder must measure a tcgen05 kind::i8 peak"): the int8 tensor-pipe ceiling of
// this B200 at the clock it actually sustains.  MEASURED_PEAKS.json holds HBM and cuBLAS bf16 only; the prefill roofline
// of the SmoothQuant GEMMs needs an int8 denominator that is neither "2 x bf16" nor the 4.5 POPS datasheet figure.
//
// One CTA per SM; one thread issues `iters` x 4 back-to-back tcgen05.mma.kind::i8 (M = 128, N = 256, K = 32) on fixed
// shared-memory operand tiles (no loads in the loop), alternating between two TMEM accumulators, then one commit.
// ops = gridDim.x * iters * 4 * 2 * 128 * 256 * 32.  kind = 1 measures kind::f16 (K = 16) the same way.
#include "common.cuh"
#include "kernels.h"

namespace tb {

template <bool I8>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int iters, int* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint8_t* sA = smem;                  // 128 rows x 128 B, SWIZZLE_128B K-major
  uint8_t* sB = smem + 128 * 128;      // 256 rows x 128 B
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (tid == 0) {
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 32) {
    constexpr uint32_t idesc = I8 ? kIdescI8(128, 256) : kIdescF16(128, 256);
    const uint64_t ad = umma_desc_sw128(smem_u32(sA)), bd = umma_desc_sw128(smem_u32(sB));
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + (uint32_t) ((it & 1) * 256);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (I8) umma_i8(d, ad + 2 * k, bd + 2 * k, idesc, (it > 1 || k > 0) ? 1u : 0u);
        else umma_f16(d, ad + 2 * k, bd + 2 * k, idesc, (it > 1 || k > 0) ? 1u : 0u);
      }
    }
    umma_commit(&done);
  }
  mbar_wait(&done, 0);
  tc_fence_after();
  if (warp == 0) {
    uint32_t v[16];
    tmem_ld16(tmem_base, v);
    tmem_ld_wait();
    if (sink && v[0] == 0x7fffffffu) sink[0] = (int) v[1];   // keeps the accumulator observable; never true
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace tb

using namespace tb;

extern "C" int tb_mma_peak(int kind, int iters, int ctas, int* sink, double* ops_out, cudaStream_t stream) {
  if (iters < 2 || ctas < 1 || kind < 0 || kind > 1) return -1;
  const size_t smem = (128 + 256) * 128 + 1024;
  auto kern = kind == 0 ? mma_peak_kernel<true> : mma_peak_kernel<false>;
  TB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  kern<<<ctas, 128, smem, stream>>>(iters, sink);
  if (ops_out) *ops_out = (double) ctas * iters * 4.0 * 2.0 * 128.0 * 256.0 * (kind == 0 ? 32.0 : 16.0);
  return (int) cudaGetLastError();
}
