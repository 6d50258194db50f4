// Shared device helpers for the sm_100a kernels: error macros, rounding primitives that pin the
// reference's numerics, mbarrier / TMA / tcgen05 PTX wrappers.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define TB_CHECK_CUDA(expr)                                                                        \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      fprintf(stderr, "[trtllm_b200] CUDA error %s at %s:%d: %s\n", #expr, __FILE__, __LINE__,    \
              cudaGetErrorString(e_));                                                             \
      return (int) e_;                                                                             \
    }                                                                                              \
  } while (0)

namespace tb {

constexpr int kNumSMs = 148;

// ---------------------------------------------------------------------------------------------
// numerics shared with the reference
// ---------------------------------------------------------------------------------------------
// cvt.rni.sat.s8.f32 — T/cpp/tensorrt_llm/common/cudaTypeUtils.cuh:327-371,
// K/decoderMaskedMultiheadAttentionUtils.h:2276-2286
__device__ __forceinline__ int8_t f2i8(float v) {
  int r;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(r) : "f"(v));
  return (int8_t) r;
}
__device__ __forceinline__ uint32_t pack4_i8(float a, float b, float c, float d) {
  return (uint32_t)(uint8_t) f2i8(a) | ((uint32_t)(uint8_t) f2i8(b) << 8) | ((uint32_t)(uint8_t) f2i8(c) << 16) |
         ((uint32_t)(uint8_t) f2i8(d) << 24);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// SiLU of the prefill-size passes (tb_swiglu, tb_swiglu_quant, the tcgen05 SwiGLU epilogue — one definition, so the fused
// and unfused forms stay bit-identical): x * rcp(1 + exp(-x)) with the approximate reciprocal.  The IEEE division these
// passes used before compiles to ~12 instructions + a slow-path call per element and made the 180 M-element SwiGLU +
// quantise pass of a cfg4 layer instruction-bound (40 instructions per element, 245 us against ~150 us of HBM time).  The
// result is rounded to fp16 right after (TRT fp16 activation), which absorbs the 2-ulp fp32 difference except within
// 2^-11 of a rounding boundary; 1 + exp(-x) beyond 2^126 (x < -87.3) returns -0 instead of a denormal-sized quotient.
__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.f + __expf(-v)); }

// exact int8 -> fp16 of 4 packed signed bytes: (b ^ 0x80) | 0x6400 is the fp16 1024 + (b + 128);
// subtracting 1152 gives b.  lo = bytes {0,1}, hi = bytes {2,3}.
__device__ __forceinline__ void i8x4_to_h2x2(uint32_t w, __half2& lo, __half2& hi) {
  uint32_t u = w ^ 0x80808080u, l, h;
  asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(l) : "r"(u), "r"(0x64646464u));
  asm("prmt.b32 %0, %1, %2, 0x4342;" : "=r"(h) : "r"(u), "r"(0x64646464u));
  const __half2 bias = __halves2half2(__ushort_as_half(0x6480), __ushort_as_half(0x6480));
  lo = __hsub2(*reinterpret_cast<__half2*>(&l), bias);
  hi = __hsub2(*reinterpret_cast<__half2*>(&h), bias);
}
// 8 packed signed int4 in this library's processed order (nibble positions 0..7 hold elements 0,2,4,6,1,3,5,7;
// quantization.py pack_processed_int4) -> 4 half2 {e0,e1},{e2,e3},{e4,e5},{e6,e7}, exact.
// n ^ 8 is the offset-binary nibble; (x & 0x000F000F) | 0x64006400 is the half2 {1024 + n_lo, 1024 + n_hi}: subtract
// 1032.  The nibbles one position up come out as 1024 + 16 n: one HFMA2 with 1/16 and -(64 + 8) (exact: <= 11 bits).
__device__ __forceinline__ void i4x8_to_h2x4(uint32_t w, __half2 out[4]) {
  const uint32_t u = w ^ 0x88888888u;
  const __half2 bias = __halves2half2(__ushort_as_half(0x6408), __ushort_as_half(0x6408));         // 1032
  const __half2 sixteenth = __halves2half2(__ushort_as_half(0x2C00), __ushort_as_half(0x2C00));    // 1/16
  const __half2 bias16 = __halves2half2(__ushort_as_half(0xD480), __ushort_as_half(0xD480));       // -72
  uint32_t t0, t1, t2, t3;
  asm("lop3.b32 %0, %1, 0x000F000F, 0x64006400, 0xEA;" : "=r"(t0) : "r"(u));        // (u & m) | c : elements {e0, e1}
  asm("lop3.b32 %0, %1, 0x00F000F0, 0x64006400, 0xEA;" : "=r"(t1) : "r"(u));        // {e2, e3} * 16
  const uint32_t v = u >> 8;
  asm("lop3.b32 %0, %1, 0x000F000F, 0x64006400, 0xEA;" : "=r"(t2) : "r"(v));        // {e4, e5}
  asm("lop3.b32 %0, %1, 0x00F000F0, 0x64006400, 0xEA;" : "=r"(t3) : "r"(v));        // {e6, e7} * 16
  out[0] = __hsub2(*reinterpret_cast<__half2*>(&t0), bias);
  out[1] = __hfma2(*reinterpret_cast<__half2*>(&t1), sixteenth, bias16);
  out[2] = __hsub2(*reinterpret_cast<__half2*>(&t2), bias);
  out[3] = __hfma2(*reinterpret_cast<__half2*>(&t3), sixteenth, bias16);
}

// ---------------------------------------------------------------------------------------------
// shared-memory address / mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (cudaErrorLaunchFailure from the
// C ABI), never as a hung GPU.  ~4 s at 2 GHz is far beyond any legitimate wait.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) {
      printf("[trtllm_b200] mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2-D tiles, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// 1-D bulk copy global -> shared (no tensor map), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM alloc, MMA, commit, ld
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major tile whose rows are exactly 128 bytes, stored with
// the 128-byte swizzle (what TMA SWIZZLE_128B writes): 8-row groups are 1024 B apart (SBO), the
// leading-byte-offset is unused for swizzled K-major, descriptor version 1 (Blackwell), layout 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
  d |= (uint64_t) 1 << 16;                        // LBO (ignored), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;               // SBO = 1024 B, bits [32,46)
  d |= (uint64_t) 1 << 46;                        // version = 1
  d |= (uint64_t) 2 << 61;                        // SWIZZLE_128B
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor layout): c_format [4,6), a_format [7,10),
// b_format [10,13), a_major bit 15, b_major bit 16 (0 = K-major), n>>3 [17,23), m>>4 [24,29).
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t c_fmt, uint32_t ab_fmt, uint32_t m, uint32_t n) {
  return (c_fmt << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t kIdescF16(uint32_t m, uint32_t n) { return umma_idesc(1u /*F32*/, 0u /*F16*/, m, n); }
__host__ __device__ constexpr uint32_t kIdescI8(uint32_t m, uint32_t n) { return umma_idesc(2u /*S32*/, 1u /*INT8*/, m, n); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// tcgen05.commit: arrive on an mbarrier when all previously issued MMAs of this thread completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t v[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// 16-byte streaming global load that does not pollute L1
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // namespace tb
