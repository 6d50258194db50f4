// RMSNorm / LayerNorm with fused residual-add and int8 quantisation, per-token and per-tensor
// quantisers.  One CTA per token row, the row lives in registers between the passes (one HBM read,
// one HBM write per element).
//
// Replaces (reference, T/ = tensorrt_llm_july-release-v1):
//   T/cpp/tensorrt_llm/kernels/layernormKernels.cu:60-194   generalLayerNorm (+ static / dynamic quant)
//   T/cpp/tensorrt_llm/kernels/quantization.cu:31-65         quantizedKernel (per-tensor)
//   T/cpp/tensorrt_llm/kernels/quantization.cu:93-117        perTokenQuantization
//   T/tensorrt_llm/functional.py:3195-3219                   rms_norm (TRT-native glue, k14)
// HBM-bound: algorithmic bytes per row = hidden * (2 in + 2 out) for fp16->fp16,
// hidden * (2 + 1) + 4 for fp16->int8 (+2*hidden each for the fused residual read and write).
#include "common.cuh"
#include "kernels.h"

namespace tb {

constexpr int kNormThreads = 512;
constexpr int kNormMaxIter = 4;  // hidden <= 512 * 8 * 4 = 16384

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // red[] may still be read from a previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float r = lane < nw ? red[lane] : (is_max ? -3.0e38f : 0.f);
  r = is_max ? warp_max(r) : warp_sum(r);
  return r;
}

struct NormParams {
  const __half* x;         // [rows, hidden]
  const __half* residual;  // optional: h = x + residual (fp16 add), normalise h
  __half* sum_out;         // optional: write h
  const __half* gamma;     // [hidden]
  const __half* beta;      // optional [hidden]
  __half* out;             // fp16 output (mode 0)
  int8_t* out_q;           // int8 output (mode 1, 2)
  const float* scale_in;   // static per-tensor scale (mode 1)
  float* scale_out;        // dynamic per-token scale [rows] (mode 2)
  float eps;
  int hidden;
  int mode;                // 0 fp16, 1 static int8, 2 dynamic int8
  int layernorm;           // 0 = RMSNorm, 1 = LayerNorm (mean subtracted, two-pass variance)
};

template <int ITER>
__global__ void __launch_bounds__(kNormThreads) norm_quant_kernel(NormParams p) {
  __shared__ float red[32];
  // programmatic dependent launch (decode steps: this kernel sits between two projection kernels of a CUDA graph): start as
  // the previous kernel drains, let the next one (which prefetches weights before it waits) start at once
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int row = blockIdx.x;
  const int nvec = p.hidden >> 3;
  const size_t base = (size_t) row * p.hidden;
  __half2 v[ITER][4];
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int i = threadIdx.x + it * kNormThreads;
    if (i < nvec) {
      uint4 raw = *reinterpret_cast<const uint4*>(p.x + base + (size_t) i * 8);
      __half2* h = reinterpret_cast<__half2*>(&raw);
      if (p.residual) {
        uint4 rr = *reinterpret_cast<const uint4*>(p.residual + base + (size_t) i * 8);
        const __half2* r2 = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 a = __half22float2(h[j]), b = __half22float2(r2[j]);
          h[j] = __floats2half2_rn(a.x + b.x, a.y + b.y);
        }
        if (p.sum_out) *reinterpret_cast<uint4*>(p.sum_out + base + (size_t) i * 8) = raw;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[it][j] = h[j];
        float2 f = __half22float2(h[j]);
        sum += f.x + f.y;
        sq += f.x * f.x + f.y * f.y;
      }
    }
  }
  float mean = 0.f, inv;
  if (p.layernorm) {
    mean = block_reduce(sum, red, false) / p.hidden;
    float var = 0.f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int i = threadIdx.x + it * kNormThreads;
      if (i < nvec) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __half22float2(v[it][j]);
          var += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
        }
      }
    }
    var = block_reduce(var, red, false);
    inv = rsqrtf(var / p.hidden + p.eps);
  } else {
    sq = block_reduce(sq, red, false);
    inv = rsqrtf(sq / p.hidden + p.eps);
  }

  // normalise -> fp16 (the reference rounds the normalised value to T before quantising)
  float amax = 0.f;
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int i = threadIdx.x + it * kNormThreads;
    if (i < nvec) {
      uint4 g4 = *reinterpret_cast<const uint4*>(p.gamma + (size_t) i * 8);
      const __half2* g = reinterpret_cast<const __half2*>(&g4);
      uint4 b4 = make_uint4(0, 0, 0, 0);
      if (p.beta) b4 = *reinterpret_cast<const uint4*>(p.beta + (size_t) i * 8);
      const __half2* bt = reinterpret_cast<const __half2*>(&b4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(v[it][j]), gg = __half22float2(g[j]), bb = __half22float2(bt[j]);
        float y0 = (f.x - mean) * inv * gg.x + bb.x;
        float y1 = (f.y - mean) * inv * gg.y + bb.y;
        v[it][j] = __floats2half2_rn(y0, y1);
        float2 r = __half22float2(v[it][j]);
        amax = fmaxf(amax, fmaxf(fabsf(r.x), fabsf(r.y)));
      }
      if (p.mode == 0) *reinterpret_cast<uint4*>(p.out + base + (size_t) i * 8) = *reinterpret_cast<uint4*>(v[it]);
    }
  }
  if (p.mode == 0) return;

  float qs;
  if (p.mode == 1) {
    qs = *p.scale_in;
  } else {
    // amax starts at fp16(1e-6) in the reference (T_scalar amax = 1e-6f)
    amax = fmaxf(block_reduce(amax, red, true), __half2float(__float2half_rn(1e-6f)));
    qs = 127.f / amax;
    if (threadIdx.x == 0) p.scale_out[row] = amax / 127.f;
  }
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int i = threadIdx.x + it * kNormThreads;
    if (i < nvec) {
      float2 a = __half22float2(v[it][0]), b = __half22float2(v[it][1]);
      float2 c = __half22float2(v[it][2]), d = __half22float2(v[it][3]);
      uint2 o;
      o.x = pack4_i8(a.x * qs, a.y * qs, b.x * qs, b.y * qs);
      o.y = pack4_i8(c.x * qs, c.y * qs, d.x * qs, d.y * qs);
      *reinterpret_cast<uint2*>(p.out_q + base + (size_t) i * 8) = o;
    }
  }
}

static int launch_norm(const NormParams& p, int rows, cudaStream_t stream) {
  if (p.hidden % 8 != 0 || p.hidden > kNormThreads * 8 * kNormMaxIter || rows <= 0) return -1;
  const int iters = (p.hidden / 8 + kNormThreads - 1) / kNormThreads;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(rows);
  cfg.blockDim = dim3(kNormThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = rows <= 64 ? 1 : 0;      // decode shapes only
  switch (iters) {
    case 1: return (int) cudaLaunchKernelEx(&cfg, norm_quant_kernel<1>, p);
    case 2: return (int) cudaLaunchKernelEx(&cfg, norm_quant_kernel<2>, p);
    default: return (int) cudaLaunchKernelEx(&cfg, norm_quant_kernel<4>, p);
  }
}

// ---------------------------------------------------------------------------------------------
// per-token quantiser (fp16 / fp32 in)
// ---------------------------------------------------------------------------------------------
// fp16 rows of up to ITER * 4096 columns: read ONCE with 16-byte loads and kept in registers between the amax and the
// quantise phase.  One row per CTA is latency-bound (load -> block reduce -> store), so what counts is resident CTAs: ITER
// is a template parameter and the bound below keeps the kernel at 32 registers = 4 CTAs per SM (a first version with a
// 4-deep register array for every width ran at 3 CTAs per SM and was slower than the two-pass kernel, 90 vs 65 us).
template <int ITER>
__global__ void __launch_bounds__(kNormThreads, ITER == 1 ? 4 : 2) per_token_quant_half_kernel(int8_t* dst, const __half* src,
                                                                                              int cols, float* scales) {
  __shared__ float red[32];
  const __half* s = src + (size_t) blockIdx.x * cols;
  int8_t* d = dst + (size_t) blockIdx.x * cols;
  uint4 v[ITER];
  float amax = 0.f;
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int i = (it * kNormThreads + threadIdx.x) * 8;
    v[it] = make_uint4(0, 0, 0, 0);
    if (i < cols) v[it] = *reinterpret_cast<const uint4*>(s + i);
    const __half2* h = reinterpret_cast<const __half2*>(&v[it]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
  amax = fmaxf(block_reduce(amax, red, true), __half2float(__float2half_rn(1e-6f)));   // localMax = T(1e-6f) in the reference
  if (threadIdx.x == 0) scales[blockIdx.x] = amax / 127.f;
  const float qs = 127.f / amax;
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int i = (it * kNormThreads + threadIdx.x) * 8;
    if (i < cols) {
      const __half2* h = reinterpret_cast<const __half2*>(&v[it]);
      const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), e = __half22float2(h[3]);
      uint2 q;
      q.x = pack4_i8(a.x * qs, a.y * qs, b.x * qs, b.y * qs);
      q.y = pack4_i8(c.x * qs, c.y * qs, e.x * qs, e.y * qs);
      *reinterpret_cast<uint2*>(d + i) = q;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kNormThreads) per_token_quant_kernel(int8_t* dst, const T* src, int cols,
                                                                       float* scales) {
  __shared__ float red[32];
  const T* s = src + (size_t) blockIdx.x * cols;
  int8_t* d = dst + (size_t) blockIdx.x * cols;
  float amax = 0.f;
  for (int i = threadIdx.x * 4; i < cols; i += kNormThreads * 4) {
    float f[4];
    if constexpr (sizeof(T) == 2) {
      uint2 raw = *reinterpret_cast<const uint2*>(s + i);
      float2 a = __half22float2(*reinterpret_cast<__half2*>(&raw.x)), b = __half22float2(*reinterpret_cast<__half2*>(&raw.y));
      f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    } else {
      float4 raw = *reinterpret_cast<const float4*>(s + i);
      f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w;
    }
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(f[0]), fabsf(f[1])), fmaxf(fabsf(f[2]), fabsf(f[3]))));
  }
  // localMax = T(1e-6f) in the reference
  const float floor_v = sizeof(T) == 2 ? __half2float(__float2half_rn(1e-6f)) : 1e-6f;
  amax = fmaxf(block_reduce(amax, red, true), floor_v);
  if (threadIdx.x == 0) scales[blockIdx.x] = amax / 127.f;
  const float qs = 127.f / amax;
  for (int i = threadIdx.x * 4; i < cols; i += kNormThreads * 4) {
    float f[4];
    if constexpr (sizeof(T) == 2) {
      uint2 raw = *reinterpret_cast<const uint2*>(s + i);
      float2 a = __half22float2(*reinterpret_cast<__half2*>(&raw.x)), b = __half22float2(*reinterpret_cast<__half2*>(&raw.y));
      f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    } else {
      float4 raw = *reinterpret_cast<const float4*>(s + i);
      f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w;
    }
    *reinterpret_cast<uint32_t*>(d + i) = pack4_i8(f[0] * qs, f[1] * qs, f[2] * qs, f[3] * qs);
  }
}

// SwiGLU + per-token int8 quantisation in one pass over the gate/up GEMM output (prefill: saves writing and re-reading
// the fp16 activation, 540 MB per layer at 16384 rows).  Same arithmetic as swiglu_kernel followed by
// per_token_quant_kernel<__half>: act = fp16(fp16(silu(g)) * u), amax over the row, q = rni_sat(act * (127 / amax)).
constexpr int kSqMaxIter = 4;   // row chunks of 8 held in registers per thread: inter <= 4 * 512 * 8
template <int kSqIter>
__global__ void __launch_bounds__(kNormThreads, kSqIter <= 3 ? 4 : 3) swiglu_quant_kernel(int8_t* dst, float* scales, const __half* gate,
                                                                    const __half* up, int inter, int in_stride) {
  __shared__ float red[32];
  const size_t row = blockIdx.x;
  uint4 act[kSqIter];
  float amax = 0.f;
#pragma unroll
  for (int it = 0; it < kSqIter; ++it) {
    const int i = (it * kNormThreads + threadIdx.x) * 8;
    if (i < inter) {
      const uint4 g4 = *reinterpret_cast<const uint4*>(gate + row * in_stride + i);
      const uint4 u4 = *reinterpret_cast<const uint4*>(up + row * in_stride + i);
      const __half2* g = reinterpret_cast<const __half2*>(&g4);
      const __half2* u = reinterpret_cast<const __half2*>(&u4);
      __half2* o = reinterpret_cast<__half2*>(&act[it]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 gf = __half22float2(g[j]), uf = __half22float2(u[j]);
        const __half2 a = __floats2half2_rn(silu_fast(gf.x), silu_fast(gf.y));
        const float2 af = __half22float2(a);
        o[j] = __floats2half2_rn(af.x * uf.x, af.y * uf.y);
        const float2 of = __half22float2(o[j]);
        amax = fmaxf(amax, fmaxf(fabsf(of.x), fabsf(of.y)));
      }
    }
  }
  amax = fmaxf(block_reduce(amax, red, true), __half2float(__float2half_rn(1e-6f)));
  if (threadIdx.x == 0) scales[row] = amax / 127.f;
  const float qs = 127.f / amax;
#pragma unroll
  for (int it = 0; it < kSqIter; ++it) {
    const int i = (it * kNormThreads + threadIdx.x) * 8;
    if (i < inter) {
      const __half2* o = reinterpret_cast<const __half2*>(&act[it]);
      const float2 a = __half22float2(o[0]), b = __half22float2(o[1]), c = __half22float2(o[2]), d = __half22float2(o[3]);
      uint2 q;
      q.x = pack4_i8(a.x * qs, a.y * qs, b.x * qs, b.y * qs);
      q.y = pack4_i8(c.x * qs, c.y * qs, d.x * qs, d.y * qs);
      *reinterpret_cast<uint2*>(dst + row * inter + i) = q;
    }
  }
}

template <typename T>
__global__ void quantize_tensor_kernel(int8_t* dst, const T* src, int64_t n4, const float* scale) {
  const float qs = __ldg(scale);
  for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n4; i += (int64_t) gridDim.x * blockDim.x) {
    float f[4];
    if constexpr (sizeof(T) == 2) {
      uint2 raw = reinterpret_cast<const uint2*>(src)[i];
      float2 a = __half22float2(*reinterpret_cast<__half2*>(&raw.x)), b = __half22float2(*reinterpret_cast<__half2*>(&raw.y));
      f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    } else {
      float4 raw = reinterpret_cast<const float4*>(src)[i];
      f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w;
    }
    reinterpret_cast<uint32_t*>(dst)[i] = pack4_i8(f[0] * qs, f[1] * qs, f[2] * qs, f[3] * qs);
  }
}

}  // namespace tb

using namespace tb;

extern "C" {

int tb_rmsnorm(void* out, const void* x, const void* residual, void* sum_out, const void* gamma, float eps,
               int rows, int hidden, cudaStream_t stream) {
  NormParams p{};
  p.x = (const __half*) x; p.residual = (const __half*) residual; p.sum_out = (__half*) sum_out;
  p.gamma = (const __half*) gamma; p.out = (__half*) out; p.eps = eps; p.hidden = hidden; p.mode = 0;
  return launch_norm(p, rows, stream);
}

int tb_rmsnorm_quant(int8_t* out_q, float* scale_out, const void* x, const void* residual, void* sum_out,
                     const void* gamma, const void* beta, const float* scale_in, float eps, int rows, int hidden,
                     int dynamic, int layernorm, cudaStream_t stream) {
  NormParams p{};
  p.x = (const __half*) x; p.residual = (const __half*) residual; p.sum_out = (__half*) sum_out;
  p.gamma = (const __half*) gamma; p.beta = (const __half*) beta; p.out_q = out_q; p.scale_in = scale_in;
  p.scale_out = scale_out; p.eps = eps; p.hidden = hidden; p.mode = dynamic ? 2 : 1; p.layernorm = layernorm;
  if (dynamic && !scale_out) return -1;
  if (!dynamic && !scale_in) return -1;
  return launch_norm(p, rows, stream);
}

int tb_quantize_per_token(int8_t* dst, float* scales, const void* src, int rows, int cols, int src_is_fp32,
                          cudaStream_t stream) {
  if (cols % 4 != 0 || rows <= 0) return -1;
  if (src_is_fp32) per_token_quant_kernel<float><<<rows, kNormThreads, 0, stream>>>(dst, (const float*) src, cols, scales);
  else if (cols % 8 == 0 && cols <= kNormThreads * 8)
    per_token_quant_half_kernel<1><<<rows, kNormThreads, 0, stream>>>(dst, (const __half*) src, cols, scales);
  else if (cols % 8 == 0 && cols <= kNormThreads * 8 * 3)
    per_token_quant_half_kernel<3><<<rows, kNormThreads, 0, stream>>>(dst, (const __half*) src, cols, scales);
  else per_token_quant_kernel<__half><<<rows, kNormThreads, 0, stream>>>(dst, (const __half*) src, cols, scales);
  return (int) cudaGetLastError();
}

int tb_swiglu_quant(int8_t* dst, float* scales, const void* gate, const void* up, int rows, int inter, int in_stride,
                    cudaStream_t stream) {
  if (inter % 8 || in_stride % 8 || rows <= 0 || inter > kSqMaxIter * kNormThreads * 8) return -1;
  if (inter <= 3 * kNormThreads * 8)
    swiglu_quant_kernel<3><<<rows, kNormThreads, 0, stream>>>(dst, scales, (const __half*) gate, (const __half*) up, inter,
                                                         in_stride);
  else
    swiglu_quant_kernel<kSqMaxIter><<<rows, kNormThreads, 0, stream>>>(dst, scales, (const __half*) gate, (const __half*) up, inter,
                                                         in_stride);
  return (int) cudaGetLastError();
}

int tb_quantize_tensor(int8_t* dst, const void* src, int64_t size, const float* scale, int src_is_fp32,
                       cudaStream_t stream) {
  if (size % 4 != 0 || size <= 0) return -1;
  const int64_t n4 = size / 4;
  const int blocks = (int) ((n4 + 255) / 256 < (int64_t) kNumSMs * 8 ? (n4 + 255) / 256 : kNumSMs * 8);
  if (src_is_fp32) quantize_tensor_kernel<float><<<blocks, 256, 0, stream>>>(dst, (const float*) src, n4, scale);
  else quantize_tensor_kernel<__half><<<blocks, 256, 0, stream>>>(dst, (const __half*) src, n4, scale);
  return (int) cudaGetLastError();
}
}
