// Decode-shape (M <= 4 token rows) matrix-vector path: Y[M,N] = X[M,K] . W[N,K]^T with W streamed
// exactly once from HBM in 16-byte loads, X staged in shared memory, fp32 (int32 for W8A8)
// accumulation, warp-shuffle reduction and a fused epilogue (per-channel / per-token scales,
// SwiGLU, residual add).  HBM-bound: algorithmic bytes = N*K*bytes_per_weight (+ M*(K+N)*2).
//
// Replaces (reference):
//   T/cpp/tensorrt_llm/kernels/weightOnlyMatrixVectorMultiplication.cu:136-277,371-378 (int8/int4 GEMV)
//   the M<=4 calls of CutlassInt8GemmRunner::gemm (int8_gemm_template.h:356-369) and of
//   GemmPlugin/cuBLAS (P/gemmPlugin/gemmPlugin.cpp:121-230) made by the decode step.
// Weight layouts (this repo's "processed" layouts, produced by preprocess.cpp / quantize ops):
//   fp16: [N, K] (torch Linear)      int8: [N, K]      int4: [N, K/2], low nibble = even k.
#include "common.cuh"
#include "kernels.h"

namespace tb {

enum GemvKind { kF16 = 0, kW8 = 1, kW4 = 2, kA8W8 = 3 };

struct GemvParams {
  const void* x;          // [M, K] fp16 (int8 for kA8W8)
  const void* w;          // see layouts above
  const __half* w_scale;  // [N] fp16 per-channel (kW8/kW4)
  const float* sc;        // kA8W8: per-channel [N] or [1]
  const float* sr;        // kA8W8: per-token [M] or [1]
  int sc_per_channel, sr_per_token;
  const __half* residual;  // optional [M, N_out]
  __half* y;               // [M, N_out]
  float* y_f32;            // optional fp32 output instead of fp16 (lm_head logits)
  int M, N, K;
  int swiglu;              // W holds [2*N_out, K]: rows [0,N_out) = gate(fc), [N_out, 2N_out) = up
  int n_out;
};

constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = kGemvThreads / 32;

template <int KIND> struct KTraits;
template <> struct KTraits<kF16>  { static constexpr int kElemsPer16B = 8;  };
template <> struct KTraits<kW8>   { static constexpr int kElemsPer16B = 16; };
template <> struct KTraits<kW4>   { static constexpr int kElemsPer16B = 32; };
template <> struct KTraits<kA8W8> { static constexpr int kElemsPer16B = 16; };

__device__ __forceinline__ float silu_f(float v) { return v / (1.f + __expf(-v)); }

// dot of one 16-byte weight chunk with the matching activation chunk(s) for MB rows
template <int KIND, int MB>
__device__ __forceinline__ void chunk_fma(const uint4& wq, const uint8_t* xs, int k0, int K, float (&acc)[MB],
                                          int (&iacc)[MB]) {
  if constexpr (KIND == kF16) {
    const __half2* w2 = reinterpret_cast<const __half2*>(&wq);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      uint4 xv = *reinterpret_cast<const uint4*>(xs + ((size_t) m * K + k0) * 2);
      const __half2* x2 = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 a = __half22float2(w2[j]), b = __half22float2(x2[j]);
        acc[m] = fmaf(a.x, b.x, acc[m]);
        acc[m] = fmaf(a.y, b.y, acc[m]);
      }
    }
  } else if constexpr (KIND == kW8) {
    __half2 wh[8];
    i8x4_to_h2x2(wq.x, wh[0], wh[1]);
    i8x4_to_h2x2(wq.y, wh[2], wh[3]);
    i8x4_to_h2x2(wq.z, wh[4], wh[5]);
    i8x4_to_h2x2(wq.w, wh[6], wh[7]);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const uint4* xp = reinterpret_cast<const uint4*>(xs + ((size_t) m * K + k0) * 2);
      uint4 xa = xp[0], xb = xp[1];
      const __half2* x2a = reinterpret_cast<const __half2*>(&xa);
      const __half2* x2b = reinterpret_cast<const __half2*>(&xb);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 a = __half22float2(wh[j]), b = __half22float2(x2a[j]);
        acc[m] = fmaf(a.x, b.x, acc[m]);
        acc[m] = fmaf(a.y, b.y, acc[m]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 a = __half22float2(wh[4 + j]), b = __half22float2(x2b[j]);
        acc[m] = fmaf(a.x, b.x, acc[m]);
        acc[m] = fmaf(a.y, b.y, acc[m]);
      }
    }
  } else if constexpr (KIND == kW4) {
    __half2 wh[16];
    i4x8_to_h2x4(wq.x, wh + 0);
    i4x8_to_h2x4(wq.y, wh + 4);
    i4x8_to_h2x4(wq.z, wh + 8);
    i4x8_to_h2x4(wq.w, wh + 12);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const uint4* xp = reinterpret_cast<const uint4*>(xs + ((size_t) m * K + k0) * 2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 xv = xp[q];
        const __half2* x2 = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 a = __half22float2(wh[q * 4 + j]), b = __half22float2(x2[j]);
          acc[m] = fmaf(a.x, b.x, acc[m]);
          acc[m] = fmaf(a.y, b.y, acc[m]);
        }
      }
    }
  } else {  // kA8W8
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t) m * K + k0);
      iacc[m] = __dp4a((int) wq.x, (int) xv.x, iacc[m]);
      iacc[m] = __dp4a((int) wq.y, (int) xv.y, iacc[m]);
      iacc[m] = __dp4a((int) wq.z, (int) xv.z, iacc[m]);
      iacc[m] = __dp4a((int) wq.w, (int) xv.w, iacc[m]);
    }
  }
}

// Each warp owns one output column n (two weight rows when SwiGLU is fused).  Persistent-style:
// a CTA walks columns n = warp_global + i * total_warps so X is staged once per CTA.
template <int KIND, int MB, bool SWIGLU>
__global__ void __launch_bounds__(kGemvThreads) gemv_kernel(GemvParams p) {
  extern __shared__ __align__(16) uint8_t xs[];
  constexpr int EPC = KTraits<KIND>::kElemsPer16B;   // k elements per 16-byte weight chunk
  const int K = p.K;
  const int x_bytes = p.M * K * (KIND == kA8W8 ? 1 : 2);
  const int xs_bytes = MB * K * (KIND == kA8W8 ? 1 : 2);    // rows [M, MB) are zero-filled
  for (int i = threadIdx.x * 16; i < xs_bytes; i += kGemvThreads * 16)
    *reinterpret_cast<uint4*>(xs + i) =
        i < x_bytes ? *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.x) + i) : make_uint4(0, 0, 0, 0);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int total_warps = gridDim.x * kGemvWarps;
  const int chunks = K / EPC;                         // 16-byte chunks per weight row
  const size_t row_bytes = (size_t) chunks * 16;
  const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w);
  constexpr int R = SWIGLU ? 2 : 1;

  for (int n = blockIdx.x * kGemvWarps + warp; n < p.n_out; n += total_warps) {
    float acc[R][MB];
    int iacc[R][MB];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int m = 0; m < MB; ++m) { acc[r][m] = 0.f; iacc[r][m] = 0; }
    const uint8_t* wr[R];
    wr[0] = wbase + (size_t) n * row_bytes;
    if constexpr (SWIGLU) wr[1] = wbase + (size_t) (n + p.n_out) * row_bytes;

    constexpr int U = SWIGLU ? 4 : 8;                 // chunks in flight per lane per row
    int c = lane;
    for (; c + 32 * (U - 1) < chunks; c += 32 * U) {
      uint4 wq[R][U];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int u = 0; u < U; ++u) wq[r][u] = ldg_nc_v4(wr[r] + (size_t) (c + 32 * u) * 16);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int u = 0; u < U; ++u) chunk_fma<KIND, MB>(wq[r][u], xs, (c + 32 * u) * EPC, K, acc[r], iacc[r]);
    }
    for (; c < chunks; c += 32) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        uint4 wq = ldg_nc_v4(wr[r] + (size_t) c * 16);
        chunk_fma<KIND, MB>(wq, xs, c * EPC, K, acc[r], iacc[r]);
      }
    }
    // reduce across the warp, lane 0 applies the epilogue
    float res[R][MB];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        if constexpr (KIND == kA8W8) {
          int v = iacc[r][m];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          res[r][m] = (float) v;
        } else {
          res[r][m] = warp_sum(acc[r][m]);
        }
      }
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        if (m >= p.M) break;
        float v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int nr = n + r * p.n_out;
          v[r] = res[r][m];
          if constexpr (KIND == kW8 || KIND == kW4) v[r] *= __half2float(p.w_scale[nr]);
          if constexpr (KIND == kA8W8) {
            // reference grouping: accum * (scale_col * scale_row)  (epilogue_per_row_per_col_scale.h:325,341)
            const float scv = p.sc[p.sc_per_channel ? nr : 0], srv = p.sr[p.sr_per_token ? m : 0];
            v[r] = v[r] * (scv * srv);
          }
        }
        float o;
        if constexpr (SWIGLU) {
          const float g = __half2float(__float2half_rn(v[0])), u = __half2float(__float2half_rn(v[1]));
          o = __half2float(__float2half_rn(silu_f(g))) * u;
        } else {
          o = v[0];
        }
        const size_t oi = (size_t) m * p.n_out + n;
        if (p.y_f32) {
          p.y_f32[oi] = o;
        } else {
          __half oh = __float2half_rn(o);
          if (p.residual) oh = __float2half_rn(__half2float(oh) + __half2float(p.residual[oi]));
          p.y[oi] = oh;
        }
      }
    }
  }
}

template <int KIND, int MB, bool SWIGLU>
static int launch_gemv_t(const GemvParams& p, cudaStream_t stream) {
  const size_t smem = (size_t) MB * p.K * (KIND == kA8W8 ? 1 : 2);
  auto kern = gemv_kernel<KIND, MB, SWIGLU>;
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) return -2;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
  }
  // persistent grid: a multiple of the SM count, capped by the number of columns
  int per_sm = smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4);
  int grid = kNumSMs * per_sm;
  const int need = (p.n_out + kGemvWarps - 1) / kGemvWarps;
  if (grid > need) grid = need;
  kern<<<grid, kGemvThreads, smem, stream>>>(p);
  return (int) cudaGetLastError();
}

template <int KIND, bool SWIGLU>
static int launch_gemv_m(const GemvParams& p, cudaStream_t stream) {
  if (p.M == 1) return launch_gemv_t<KIND, 1, SWIGLU>(p, stream);
  if (p.M == 2) return launch_gemv_t<KIND, 2, SWIGLU>(p, stream);
  if (p.M <= 4) return launch_gemv_t<KIND, 4, SWIGLU>(p, stream);
  return -3;
}

}  // namespace tb

using namespace tb;

extern "C" int tb_gemv(int kind, void* y, float* y_f32, const void* x, const void* w, const void* w_scale,
                       const float* sc, const float* sr, int sc_per_channel, int sr_per_token, const void* residual,
                       int M, int N, int K, int swiglu, cudaStream_t stream) {
  GemvParams p{};
  p.x = x; p.w = w; p.w_scale = (const __half*) w_scale; p.sc = sc; p.sr = sr;
  p.sc_per_channel = sc_per_channel; p.sr_per_token = sr_per_token; p.residual = (const __half*) residual;
  p.y = (__half*) y; p.y_f32 = y_f32; p.M = M; p.N = N; p.K = K; p.swiglu = swiglu;
  p.n_out = swiglu ? N / 2 : N;
  const int epc = kind == kF16 ? 8 : (kind == kW4 ? 32 : 16);
  if (M < 1 || M > 4 || K % epc != 0 || (swiglu && (N & 1))) return -1;
  if ((kind == kW8 || kind == kW4) && !w_scale) return -1;
  if (kind == kA8W8 && (!sc || !sr)) return -1;
  switch (kind) {
    case kF16:  return swiglu ? launch_gemv_m<kF16, true>(p, stream)  : launch_gemv_m<kF16, false>(p, stream);
    case kW8:   return swiglu ? launch_gemv_m<kW8, true>(p, stream)   : launch_gemv_m<kW8, false>(p, stream);
    case kW4:   return swiglu ? launch_gemv_m<kW4, true>(p, stream)   : launch_gemv_m<kW4, false>(p, stream);
    case kA8W8: return swiglu ? launch_gemv_m<kA8W8, true>(p, stream) : launch_gemv_m<kA8W8, false>(p, stream);
  }
  return -1;
}
