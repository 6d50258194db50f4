// Decode-shape (M <= 4 token rows) projection: Y[M,N] = X[M,K] . W[N,K]^T, the kernel that streams the
// model's weights once per generated token and therefore bounds decode tokens/s (HBM roofline).
//
// sm_100a design (measured with tools/membench.cu on this pool's B200s: a ring of cp.async.bulk copies tops out
// at ~6.1 TB/s with a ~4 us ramp because every stage is a dependent round trip, while plain 16-byte LDG streams
// with >= 128 KB in flight per SM reach 7.3 TB/s with no ramp): one warp per output row, rows dealt round-robin
// so that at any instant the whole chip reads one contiguous window of the weight matrix; each lane keeps eight
// 16-byte loads in flight (4 KB per warp, 4 CTAs x 8 warps per SM = 128 KB per SM); activations are staged once per
// CTA in shared memory; products are exact (fp16 x fp16 in fp32), accumulation fp32 (int32 dp4a for W8A8); fused
// epilogue (per-channel / per-token scales, SwiGLU, residual add).  Variants that measured slower on LLaMA-7B
// decode steps are listed in DESIGN.md section 5.
// Fused prologues (the TensorRT-native glue / extra plugins of the reference):
//   RMSNorm of the residual stream, RMSNorm + dynamic per-token int8 quantisation (RmsnormQuantization),
//   plain dynamic per-token quantisation (QuantizePerToken) — each CTA recomputes the row statistics of the
//   8-22 KB activation it stages anyway (L2-resident), removing two to four launches per layer.
// Programmatic dependent launch: griddepcontrol.launch_dependents is issued at once, so the next kernel's CTAs
// become resident as this grid drains; griddepcontrol.wait guards the first read of upstream activations.
// 5..8 token rows, and int4 weights at any M, go to the tensor-core kernel in gemv_mma.cu (see tb_gemv_fused).
//
// Replaces (reference):
//   T/cpp/tensorrt_llm/kernels/weightOnlyMatrixVectorMultiplication.cu:136-277,371-378 (int8/int4 GEMV)
//   the M<=4 calls of CutlassInt8GemmRunner::gemm (int8_gemm_template.h:356-369) and of
//   GemmPlugin/cuBLAS (P/gemmPlugin/gemmPlugin.cpp:121-230) made by the decode step.
// Weight layouts (this library's processed layouts, quantization.py):
//   fp16: [N, K] (torch Linear)      int8: [N, K]      int4: [N, K/2], nibbles interleaved per 8 k (pack_processed_int4).
// Algorithmic bytes per launch = N*K*bytes_per_weight (+ M*(K+N)*2, < 0.1 %).
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace tb {

enum GemvKind { kF16 = 0, kW8 = 1, kW4 = 2, kA8W8 = 3 };
enum GemvPrologue { kProNone = 0, kProRms = 1, kProRmsQuant = 2, kProQuant = 3 };

struct GemvParams {
  const void* x;          // [M, K] fp16 (int8 for kA8W8 without a quantising prologue)
  const void* w;          // see layouts above
  const __half* w_scale;  // [N] fp16 per-channel (kW8/kW4)
  const float* sc;        // kA8W8: per-channel [N] or [1]
  const float* sr;        // kA8W8: per-token [M] or [1] (ignored when the prologue quantises)
  int sc_per_channel, sr_per_token;
  const __half* residual;  // optional [M, N_out]
  __half* y;               // [M, N_out]
  float* y_f32;            // optional fp32 output instead of fp16 (lm_head logits)
  int M, N, K;
  int swiglu;              // W holds [2*N_out, K]: rows [0,N_out) = gate(fc), [N_out, 2N_out) = up
  int n_out;
  int prologue;            // GemvPrologue
  const __half* gamma;     // [K] RMSNorm weight (prologue 1, 2)
  float eps;
  // optional: the first window(s) of the weights the NEXT projection of the step will stream (tb_gemv_hint_next); every
  // warp requests its share into L2 when it runs out of rows, so HBM keeps streaming through this grid's tail, the launch
  // gap and the next kernel's activation prologue.  Measured (B200, graph replay): 8-12 MB windows take the cfg2 step from
  // 2.68 to 2.57 ms and the SmoothQuant step from 1.71 to 1.64 ms; 48 MB windows LOSE (2.74 / 1.82 ms: the requests queue in
  // front of the stragglers' demand loads), and a warp requesting its OWN first rows right after griddepcontrol.wait loses
  // too (2.75 / 1.74 ms: a demand load that meets an in-flight prefetch of the same line is slower than the load alone)
  const uint8_t* pf[2];
  unsigned pf_lines[2];    // 128-byte lines per region
};

struct GemvNextHint { const void* p[2]; size_t bytes[2]; };
extern thread_local GemvNextHint g_gemv_next;

constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = kGemvThreads / 32;
constexpr int kGemvU = 8;                                     // 16-byte weight loads in flight per lane
constexpr int kProRegIters = 6;                               // prologue: a row of up to 6 x 2048 halves stays in registers

template <int KIND> struct KTraits;
template <> struct KTraits<kF16>  { static constexpr int kElemsPer16B = 8;  };
template <> struct KTraits<kW8>   { static constexpr int kElemsPer16B = 16; };
template <> struct KTraits<kW4>   { static constexpr int kElemsPer16B = 32; };
template <> struct KTraits<kA8W8> { static constexpr int kElemsPer16B = 16; };

__device__ __forceinline__ float silu_f(float v) { return v / (1.f + __expf(-v)); }

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// dot of one 16-byte weight chunk with the matching activation chunk(s) for MB rows.
// xs: staged activations, row stride `xstride` bytes; k0: first k element of this chunk.
template <int KIND, int MB>
__device__ __forceinline__ void chunk_fma(const uint4& wq, const uint8_t* xs, int k0, int xstride, float (&acc)[MB],
                                          int (&iacc)[MB]) {
  if constexpr (KIND == kF16) {
    const __half2* w2 = reinterpret_cast<const __half2*>(&wq);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t) m * xstride + (size_t) k0 * 2);
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
      const uint4* xp = reinterpret_cast<const uint4*>(xs + (size_t) m * xstride + (size_t) k0 * 2);
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
      const uint4* xp = reinterpret_cast<const uint4*>(xs + (size_t) m * xstride + (size_t) k0 * 2);
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
      uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t) m * xstride + k0);
      iacc[m] = __dp4a((int) wq.x, (int) xv.x, iacc[m]);
      iacc[m] = __dp4a((int) wq.y, (int) xv.y, iacc[m]);
      iacc[m] = __dp4a((int) wq.z, (int) xv.z, iacc[m]);
      iacc[m] = __dp4a((int) wq.w, (int) xv.w, iacc[m]);
    }
  }
}

// sum / max over the CTA; red has kGemvWarps floats, safe to call back to back
__device__ __forceinline__ float cta_reduce(float v, float* red, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kGemvWarps; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

template <int KIND, int MB, bool SWIGLU>
__global__ void __launch_bounds__(kGemvThreads, (MB >= 4 ? 2 : 4)) gemv_kernel(const GemvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int EPC = KTraits<KIND>::kElemsPer16B;     // k elements per 16-byte weight chunk
  constexpr int XB = KIND == kA8W8 ? 1 : 2;            // bytes per staged activation element
  constexpr int U = kGemvU;
  float* red = reinterpret_cast<float*>(smem);          // [8] reduction scratch
  float* srow = red + 8;                                // [4] per-token scales (W8A8)
  uint8_t* xs = reinterpret_cast<uint8_t*>(srow + 8);   // [MB][K * XB] staged activations

  const int K = p.K;
  const int xstride = K * XB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cpr = K / EPC;                              // 16-byte chunks per weight row
  const size_t row_bytes = (size_t) cpr * 16;
  const int gw = blockIdx.x * kGemvWarps + warp, tw = gridDim.x * kGemvWarps;
  const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w);

  // activations come from the upstream kernel (programmatic dependent launch: this grid may already be
  // resident while it drains); let the downstream kernel start its own launch as early as resources allow
  pdl_wait();
  pdl_launch_dependents();

  const int tid = threadIdx.x;
  if (p.prologue == kProNone) {
    const int x_bytes = p.M * xstride, xs_bytes = MB * xstride;     // rows [M, MB) are zero-filled
    for (int i = tid * 16; i < xs_bytes; i += kGemvThreads * 16)
      *reinterpret_cast<uint4*>(xs + i) =
          i < x_bytes ? *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.x) + i) : make_uint4(0, 0, 0, 0);
    if (KIND == kA8W8 && tid < 4) srow[tid] = tid < p.M ? p.sr[p.sr_per_token ? tid : 0] : 0.f;
  } else {
    // x is fp16 [M, K]; per row: (RMSNorm ->) fp16 (-> dynamic int8).  Same arithmetic as norm_quant.cu.
    const __half* xin = reinterpret_cast<const __half*>(p.x);
    for (int m = 0; m < MB; ++m) {
      if (m >= p.M) {
        for (int i = tid * 16; i < xstride; i += kGemvThreads * 16)
          *reinterpret_cast<uint4*>(xs + (size_t) m * xstride + i) = make_uint4(0, 0, 0, 0);
        if (tid == 0) srow[m] = 0.f;
        continue;
      }
      const __half* xr = xin + (size_t) m * K;
      if (K <= kProRegIters * kGemvThreads * 8) {
        // the row in registers: ONE trip to L2 for the whole prologue instead of one per pass (norm: 2, norm + quantise:
        // 3); same arithmetic in the same order as the streaming version below, which stays for longer rows
        uint4 raw[kProRegIters];
#pragma unroll
        for (int it = 0; it < kProRegIters; ++it) {
          const int i = (it * kGemvThreads + tid) * 8;
          if (i < K) raw[it] = *reinterpret_cast<const uint4*>(xr + i);
        }
        if (p.prologue != kProQuant) {
          float sq = 0.f;
#pragma unroll
          for (int it = 0; it < kProRegIters; ++it) {
            if ((it * kGemvThreads + tid) * 8 < K) {
              const __half2* h = reinterpret_cast<const __half2*>(&raw[it]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 f = __half22float2(h[j]);
                sq += f.x * f.x + f.y * f.y;
              }
            }
          }
          sq = cta_reduce(sq, red, false);
          const float inv = rsqrtf(sq / K + p.eps);
#pragma unroll
          for (int it = 0; it < kProRegIters; ++it) {
            const int i = (it * kGemvThreads + tid) * 8;
            if (i < K) {
              const uint4 g4 = *reinterpret_cast<const uint4*>(p.gamma + i);
              const __half2* g = reinterpret_cast<const __half2*>(&g4);
              __half2* h = reinterpret_cast<__half2*>(&raw[it]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 f = __half22float2(h[j]), gg = __half22float2(g[j]);
                h[j] = __floats2half2_rn(f.x * inv * gg.x, f.y * inv * gg.y);
              }
            }
          }
        }
        if constexpr (KIND == kA8W8) {
          float amax = 0.f;
#pragma unroll
          for (int it = 0; it < kProRegIters; ++it) {
            if ((it * kGemvThreads + tid) * 8 < K) {
              const __half2* h = reinterpret_cast<const __half2*>(&raw[it]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 f = __half22float2(h[j]);
                amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
              }
            }
          }
          amax = fmaxf(cta_reduce(amax, red, true), __half2float(__float2half_rn(1e-6f)));
          const float qs = 127.f / amax;
          if (tid == 0) srow[m] = amax / 127.f;
#pragma unroll
          for (int it = 0; it < kProRegIters; ++it) {
            const int i = (it * kGemvThreads + tid) * 8;
            if (i < K) {
              const __half2* h = reinterpret_cast<const __half2*>(&raw[it]);
              float f[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 t = __half22float2(h[j]);
                f[2 * j] = t.x * qs;
                f[2 * j + 1] = t.y * qs;
              }
              uint2 o;
              o.x = pack4_i8(f[0], f[1], f[2], f[3]);
              o.y = pack4_i8(f[4], f[5], f[6], f[7]);
              *reinterpret_cast<uint2*>(xs + (size_t) m * xstride + i) = o;
            }
          }
        } else {
#pragma unroll
          for (int it = 0; it < kProRegIters; ++it) {
            const int i = (it * kGemvThreads + tid) * 8;
            if (i < K) *reinterpret_cast<uint4*>(xs + (size_t) m * xstride + (size_t) i * 2) = raw[it];
          }
        }
        continue;
      }
      float inv = 1.f;
      if (p.prologue != kProQuant) {
        float sq = 0.f;
        for (int i = tid * 8; i < K; i += kGemvThreads * 8) {
          uint4 raw = *reinterpret_cast<const uint4*>(xr + i);
          const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(h[j]);
            sq += f.x * f.x + f.y * f.y;
          }
        }
        sq = cta_reduce(sq, red, false);
        inv = rsqrtf(sq / K + p.eps);
      }
      float amax = 0.f;
      for (int i = tid * 8; i < K; i += kGemvThreads * 8) {
        uint4 raw = *reinterpret_cast<const uint4*>(xr + i);
        __half2* h = reinterpret_cast<__half2*>(&raw);
        if (p.prologue != kProQuant) {
          uint4 g4 = *reinterpret_cast<const uint4*>(p.gamma + i);
          const __half2* g = reinterpret_cast<const __half2*>(&g4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(h[j]), gg = __half22float2(g[j]);
            h[j] = __floats2half2_rn(f.x * inv * gg.x, f.y * inv * gg.y);
          }
        }
        if constexpr (KIND == kA8W8) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(h[j]);
            amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
          }
        } else {
          *reinterpret_cast<uint4*>(xs + (size_t) m * xstride + (size_t) i * 2) = raw;
        }
      }
      if constexpr (KIND == kA8W8) {
        amax = fmaxf(cta_reduce(amax, red, true), __half2float(__float2half_rn(1e-6f)));
        const float qs = 127.f / amax;
        if (tid == 0) srow[m] = amax / 127.f;
        for (int i = tid * 8; i < K; i += kGemvThreads * 8) {
          uint4 raw = *reinterpret_cast<const uint4*>(xr + i);
          __half2* h = reinterpret_cast<__half2*>(&raw);
          float f[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 t = __half22float2(h[j]);
            if (p.prologue != kProQuant) {
              float2 gg = __half22float2(reinterpret_cast<const __half2*>(p.gamma + i)[j]);
              t = __half22float2(__floats2half2_rn(t.x * inv * gg.x, t.y * inv * gg.y));
            }
            f[2 * j] = t.x * qs;
            f[2 * j + 1] = t.y * qs;
          }
          uint2 o;
          o.x = pack4_i8(f[0], f[1], f[2], f[3]);
          o.y = pack4_i8(f[4], f[5], f[6], f[7]);
          *reinterpret_cast<uint2*>(xs + (size_t) m * xstride + i) = o;
        }
      }
    }
  }
  __syncthreads();

  constexpr int R = SWIGLU ? 2 : 1;                  // weight rows per output (gate row, up row)
  constexpr int UR = U / R;                            // 16-byte loads in flight per lane per row
  for (int n = gw; n < p.n_out; n += tw) {
    float acc[R][MB];
    int iacc[R][MB];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int m = 0; m < MB; ++m) { acc[r][m] = 0.f; iacc[r][m] = 0; }
    const uint8_t* wr[R];
    wr[0] = wbase + (size_t) n * row_bytes;
    if constexpr (SWIGLU) wr[1] = wbase + (size_t) (n + p.n_out) * row_bytes;

    int c = lane;
    for (; c + 32 * (UR - 1) < cpr; c += 32 * UR) {
      uint4 wq[R][UR];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int u = 0; u < UR; ++u) wq[r][u] = ldg_nc_v4(wr[r] + (size_t) (c + 32 * u) * 16);
#pragma unroll
      for (int u = 0; u < UR; ++u)
#pragma unroll
        for (int r = 0; r < R; ++r) chunk_fma<KIND, MB>(wq[r][u], xs, (c + 32 * u) * EPC, xstride, acc[r], iacc[r]);
    }
    for (; c < cpr; c += 32) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint4 wq = ldg_nc_v4(wr[r] + (size_t) c * 16);
        chunk_fma<KIND, MB>(wq, xs, c * EPC, xstride, acc[r], iacc[r]);
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
          // reference grouping: accum * (scale_col * scale_row)  (epilogue_per_row_per_col_scale.h:325,341)
          if constexpr (KIND == kA8W8) v[r] = v[r] * (p.sc[p.sc_per_channel ? nr : 0] * srow[m]);
        }
        float o;
        if constexpr (SWIGLU) {
          const float gte = __half2float(__float2half_rn(v[0])), up = __half2float(__float2half_rn(v[1]));
          o = __half2float(__float2half_rn(silu_f(gte))) * up;
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
  // out of rows: request this warp's share of the next projection's first weight window into L2 (lines dealt round-robin
  // over all warps of the grid, so the chip requests one contiguous window front to back)
  {
    const unsigned total = p.pf_lines[0] + p.pf_lines[1];
    for (unsigned l = (unsigned) gw * 32 + lane; l < total; l += (unsigned) tw * 32) {
      const uint8_t* a = l < p.pf_lines[0] ? p.pf[0] + (size_t) l * 128 : p.pf[1] + (size_t) (l - p.pf_lines[0]) * 128;
      prefetch_l2_line(a);
    }
  }
}

template <int KIND, int MB, bool SWIGLU>
static int launch_gemv_t(const GemvParams& p, cudaStream_t stream) {
  const size_t xs_bytes = (size_t) MB * p.K * (KIND == kA8W8 ? 1 : 2);
  const size_t smem = 16 * sizeof(float) + xs_bytes;
  if (smem > 200 * 1024) return -2;
  auto kern = gemv_kernel<KIND, MB, SWIGLU>;
  static bool attr_done = false;   // per template instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int) e;
    attr_done = true;
  }
  // one warp per output row, dealt round-robin: size the grid to a whole number of resident CTAs per SM
  int per_sm = smem > 100 * 1024 ? 1 : (smem > 72 * 1024 ? 2 : (smem > 54 * 1024 ? 3 : 4));
  if (MB >= 4 && per_sm > 2) per_sm = 2;     // register budget of the MB = 4 variants
  int grid = kNumSMs * per_sm;
  const int need = (p.n_out + kGemvWarps - 1) / kGemvWarps;
  if (grid > need) grid = need;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int) cudaLaunchKernelEx(&cfg, kern, p);
}

template <int KIND, bool SWIGLU>
static int launch_gemv_m(const GemvParams& p, cudaStream_t stream) {
  if (p.M == 1) return launch_gemv_t<KIND, 1, SWIGLU>(p, stream);
  if (p.M == 2) return launch_gemv_t<KIND, 2, SWIGLU>(p, stream);
  if (p.M <= 4) return launch_gemv_t<KIND, 4, SWIGLU>(p, stream);
  return -3;
}

}  // namespace tb

namespace tb { thread_local GemvNextHint g_gemv_next{}; }

using namespace tb;

// Hint (optional, one-shot, per calling thread): the next tb_gemv / tb_gemv_fused launch also requests these byte ranges —
// the head of the weights the FOLLOWING projection will stream — into L2 as its warps finish.  NULL / 0 clears it.
extern "C" int tb_gemv_hint_next(const void* a, size_t a_bytes, const void* b, size_t b_bytes) {
  g_gemv_next = GemvNextHint{{a, b}, {a ? a_bytes : 0, b ? b_bytes : 0}};
  return 0;
}

// rows the decode-shape path accepts for this weight kind and K: 8 on the tensor-core kernel, 4 on the FMA kernel
extern "C" int tb_gemv_max_rows(int kind, int K) { return gemv_mma_eligible(kind, 8, K) ? 8 : 4; }

// 1: tb_gemv / tb_gemv_fused run this problem on the tensor-core kernel (gemv_mma.cu), 0: on the FMA kernel of this file
extern "C" int tb_gemv_on_tensor_cores(int kind, int M, int K) {
  static const int mma_min_m_env = getenv("TB_GEMV_MMA_MIN_M") ? atoi(getenv("TB_GEMV_MMA_MIN_M")) : 0;   // A/B switch (5: FMA at <= 4 rows)
  const int mma_min_m = mma_min_m_env > 0 ? mma_min_m_env : (kind == kA8W8 ? 5 : 1);
  return M >= mma_min_m && gemv_mma_eligible(kind, M, K) ? 1 : 0;
}

extern "C" int tb_gemv_fused(int kind, void* y, float* y_f32, const void* x, const void* w, const void* w_scale,
                             const float* sc, const float* sr, int sc_per_channel, int sr_per_token, const void* residual,
                             int M, int N, int K, int swiglu, int prologue, const void* gamma, float eps,
                             cudaStream_t stream) {
  GemvParams p{};
  p.x = x; p.w = w; p.w_scale = (const __half*) w_scale; p.sc = sc; p.sr = sr;
  p.sc_per_channel = sc_per_channel; p.sr_per_token = sr_per_token; p.residual = (const __half*) residual;
  p.y = (__half*) y; p.y_f32 = y_f32; p.M = M; p.N = N; p.K = K; p.swiglu = swiglu;
  p.n_out = swiglu ? N / 2 : N;
  p.prologue = prologue; p.gamma = (const __half*) gamma; p.eps = eps;
  for (int i = 0; i < 2; ++i) {
    p.pf[i] = static_cast<const uint8_t*>(g_gemv_next.p[i]);
    p.pf_lines[i] = (unsigned) (g_gemv_next.bytes[i] / 128);
  }
  g_gemv_next = GemvNextHint{};
  const int epc = kind == kF16 ? 8 : (kind == kW4 ? 32 : 16);
  if (kind < 0 || kind > 3 || M < 1 || M > tb_gemv_max_rows(kind, K) || K % epc != 0 || (swiglu && (N & 1))) return -1;
  if ((kind == kW8 || kind == kW4) && !w_scale) return -1;
  if (kind == kA8W8 && (!sc || (prologue < kProRmsQuant && !sr))) return -1;
  if (prologue < 0 || prologue > 3) return -1;
  if ((prologue == kProRms || prologue == kProRmsQuant) && !gamma) return -1;
  if ((prologue >= kProRmsQuant) != (kind == kA8W8) && prologue != kProNone && prologue != kProRms) return -1;
  if (prologue == kProRms && kind == kA8W8) return -1;
  if (swiglu && residual) return -1;
  // Which kernel: measured on LLaMA-7B decode steps (B200, CUDA-graph replay, tools/mma_minm_ab.sh, same-run A/B).  The
  // tensor-core kernel (per-lane cp.async weight ring) takes every 5..8-row problem, and at 1..4 rows: weight-only int8 and
  // int4, which are conversion-bound on FMAs (W8: 1.87 vs 2.35 ms per step, int4: 1.66 vs 2.26), and fp16 (16.5 vs 18.6 us per
  // launch timed alone = 0.93 vs 0.83 of the HBM peak; 2.573 vs 2.592 ms per step once the attention kernel triggers its
  // dependents at entry — before that the one-CTA-per-SM kernel lost the step, 2.62 vs 2.60, to the ramp behind attention).
  // W8A8 at 1..4 rows stays on the FMA (dp4a) kernel: 1.657 vs 1.687 ms per step.
  if (tb_gemv_on_tensor_cores(kind, M, K)) {
    // the next-weights window measured slower on this kernel's workloads (cfg3 int8-KV 3.32 -> 3.40 ms, int4 2.10 -> 2.26 ms)
    static const bool mma_pf = getenv("TB_MMA_PF") && atoi(getenv("TB_MMA_PF")) != 0;
    if (!mma_pf) p.pf_lines[0] = p.pf_lines[1] = 0;
    return gemv_mma_launch(kind, y, y_f32, x, w, w_scale, sc, sr, sc_per_channel, sr_per_token, residual, M, N, K, swiglu,
                           prologue, gamma, eps, reinterpret_cast<const void* const*>(p.pf), p.pf_lines, stream);
  }
  switch (kind) {
    case kF16:  return swiglu ? launch_gemv_m<kF16, true>(p, stream)  : launch_gemv_m<kF16, false>(p, stream);
    case kW8:   return swiglu ? launch_gemv_m<kW8, true>(p, stream)   : launch_gemv_m<kW8, false>(p, stream);
    case kW4:   return swiglu ? launch_gemv_m<kW4, true>(p, stream)   : launch_gemv_m<kW4, false>(p, stream);
    case kA8W8: return swiglu ? launch_gemv_m<kA8W8, true>(p, stream) : launch_gemv_m<kA8W8, false>(p, stream);
  }
  return -1;
}

extern "C" int tb_gemv(int kind, void* y, float* y_f32, const void* x, const void* w, const void* w_scale,
                       const float* sc, const float* sr, int sc_per_channel, int sr_per_token, const void* residual,
                       int M, int N, int K, int swiglu, cudaStream_t stream) {
  return tb_gemv_fused(kind, y, y_f32, x, w, w_scale, sc, sr, sc_per_channel, sr_per_token, residual, M, N, K, swiglu,
                       kProNone, nullptr, 0.f, stream);
}
