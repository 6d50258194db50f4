// Decode-shape projection on tensor cores: Y[M,N] = X[M,K] . W[N,K]^T for M <= 8 token rows.
//
// Why: the FMA GEMV (gemv.cu) spends ~40 (fp16) to ~140 (int4) instructions per 16 weight bytes and is
// issue-bound long before HBM (ncu: profiles/r01_gemv_full.txt).  Here the streamed weight rows are the A operand
// of mma.sync m16n8k16 (fp16, fp32 accumulate) / m16n8k32 (s8, s32 accumulate) and the <= 8 token rows are the
// N = 8 operand, so a lane issues 2 HMMA per 32 weight bytes, products stay exact and accumulation stays fp32 / int32.
// (tcgen05 would need the weights in shared memory; for a stream that is read exactly once the registers are
// the shorter path — gemm_tc.cu is the tcgen05 kernel for M > 8.)
//
// Mapping: a CTA (8 warps) owns a tile of 16 weight rows (SwiGLU: 8 gate rows + the 8 matching up rows).  K is cut
// into 8 x S slabs: one per warp of each CTA of a thread-block cluster of S CTAs (S = 1, 2, 4 chosen by the host so
// that small-N projections still cover the chip).  In a k-step lane (g = lane / 4, t = lane % 4) loads 16 bytes of
// row g and 16 bytes of row g + 8 at byte offset 64 * step + 16 * t: every request is 64 contiguous bytes per row,
// consecutive steps continue the same rows.  The logical k order inside an MMA is a fixed permutation applied to
// both operands (a lane's 8 halves feed k-slots {2t,2t+1,2t+8,2t+9} of two MMAs), which a dot product allows.
// Partial accumulators are reduced in warp order through shared memory, then in CTA-rank order through
// distributed shared memory: deterministic, no atomics.
// Fused prologues / epilogues and programmatic dependent launch as in gemv.cu.
#include "common.cuh"
#include "kernels.h"

namespace tb {

enum { kMF16 = 0, kMW8 = 1, kMW4 = 2, kMA8W8 = 3 };
enum { kMProNone = 0, kMProRms = 1, kMProRmsQuant = 2, kMProQuant = 3 };

struct GemvMmaParams {
  const void* x;
  const void* w;
  const __half* w_scale;
  const float* sc;
  const float* sr;
  int sc_per_channel, sr_per_token;
  const __half* residual;
  __half* y;
  float* y_f32;
  int M, N, K, n_out, swiglu;
  int prologue;
  const __half* gamma;
  float eps;
  int S;   // cluster size (K split across CTAs)
  const uint8_t* pf[2];    // tb_gemv_hint_next: head of the next projection's weights, requested into L2 at the end (gemv.cu;
  unsigned pf_lines[2];    // off by default on this kernel: measured slower on its workloads)
};

constexpr int kMmaThreads = 256;
constexpr int kMmaWarps = 8;
constexpr int kMmaU = 8;   // k-steps in flight per lane (2 x 16-byte loads each)

__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f32(float* local_smem_ptr, uint32_t target_rank, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem_ptr)), "r"(target_rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}
__device__ __forceinline__ void mma_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void mma_pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ float mma_silu(float v) { return v / (1.f + __expf(-v)); }

// 16-byte activation load through the generic address space (shared when a prologue staged x, else global / L2)
__device__ __forceinline__ uint4 ld_x16(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }

template <int KIND> struct MmaTraits;
template <> struct MmaTraits<kMF16>  { static constexpr int kStepElems = 32,  kXBytesPerLane = 16; };
template <> struct MmaTraits<kMW8>   { static constexpr int kStepElems = 64,  kXBytesPerLane = 32; };
template <> struct MmaTraits<kMW4>   { static constexpr int kStepElems = 128, kXBytesPerLane = 64; };
template <> struct MmaTraits<kMA8W8> { static constexpr int kStepElems = 64,  kXBytesPerLane = 16; };

__device__ __forceinline__ float cta_reduce_mma(float v, float* red, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kMmaWarps; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

template <int KIND, bool SWIGLU>
__global__ void __launch_bounds__(kMmaThreads, 2) gemv_mma_kernel(const GemvMmaParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  using TR = MmaTraits<KIND>;
  constexpr bool INT = KIND == kMA8W8;
  constexpr int XB = INT ? 1 : 2;
  float* red = reinterpret_cast<float*>(smem);             // [8]
  float* srow = red + 8;                                    // [8] per-token scales (W8A8)
  float* wpart = srow + 8;                                  // [8 warps][16][8] per-warp partial tiles
  float* cpart = wpart + kMmaWarps * 128;                   // [2 parities][4 ranks][16][8] per-CTA partial tiles (cluster reduce)
  uint8_t* xs = reinterpret_cast<uint8_t*>(cpart + 2 * 4 * 128);   // [M][K * XB] only when a prologue transforms x

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int K = p.K, M = p.M;
  const uint32_t crank = p.S > 1 ? cluster_rank() : 0;
  const int row_bytes = K / TR::kStepElems * 64;
  const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w);
  const int tiles = SWIGLU ? (p.n_out + 7) / 8 : (p.N + 15) / 16;
  const int rows_real = SWIGLU ? 2 * p.n_out : p.N;
  // this warp's K slab
  const int ksteps = K / TR::kStepElems;
  const int slabs = kMmaWarps * p.S, slab = crank * kMmaWarps + warp;
  const int ks0 = (int) ((long long) ksteps * slab / slabs), ks1 = (int) ((long long) ksteps * (slab + 1) / slabs);

  mma_pdl_wait();
  mma_pdl_launch();

  // ---- activations: plain -> read in place (global/L2); prologue -> transformed copy in shared memory -------------
  const uint8_t* xsrc = reinterpret_cast<const uint8_t*>(p.x);
  const int xstride = K * XB;
  const int tid = threadIdx.x;
  if (p.prologue == kMProNone) {
    if (INT && tid < 8) srow[tid] = tid < M ? p.sr[p.sr_per_token ? tid : 0] : 0.f;
    if (INT) __syncthreads();
  } else {
    // one warp per token row (M <= 8 rows, 8 warps): every row's statistics are warp-local reductions, so the rows
    // proceed in parallel and the whole prologue costs two L2 round trips instead of 2 x M block-wide ones
    const __half* xin = reinterpret_cast<const __half*>(p.x);
    // (a register-resident row for M <= 4 was measured and removed: int4 B=1 step 2.18 -> 2.03 ms without it; requesting the
    // first weight batch above griddepcontrol.wait, or every tile's first batch one tile ahead, measured slower at step
    // level: cfg3 int8-KV 3.15 -> 3.16 / 3.34 ms, tools/mma_ab.sh)
    for (int m = warp; m < M; m += kMmaWarps) {
      const __half* xr = xin + (size_t) m * K;
      float inv = 1.f;
      if (p.prologue != kMProQuant) {
        float sq = 0.f;
        for (int i = lane * 8; i < K; i += 32 * 8) {
          uint4 raw = *reinterpret_cast<const uint4*>(xr + i);
          const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(h[j]);
            sq += f.x * f.x + f.y * f.y;
          }
        }
        sq = warp_sum(sq);
        inv = rsqrtf(sq / K + p.eps);
      }
      float amax = 0.f;
      for (int i = lane * 8; i < K; i += 32 * 8) {
        uint4 raw = *reinterpret_cast<const uint4*>(xr + i);
        __half2* h = reinterpret_cast<__half2*>(&raw);
        if (p.prologue != kMProQuant) {
          uint4 g4 = *reinterpret_cast<const uint4*>(p.gamma + i);
          const __half2* gm = reinterpret_cast<const __half2*>(&g4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(h[j]), gg = __half22float2(gm[j]);
            h[j] = __floats2half2_rn(f.x * inv * gg.x, f.y * inv * gg.y);
          }
        }
        if constexpr (INT) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(h[j]);
            amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
          }
        } else {
          *reinterpret_cast<uint4*>(xs + (size_t) m * xstride + (size_t) i * 2) = raw;
        }
      }
      if constexpr (INT) {
        amax = fmaxf(warp_max(amax), __half2float(__float2half_rn(1e-6f)));
        const float qs = 127.f / amax;
        if (lane == 0) srow[m] = amax / 127.f;
        for (int i = lane * 8; i < K; i += 32 * 8) {
          uint4 raw = *reinterpret_cast<const uint4*>(xr + i);
          __half2* h = reinterpret_cast<__half2*>(&raw);
          float f[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 v = __half22float2(h[j]);
            if (p.prologue != kMProQuant) {
              float2 gg = __half22float2(reinterpret_cast<const __half2*>(p.gamma + i)[j]);
              v = __half22float2(__floats2half2_rn(v.x * inv * gg.x, v.y * inv * gg.y));
            }
            f[2 * j] = v.x * qs;
            f[2 * j + 1] = v.y * qs;
          }
          uint2 o;
          o.x = pack4_i8(f[0], f[1], f[2], f[3]);
          o.y = pack4_i8(f[4], f[5], f[6], f[7]);
          *reinterpret_cast<uint2*>(xs + (size_t) m * xstride + i) = o;
        }
      }
    }
    __syncthreads();
    xsrc = xs;
  }

  // ---- persistent loop over 16-row tiles: the prologue above is paid once per CTA --------------------------------------
  const bool x_ok = g < M;                                   // fragment column g = token g
  const uint8_t* xrow = xsrc + (size_t) (x_ok ? g : 0) * xstride + (size_t) t * TR::kXBytesPerLane;
  constexpr int XSTEP = TR::kStepElems * XB;                 // activation bytes per k-step per token
  const int nclusters = gridDim.x / p.S;
  int parity = 0;
  for (int tile = blockIdx.x / p.S; tile < tiles; tile += nclusters, parity ^= 1) {
    // rows of this tile: fragment row g and g + 8
    const int r_lo = SWIGLU ? tile * 8 + g : tile * 16 + g;
    const int r_hi = SWIGLU ? p.n_out + tile * 8 + g : tile * 16 + 8 + g;
    const bool lo_ok = SWIGLU ? (tile * 8 + g < p.n_out) : (r_lo < rows_real);
    const bool hi_ok = SWIGLU ? lo_ok : (r_hi < rows_real);
    const uint8_t* p_lo = wbase + (size_t) (lo_ok ? r_lo : 0) * row_bytes + t * 16;
    const uint8_t* p_hi = wbase + (size_t) (hi_ok ? r_hi : 0) * row_bytes + t * 16;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int iacc[4] = {0, 0, 0, 0};

    auto consume = [&](const uint4& lo, const uint4& hi, int ks) {
      const uint8_t* xp = xrow + (size_t) ks * XSTEP;
      if constexpr (KIND == kMF16) {
        uint4 xv = x_ok ? ld_x16(xp) : make_uint4(0, 0, 0, 0);
        mma_f16(acc, lo.x, hi.x, lo.y, hi.y, xv.x, xv.y);
        mma_f16(acc, lo.z, hi.z, lo.w, hi.w, xv.z, xv.w);
      } else if constexpr (KIND == kMA8W8) {
        uint4 xv = x_ok ? ld_x16(xp) : make_uint4(0, 0, 0, 0);
        mma_s8(iacc, lo.x, hi.x, lo.y, hi.y, xv.x, xv.y);
        mma_s8(iacc, lo.z, hi.z, lo.w, hi.w, xv.z, xv.w);
      } else if constexpr (KIND == kMW8) {
        uint4 xa = make_uint4(0, 0, 0, 0), xb = xa;
        if (x_ok) { xa = ld_x16(xp); xb = ld_x16(xp + 16); }
        const uint32_t wl[4] = {lo.x, lo.y, lo.z, lo.w}, wh[4] = {hi.x, hi.y, hi.z, hi.w};
        const uint32_t xw[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {       // 4 int8 of each row -> one MMA
          __half2 l0, l1, h0, h1;
          i8x4_to_h2x2(wl[j], l0, l1);
          i8x4_to_h2x2(wh[j], h0, h1);
          mma_f16(acc, h2u(l0), h2u(h0), h2u(l1), h2u(h1), xw[2 * j], xw[2 * j + 1]);
        }
      } else {  // kMW4
        const uint32_t wl[4] = {lo.x, lo.y, lo.z, lo.w}, wh[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {       // 8 int4 of each row -> two MMAs
          __half2 l[4], h[4];
          i4x8_to_h2x4(wl[j], l);
          i4x8_to_h2x4(wh[j], h);
          uint4 xv = x_ok ? ld_x16(xp + 16 * j) : make_uint4(0, 0, 0, 0);
          mma_f16(acc, h2u(l[0]), h2u(h[0]), h2u(l[1]), h2u(h[1]), xv.x, xv.y);
          mma_f16(acc, h2u(l[2]), h2u(h[2]), h2u(l[3]), h2u(h[3]), xv.z, xv.w);
        }
      }
    };

    int ks = ks0;
    for (; ks + kMmaU <= ks1; ks += kMmaU) {
      uint4 lo[kMmaU], hi[kMmaU];
#pragma unroll
      for (int u = 0; u < kMmaU; ++u) {
        lo[u] = ldg_nc_v4(p_lo + (size_t) (ks + u) * 64);
        hi[u] = ldg_nc_v4(p_hi + (size_t) (ks + u) * 64);
      }
#pragma unroll
      for (int u = 0; u < kMmaU; ++u) consume(lo[u], hi[u], ks + u);
    }
    if (ks < ks1) {
      // tail (< kMmaU steps; the whole slab for int4 at K = 4096): still ONE batch of requests, predicated warp-uniformly —
      // a step-at-a-time tail cost one full memory round trip per step
      uint4 lo[kMmaU], hi[kMmaU];
#pragma unroll
      for (int u = 0; u < kMmaU; ++u) {
        if (ks + u < ks1) {
          lo[u] = ldg_nc_v4(p_lo + (size_t) (ks + u) * 64);
          hi[u] = ldg_nc_v4(p_hi + (size_t) (ks + u) * 64);
        }
      }
#pragma unroll
      for (int u = 0; u < kMmaU; ++u) {
        if (ks + u < ks1) consume(lo[u], hi[u], ks + u);
      }
    }

    // ---- reduce: warps (shared memory, warp order) then cluster ranks (DSMEM, rank order) -----------------------------
    // fragment: c0,c1 = (row g, tokens 2t, 2t+1), c2,c3 = (row g + 8, tokens 2t, 2t+1)
    {
      float* wp = wpart + warp * 128;
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = INT ? __int_as_float(iacc[i]) : acc[i];
      wp[g * 8 + 2 * t] = v[0];
      wp[g * 8 + 2 * t + 1] = v[1];
      wp[(g + 8) * 8 + 2 * t] = v[2];
      wp[(g + 8) * 8 + 2 * t + 1] = v[3];
    }
    __syncthreads();
    float tot = 0.f;
    int itot = 0;
    if (tid < 128) {
#pragma unroll
      for (int w = 0; w < kMmaWarps; ++w) {
        if constexpr (INT) itot += __float_as_int(wpart[w * 128 + tid]); else tot += wpart[w * 128 + tid];
      }
    }
    bool finisher = true;                                  // this CTA applies the epilogue for the tile
    if (p.S > 1) {
      float* cp = cpart + parity * (4 * 128);              // double-buffered by tile parity: one cluster barrier per tile
      if (tid < 128) st_cluster_f32(cp + crank * 128 + tid, 0, INT ? __int_as_float(itot) : tot);
      cluster_sync_all();
      finisher = crank == 0;
      if (finisher && tid < 128) {
        tot = 0.f;
        itot = 0;
        for (int c = 0; c < p.S; ++c) {
          if constexpr (INT) itot += __float_as_int(cp[c * 128 + tid]); else tot += cp[c * 128 + tid];
        }
      }
    }

    // ---- epilogue: thread = (fragment row r, token m) -------------------------------------------------------------------
    if (finisher && tid < 128) {
      const int r = tid >> 3, m = tid & 7;
      float v = INT ? (float) itot : tot;
      const int wrow = SWIGLU ? (r < 8 ? tile * 8 + r : p.n_out + tile * 8 + (r - 8)) : tile * 16 + r;
      const bool row_ok = SWIGLU ? (tile * 8 + (r & 7) < p.n_out) : (wrow < p.N);
      if (row_ok) {
        if constexpr (KIND == kMW8 || KIND == kMW4) v *= __half2float(p.w_scale[wrow]);
        // reference grouping: accum * (scale_col * scale_row)  (epilogue_per_row_per_col_scale.h:325,341)
        if constexpr (INT) v = v * (p.sc[p.sc_per_channel ? wrow : 0] * srow[m]);
      }
      if constexpr (SWIGLU) {
        // rows 0-7 hold gate, rows 8-15 the matching up projection: exchange through shared memory (slot 0 of wpart is
        // only read by the thread that now overwrites it)
        float* ex = wpart;
        ex[tid] = v;
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (r < 8 && row_ok && m < M) {
          const float gte = __half2float(__float2half_rn(ex[r * 8 + m])), up = __half2float(__float2half_rn(ex[(r + 8) * 8 + m]));
          const float o = __half2float(__float2half_rn(mma_silu(gte))) * up;
          const size_t oi = (size_t) m * p.n_out + tile * 8 + r;
          if (p.y_f32) p.y_f32[oi] = o; else p.y[oi] = __float2half_rn(o);
        }
      } else if (row_ok && m < M) {
        const size_t oi = (size_t) m * p.n_out + wrow;
        if (p.y_f32) {
          p.y_f32[oi] = v;
        } else {
          __half oh = __float2half_rn(v);
          if (p.residual) oh = __float2half_rn(__half2float(oh) + __half2float(p.residual[oi]));
          p.y[oi] = oh;
        }
      }
    }
    __syncthreads();   // wpart is rewritten by the next tile
  }
  {
    const unsigned total = p.pf_lines[0] + p.pf_lines[1];
    const unsigned gw = blockIdx.x * kMmaWarps + warp, tw = gridDim.x * kMmaWarps;
    for (unsigned l = gw * 32 + lane; l < total; l += tw * 32) {
      const uint8_t* a = l < p.pf_lines[0] ? p.pf[0] + (size_t) l * 128 : p.pf[1] + (size_t) (l - p.pf_lines[0]) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
  }
}

template <int KIND, bool SWIGLU>
static int launch_gemv_mma(GemvMmaParams p, cudaStream_t stream) {
  const size_t xs_bytes = p.prologue ? (size_t) p.M * p.K * (KIND == kMA8W8 ? 1 : 2) : 0;
  const size_t smem = (16 + kMmaWarps * 128 + 2 * 4 * 128) * sizeof(float) + xs_bytes;
  if (smem > 200 * 1024) return -2;
  auto kern = gemv_mma_kernel<KIND, SWIGLU>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int) e;
    attr_done = true;
  }
  const int tiles = SWIGLU ? (p.n_out + 7) / 8 : (p.N + 15) / 16;
  const int ksteps = p.K / MmaTraits<KIND>::kStepElems;
  int S = 1;
  while (S < 4 && tiles * S < 3 * kNumSMs && ksteps / (kMmaWarps * S * 2) >= 1) S *= 2;
  p.S = S;
  cudaLaunchConfig_t cfg{};
  // persistent: at most three resident CTAs per SM (register budget), whole clusters
  int clusters = tiles;
  int per_sm = (int) ((220 * 1024) / (smem + 1024));
  per_sm = per_sm > 2 ? 2 : (per_sm < 1 ? 1 : per_sm);
  const int max_clusters = (per_sm * kNumSMs) / S;
  if (clusters > max_clusters) clusters = max_clusters;
  cfg.gridDim = dim3(clusters * S);
  cfg.blockDim = dim3(kMmaThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (S > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = S;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return (int) cudaLaunchKernelEx(&cfg, kern, p);
}

// true when the tensor-core GEMV handles this problem (else gemv.cu's FMA kernel, M <= 4)
bool gemv_mma_eligible(int kind, int M, int K) {
  if (M < 1 || M > 8) return false;
  const int step = kind == kMF16 ? 32 : (kind == kMW4 ? 128 : 64);
  return K % step == 0 && K / step >= kMmaWarps;
}

int gemv_mma_launch(int kind, void* y, float* y_f32, const void* x, const void* w, const void* w_scale, const float* sc,
                    const float* sr, int sc_per_channel, int sr_per_token, const void* residual, int M, int N, int K,
                    int swiglu, int prologue, const void* gamma, float eps, const void* const* pf, const unsigned* pf_lines,
                    cudaStream_t stream) {
  GemvMmaParams p{};
  for (int i = 0; i < 2; ++i) { p.pf[i] = static_cast<const uint8_t*>(pf[i]); p.pf_lines[i] = pf_lines[i]; }
  p.x = x; p.w = w; p.w_scale = (const __half*) w_scale; p.sc = sc; p.sr = sr;
  p.sc_per_channel = sc_per_channel; p.sr_per_token = sr_per_token; p.residual = (const __half*) residual;
  p.y = (__half*) y; p.y_f32 = y_f32; p.M = M; p.N = N; p.K = K; p.swiglu = swiglu; p.n_out = swiglu ? N / 2 : N;
  p.prologue = prologue; p.gamma = (const __half*) gamma; p.eps = eps;
  switch (kind) {
    case kMF16:  return swiglu ? launch_gemv_mma<kMF16, true>(p, stream)  : launch_gemv_mma<kMF16, false>(p, stream);
    case kMW8:   return swiglu ? launch_gemv_mma<kMW8, true>(p, stream)   : launch_gemv_mma<kMW8, false>(p, stream);
    case kMW4:   return swiglu ? launch_gemv_mma<kMW4, true>(p, stream)   : launch_gemv_mma<kMW4, false>(p, stream);
    case kMA8W8: return swiglu ? launch_gemv_mma<kMA8W8, true>(p, stream) : launch_gemv_mma<kMA8W8, false>(p, stream);
  }
  return -1;
}

}  // namespace tb
