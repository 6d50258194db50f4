// Decode-shape projection on tensor cores: Y[M,N] = X[M,K] . W[N,K]^T for M <= 8 token rows.
//
// Why: the FMA GEMV (gemv.cu) spends ~40 (fp16) to ~140 (int4) instructions per 16 weight bytes and is
// issue-bound long before HBM (ncu: profiles/r01_gemv_full.txt).  Here the streamed weight rows are the A operand
// of mma.sync m16n8k16 (fp16, fp32 accumulate) / m16n8k32 (s8, s32 accumulate) and the <= 8 token rows are the
// N = 8 operand, so a lane issues 2 HMMA per 32 weight bytes, products stay exact and accumulation stays fp32 / int32.
// (tcgen05 would need the weights in shared memory; for a stream that is read exactly once the registers are
// the shorter path — gemm_tc.cu is the tcgen05 kernel for M > 8.)
//
// Mapping: a CTA (16 warps, one CTA per SM, persistent) owns a tile of 16 weight rows at a time (SwiGLU: 8 gate rows + the
// 8 matching up rows).  K is cut into 16 x S slabs: one per warp of each CTA of a thread-block cluster of S CTAs (S > 1 only
// when there are fewer tiles than SMs, or when the staged activations of the whole K would not fit beside the weight ring).
// In a k-step lane (g = lane / 4, t = lane % 4) owns 16 bytes of row g and 16 bytes of row g + 8 at byte offset
// 64 * step + 16 * t: every request is 64 contiguous bytes per row, consecutive steps continue the same rows.  The logical
// k order inside an MMA is a fixed permutation applied to both operands (a lane's 8 halves feed k-slots {2t,2t+1,2t+8,2t+9}
// of two MMAs), which a dot product allows.
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
  int xrs; // bytes per staged activation row in shared memory
  const uint8_t* pf[2];    // tb_gemv_hint_next: head of the next projection's weights, requested into L2 at the end (gemv.cu;
  unsigned pf_lines[2];    // off by default on this kernel: measured slower on its workloads)
};

constexpr int kMmaThreads = 512;
constexpr int kMmaWarps = 16;
constexpr int kMmaMinSteps = 8;    // eligibility: at least this many k-steps in K (warps without a k-step just contribute zeros)
constexpr int kMmaC = 2;           // k-steps per ring chunk: 2 x (2 rows x 16 bytes) per lane = 2 KB per warp
constexpr int kMmaChunk = kMmaC * 1024;
constexpr int kMmaRing = 4;        // ring chunks per warp (3 in flight while one is consumed): 16 warps x 8 KB = 128 KB per SM
constexpr int kMmaRingSmall = 2;   // when the staged activations leave no room for the deep ring

__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
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

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Weight stream: every lane owns a private strip of a per-warp shared-memory ring and fills it with 16-byte cp.async
// copies (no registers held by bytes in flight), R - 1 chunks of 2 k-steps ahead of the chunk it is consuming.  The
// sequence a warp walks is (its tiles) x (its K slab) flattened, so the requests for the next tile — and, at kernel
// entry, the first tiles' requests, issued ABOVE griddepcontrol.wait and the activation prologue (weights do not depend
// on the previous kernel) — are in flight while the CTA reduces and writes the current one.  A lane only ever reads
// bytes it copied itself: cp.async.wait_group is the only synchronisation the ring needs.
// (The previous form of this kernel loaded a tile's slab into registers, consumed it, reduced, and only then requested the
// next tile: 2.1-3.5 TB/s; ncu of this form: profiles/r02_gemv_mma_*.)
// Measured on top of this form and not kept (same-run A/B at step level, tools/mma_ab.sh): CTA pairs sharing the norm
// prologue through distributed shared memory, each normalising half of the token rows and storing into both (cfg3 int8-KV
// 3.00 vs 3.00 ms, int4 B=8 2.17 vs 2.18: two cluster barriers cost what the halved work saves); RMSNorm as its own
// PDL-chained kernel from 5 rows on (2.934 vs 2.941 ms, kept as TB_FUSE_NORM_ROWS); the next-projection L2 window
// (TB_MMA_PF=1: 3.05 vs 3.17 ms).
//
// Activations: always staged in shared memory (the K range of this CTA only), transformed by the fused prologue or copied
// as they are, in a layout whose 16-byte chunks are permuted so that the B-fragment loads of a quarter-warp (4 k-groups
// x 2 token rows) fall into 8 different bank groups:
//   chunk c of token row m is stored at chunk c ^ ((m & 1) * RX) ^ (int4 weights: ((c >> 3) & 1) << 1),
//   RX = 4 when a lane reads 16 bytes per k-step (fp16 / int8 activations), 1 when it reads 32 or 64 (W8 / W4).
// (Before: read in place through L1 when no prologue ran — 176 KB for the down projection at 8 rows, more than L1 holds
// next to the ring — and 2- to 4-way bank conflicts on the staged copy.)
template <int KIND> struct XSwz {
  static constexpr int RX = (KIND == kMW8 || KIND == kMW4) ? 1 : 4;
  __device__ static __forceinline__ uint32_t chunk(uint32_t c, int row) {
    uint32_t r = c ^ (uint32_t) ((row & 1) * RX);
    if constexpr (KIND == kMW4) r ^= ((c >> 3) & 1u) << 1;
    return r;
  }
};

template <int KIND, bool SWIGLU, int R>
__global__ void __launch_bounds__(kMmaThreads, 1) gemv_mma_kernel(const GemvMmaParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  using TR = MmaTraits<KIND>;
  using SW = XSwz<KIND>;
  constexpr bool INT = KIND == kMA8W8;
  constexpr int XB = INT ? 1 : 2;
  constexpr int XSTEP = TR::kStepElems * XB;                // activation bytes per k-step per token
  float* srow = reinterpret_cast<float*>(smem);             // [8] per-token scales (W8A8)
  float* rinv = srow + 8;                                   // [8] per-row 1/rms, then 127/amax
  float* red2 = rinv + 8;                                   // [16 warps][8] prologue reductions
  float* ex = red2 + kMmaWarps * 8;                         // [128] SwiGLU gate/up exchange
  float* zs = ex + 128;                                     // [64] zeros: the "activations" of fragment columns >= M
  float* wpart = zs + 64;                                   // [2 parities][16 warps][16][8] per-warp partial tiles
  float* cpart = wpart + 2 * kMmaWarps * 128;               // [2 parities][4 ranks][16][8] per-CTA partial tiles (cluster reduce)
  uint8_t* ring = reinterpret_cast<uint8_t*>(cpart + 2 * 4 * 128);     // [16 warps][R][2 KB]
  uint8_t* xs = ring + kMmaWarps * R * kMmaChunk;           // [M][xrs] staged activations of k-steps [kb, ke)

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int K = p.K, M = p.M;
  const uint32_t crank = p.S > 1 ? cluster_rank() : 0;
  const int row_bytes = K / TR::kStepElems * 64;
  const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w);
  const int tiles = SWIGLU ? (p.n_out + 7) / 8 : (p.N + 15) / 16;
  const int rows_real = SWIGLU ? 2 * p.n_out : p.N;
  // K slabs: one per warp of every CTA of the cluster; this CTA stages the activations of k-steps [kb, ke)
  const int ksteps = K / TR::kStepElems;
  const int slabs = kMmaWarps * p.S, slab = crank * kMmaWarps + warp;
  const int ks0 = (int) ((long long) ksteps * slab / slabs), ks1 = (int) ((long long) ksteps * (slab + 1) / slabs);
  const int kb = (int) ((long long) ksteps * (crank * kMmaWarps) / slabs);
  const int ke = (int) ((long long) ksteps * ((crank + 1) * kMmaWarps) / slabs);
  const int xrs = p.xrs;                                    // bytes per staged token row (multiple of 128)
  const int nks = ks1 - ks0;
  const int cpt = (nks + kMmaC - 1) / kMmaC;                 // ring chunks per tile of this warp (0: nothing to stream)
  const int nclusters = gridDim.x / p.S;
  const int tile0 = blockIdx.x / p.S;
  const int my_tiles = tile0 < tiles ? (tiles - tile0 + nclusters - 1) / nclusters : 0;
  const int total = my_tiles * cpt;

  // rows of a tile held by this lane: fragment rows g and g + 8 (rows past the matrix are read as row 0 and masked later)
  auto row_ptrs = [&](int tile, const uint8_t*& lo, const uint8_t*& hi) {
    const int r_lo = SWIGLU ? tile * 8 + g : tile * 16 + g;
    const int r_hi = SWIGLU ? p.n_out + tile * 8 + g : tile * 16 + 8 + g;
    const bool lo_ok = SWIGLU ? (tile * 8 + g < p.n_out) : (r_lo < rows_real);
    const bool hi_ok = SWIGLU ? lo_ok : (r_hi < rows_real);
    lo = wbase + (size_t) (lo_ok ? r_lo : 0) * row_bytes + t * 16 + (size_t) ks0 * 64;
    hi = wbase + (size_t) (hi_ok ? r_hi : 0) * row_bytes + t * 16 + (size_t) ks0 * 64;
  };
  const uint8_t* myring_p = ring + (size_t) warp * (R * kMmaChunk) + (size_t) lane * 16;
  const uint32_t myring = smem_u32(myring_p);
  int iq = 0, ic = 0, itile = tile0, islot = 0;
  const uint8_t *ilo, *ihi;                                  // running request pointers (advance 128 bytes per chunk)
  row_ptrs(itile, ilo, ihi);
  auto issue = [&]() {                                       // request the next chunk of the flat sequence; always one group
    if (iq < total) {
      const uint32_t dst = myring + (uint32_t) islot * kMmaChunk;
      const int left = nks - ic * kMmaC;
#pragma unroll
      for (int u = 0; u < kMmaC; ++u) {
        if (u < left) {
          cp_async16(dst + (2 * u) * 512, ilo + u * 64);
          cp_async16(dst + (2 * u + 1) * 512, ihi + u * 64);
        }
      }
      ilo += kMmaC * 64;
      ihi += kMmaC * 64;
      islot = islot + 1 == R ? 0 : islot + 1;
      ++iq;
      if (++ic == cpt) {
        ic = 0;
        itile += nclusters;
        row_ptrs(itile, ilo, ihi);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < R - 1; ++i) issue();

  mma_pdl_wait();
  mma_pdl_launch();

  // ---- activations -> shared memory (k-steps [kb, ke) of every token row), chunk-permuted ------------------------------
  const int e0 = kb * TR::kStepElems, e1 = ke * TR::kStepElems;      // element range this CTA stages
  auto xs_at = [&](int m, int i) -> uint8_t* {               // address of element i (e0 <= i < e1, 16-byte piece aligned) of row m
    const uint32_t b = (uint32_t) (i - e0) * XB;
    return xs + (size_t) m * xrs + ((size_t) SW::chunk(b >> 4, m) << 4) + (b & 15u);
  };
  if (tid < 64) zs[tid] = 0.f;
  if (p.prologue == kMProNone) {
    if (INT && tid < 8) srow[tid] = tid < M ? p.sr[p.sr_per_token ? tid : 0] : 0.f;
    const uint8_t* xg = reinterpret_cast<const uint8_t*>(p.x);
    const int ppr = (e1 - e0) * XB / 16;                     // 16-byte pieces per row
    for (int m = 0; m < M; ++m) {
      const uint8_t* src = xg + ((size_t) m * K + e0) * XB;
      const uint32_t dst = smem_u32(xs + (size_t) m * xrs);
      for (int c = tid; c < ppr; c += kMmaThreads) cp_async16(dst + (SW::chunk((uint32_t) c, m) << 4), src + (size_t) c * 16);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  } else {
    const __half* xin = reinterpret_cast<const __half*>(p.x);
    const int iters = (K + 4095) >> 12;                      // 512 threads x 8 elements per pass over a row
    if (M * iters <= 8) {
      // whole CTA, every row at once: each thread holds <= 8 16-byte pieces (piece s = row s / iters, pass s % iters); the
      // statistics are one warp reduction + one 16 x 8 table in shared memory per pass — one L2 round trip for the lot
      uint4 raw[8];
      const int npieces = M * iters;                         // pieces s >= npieces do not exist: skipped warp-uniformly
      int pm[8], pi[8];                                      // piece -> (row, first element of this thread), no division
      {
        int m = 0, it = 0;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          pm[s] = m;
          pi[s] = tid * 8 + (it << 12);
          if (++it == iters) { it = 0; ++m; }
        }
      }
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        raw[s] = make_uint4(0, 0, 0, 0);
        if (s < npieces && pi[s] < K) raw[s] = *reinterpret_cast<const uint4*>(xin + (size_t) pm[s] * K + pi[s]);
      }
      if (p.prologue != kMProQuant) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s >= npieces) break;
          const __half2* h = reinterpret_cast<const __half2*>(&raw[s]);
          float sq = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(h[j]);
            sq += f.x * f.x + f.y * f.y;
          }
          sq = warp_sum(sq);
          if (lane == 0) red2[warp * 8 + s] = sq;
        }
        __syncthreads();
        if (tid < M) {
          float tot = 0.f;
          for (int s = tid * iters; s < (tid + 1) * iters && s < 8; ++s)
            for (int w = 0; w < kMmaWarps; ++w) tot += red2[w * 8 + s];
          rinv[tid] = rsqrtf(tot / K + p.eps);
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int m = pm[s], i = pi[s];
          if (s < npieces && i < K) {
            const float inv = rinv[m];
            __half2* h = reinterpret_cast<__half2*>(&raw[s]);
            uint4 g4 = *reinterpret_cast<const uint4*>(p.gamma + i);
            const __half2* gm = reinterpret_cast<const __half2*>(&g4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f = __half22float2(h[j]), gg = __half22float2(gm[j]);
              h[j] = __floats2half2_rn(f.x * inv * gg.x, f.y * inv * gg.y);
            }
          }
        }
      }
      if constexpr (INT) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s >= npieces) break;
          const __half2* h = reinterpret_cast<const __half2*>(&raw[s]);
          float amax = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(h[j]);
            amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
          }
          amax = warp_max(amax);
          if (lane == 0) red2[warp * 8 + s] = amax;
        }
        __syncthreads();
        if (tid < M) {
          float amax = 0.f;
          for (int s = tid * iters; s < (tid + 1) * iters && s < 8; ++s)
            for (int w = 0; w < kMmaWarps; ++w) amax = fmaxf(amax, red2[w * 8 + s]);
          amax = fmaxf(amax, __half2float(__float2half_rn(1e-6f)));
          srow[tid] = amax / 127.f;
          rinv[tid] = 127.f / amax;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int m = pm[s], i = pi[s];
          if (s < npieces && i >= e0 && i < e1) {
            const float qs = rinv[m];
            const __half2* h = reinterpret_cast<const __half2*>(&raw[s]);
            float f[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 v = __half22float2(h[j]);
              f[2 * j] = v.x * qs;
              f[2 * j + 1] = v.y * qs;
            }
            uint2 o;
            o.x = pack4_i8(f[0], f[1], f[2], f[3]);
            o.y = pack4_i8(f[4], f[5], f[6], f[7]);
            *reinterpret_cast<uint2*>(xs_at(m, i)) = o;
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < npieces && pi[s] >= e0 && pi[s] < e1) *reinterpret_cast<uint4*>(xs_at(pm[s], pi[s])) = raw[s];
        }
      }
    } else {
      // long rows: one warp per token row, statistics are warp-local reductions
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
            if (i >= e0 && i < e1) *reinterpret_cast<uint4*>(xs_at(m, i)) = raw;
          }
        }
        if constexpr (INT) {
          amax = fmaxf(warp_max(amax), __half2float(__float2half_rn(1e-6f)));
          const float qs = 127.f / amax;
          if (lane == 0) srow[m] = amax / 127.f;
          for (int i = e0 + lane * 8; i < e1; i += 32 * 8) {
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
            *reinterpret_cast<uint2*>(xs_at(m, i)) = o;
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- persistent loop over 16-row tiles ---------------------------------------------------------------------------------
  // fragment column g = token g; columns >= M read a block of zeros at stride 0.  Within a k-step lane t reads the 16-byte
  // chunks t * CPL + j (CPL = 1, 2, 4); with the permutation above their positions inside the k-step are lane constants.
  const bool x_ok = g < M;
  constexpr int CPL = TR::kXBytesPerLane / 16;
  constexpr int CPS = XSTEP / 16;                            // chunks per k-step per token: 4 (fp16 w, int8 x), 8 (W8), 16 (W4)
  uint32_t xoff[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const uint32_t c = (uint32_t) (t * CPL + j);             // chunk inside the k-step
    if constexpr (CPS >= 8) xoff[j] = x_ok ? (SW::chunk(c, g) << 4) : 0u;      // k-step = whole 128-byte groups: lane constant
    else xoff[j] = x_ok ? (c << 4) : 0u;                     // CPS = 4: the row term flips the k-step parity bit (below)
  }
  const uint8_t* xrow = x_ok ? xs + (size_t) g * xrs : reinterpret_cast<const uint8_t*>(zs);
  const int xstep = x_ok ? XSTEP : 0;
  const int kflip = (CPS == 4 && x_ok) ? (g & 1) : 0;        // CPS = 4: chunk ^ 4 == k-step ^ 1 inside a 128-byte group
  int parity = 0, cslot = 0, fin = 0;
  for (int tile = tile0; tile < tiles; tile += nclusters, parity ^= 1, fin = (fin + 1) & 3) {
    // the 4 warps fin * 4 .. fin * 4 + 3 finish this tile (the role rotates so that no warp is always the late one); their
    // epilogue operands are requested now, a whole tile ahead of their use
    const int etid = tid - fin * 128;                        // 0 .. 127 in the finishing warps
    const bool fin_warp = etid >= 0 && etid < 128;
    const int er = etid >> 3, em = etid & 7;
    int wrow = 0;
    bool row_ok = false;
    float e_wscale = 1.f, e_sc = 1.f, e_res = 0.f;
    if (fin_warp) {
      wrow = SWIGLU ? (er < 8 ? tile * 8 + er : p.n_out + tile * 8 + (er - 8)) : tile * 16 + er;
      row_ok = SWIGLU ? (tile * 8 + (er & 7) < p.n_out) : (wrow < p.N);
      if (row_ok) {
        if constexpr (KIND == kMW8 || KIND == kMW4) e_wscale = __half2float(p.w_scale[wrow]);
        if constexpr (INT) e_sc = p.sc[p.sc_per_channel ? wrow : 0];
        if (!SWIGLU && p.residual && !p.y_f32 && em < M) e_res = __half2float(p.residual[(size_t) em * p.n_out + wrow]);
      }
    }
    // two accumulator sets (k-steps alternate): two independent MMA dependency chains per warp
    float acc2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    int iacc2[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};

    auto consume = [&](const uint4& lo, const uint4& hi, int ks, float (&acc)[4], int (&iacc)[4]) {
      const uint8_t* xp = xrow + (size_t) ((ks - kb) ^ kflip) * xstep;
      if constexpr (KIND == kMF16) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xp + xoff[0]);
        mma_f16(acc, lo.x, hi.x, lo.y, hi.y, xv.x, xv.y);
        mma_f16(acc, lo.z, hi.z, lo.w, hi.w, xv.z, xv.w);
      } else if constexpr (KIND == kMA8W8) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xp + xoff[0]);
        mma_s8(iacc, lo.x, hi.x, lo.y, hi.y, xv.x, xv.y);
        mma_s8(iacc, lo.z, hi.z, lo.w, hi.w, xv.z, xv.w);
      } else if constexpr (KIND == kMW8) {
        const uint4 xa = *reinterpret_cast<const uint4*>(xp + xoff[0]), xb = *reinterpret_cast<const uint4*>(xp + xoff[1]);
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
          const uint4 xv = *reinterpret_cast<const uint4*>(xp + xoff[j]);
          mma_f16(acc, h2u(l[0]), h2u(h[0]), h2u(l[1]), h2u(h[1]), xv.x, xv.y);
          mma_f16(acc, h2u(l[2]), h2u(h[2]), h2u(l[3]), h2u(h[3]), xv.z, xv.w);
        }
      }
    };

    for (int c = 0; c < cpt; ++c) {
      issue();
      cp_async_wait<R - 1>();                                // this lane's copies of the oldest chunk have landed
      // plain loads (ordered behind the wait by its memory clobber, free to move above the MMAs of the previous k-step)
      const uint8_t* src = myring_p + (size_t) cslot * kMmaChunk;
      uint4 lo[kMmaC], hi[kMmaC];
#pragma unroll
      for (int u = 0; u < kMmaC; ++u) {
        lo[u] = *reinterpret_cast<const uint4*>(src + (2 * u) * 512);
        hi[u] = *reinterpret_cast<const uint4*>(src + (2 * u + 1) * 512);
      }
#pragma unroll
      for (int u = 0; u < kMmaC; ++u) {
        const int kk = c * kMmaC + u;
        if (kk < nks) consume(lo[u], hi[u], ks0 + kk, acc2[u & 1], iacc2[u & 1]);
      }
      cslot = cslot + 1 == R ? 0 : cslot + 1;
    }

    // ---- reduce: warps (shared memory, warp order) then cluster ranks (DSMEM, rank order) -----------------------------
    // fragment: c0,c1 = (row g, tokens 2t, 2t+1), c2,c3 = (row g + 8, tokens 2t, 2t+1).  wpart / cpart are double-buffered
    // by tile parity: a buffer is rewritten two tiles later, behind the next tile's barrier, so one barrier per tile does.
    float* wpar = wpart + parity * (kMmaWarps * 128);
    {
      float* wp = wpar + warp * 128;
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = INT ? __int_as_float(iacc2[0][i] + iacc2[1][i]) : acc2[0][i] + acc2[1][i];
      wp[g * 8 + 2 * t] = v[0];
      wp[g * 8 + 2 * t + 1] = v[1];
      wp[(g + 8) * 8 + 2 * t] = v[2];
      wp[(g + 8) * 8 + 2 * t + 1] = v[3];
    }
    __syncthreads();
    float tot = 0.f;
    int itot = 0;
    if (fin_warp) {
#pragma unroll
      for (int w = 0; w < kMmaWarps; ++w) {
        if constexpr (INT) itot += __float_as_int(wpar[w * 128 + etid]); else tot += wpar[w * 128 + etid];
      }
    }
    bool finisher = true;                                  // this CTA applies the epilogue for the tile
    if (p.S > 1) {
      float* cp = cpart + parity * (4 * 128);
      if (fin_warp) st_cluster_f32(cp + crank * 128 + etid, 0, INT ? __int_as_float(itot) : tot);
      cluster_sync_all();
      finisher = crank == 0;
      if (finisher && fin_warp) {
        tot = 0.f;
        itot = 0;
        for (int c = 0; c < p.S; ++c) {
          if constexpr (INT) itot += __float_as_int(cp[c * 128 + etid]); else tot += cp[c * 128 + etid];
        }
      }
    }

    // ---- epilogue: thread = (fragment row er, token em) ------------------------------------------------------------------
    if (finisher && fin_warp) {
      float v = INT ? (float) itot : tot;
      if (row_ok) {
        if constexpr (KIND == kMW8 || KIND == kMW4) v *= e_wscale;
        // reference grouping: accum * (scale_col * scale_row)  (epilogue_per_row_per_col_scale.h:325,341)
        if constexpr (INT) v = v * (e_sc * srow[em]);
      }
      if constexpr (SWIGLU) {
        // rows 0-7 hold gate, rows 8-15 the matching up projection: exchange through shared memory
        ex[etid] = v;
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (er < 8 && row_ok && em < M) {
          const float gte = __half2float(__float2half_rn(ex[er * 8 + em])), up = __half2float(__float2half_rn(ex[(er + 8) * 8 + em]));
          const float o = __half2float(__float2half_rn(mma_silu(gte))) * up;
          const size_t oi = (size_t) em * p.n_out + tile * 8 + er;
          if (p.y_f32) p.y_f32[oi] = o; else p.y[oi] = __float2half_rn(o);
        }
      } else if (row_ok && em < M) {
        const size_t oi = (size_t) em * p.n_out + wrow;
        if (p.y_f32) {
          p.y_f32[oi] = v;
        } else {
          __half oh = __float2half_rn(v);
          if (p.residual) oh = __float2half_rn(__half2float(oh) + e_res);
          p.y[oi] = oh;
        }
      }
    }
  }
  cp_async_wait<0>();
  {
    const unsigned total_pf = p.pf_lines[0] + p.pf_lines[1];
    const unsigned gw = blockIdx.x * kMmaWarps + warp, tw = gridDim.x * kMmaWarps;
    for (unsigned l = gw * 32 + lane; l < total_pf; l += tw * 32) {
      const uint8_t* a = l < p.pf_lines[0] ? p.pf[0] + (size_t) l * 128 : p.pf[1] + (size_t) (l - p.pf_lines[0]) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
  }
}

constexpr size_t kMmaFixedSmem = (8 + 8 + kMmaWarps * 8 + 128 + 64 + 2 * kMmaWarps * 128 + 2 * 4 * 128) * sizeof(float);
constexpr size_t kMmaSmemMax = 227 * 1024;

template <int KIND, bool SWIGLU, int R>
static int launch_gemv_mma_r(GemvMmaParams p, size_t xs_bytes, cudaStream_t stream) {
  const size_t smem = kMmaFixedSmem + (size_t) kMmaWarps * R * kMmaChunk + xs_bytes;
  if (smem > kMmaSmemMax) return -2;
  auto kern = gemv_mma_kernel<KIND, SWIGLU, R>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMmaSmemMax);
    if (e != cudaSuccess) return (int) e;
    attr_done = true;
  }
  const int tiles = SWIGLU ? (p.n_out + 7) / 8 : (p.N + 15) / 16;
  const int S = p.S;
  cudaLaunchConfig_t cfg{};
  // persistent: one CTA per SM, whole clusters
  int clusters = tiles;
  const int max_clusters = kNumSMs / S;
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

// Cluster size S (K split across CTAs) and ring depth: S grows while there are too few tiles to give every SM one, and
// further until the staged activations of one CTA's K range fit beside a ring of at least 3 chunks per warp.
template <int KIND, bool SWIGLU>
static int launch_gemv_mma(GemvMmaParams p, cudaStream_t stream) {
  constexpr int XB = KIND == kMA8W8 ? 1 : 2;
  constexpr int XSTEP = MmaTraits<KIND>::kStepElems * XB;
  const int tiles = SWIGLU ? (p.n_out + 7) / 8 : (p.N + 15) / 16;
  const int ksteps = p.K / MmaTraits<KIND>::kStepElems;
  auto xs_of = [&](int S, int& xrs) -> size_t {
    const int per = S == 1 ? ksteps : (ksteps + S - 1) / S + 1;          // an uneven split gives some CTA one more k-step
    xrs = ((per * XSTEP + 127) / 128) * 128;
    return (size_t) p.M * xrs;
  };
  auto fits = [&](int S, int R) {
    int xrs;
    return kMmaFixedSmem + (size_t) kMmaWarps * R * kMmaChunk + xs_of(S, xrs) <= kMmaSmemMax;
  };
  int S = 1;
  while (S < 4 && tiles * S < kNumSMs && ksteps / (kMmaWarps * S * 2) >= 1) S *= 2;
  while (S < 4 && !fits(S, 3) && ksteps / (kMmaWarps * S * 2) >= 1) S *= 2;
  p.S = S;
  int xrs;
  const size_t xs_bytes = xs_of(S, xrs);
  p.xrs = xrs;
  if (fits(S, kMmaRing)) return launch_gemv_mma_r<KIND, SWIGLU, kMmaRing>(p, xs_bytes, stream);
  if (fits(S, 3)) return launch_gemv_mma_r<KIND, SWIGLU, 3>(p, xs_bytes, stream);
  return launch_gemv_mma_r<KIND, SWIGLU, kMmaRingSmall>(p, xs_bytes, stream);
}

// true when the tensor-core GEMV handles this problem (else gemv.cu's FMA kernel, M <= 4): K a whole number of k-steps, and
// the staged activations of the deepest K split (4 CTAs) fit beside the shallowest weight ring
bool gemv_mma_eligible(int kind, int M, int K) {
  if (M < 1 || M > 8) return false;
  const int step = kind == kMF16 ? 32 : (kind == kMW4 ? 128 : 64);
  if (K % step != 0 || K / step < kMmaMinSteps) return false;
  const int xstep = step * (kind == kMA8W8 ? 1 : 2);
  const int ksteps = K / step;
  const int s_max = ksteps / (kMmaWarps * 2) >= 2 ? 4 : (ksteps / (kMmaWarps * 2) >= 1 ? 2 : 1);   // launch_gemv_mma's cap on S
  const int per = s_max == 1 ? ksteps : (ksteps + s_max - 1) / s_max + 1;
  const size_t xs = (size_t) M * (((size_t) per * xstep + 127) / 128 * 128);
  return kMmaFixedSmem + (size_t) kMmaWarps * kMmaRingSmall * kMmaChunk + xs <= kMmaSmemMax;
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
