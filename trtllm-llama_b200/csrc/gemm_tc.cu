// tcgen05 / TMEM / TMA GEMM for every projection of the LLaMA decoder:  C[M,N] = epi( X[M,K] . W[N,K]^T )
//
// "Weights-as-A": the weight matrix W [N, K] (K contiguous) is the UMMA A operand (UMMA_M = 128
// output channels per tile), the activations X [M, K] are the B operand (UMMA_N = NT token rows,
// 16..256), so decode (M = 1..16) and prefill (M = 16384) run the same kernel with different NT and
// the weights are always streamed exactly once per m-tile column.  The accumulator D[n, m] lives in
// TMEM (double-buffered, 2*NT columns) and is read back with tcgen05.ld by four epilogue warps.
//
//   KIND 0  fp16 x fp16  -> kind::f16, fp32 accumulate        (GemmPlugin / lm_head)
//   KIND 3  int8 x int8  -> kind::i8,  int32 accumulate, epilogue float(acc) * (sc[n]*sr[m])
//           replaces CutlassInt8GemmRunner<T>::gemm, K/cutlass_kernels/int8_gemm/int8_gemm_template.h:56-172,
//           epilogue CE/epilogue/threadblock/epilogue_per_row_per_col_scale.h:279-349
//   KIND 1/2 fp16 x int8/int4 weight-only: raw weight bytes arrive by TMA, four converter warps expand
//           them exactly to fp16 into the 128B-swizzled UMMA tile, per-channel scale in the epilogue
//           replaces CutlassFpAIntBGemmRunner::gemm, K/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:49-175
//
// Warp roles (persistent CTA, static round-robin over work items = n_tile x m_tile x k_split):
//   warp 0   TMA producer (one elected lane)          warp 1   MMA issuer (one elected lane) + TMEM alloc
//   warps 2-5 epilogue (TMEM lane quarter = warp % 4)  warps 6-9 weight converters (KIND 2/3 only)
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
// Split-K (small M, to give all 148 SMs work): fp32/int32 partials to a workspace, the last-arriving
// CTA of a tile sums them in split order (deterministic) and applies the epilogue.
//
// Roofline: decode shapes are HBM-bound (algorithmic bytes = N*K*bytes_per_weight); prefill shapes
// are tensor-bound (2*M*N*K ops).
#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "tmap_host.h"

namespace tb {

enum { kGF16 = 0, kGW8 = 1, kGW4 = 2, kGI8 = 3 };   // same numbering as tb_gemv

struct GemmTcParams {
  void* c;                 // [M, N] fp16 / fp32 / int32
  int out_type;            // 0 fp16, 1 fp32, 2 int32
  const __half* residual;  // optional [M, N] (fp16 out only)
  const __half* w_scale;   // KIND 2/3: [N] fp16
  const float* sc;         // KIND 1: per-channel [N] or [1]
  const float* sr;         // KIND 1: per-token [M] or [1]
  int sc_per_channel, sr_per_token;
  int M, N, K;
  int n_tiles, m_tiles, splits, kb_total, kb_per_split;
  int band;                // m-tiles per L2 band (rasterisation, see item_coords)
  float* partial;          // [items][NT][128] fp32 (int32 bit patterns for KIND 1)
  int* counters;           // [n_tiles * m_tiles]
};

constexpr int kTileN = 128;      // output channels per tile (UMMA_M)
constexpr int kGemmMaxCounters = 4096;   // tiles that may be split (split-K only when tiles < #SMs)
constexpr int kStageKBytes = 128; // bytes of K per row per stage (one 128B swizzle atom)

// Work item -> (n-tile, m-tile, k-split).  Rasterised in bands of `band` m-tiles: consecutive CTAs share one weight
// tile and walk the band's token tiles, and a band finishes every n-tile before the next band starts, so the band's
// activations (band x NT rows x K) stay L2-resident while the weights stream (ncu before this: the 84 MB dense prefill
// GEMM read 493 MB from DRAM because every wave touched all of X).
__device__ __forceinline__ void item_coords(const GemmTcParams& p, int it, int& nt, int& mt, int& split) {
  split = it % p.splits;
  const int t = it / p.splits;
  const int per_band = p.band * p.n_tiles;
  const int b = t / per_band, r = t - b * per_band;
  const int bw = min(p.band, p.m_tiles - b * p.band);      // the last band may be narrower
  nt = r / bw;
  mt = b * p.band + (r - nt * bw);
}

template <int KIND, int NT>
struct GemmCfg {
  static constexpr bool kWO = (KIND == kGW8 || KIND == kGW4);
  static constexpr int kKElems = (KIND == kGI8) ? 128 : 64;         // K elements per stage
  static constexpr int kABytes = kTileN * kStageKBytes;              // 16 KB
  static constexpr int kBBytes = NT * kStageKBytes;
  static constexpr int kRawRowBytes = KIND == kGW8 ? 64 : (KIND == kGW4 ? 32 : 0);
  static constexpr int kRawBytes = kTileN * kRawRowBytes;
  static constexpr int kStageBytes = kABytes + kBBytes + kRawBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kThreads = kWO ? 320 : 192;
  static constexpr int kTmemCols = (2 * NT) < 32 ? 32 : 2 * NT;
  static constexpr size_t kSmemBytes =
      (size_t) kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 4 * NT * sizeof(float) /*sr stash*/;
};

template <int KIND, int NT>
__global__ void __launch_bounds__(GemmCfg<KIND, NT>::kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
               const GemmTcParams p) {
  using Cfg = GemmCfg<KIND, NT>;
  constexpr int ST = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t) ST * Cfg::kABytes;
  uint8_t* sRaw = sB + (size_t) ST * Cfg::kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRaw + (size_t) ST * Cfg::kRawBytes);
  uint64_t* full = bars;                 // [ST]  TMA (+converters) -> MMA
  uint64_t* empty = bars + ST;           // [ST]  MMA -> TMA
  uint64_t* rawfull = bars + 2 * ST;     // [ST]  TMA -> converters (KIND 2/3)
  uint64_t* tfull = bars + 3 * ST;       // [2]   MMA -> epilogue
  uint64_t* tempty = bars + 3 * ST + 2;  // [2]   epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * ST + 4);
  float* sr_stash = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [4 epilogue warps][NT]
  __shared__ int s_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.n_tiles * p.m_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full[s], Cfg::kWO ? 1 + 4 : 1);
      mbar_init(&empty[s], 1);
      mbar_init(&rawfull[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0, phase = 0;
      const uint64_t pol_w = p.m_tiles > 1 ? policy_evict_last() : policy_evict_first();
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        int nt, mt, split;
        item_coords(p, it, nt, mt, split);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if constexpr (Cfg::kWO) {
            mbar_expect_tx(&rawfull[stage], Cfg::kRawBytes);
            tma_load_2d_hint(sRaw + (size_t) stage * Cfg::kRawBytes, &tmap_w, &rawfull[stage],
                             kb * Cfg::kRawRowBytes, nt * kTileN, pol_w);
            mbar_expect_tx(&full[stage], Cfg::kBBytes);
          } else {
            mbar_expect_tx(&full[stage], Cfg::kABytes + Cfg::kBBytes);
            tma_load_2d_hint(sA + (size_t) stage * Cfg::kABytes, &tmap_w, &full[stage], kb * Cfg::kKElems,
                             nt * kTileN, pol_w);
          }
          tma_load_2d(sB + (size_t) stage * Cfg::kBBytes, &tmap_x, &full[stage], kb * Cfg::kKElems, mt * NT);
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    if (lane == 0) {
      constexpr uint32_t idesc = KIND == kGI8 ? kIdescI8(kTileN, NT) : kIdescF16(kTileN, NT);
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int split = it % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t) (acc * NT);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t ad = umma_desc_sw128(smem_u32(sA + (size_t) stage * Cfg::kABytes));
          const uint64_t bd = umma_desc_sw128(smem_u32(sB + (size_t) stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < kStageKBytes / 32; ++k) {
            // advance 32 bytes (UMMA_K = 16 fp16 / 32 int8) inside the swizzle atom: +2 in 16-byte units
            if constexpr (KIND == kGI8) umma_i8(d_addr, ad + 2 * k, bd + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_f16(d_addr, ad + 2 * k, bd + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 6) {
    // =========================== epilogue ===============================
    const int q = warp & 3;                       // TMEM lane quarter owned by this warp
    int acc = 0, acc_phase = 0;
    float* s_sr = sr_stash + (warp - 2) * NT;     // this warp's copy of the tile's per-token scales
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      int nt, mt, split;
      item_coords(p, it, nt, mt, split);
      const int n = nt * kTileN + q * 32 + lane;
      const int m0 = mt * NT;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (acc * NT);

      float chan = 1.f;
      const bool f16_scaled = KIND == kGF16 && p.w_scale != nullptr;   // weight-only weights dequantised to fp16 up front
      if (n < p.N) {
        if (Cfg::kWO || f16_scaled) chan = __half2float(p.w_scale[n]);
        if constexpr (KIND == kGI8) chan = p.sc[p.sc_per_channel ? n : 0];
      }
      auto finish = [&](float v, int m) {   // v: accumulated value as float, before scaling
        if (n >= p.N || m >= p.M) return;
        if constexpr (KIND == kGI8) v = v * (chan * p.sr[p.sr_per_token ? m : 0]);
        if (Cfg::kWO || f16_scaled) v = v * chan;
        const size_t oi = (size_t) m * p.N + n;
        if (p.out_type == 0) {
          __half h = __float2half_rn(v);
          if (p.residual) h = __float2half_rn(__half2float(h) + __half2float(p.residual[oi]));
          reinterpret_cast<__half*>(p.c)[oi] = h;
        } else if (p.out_type == 1) {
          reinterpret_cast<float*>(p.c)[oi] = v;
        } else {
          reinterpret_cast<int*>(p.c)[oi] = __float2int_rn(v);
        }
      };

      if (p.splits == 1) {
        // per-token scales of this tile into a warp-private stash: a dependent global load per token row inside the
        // drain loop serialised the epilogue (ncu: 170 cycles/row, the K=4096 int8 GEMMs were epilogue-bound)
        if constexpr (KIND == kGI8) {
          __syncwarp();
          for (int i = lane; i < NT; i += 32)
            s_sr[i] = p.sr_per_token ? (m0 + i < p.M ? p.sr[m0 + i] : 0.f) : p.sr[0];
          __syncwarp();
        }
#pragma unroll 1
        for (int c = 0; c < NT / 16; ++c) {
          uint32_t v[16];
          tmem_ld16(t_addr + c * 16, v);
          tmem_ld_wait();
          const int mc = m0 + c * 16;
          if (n < p.N && mc + 16 <= p.M && p.out_type == 0) {
            // fast path: a full chunk of fp16 outputs, no per-element branches
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if constexpr (KIND == kGI8) f[j] = (float) (int) v[j] * (chan * s_sr[c * 16 + j]);
              else if constexpr (Cfg::kWO) f[j] = __uint_as_float(v[j]) * chan;
              else f[j] = f16_scaled ? __uint_as_float(v[j]) * chan : __uint_as_float(v[j]);
            }
            __half* cp = reinterpret_cast<__half*>(p.c) + (size_t) mc * p.N + n;
            if (p.residual) {
              const __half* rp = p.residual + (size_t) mc * p.N + n;
              __half r[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) r[j] = rp[(size_t) j * p.N];
#pragma unroll
              for (int j = 0; j < 16; ++j)
                cp[(size_t) j * p.N] = __float2half_rn(__half2float(__float2half_rn(f[j])) + __half2float(r[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) cp[(size_t) j * p.N] = __float2half_rn(f[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float f = KIND == kGI8 ? (float) (int) v[j] : __uint_as_float(v[j]);
              finish(f, mc + j);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      } else {
        // split-K: store the raw partial [m][n-in-tile], then the last CTA of the tile reduces
        float* part = p.partial + ((size_t) it) * NT * kTileN;
#pragma unroll 1
        for (int c = 0; c < NT / 16; ++c) {
          uint32_t v[16];
          tmem_ld16(t_addr + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            reinterpret_cast<uint32_t*>(part)[(size_t) (c * 16 + j) * kTileN + q * 32 + lane] = v[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        __threadfence();
        // the four epilogue warps synchronise on named barrier 1 (128 threads)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) {
          const int prev = atomicAdd(&p.counters[nt * p.m_tiles + mt], 1);
          s_last = (prev == p.splits - 1);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (s_last) {
          __threadfence();
          const size_t tile_first = (size_t) (it - split);   // the splits of one tile are consecutive items
          for (int mm = 0; mm < NT; ++mm) {
            if (m0 + mm >= p.M) break;
            float fsum = 0.f;
            int isum = 0;
            for (int s = 0; s < p.splits; ++s) {
              const uint32_t u = __ldcg(reinterpret_cast<const uint32_t*>(p.partial) +
                                        ((tile_first + s) * NT + mm) * kTileN + q * 32 + lane);
              if constexpr (KIND == kGI8) isum += (int) u; else fsum += __uint_as_float(u);
            }
            finish(KIND == kGI8 ? (float) isum : fsum, m0 + mm);
          }
          if (warp == 2 && lane == 0) p.counters[nt * p.m_tiles + mt] = 0;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // s_last is reused by the next item
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // =========================== weight converters (KIND 2/3) ============
    if constexpr (Cfg::kWO) {
      const int ct = threadIdx.x - 6 * 32;          // 0..127
      int stage = 0, phase = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int split = it % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&rawfull[stage], phase);
          const uint8_t* raw = sRaw + (size_t) stage * Cfg::kRawBytes;
          uint8_t* dst = sA + (size_t) stage * Cfg::kABytes;
          if constexpr (KIND == kGW8) {
            // 64 raw bytes per row: thread -> (row = ct/4 + 32*i, 16-byte chunk = ct%4)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = (ct >> 2) + 32 * i, ch = ct & 3;
              const uint4 w = *reinterpret_cast<const uint4*>(raw + r * 64 + ch * 16);
              __half2 h[8];
              i8x4_to_h2x2(w.x, h[0], h[1]);
              i8x4_to_h2x2(w.y, h[2], h[3]);
              i8x4_to_h2x2(w.z, h[4], h[5]);
              i8x4_to_h2x2(w.w, h[6], h[7]);
              // 16 fp16 = output chunks 2ch, 2ch+1 of the 128-byte row, XOR-swizzled by (row % 8)
              uint8_t* rowp = dst + r * 128;
              *reinterpret_cast<uint4*>(rowp + (((2 * ch) ^ (r & 7)) << 4)) = *reinterpret_cast<uint4*>(&h[0]);
              *reinterpret_cast<uint4*>(rowp + (((2 * ch + 1) ^ (r & 7)) << 4)) = *reinterpret_cast<uint4*>(&h[4]);
            }
          } else {
            // 32 raw bytes per row (64 int4): thread -> (row = ct/2 + 64*i, 16-byte chunk = ct%2)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int r = (ct >> 1) + 64 * i, ch = ct & 1;
              const uint4 w = *reinterpret_cast<const uint4*>(raw + r * 32 + ch * 16);
              __half2 h[16];
              i4x8_to_h2x4(w.x, h + 0);
              i4x8_to_h2x4(w.y, h + 4);
              i4x8_to_h2x4(w.z, h + 8);
              i4x8_to_h2x4(w.w, h + 12);
              uint8_t* rowp = dst + r * 128;
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4)
                *reinterpret_cast<uint4*>(rowp + (((4 * ch + c4) ^ (r & 7)) << 4)) = *reinterpret_cast<uint4*>(&h[4 * c4]);
            }
          }
          fence_proxy_async();      // make the generic-proxy smem writes visible to the tensor core
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[stage]);
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int KIND, int NT>
static int launch_gemm_tc(GemmTcParams p, const void* x, const void* w, void* workspace, size_t workspace_bytes,
                          int* counters, int force_splits, cudaStream_t stream) {
  using Cfg = GemmCfg<KIND, NT>;
  CUtensorMap tw, tx;
  int rc;
  if constexpr (KIND == kGF16) {
    rc = make_tmap(&tw, w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p.N, p.K, kTileN, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap(&tx, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p.M, p.K, NT, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  } else if constexpr (KIND == kGI8) {
    rc = make_tmap(&tw, w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, p.N, p.K, kTileN, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap(&tx, x, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, p.M, p.K, NT, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  } else if constexpr (KIND == kGW8) {
    rc = make_tmap(&tw, w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, p.N, p.K, kTileN, 64, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = make_tmap(&tx, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p.M, p.K, NT, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    rc = make_tmap(&tw, w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, p.N, p.K / 2, kTileN, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = make_tmap(&tx, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p.M, p.K, NT, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (rc) return rc;

  p.n_tiles = (p.N + kTileN - 1) / kTileN;
  p.m_tiles = (p.M + NT - 1) / NT;
  p.band = p.m_tiles < 16 ? p.m_tiles : 16;
  p.kb_total = (p.K + Cfg::kKElems - 1) / Cfg::kKElems;
  // split K until there are >= 2 work items per SM (HBM-bound small-M shapes), >= 4 k-blocks per split
  int splits = 1;
  const int base_items = p.n_tiles * p.m_tiles;
  if (force_splits > 0) {
    splits = force_splits;
  } else if (base_items < kNumSMs && NT <= 16) {
    // split-K only for decode-like shapes: the in-order partial-sum reduction is serial in the token dimension
    splits = (2 * kNumSMs + base_items - 1) / base_items;
    const int max_by_k = p.kb_total / 4 > 0 ? p.kb_total / 4 : 1;
    if (splits > max_by_k) splits = max_by_k;
  }
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;   // no empty splits
  p.splits = splits;
  if (splits > 1) {
    const size_t need = (size_t) base_items * splits * NT * kTileN * sizeof(float);
    if (!workspace || workspace_bytes < need || !counters || base_items > kGemmMaxCounters) return -12;
    p.counters = counters;
    p.partial = reinterpret_cast<float*>(workspace);
  }
  const int items = base_items * splits;
  int grid = items < kNumSMs ? items : kNumSMs;
  auto kern = gemm_tc_kernel<KIND, NT>;
  TB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::kSmemBytes));
  kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(tw, tx, p);
  return (int) cudaGetLastError();
}

// Weight-only weights -> plain fp16 [N, K] (exact: |w| <= 127 / 7), unscaled; the per-channel scale stays in the GEMM
// epilogue.  One thread per 16 raw bytes, same element order as the converter warps of the fused kernel.
template <int KIND>
__global__ void dequant_weights_kernel(__half* __restrict__ out, const uint8_t* __restrict__ w, int64_t chunks) {
  for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < chunks; i += (int64_t) gridDim.x * blockDim.x) {
    const uint4 raw = ldg_nc_v4(w + i * 16);
    if constexpr (KIND == kGW8) {
      __half2 h[8];
      i8x4_to_h2x2(raw.x, h[0], h[1]);
      i8x4_to_h2x2(raw.y, h[2], h[3]);
      i8x4_to_h2x2(raw.z, h[4], h[5]);
      i8x4_to_h2x2(raw.w, h[6], h[7]);
      uint4* o = reinterpret_cast<uint4*>(out + i * 16);
      o[0] = *reinterpret_cast<uint4*>(&h[0]);
      o[1] = *reinterpret_cast<uint4*>(&h[4]);
    } else {
      __half2 h[16];
      i4x8_to_h2x4(raw.x, h + 0);
      i4x8_to_h2x4(raw.y, h + 4);
      i4x8_to_h2x4(raw.z, h + 8);
      i4x8_to_h2x4(raw.w, h + 12);
      uint4* o = reinterpret_cast<uint4*>(out + i * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) o[c] = *reinterpret_cast<uint4*>(&h[4 * c]);
    }
  }
}

// Prefill-size weight-only problems: the fused kernel's converter stage costs 20-30 % against the plain fp16 kernel
// (1.0-1.2 vs 1.3-1.4 PFLOP/s at M = 15360: one pipeline stage less and a longer stage latency), while dequantising the
// whole matrix once is a 25-60 us HBM-bound pass.  Same MMA inputs in the same order -> bit-identical results.
constexpr int kDequantMinM = 2048;
static bool wo_dequant_enabled() {
  static const bool on = [] { const char* e = getenv("TB_GEMM_WO_DEQUANT"); return !(e && e[0] == '0'); }();
  return on;
}

// token-tile width: the smallest of 16..256 covering M; prefill-like shapes shrink it until the grid covers the SMs
static int select_nt(int M, int N) {
  int nt = M <= 16 ? 16 : (M <= 32 ? 32 : (M <= 64 ? 64 : (M <= 128 ? 128 : 256)));
  const int n_tiles = (N + kTileN - 1) / kTileN;
  while (nt > 32 && n_tiles * ((M + nt - 1) / nt) < kNumSMs) nt >>= 1;
  return nt;
}

template <int KIND>
static int dispatch_nt(const GemmTcParams& p, const void* x, const void* w, void* ws, size_t ws_bytes, int* counters,
                       int force_splits, int force_nt, cudaStream_t stream) {
  int nt = force_nt;
  if (nt <= 0) nt = select_nt(p.M, p.N);
  switch (nt) {
    case 16:  return launch_gemm_tc<KIND, 16>(p, x, w, ws, ws_bytes, counters, force_splits, stream);
    case 32:  return launch_gemm_tc<KIND, 32>(p, x, w, ws, ws_bytes, counters, force_splits, stream);
    case 64:  return launch_gemm_tc<KIND, 64>(p, x, w, ws, ws_bytes, counters, force_splits, stream);
    case 128: return launch_gemm_tc<KIND, 128>(p, x, w, ws, ws_bytes, counters, force_splits, stream);
    case 256: return launch_gemm_tc<KIND, 256>(p, x, w, ws, ws_bytes, counters, force_splits, stream);
  }
  return -1;
}

// CTA-pair kernel (gemm_tc2.cu)
int gemm_tc_pair(int kind, void* c, int out_type, const void* x, const void* w, const float* sc, const float* sr,
                 int sc_per_channel, int sr_per_token, const void* residual, int M, int N, int K, cudaStream_t stream);

int gemm_tc_pair_swiglu(int kind, void* c, const void* x, const void* w, const float* sc, const float* sr, int sc_per_channel,
                        int sr_per_token, int M, int N, int K, cudaStream_t stream);

// TB_GEMM_TC_PAIR=0 keeps every shape on the one-CTA kernel (A/B measurements)
static bool pair_enabled() {
  static const bool on = [] { const char* e = getenv("TB_GEMM_TC_PAIR"); return !(e && e[0] == '0'); }();
  return on;
}

}  // namespace tb

using namespace tb;

extern "C" {

size_t tb_gemm_tc_workspace_bytes(int M, int N, int K) {
  // upper bound of what launch_gemm_tc uses: same tile selection, split count before the k-block caps
  (void) K;
  const int n_tiles = (N + kTileN - 1) / kTileN;
  const int nt = select_nt(M, N);
  const int m_tiles = (M + nt - 1) / nt;
  const size_t base = (size_t) n_tiles * m_tiles;
  size_t splits = (base < (size_t) kNumSMs && nt <= 16) ? (2 * kNumSMs + base - 1) / base : 1;
  size_t bytes = base * splits * nt * kTileN * sizeof(float) + 256;
  // prefill sizes: room for a weight-only matrix dequantised to fp16 (the function does not know the weight kind)
  if (M >= kDequantMinM) bytes = std::max(bytes, (size_t) N * K * sizeof(__half) + 256);
  return bytes;
}
size_t tb_gemm_tc_counter_bytes(void) { return (size_t) kGemmMaxCounters * sizeof(int); }

int tb_gemm_tc(int kind, void* c, int out_type, const void* x, const void* w, const void* w_scale, const float* sc,
               const float* sr, int sc_per_channel, int sr_per_token, const void* residual, int M, int N, int K,
               void* workspace, size_t workspace_bytes, int* counters, int force_splits, int force_nt,
               cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return -1;
  if (kind == kGF16 || kind == kGW8) { if (K % 8) return -1; }   // TMA: 16-byte aligned row pitch
  if (kind == kGI8 && K % 16) return -1;
  if (kind == kGW4 && K % 32) return -1;
  if ((kind == kGW8 || kind == kGW4) && !w_scale) return -1;
  if (kind == kGI8 && (!sc || !sr)) return -1;
  if (residual && out_type != 0) return -1;
  if ((kind == kGW8 || kind == kGW4) && M >= kDequantMinM && force_splits <= 0 && force_nt <= 0 && wo_dequant_enabled() &&
      workspace && workspace_bytes >= (size_t) N * K * sizeof(__half)) {
    __half* w16 = static_cast<__half*>(workspace);
    const int64_t chunks = (int64_t) N * K / (kind == kGW8 ? 16 : 32);
    const int blocks = (int) std::min<int64_t>((chunks + 255) / 256, (int64_t) kNumSMs * 16);
    if (kind == kGW8) dequant_weights_kernel<kGW8><<<blocks, 256, 0, stream>>>(w16, static_cast<const uint8_t*>(w), chunks);
    else dequant_weights_kernel<kGW4><<<blocks, 256, 0, stream>>>(w16, static_cast<const uint8_t*>(w), chunks);
    if (cudaGetLastError() != cudaSuccess) return -13;
    GemmTcParams pd{};
    pd.c = c; pd.out_type = out_type; pd.residual = (const __half*) residual; pd.w_scale = (const __half*) w_scale;
    pd.M = M; pd.N = N; pd.K = K;
    return dispatch_nt<kGF16>(pd, x, w16, nullptr, 0, counters, 0, 0, stream);
  }
  // prefill-size fp16 / int8 problems (>= 2 full-size tiles per SM): CTA-pair kernel, a third less L2 traffic per MAC
  if ((kind == kGF16 || kind == kGI8) && force_splits <= 0) {
    const long tiles256 = (long) ((N + kTileN - 1) / kTileN) * ((M + 255) / 256);
    // (measured: int8 +7..16 % over the one-CTA kernel at M = 16384; fp16 is no faster, so it stays opt-in)
    const bool auto_pair =
        kind == kGI8 && force_nt <= 0 && pair_enabled() && select_nt(M, N) == 256 && tiles256 >= 2 * kNumSMs;
    if (force_nt == 512 || auto_pair)
      return gemm_tc_pair(kind, c, out_type, x, w, sc, sr, sc_per_channel, sr_per_token, residual, M, N, K, stream);
  }
  GemmTcParams p{};
  p.c = c; p.out_type = out_type; p.residual = (const __half*) residual; p.w_scale = (const __half*) w_scale;
  p.sc = sc; p.sr = sr; p.sc_per_channel = sc_per_channel; p.sr_per_token = sr_per_token;
  p.M = M; p.N = N; p.K = K;
  switch (kind) {
    case kGF16: return dispatch_nt<kGF16>(p, x, w, workspace, workspace_bytes, counters, force_splits, force_nt, stream);
    case kGI8:  return dispatch_nt<kGI8>(p, x, w, workspace, workspace_bytes, counters, force_splits, force_nt, stream);
    case kGW8:  return dispatch_nt<kGW8>(p, x, w, workspace, workspace_bytes, counters, force_splits, force_nt, stream);
    case kGW4:  return dispatch_nt<kGW4>(p, x, w, workspace, workspace_bytes, counters, force_splits, force_nt, stream);
  }
  return -1;
}
}

// Gate / up projection with SwiGLU in the tcgen05 epilogue (prefill shapes): w holds [N = 2 * inter, K] (gate rows, then up
// rows), c is fp16 [M, inter] = silu(x . gate^T) * (x . up^T).  kind 0 (fp16) or 3 (SmoothQuant int8 with per-token /
// per-channel scales).  Bit-identical to tb_gemm_tc followed by tb_swiglu.
extern "C" int tb_gemm_tc_swiglu(int kind, void* c, const void* x, const void* w, const float* sc, const float* sr,
                                 int sc_per_channel, int sr_per_token, int M, int N, int K, cudaStream_t stream) {
  if (M < 1 || N < 2 || (N & 1) || K < 1) return -1;
  if (kind == kGI8 && (!sc || !sr || K % 16)) return -1;
  if (kind == kGF16 && K % 8) return -1;
  const int rc = gemm_tc_pair_swiglu(kind, c, x, w, sc, sr, sc_per_channel, sr_per_token, M, N, K, stream);
  return rc == -100 ? -1 : rc;
}
