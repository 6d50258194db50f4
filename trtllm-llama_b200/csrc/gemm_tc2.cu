// CTA-pair (tcgen05 cta_group::2) GEMM for prefill-size projections:  C[M,N] = epi( X[M,K] . W[N,K]^T )
//
// Why a second kernel: the one-CTA kernel (gemm_tc.cu) at its largest tile (128 channels x 256 tokens) pulls
// 48 KB from L2 per 512 tensor-core cycles = 96 B/clk/SM, and the measured chip-wide L2 -> SM throughput is
// ~6000 B/clk (~41 B/clk/SM): ncu showed 8.9-11 TB/s of TMA traffic and a tensor pipe stuck at 37-42 %.  Here two
// CTAs of one TPC compute one 256-channel x 256-token tile together: each CTA loads ITS 128 weight rows and ITS half
// (128 rows) of the token tile, the tensor cores read the peer's half of B through the pair's shared-memory path, so
// the L2 traffic per MAC drops by a third (32 KB per CTA per 512 cycles = 64 B/clk/SM).
//
//   KIND 0  fp16 x fp16 (kind::f16, fp32 accumulate)     KIND 3  int8 x int8 (kind::i8, int32 accumulate, SmoothQuant epilogue)
//
// Pair protocol (leader = cluster rank 0):
//   full[s]   (leader's smem, count 2)   each CTA's producer arrives with expect_tx for its own 32 KB; both CTAs' TMA
//                                        loads complete_tx on the LEADER's barrier (cp.async.bulk.tensor .cta_group::2)
//   empty[s]  (both CTAs, count 1)       the leader's tcgen05.commit multicasts the arrive to both CTAs
//   tfull[a]  (both CTAs, count 1)       leader's commit, multicast: each CTA's epilogue drains its own 128 TMEM lanes
//   tempty[a] (leader's smem, count 8)   four epilogue warps of each CTA arrive (the peer's remotely)
// Same replacement targets as gemm_tc.cu (CutlassInt8GemmRunner / GemmPlugin).
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "tmap_host.h"

namespace tb {

enum { k2F16 = 0, k2I8 = 3 };   // same numbering as tb_gemv / gemm_tc.cu

struct GemmTc2Params {
  void* c;
  int out_type;            // 0 fp16, 1 fp32, 2 int32
  const __half* residual;
  const float* sc;
  const float* sr;
  int sc_per_channel, sr_per_token;
  int M, N, K;
  int n_pairs, m_tiles, kb_total, band;
  int n_out;               // SWIGLU: W holds [2 * n_out, K] (gate rows, then up rows), C is [M, n_out] = silu(gate) * up
};

constexpr int k2TileN = 128;     // output channels per CTA (UMMA_M = 256 over the pair)
constexpr int k2NT = 256;        // tokens per tile (UMMA_N), 128 of them loaded by each CTA
constexpr int k2Stages = 6;
constexpr int k2ABytes = k2TileN * 128;        // 16 KB
constexpr int k2BBytes = (k2NT / 2) * 128;     // 16 KB
constexpr int k2Threads = 192;
constexpr size_t k2SmemBytes = (size_t) k2Stages * (k2ABytes + k2BBytes) + 1024 + 256 + 4 * k2NT * sizeof(float);
// SWIGLU variant (the gate / up projection with silu(gate) * up computed in the TMEM drain): a stage holds the pair's gate
// rows AND the matching up rows (2 x 16 KB per CTA) and a 128-token tile (8 KB per CTA); two MMAs per k-step accumulate
// gate and up side by side in TMEM (2 x 128 columns per buffer, still double-buffered), so the thread that owns channel n
// holds both values of every token.  40 KB per stage per CTA for the same MMA work as the plain kernel's 32 KB.
template <bool SWIGLU> struct Tc2Cfg {
  static constexpr int NT = SWIGLU ? 128 : k2NT;
  static constexpr int ABytes = SWIGLU ? 2 * k2ABytes : k2ABytes;
  static constexpr int BBytes = (NT / 2) * 128;
  static constexpr int Stages = SWIGLU ? 5 : k2Stages;
  static constexpr size_t Smem = (size_t) Stages * (ABytes + BBytes) + 1024 + 256 + 4 * NT * sizeof(float);
};

__device__ __forceinline__ uint32_t cta_rank_in_cluster() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(const void* local_smem_ptr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
  return remote;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;"
               ::"r"(bar_cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA tile load into THIS CTA's shared memory whose completion is signalled on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0,
                                                 int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_pair_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_pair_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive (once every MMA issued so far has completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t) 3)
      : "memory");
}

// pair-item -> (channel pair, token tile); same band rasterisation as gemm_tc.cu item_coords
__device__ __forceinline__ void pair_coords(const GemmTc2Params& p, int it, int& np, int& mt) {
  const int per_band = p.band * p.n_pairs;
  const int b = it / per_band, r = it - b * per_band;
  const int bw = min(p.band, p.m_tiles - b * p.band);
  np = r / bw;
  mt = b * p.band + (r - np * bw);
}

__device__ __forceinline__ float tc2_silu(float v) { return silu_fast(v); }   // same definition as tb_swiglu: bit-identical

template <int KIND, bool SWIGLU = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                const GemmTc2Params p) {
  using CF = Tc2Cfg<SWIGLU>;
  constexpr int ST = CF::Stages;
  constexpr int k2NT = CF::NT, k2ABytes = CF::ABytes, k2BBytes = CF::BBytes;   // shadow the plain kernel's constants
  constexpr int kKElems = KIND == k2I8 ? 128 : 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t) ST * k2ABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t) ST * k2BBytes);
  uint64_t* full = bars;                 // [ST]  used in the leader only
  uint64_t* empty = bars + ST;           // [ST]
  uint64_t* tfull = bars + 2 * ST;       // [2]
  uint64_t* tempty = bars + 2 * ST + 2;  // [2]   used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);
  float* sr_stash = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [4 epilogue warps][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cta_rank_in_cluster();
  const int items = p.n_pairs * p.m_tiles;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full[s], 2);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // the peer's barriers are initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    if (lane == 0) {
      int stage = 0, phase = 0;
      const uint64_t pol_w = policy_evict_last(), pol_x = policy_evict_last();
      for (int it = cluster_id; it < items; it += n_clusters) {
        int np, mt;
        pair_coords(p, it, np, mt);
        const int nt = 2 * np + (int) rank;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t fb = map_to_cta(&full[stage], 0);
          mbar_expect_tx_cluster(fb, k2ABytes + k2BBytes);
          tma_load_2d_pair(sA + (size_t) stage * k2ABytes, &tmap_w, fb, kb * kKElems, nt * k2TileN, pol_w);
          if constexpr (SWIGLU)
            tma_load_2d_pair(sA + (size_t) stage * k2ABytes + k2TileN * 128, &tmap_w, fb, kb * kKElems,
                             p.n_out + nt * k2TileN, pol_w);
          tma_load_2d_pair(sB + (size_t) stage * k2BBytes, &tmap_x, fb, kb * kKElems,
                           mt * k2NT + (int) rank * (k2NT / 2), pol_x);
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA) =============================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = KIND == k2I8 ? kIdescI8(2 * k2TileN, k2NT) : kIdescF16(2 * k2TileN, k2NT);
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int it = cluster_id; it < items; it += n_clusters) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t) (acc * 256);
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t ad = umma_desc_sw128(smem_u32(sA + (size_t) stage * k2ABytes));
          const uint64_t bd = umma_desc_sw128(smem_u32(sB + (size_t) stage * k2BBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if constexpr (KIND == k2I8) umma_pair_i8(d_addr, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_pair_f16(d_addr, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if constexpr (SWIGLU) {   // the up rows of the same channels against the same token tile, next to gate in TMEM
            const uint64_t au = umma_desc_sw128(smem_u32(sA + (size_t) stage * k2ABytes + k2TileN * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if constexpr (KIND == k2I8) umma_pair_i8(d_addr + k2NT, au + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              else umma_pair_f16(d_addr + k2NT, au + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit_pair(&empty[stage]);
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // =========================== epilogue (both CTAs, own 128 TMEM lanes) ============
    const int q = warp & 3;
    int acc = 0, acc_phase = 0;
    const uint32_t tempty_leader[2] = {map_to_cta(&tempty[0], 0), map_to_cta(&tempty[1], 0)};
    float* s_sr = sr_stash + (warp - 2) * k2NT;
    for (int it = cluster_id; it < items; it += n_clusters) {
      int np, mt;
      pair_coords(p, it, np, mt);
      const int n = (2 * np + (int) rank) * k2TileN + q * 32 + lane;
      const int m0 = mt * k2NT;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (acc * 256);
      float chan = 1.f, chan_up = 1.f;
      if constexpr (KIND == k2I8) {
        if (n < p.N) chan = p.sc[p.sc_per_channel ? n : 0];
        if (SWIGLU && n < p.n_out) chan_up = p.sc[p.sc_per_channel ? p.n_out + n : 0];
        // the tile's per-token scales into a warp-private stash (no dependent global load inside the drain loop)
        __syncwarp();
        for (int i = lane; i < k2NT; i += 32)
          s_sr[i] = p.sr_per_token ? (m0 + i < p.M ? p.sr[m0 + i] : 0.f) : p.sr[0];
        __syncwarp();
      }
      if constexpr (SWIGLU) {
        // C[m, n] = fp16( fp16(silu(g)) * u ) with g, u the fp16-rounded gate / up projections: the arithmetic of the plain
        // kernel followed by tb_swiglu, bit for bit (mlp.py:68-73)
#pragma unroll 1
        for (int c = 0; c < k2NT / 16; ++c) {
          uint32_t vg[16], vu[16];
          tmem_ld16(t_addr + c * 16, vg);
          tmem_ld16(t_addr + k2NT + c * 16, vu);
          tmem_ld_wait();
          const int mc = m0 + c * 16;
          if (n < p.n_out) {
            __half* cp = reinterpret_cast<__half*>(p.c) + (size_t) mc * p.n_out + n;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float g, u;
              if constexpr (KIND == k2I8) {
                g = (float) (int) vg[j] * (chan * s_sr[c * 16 + j]);
                u = (float) (int) vu[j] * (chan_up * s_sr[c * 16 + j]);
              } else {
                g = __uint_as_float(vg[j]);
                u = __uint_as_float(vu[j]);
              }
              g = __half2float(__float2half_rn(g));
              u = __half2float(__float2half_rn(u));
              const float o = __half2float(__float2half_rn(tc2_silu(g))) * u;
              if (mc + j < p.M) cp[(size_t) j * p.n_out] = __float2half_rn(o);
            }
          }
          __syncwarp();
        }
      } else
#pragma unroll 1
      for (int c = 0; c < k2NT / 16; ++c) {
        uint32_t v[16];
        tmem_ld16(t_addr + c * 16, v);
        tmem_ld_wait();
        const int mc = m0 + c * 16;
        if (n < p.N) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if constexpr (KIND == k2I8) f[j] = (float) (int) v[j] * (chan * s_sr[c * 16 + j]);
            else f[j] = __uint_as_float(v[j]);
          }
          if (mc + 16 <= p.M && p.out_type == 0) {
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
              const int m = mc + j;
              if (m >= p.M) break;
              const size_t oi = (size_t) m * p.N + n;
              if (p.out_type == 0) {
                __half h = __float2half_rn(f[j]);
                if (p.residual) h = __float2half_rn(__half2float(h) + __half2float(p.residual[oi]));
                reinterpret_cast<__half*>(p.c)[oi] = h;
              } else if (p.out_type == 1) {
                reinterpret_cast<float*>(p.c)[oi] = f[j];
              } else {
                reinterpret_cast<int*>(p.c)[oi] = __float2int_rn(f[j]);
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // neither CTA exits (or frees TMEM) while the peer can still signal or read it
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

template <int KIND, bool SWIGLU = false>
static int launch_gemm_tc2(GemmTc2Params p, const void* x, const void* w, cudaStream_t stream) {
  using CF = Tc2Cfg<SWIGLU>;
  constexpr int k2NT = CF::NT;
  constexpr size_t k2SmemBytes = CF::Smem;
  CUtensorMap tw, tx;
  int rc;
  if constexpr (KIND == k2F16) {
    rc = make_tmap(&tw, w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p.N, p.K, k2TileN, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap(&tx, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p.M, p.K, k2NT / 2, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    rc = make_tmap(&tw, w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, p.N, p.K, k2TileN, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap(&tx, x, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, p.M, p.K, k2NT / 2, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (rc) return rc;
  constexpr int kKElems = KIND == k2I8 ? 128 : 64;
  const int n_tiles = ((SWIGLU ? p.n_out : p.N) + k2TileN - 1) / k2TileN;
  p.n_pairs = (n_tiles + 1) / 2;
  p.m_tiles = (p.M + k2NT - 1) / k2NT;
  p.band = p.m_tiles < 16 ? p.m_tiles : 16;
  p.kb_total = (p.K + kKElems - 1) / kKElems;
  const int items = p.n_pairs * p.m_tiles;
  const int max_clusters = kNumSMs / 2;
  const int grid = 2 * (items < max_clusters ? items : max_clusters);
  auto kern = gemm_tc2_kernel<KIND, SWIGLU>;
  TB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) k2SmemBytes));
  kern<<<grid, k2Threads, k2SmemBytes, stream>>>(tw, tx, p);
  return (int) cudaGetLastError();
}

// Called by tb_gemm_tc for prefill-size fp16 / int8 problems.  Returns -100 when the shape is not for this kernel.
int gemm_tc_pair(int kind, void* c, int out_type, const void* x, const void* w, const float* sc, const float* sr,
                 int sc_per_channel, int sr_per_token, const void* residual, int M, int N, int K, cudaStream_t stream) {
  if (kind != k2F16 && kind != k2I8) return -100;
  GemmTc2Params p{};
  p.c = c; p.out_type = out_type; p.residual = (const __half*) residual;
  p.sc = sc; p.sr = sr; p.sc_per_channel = sc_per_channel; p.sr_per_token = sr_per_token;
  p.M = M; p.N = N; p.K = K;
  return kind == k2F16 ? launch_gemm_tc2<k2F16>(p, x, w, stream) : launch_gemm_tc2<k2I8>(p, x, w, stream);
}

// C[M, N / 2] = silu(X . Wgate^T) * (X . Wup^T) with W = [gate rows; up rows] ([N, K]) — the gate / up projection of the
// GatedMLP (T/tensorrt_llm/layers/mlp.py:43-73) with SwiGLU in the epilogue; bit-identical to gemm_tc_pair + tb_swiglu.
int gemm_tc_pair_swiglu(int kind, void* c, const void* x, const void* w, const float* sc, const float* sr, int sc_per_channel,
                        int sr_per_token, int M, int N, int K, cudaStream_t stream) {
  if ((kind != k2F16 && kind != k2I8) || (N & 1)) return -100;
  GemmTc2Params p{};
  p.c = c; p.out_type = 0; p.residual = nullptr;
  p.sc = sc; p.sr = sr; p.sc_per_channel = sc_per_channel; p.sr_per_token = sr_per_token;
  p.M = M; p.N = N; p.K = K; p.n_out = N / 2;
  return kind == k2F16 ? launch_gemm_tc2<k2F16, true>(p, x, w, stream) : launch_gemm_tc2<k2I8, true>(p, x, w, stream);
}

}  // namespace tb
