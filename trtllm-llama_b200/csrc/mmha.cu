// Decode-step fused masked multi-head attention for sm_100a: RoPE(neox) on q,k of the new token,
// in-place KV-cache append (int8 quantised or fp16), QK^T . softmax . V over the cache with the
// sequence split across CTAs (split-L) and a deterministic in-order combine by the last CTA.
//
// Replaces (reference, K/ = T/cpp/tensorrt_llm/kernels/):
//   K/decoderMaskedMultiheadAttention/decoderMaskedMultiheadAttentionTemplate.h:1195-2183
//   K/gptKernels.cu:239-253 (updatePaddingCount, folded in: pad = max_input_len - input_lengths[b])
//   P/gptAttentionCommon/gptAttentionCommon.cpp:649-780 (enqueueGeneration dispatch)
//
// HBM-bound byte streaming (one query row per head, MHA: no reuse of K/V across heads), so SIMT with
// 16-byte coalesced loads, not tensor cores.  Algorithmic bytes per step and layer:
//   2 * H * B * L * Dh * b_kv   (K and V rows read once)  + B * 3*H*Dh*2 (qkv in) + B*H*Dh*2 (out).
// Grid = (H, B, nsplit) with nsplit chosen by the host so that H*B*nsplit covers >= 2 waves of 148 SMs.
//
// Numerics (documented deviations from the reference, all inside its test tolerance 2e-3,
// T/tests/attention/test_gpt_attention.py:828-831):
//   * int8 dequantisation scale is folded: s*(q.K_int8) instead of q.fp16(s*K_int8);
//   * with nsplit > 1 the probabilities are normalised after the combine (flash-decoding) rather
//     than before P.V; with nsplit == 1 the reference order (p * 1/(sum+1e-6) -> fp16 -> P.V) is kept.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace tb {

constexpr int kMmhaThreads = 256;
constexpr int kDh = 128;
constexpr int kMaxClusterSplits = 8;                 // portable thread-block-cluster limit

struct MmhaParams {
  const __half* qkv;         // [B, 3*H*Dh]
  void* kv_cache;            // [B, 2, H, S_max, Dh] int8 or fp16
  __half* out;               // [B, H*Dh]
  const int* seq_lens;       // [B] tlength per sequence (device); nullptr -> past_len for all
  const int* input_lengths;  // [B] real prompt lengths (device); nullptr -> no padding
  const int* masked_tokens;  // [B, S_max] optional
  const int* max_in_dev;     // optional device int: overrides max_input_len (one captured graph serves any prompt length)
  // paged KV cache (K/kvCacheUtils.h:34-112 KVBlockArray): block_ptrs [B, 2, max_blocks] device pointers (beam width 1) to
  // blocks laid out [H, tokens_per_block, Dh]; NULL: contiguous kv_cache [B, 2, H, S_max, Dh]
  const long long* block_ptrs;
  int tpb_log2, max_blocks;
  // beam search (BEAMS): cache_indirection [batch, beam, S_max] — cached position t of row (batch, beam) lives in the cache
  // of row (batch, cache_indir[batch][beam][t]) (decoderMaskedMultiheadAttentionTemplate.h:1137-1146,1624-1631)
  const int* cache_indir;
  int beam_width;
  const float* kv_scale_orig_quant;
  const float* kv_scale_quant_orig;
  float* partial;            // [B*H*nsplit*(Dh+2)] fp32 workspace
  int* counters;             // [B*H], zero on entry, zero on exit
  int past_len, max_input_len, S_max, H, rotary_dim;
  float inv_sqrt_dh;
  int pdl_trigger;           // let the next kernel (a PDL projection that requests weights before it waits) become resident now
};

template <bool INT8>
struct KvTraits {
  static constexpr int kLanesPerKey = INT8 ? 8 : 16;   // 16 bytes per lane
  static constexpr int kDimsPerLane = kDh / kLanesPerKey;
  static constexpr int kKeysPerIter = kMmhaThreads / kLanesPerKey;
};

// 16 bytes of one K/V row -> floats (int8: exact integer values, scale applied by the caller)
template <bool INT8>
__device__ __forceinline__ void unpack16(const uint4& r, float* f) {
  if constexpr (INT8) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t u = w[i] ^ 0x80808080u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // (b ^ 0x80) | 0x4B000000 is the float 2^23 + (b + 128); subtracting 8388736 is exact
        uint32_t bits;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(bits) : "r"(u), "r"(0x4B000000u), "r"(0x7650u + j));
        f[i * 4 + j] = __uint_as_float(bits) - 8388736.f;
      }
    }
  } else {
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}

__device__ __forceinline__ void mmha_mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                             uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t mmha_h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ uint32_t mmha_pack_h2(float a, float b) { return mmha_h2u(__floats2half2_rn(a, b)); }

// Measured and rejected (round 2): feeding the tensor-core loops from a per-lane cp.async ring (the K blocks, then the V blocks
// of a warp as one flat sequence, 4 blocks = 8 KB per warp ahead, started above the RoPE prologue and running through the
// softmax) — the form that took the tensor-core GEMV from 3.5 to 4.3 TB/s — made this kernel SLOWER at B = 8, L = 2047, int8:
// 29.9 -> 37.8 us.  One 16-key block per iteration leaves a warp 16 MMAs of work per wait, against two blocks' worth with
// the register-resident loads below, and the extra shared-memory round trip is not free at two CTAs per SM.
//
// Launch bounds.  fp16 caches: capped at 64 registers so that four CTAs share an SM (76 registers = three CTAs: the 512
// CTAs of cfg3 no longer fit one wave, 1332 -> 1428 tokens/s); the int8 FMA variant is at 64 already and loses with an
// explicit bound.
// MMA = true (int8 caches, long contexts): the two streaming loops run on tensor cores (mma.sync m16n8k16, fp32
// accumulate).  The FMA loops spend PRMT + FSUB + FFMA per cached element and are issue-bound at 2048-token contexts
// (ncu: smsp__issue_active 63 %, 3.4 TB/s); here an element costs 1.25 instructions of exact int8 -> fp16 expansion and
// the multiply-adds ride in the MMA.  Q.K^T: a 16-key block is the A operand (lane (g, t) loads 16 contiguous bytes of
// keys g and g + 8 per 64-dim step; the k order inside an MMA is a fixed permutation applied to q as well), q sits in
// column 0 of B.  P.V: A = V^T with lane g owning dims [16g, 16g + 16) (row g of MMA m = dim 16g + m, row g + 8 = dim
// 16g + 8 + m), so a lane loads 16 contiguous bytes of 4 keys; key pairs are interleaved with PRMT before the
// expansion; the fp16 probabilities sit in column 0 of B.
template <bool INT8, bool MMA = false, bool PAGED = false, bool BEAMS = false>
__global__ void __launch_bounds__(kMmhaThreads, MMA ? 2 : (INT8 ? 0 : 4)) mmha_decode_kernel(MmhaParams p) {
  static_assert(!MMA || INT8, "the tensor-core loops are for int8 caches");
  static_assert(!BEAMS || (!MMA && !PAGED), "beam search reads the contiguous cache through the FMA loops");
  using TR = KvTraits<INT8>;
  constexpr int LPK = TR::kLanesPerKey, DPL = TR::kDimsPerLane, KPI = TR::kKeysPerIter;
  constexpr int ELT = INT8 ? 1 : 2;
  extern __shared__ __align__(16) float smem[];
  __shared__ float q_s[kDh];
  __shared__ __align__(16) __half kcur_s[kDh];
  __shared__ __align__(16) __half vcur_s[kDh];
  __shared__ float red[2 * (kMmhaThreads / 32)];
  __shared__ float c_o[kMaxClusterSplits][kDh];      // split partial outputs, written by the cluster's CTAs into rank 0
  __shared__ float c_ml[kMaxClusterSplits][2];       // split (max, sum)

  if (p.pdl_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int h = blockIdx.x, b = blockIdx.y, split = blockIdx.z, nsplit = gridDim.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, hidden = H * kDh;
  const int tlen = p.seq_lens ? p.seq_lens[b] : p.past_len;              // positions [0, tlen) are cached
  const int max_in = p.max_in_dev ? p.max_in_dev[0] : p.max_input_len;
  const int in_len = p.input_lengths ? p.input_lengths[b] : max_in;
  const int pad = max_in - in_len;
  const int pos = tlen - pad;

  // balanced split of the cached positions, multiples of KPI
  int chunk = (tlen + nsplit - 1) / nsplit;
  chunk = (chunk + KPI - 1) / KPI * KPI;
  const int l0 = min(split * chunk, tlen), l1 = min(l0 + chunk, tlen);
  const int len = l1 - l0;
  const bool has_cur = (split == nsplit - 1);

  float* s_s = smem;                                   // [chunk + 1] scores / probabilities
  float* o_red = smem + ((chunk + 1 + 3) & ~3);        // [KPI][Dh] partial outputs

  const float kv_dq = INT8 ? p.kv_scale_quant_orig[0] : 1.f;
  const size_t seq_stride = (size_t) 2 * H * p.S_max * kDh * ELT;
  uint8_t* kbase = reinterpret_cast<uint8_t*>(p.kv_cache) + (size_t) b * seq_stride + (size_t) h * p.S_max * kDh * ELT;
  uint8_t* vbase = kbase + (size_t) H * p.S_max * kDh * ELT;
  // row of cached position t of this (sequence, head): contiguous cache, or KVBlockArray addressing — block t >> log2(tpb)
  // of the sequence's K (V) table, row (h * tpb + (t & (tpb - 1))) of the block (kvCacheUtils.h:58-112 getKVLocalIdx)
  const long long* ktab = PAGED ? p.block_ptrs + (size_t) b * 2 * p.max_blocks : nullptr;
  // BEAMS: cached positions come from the row the indirection names; the appended position tlen is the row's own
  const int* indir = BEAMS ? p.cache_indir + (size_t) b * p.S_max : nullptr;
  const int beam0 = BEAMS ? b / p.beam_width * p.beam_width : 0;
  auto beam_shift = [&](int t) -> ptrdiff_t {
    return t < tlen ? ((ptrdiff_t) (beam0 + indir[t]) - b) * (ptrdiff_t) seq_stride : 0;
  };
  auto krow = [&](int t) -> uint8_t* {
    if constexpr (PAGED)
      return reinterpret_cast<uint8_t*>(ktab[t >> p.tpb_log2]) +
             ((size_t) (h << p.tpb_log2) + (t & ((1 << p.tpb_log2) - 1))) * kDh * ELT;
    else if constexpr (BEAMS)
      return kbase + beam_shift(t) + (size_t) t * kDh * ELT;
    else
      return kbase + (size_t) t * kDh * ELT;
  };
  auto vrow = [&](int t) -> uint8_t* {
    if constexpr (PAGED)
      return reinterpret_cast<uint8_t*>(ktab[p.max_blocks + (t >> p.tpb_log2)]) +
             ((size_t) (h << p.tpb_log2) + (t & ((1 << p.tpb_log2) - 1))) * kDh * ELT;
    else if constexpr (BEAMS)
      return vbase + beam_shift(t) + (size_t) t * kDh * ELT;
    else
      return vbase + (size_t) t * kDh * ELT;
  };

  // Long fp16 contexts: pull this CTA's K and V ranges into L2 up front.  The load loops keep 32 KB per CTA in flight,
  // short of what HBM needs at 2048-token contexts, and the V range is not touched until the Q.K^T pass and the softmax
  // are done (cfg3, fp16 KV: 1260 -> 1295 tokens/s).  The int8 variant is issue-bound on the dequantisation (3
  // instructions per element), not on memory: the same prefetch costs it 2 %, so it is compiled out there.
  if constexpr (!INT8 && !PAGED && !BEAMS) {
    if (len >= 256) {
      const uint32_t range = (uint32_t) len * kDh * ELT, piece = 16384;
      const uint32_t npiece = (range + piece - 1) / piece;
      if (tid < 2 * npiece) {
        const uint32_t i = tid >= npiece ? tid - npiece : tid;
        const uint8_t* src = (tid >= npiece ? vbase : kbase) + (size_t) l0 * kDh * ELT + (size_t) i * piece;
        const uint32_t bytes = min(piece, range - i * piece);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
      }
    }
  }

  // ---- RoPE on q (all CTAs) and k (last split), cache append (last split) --------------------
  const __half* qrow = p.qkv + (size_t) b * 3 * hidden + (size_t) h * kDh;
  if (tid < kDh / 2) {
    const int half_rot = p.rotary_dim / 2;
    float c = 1.f, s = 0.f;
    if (tid < half_rot) {
      // inv_freq = t / pow(10000, 2j/rot)  (decoderMaskedMultiheadAttentionUtils.h:1511-1515)
      const float ang = (float) pos / powf(10000.0f, (2 * tid) / (float) p.rotary_dim);
      c = cosf(ang);
      s = sinf(ang);
    }
    const int i0 = tid < half_rot ? tid : 2 * tid - half_rot + 0, i1 = tid < half_rot ? tid + half_rot : i0 + 1;
    // (for rotary_dim == Dh every thread owns the pair (j, j + Dh/2); otherwise the non-rotated
    //  tail dims are passed through pairwise with c = 1, s = 0)
    const float qa = __half2float(qrow[i0]), qb = __half2float(qrow[i1]);
    q_s[i0] = __half2float(__float2half_rn(c * qa - s * qb));
    q_s[i1] = __half2float(__float2half_rn(c * qb + s * qa));
    if (has_cur) {
      const float ka = __half2float(qrow[hidden + i0]), kb = __half2float(qrow[hidden + i1]);
      kcur_s[i0] = __float2half_rn(c * ka - s * kb);
      kcur_s[i1] = __float2half_rn(c * kb + s * ka);
      vcur_s[i0] = qrow[2 * hidden + i0];
      vcur_s[i1] = qrow[2 * hidden + i1];
    }
  }
  __syncthreads();
  if (has_cur && tid < kDh / 8) {
    // append 8 dims per thread: K[t] <- k, V[t] <- v (int8: cvt.rni.sat(x * scale))
    const int d0 = tid * 8;
    if constexpr (INT8) {
      const float qs = p.kv_scale_orig_quant[0];
      uint2 kq, vq;
      kq.x = pack4_i8(__half2float(kcur_s[d0]) * qs, __half2float(kcur_s[d0 + 1]) * qs,
                      __half2float(kcur_s[d0 + 2]) * qs, __half2float(kcur_s[d0 + 3]) * qs);
      kq.y = pack4_i8(__half2float(kcur_s[d0 + 4]) * qs, __half2float(kcur_s[d0 + 5]) * qs,
                      __half2float(kcur_s[d0 + 6]) * qs, __half2float(kcur_s[d0 + 7]) * qs);
      vq.x = pack4_i8(__half2float(vcur_s[d0]) * qs, __half2float(vcur_s[d0 + 1]) * qs,
                      __half2float(vcur_s[d0 + 2]) * qs, __half2float(vcur_s[d0 + 3]) * qs);
      vq.y = pack4_i8(__half2float(vcur_s[d0 + 4]) * qs, __half2float(vcur_s[d0 + 5]) * qs,
                      __half2float(vcur_s[d0 + 6]) * qs, __half2float(vcur_s[d0 + 7]) * qs);
      *reinterpret_cast<uint2*>(krow(tlen) + d0) = kq;
      *reinterpret_cast<uint2*>(vrow(tlen) + d0) = vq;
    } else {
      *reinterpret_cast<uint4*>(krow(tlen) + d0 * 2) = *reinterpret_cast<uint4*>(&kcur_s[d0]);
      *reinterpret_cast<uint4*>(vrow(tlen) + d0 * 2) = *reinterpret_cast<uint4*>(&vcur_s[d0]);
    }
  }

  // ---- Q.K^T over [l0, l1) -------------------------------------------------------------------
  const int grp = tid / LPK, gl = tid % LPK;
  float qreg[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) qreg[i] = q_s[gl * DPL + i];
  const float qk_scale = kv_dq * p.inv_sqrt_dh;
  const int* mrow = p.masked_tokens ? p.masked_tokens + (size_t) b * p.S_max : nullptr;

  float lmax = -3.0e38f;
  constexpr int UN = INT8 ? 4 : 8;   // 16-byte loads in flight per thread (the int8 variant is register-bound at 64)
  if constexpr (MMA) {
    const int g = lane >> 2, t = lane & 3;
    // q as the B operand of the 8 MMAs of a key block: column 0 only (lanes with g == 0), same k permutation as A
    uint32_t bq[2][4][2];
#pragma unroll
    for (int st = 0; st < 2; ++st)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = 64 * st + 16 * t + 4 * j;
        bq[st][j][0] = g == 0 ? mmha_pack_h2(q_s[d], q_s[d + 1]) : 0u;
        bq[st][j][1] = g == 0 ? mmha_pack_h2(q_s[d + 2], q_s[d + 3]) : 0u;
      }
    const int nblk = (len + 15) >> 4;
    for (int kb = warp; kb < nblk; kb += 2 * (kMmhaThreads / 32)) {   // two key blocks in flight per warp
      uint4 lo[2][2], hi[2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int k0 = (kb + u * (kMmhaThreads / 32)) * 16;
        const bool okl = k0 + g < len, okh = k0 + g + 8 < len;
        // keys g and g + 8 of the 16-key block (a block of the paged cache holds a multiple of 16 positions)
        const uint8_t* rl = krow(l0 + k0 + g) + t * 16;
        const uint8_t* rh = PAGED ? krow(l0 + k0 + g + 8) + t * 16 : rl + 8 * kDh;
#pragma unroll
        for (int st = 0; st < 2; ++st) {
          lo[u][st] = okl ? ldg_nc_v4(rl + 64 * st) : make_uint4(0, 0, 0, 0);
          hi[u][st] = okh ? ldg_nc_v4(rh + 64 * st) : make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int k0 = (kb + u * (kMmhaThreads / 32)) * 16;
        if (k0 < len) {                                                 // warp-uniform
          float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int st = 0; st < 2; ++st) {
            const uint32_t wl[4] = {lo[u][st].x, lo[u][st].y, lo[u][st].z, lo[u][st].w};
            const uint32_t wh[4] = {hi[u][st].x, hi[u][st].y, hi[u][st].z, hi[u][st].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __half2 l0h, l1h, h0h, h1h;
              i8x4_to_h2x2(wl[j], l0h, l1h);
              i8x4_to_h2x2(wh[j], h0h, h1h);
              mmha_mma_f16(c, mmha_h2u(l0h), mmha_h2u(h0h), mmha_h2u(l1h), mmha_h2u(h1h), bq[st][j][0], bq[st][j][1]);
            }
          }
          if (t == 0) {                                                 // column 0: c[0] = key g, c[2] = key g + 8
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int ii = k0 + g + 8 * hh;
              if (ii < len) {
                float d = c[2 * hh] * qk_scale;
                const bool masked = mrow ? (mrow[l0 + ii] != 0) : (l0 + ii >= in_len && l0 + ii < max_in);
                if (masked) d = -3.0e38f;
                s_s[ii] = d;
                lmax = fmaxf(lmax, d);
              }
            }
          }
        }
      }
    }
  } else
  for (int i = grp; i - grp < len; i += KPI * UN) {   // trip count uniform across the warp (shuffles)
    uint4 raw[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int ii = i + u * KPI;
      raw[u] = make_uint4(0, 0, 0, 0);
      if (ii < len) raw[u] = ldg_nc_v4(krow(l0 + ii) + (size_t) gl * DPL * ELT);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int ii = i + u * KPI;
      float kf[DPL];
      unpack16<INT8>(raw[u], kf);
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < DPL; ++j) d = fmaf(qreg[j], kf[j], d);
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (ii < len && gl == 0) {
        d *= qk_scale;
        const bool masked = mrow ? (mrow[l0 + ii] != 0) : (l0 + ii >= in_len && l0 + ii < max_in);
        if (masked) d = -3.0e38f;
        s_s[ii] = d;
        lmax = fmaxf(lmax, d);
      }
    }
  }
  if (has_cur && warp == 0) {
    // current token: unquantised k (decoderMaskedMultiheadAttentionTemplate.h:1511-1549)
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) d = fmaf(q_s[lane * 4 + j], __half2float(kcur_s[lane * 4 + j]), d);
    d = warp_sum(d) * p.inv_sqrt_dh;
    if (lane == 0) {
      s_s[len] = d;
      lmax = fmaxf(lmax, d);
    }
  }
  const int n_s = len + (has_cur ? 1 : 0);

  // ---- softmax statistics ---------------------------------------------------------------------
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float m_s = red[0];
#pragma unroll
  for (int w = 1; w < kMmhaThreads / 32; ++w) m_s = fmaxf(m_s, red[w]);
  float lsum = 0.f;
  for (int i = tid; i < n_s; i += kMmhaThreads) {
    const float sv = s_s[i];
    const float e = sv <= -1.0e38f ? 0.f : __expf(sv - m_s);
    s_s[i] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[kMmhaThreads / 32 + warp] = lsum;
  __syncthreads();
  float l_s = 0.f;
#pragma unroll
  for (int w = 0; w < kMmhaThreads / 32; ++w) l_s += red[kMmhaThreads / 32 + w];
  const float inv_sum = nsplit == 1 ? __fdividef(1.f, l_s + 1.e-6f) : 1.f;
  for (int i = tid; i < n_s; i += kMmhaThreads)  // p -> fp16 (Template.h:1765 / :1772)
    s_s[i] = __half2float(__float2half_rn(s_s[i] * inv_sum));
  __syncthreads();

  // ---- P.V --------------------------------------------------------------------------------------
  if constexpr (MMA) {
    const int g = lane >> 2, t = lane & 3;
    float oacc[8][4];
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) oacc[m][i] = 0.f;
    const int nblk = (len + 15) >> 4;
    // keys of this lane's k slots {2t, 2t+1, 2t+8, 2t+9}: 16 contiguous dims [16g, 16g+16) of each; the next block's
    // rows are requested before the current block is expanded (two blocks = 8 loads in flight per lane)
    auto load_block = [&](int kb, uint4 (&r)[4], float (&pk)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ii = kb * 16 + 2 * t + (i & 1) + 8 * (i >> 1);
        const bool ok = kb < nblk && ii < len;
        r[i] = ok ? ldg_nc_v4(vrow(l0 + ii) + g * 16) : make_uint4(0, 0, 0, 0);
        pk[i] = ok ? s_s[ii] : 0.f;
      }
    };
    uint4 r[4], rn[4];
    float pk[4], pkn[4];
    load_block(warp, r, pk);
    for (int kb = warp; kb < nblk; kb += kMmhaThreads / 32) {
      load_block(kb + kMmhaThreads / 32, rn, pkn);
      const uint32_t b0 = g == 0 ? mmha_pack_h2(pk[0], pk[1]) : 0u, b1 = g == 0 ? mmha_pack_h2(pk[2], pk[3]) : 0u;
      const uint32_t w0[4] = {r[0].x, r[0].y, r[0].z, r[0].w}, w1[4] = {r[1].x, r[1].y, r[1].z, r[1].w};
      const uint32_t w2[4] = {r[2].x, r[2].y, r[2].z, r[2].w}, w3[4] = {r[3].x, r[3].y, r[3].z, r[3].w};
      // word w holds dims 4w..4w+3 (rows g of MMAs 4w..4w+3), word w + 2 dims 8+4w.. (rows g + 8 of the same MMAs)
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        // e01[x][d] / e23[x][d]: half2 {V[k_even][dim], V[k_odd][dim]}, x = 0: dim 16g + 4w + d, x = 1: dim 16g + 8 + 4w + d
        uint32_t e01[2][4], e23[2][4];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const int ww = w + 2 * x;
          uint32_t ga, gb;
          __half2 x0, x1;
          asm("prmt.b32 %0, %1, %2, 0x5140;" : "=r"(ga) : "r"(w0[ww]), "r"(w1[ww]));   // {k0.b0, k1.b0, k0.b1, k1.b1}
          asm("prmt.b32 %0, %1, %2, 0x7362;" : "=r"(gb) : "r"(w0[ww]), "r"(w1[ww]));   // {k0.b2, k1.b2, k0.b3, k1.b3}
          i8x4_to_h2x2(ga, x0, x1); e01[x][0] = mmha_h2u(x0); e01[x][1] = mmha_h2u(x1);
          i8x4_to_h2x2(gb, x0, x1); e01[x][2] = mmha_h2u(x0); e01[x][3] = mmha_h2u(x1);
          asm("prmt.b32 %0, %1, %2, 0x5140;" : "=r"(ga) : "r"(w2[ww]), "r"(w3[ww]));
          asm("prmt.b32 %0, %1, %2, 0x7362;" : "=r"(gb) : "r"(w2[ww]), "r"(w3[ww]));
          i8x4_to_h2x2(ga, x0, x1); e23[x][0] = mmha_h2u(x0); e23[x][1] = mmha_h2u(x1);
          i8x4_to_h2x2(gb, x0, x1); e23[x][2] = mmha_h2u(x0); e23[x][3] = mmha_h2u(x1);
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) mmha_mma_f16(oacc[4 * w + d], e01[0][d], e01[1][d], e23[0][d], e23[1][d], b0, b1);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { r[i] = rn[i]; pk[i] = pkn[i]; }
    }
    // column 0 (lanes t == 0): oacc[m][0] = dim 16g + m, oacc[m][2] = dim 16g + 8 + m; one partial row per warp
    if (t == 0) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        o_red[warp * kDh + 16 * g + m] = oacc[m][0] * kv_dq;
        o_red[warp * kDh + 16 * g + 8 + m] = oacc[m][2] * kv_dq;
      }
    }
  } else {
  float acc[DPL];
#pragma unroll
  for (int j = 0; j < DPL; ++j) acc[j] = 0.f;
  for (int i = grp; i - grp < len; i += KPI * UN) {
    uint4 raw[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int ii = i + u * KPI;
      raw[u] = make_uint4(0, 0, 0, 0);
      if (ii < len) raw[u] = ldg_nc_v4(vrow(l0 + ii) + (size_t) gl * DPL * ELT);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int ii = i + u * KPI;
      const float pv = ii < len ? s_s[ii] : 0.f;
      float vf[DPL];
      unpack16<INT8>(raw[u], vf);
#pragma unroll
      for (int j = 0; j < DPL; ++j) acc[j] = fmaf(pv, vf[j], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < DPL; ++j) o_red[grp * kDh + gl * DPL + j] = acc[j] * kv_dq;
  }
  __syncthreads();
  constexpr int NRED = MMA ? kMmhaThreads / 32 : KPI;    // partial output rows in o_red

  if (nsplit == 1) {
    if (tid < kDh) {
      float o = 0.f;
#pragma unroll 8
      for (int g = 0; g < NRED; ++g) o += o_red[g * kDh + tid];
      if (has_cur) o = fmaf(s_s[len], __half2float(vcur_s[tid]), o);
      p.out[(size_t) b * hidden + h * kDh + tid] = __float2half_rn(o);
    }
    return;
  }

  // ---- split-L combine inside the thread-block cluster: every CTA of the (b, h) cluster stores its partial
  // (o, max, sum) into rank 0's shared memory through DSMEM; after the cluster barrier rank 0 merges the splits
  // in index order (deterministic, no atomics, no global round trip) -------------------------------------------
  if (tid < kDh) {
    float o = 0.f;
#pragma unroll 8
    for (int g = 0; g < NRED; ++g) o += o_red[g * kDh + tid];
    if (has_cur) o = fmaf(s_s[len], __half2float(vcur_s[tid]), o);
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&c_o[split][tid])), "r"(0));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(o) : "memory");
    if (tid < 2) {
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&c_ml[split][tid])), "r"(0));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(tid == 0 ? m_s : l_s) : "memory");
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (split != 0) return;
  if (tid < kDh) {
    float gm = -3.0e38f;
    for (int s = 0; s < nsplit; ++s) gm = fmaxf(gm, c_ml[s][0]);
    float o = 0.f, l = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float w = __expf(c_ml[s][0] - gm);
      o = fmaf(w, c_o[s][tid], o);
      l = fmaf(w, c_ml[s][1], l);
    }
    p.out[(size_t) b * hidden + h * kDh + tid] = __float2half_rn(o * __fdividef(1.f, l + 1.e-6f));
  }
}

}  // namespace tb

using namespace tb;

static int g_mma_mode = getenv("TB_MMHA_MMA") ? atoi(getenv("TB_MMHA_MMA")) : -1;

extern "C" {

int tb_mmha_set_mode(int mode) {
  const int prev = g_mma_mode;
  g_mma_mode = mode < -1 ? -1 : (mode > 1 ? 1 : mode);
  return prev;
}

size_t tb_mmha_workspace_bytes(int batch, int num_heads, int max_splits) {
  return (size_t) batch * num_heads * max_splits * (kDh + 2) * sizeof(float) + 256;
}
size_t tb_mmha_counter_bytes(int batch, int num_heads) { return (size_t) batch * num_heads * sizeof(int); }

// split count: enough CTAs for >= 2 waves of 148 SMs, at least 64 cached keys per split
int tb_mmha_num_splits(int batch, int num_heads, int len_hint, int max_splits) {
  static const int forced = getenv("TB_MMHA_NSPLIT") ? atoi(getenv("TB_MMHA_NSPLIT")) : 0;   // A/B switch
  if (forced > 0) return forced > kMaxClusterSplits ? kMaxClusterSplits : forced;
  const int base = batch * num_heads;
  int want = (2 * kNumSMs + base - 1) / base;
  int by_len = len_hint / 64;
  if (by_len < 1) by_len = 1;
  int n = want < by_len ? want : by_len;
  if (n > max_splits) n = max_splits;
  if (n > kMaxClusterSplits) n = kMaxClusterSplits;   // the splits of one (b, h) form a thread-block cluster
  if (n < 1) n = 1;
  return n;
}

int tb_mmha_decode(void* out, const void* qkv, void* kv_cache, const int* seq_lens, const int* input_lengths,
                   const int* masked_tokens, const float* kv_scale_orig_quant, const float* kv_scale_quant_orig,
                   void* workspace, int* counters, int batch, int num_heads, int head_size, int max_seq_len, int past_len,
                   int max_input_len, int len_cap, int rotary_dim, float q_scaling, int int8_kv, int nsplit,
                   cudaStream_t stream) {
  return tb_mmha_decode_dev(out, qkv, kv_cache, seq_lens, input_lengths, masked_tokens, nullptr, kv_scale_orig_quant,
                            kv_scale_quant_orig, workspace, counters, batch, num_heads, head_size, max_seq_len, past_len,
                            max_input_len, len_cap, rotary_dim, q_scaling, int8_kv, nsplit, stream);
}

static int mmha_launch(void* out, const void* qkv, void* kv_cache, const long long* block_ptrs, int tokens_per_block,
                       int max_blocks, const int* seq_lens, const int* input_lengths, const int* masked_tokens,
                       const int* max_input_len_dev, const float* kv_scale_orig_quant, const float* kv_scale_quant_orig,
                       void* workspace, int* counters, int batch, int num_heads, int head_size, int max_seq_len,
                       int past_len, int max_input_len, int len_cap, int rotary_dim, float q_scaling, int int8_kv,
                       int nsplit, cudaStream_t stream, const int* cache_indir = nullptr, int beam_width = 1);

int tb_mmha_decode_dev(void* out, const void* qkv, void* kv_cache, const int* seq_lens, const int* input_lengths,
                       const int* masked_tokens, const int* max_input_len_dev, const float* kv_scale_orig_quant,
                       const float* kv_scale_quant_orig, void* workspace, int* counters, int batch, int num_heads,
                       int head_size, int max_seq_len, int past_len, int max_input_len, int len_cap, int rotary_dim,
                       float q_scaling, int int8_kv, int nsplit, cudaStream_t stream) {
  return mmha_launch(out, qkv, kv_cache, nullptr, 0, 0, seq_lens, input_lengths, masked_tokens, max_input_len_dev,
                     kv_scale_orig_quant, kv_scale_quant_orig, workspace, counters, batch, num_heads, head_size, max_seq_len,
                     past_len, max_input_len, len_cap, rotary_dim, q_scaling, int8_kv, nsplit, stream);
}

int tb_mmha_decode_beams(void* out, const void* qkv, void* kv_cache, const int* cache_indirection, int beam_width,
                         const int* seq_lens, const int* input_lengths, const int* masked_tokens, const int* max_input_len_dev,
                         const float* kv_scale_orig_quant, const float* kv_scale_quant_orig, int batch, int num_heads,
                         int head_size, int max_seq_len, int past_len, int max_input_len, int len_cap, int rotary_dim,
                         float q_scaling, int int8_kv, int nsplit, cudaStream_t stream) {
  if (!cache_indirection || beam_width < 1) return -1;
  return mmha_launch(out, qkv, kv_cache, nullptr, 0, 0, seq_lens, input_lengths, masked_tokens, max_input_len_dev,
                     kv_scale_orig_quant, kv_scale_quant_orig, nullptr, nullptr, batch, num_heads, head_size, max_seq_len,
                     past_len, max_input_len, len_cap, rotary_dim, q_scaling, int8_kv, nsplit, stream, cache_indirection,
                     beam_width);
}

int tb_mmha_decode_paged(void* out, const void* qkv, const int64_t* block_pointers, int tokens_per_block,
                         int max_blocks_per_seq, const int* seq_lens, const int* input_lengths, const int* masked_tokens,
                         const int* max_input_len_dev, const float* kv_scale_orig_quant, const float* kv_scale_quant_orig,
                         int batch, int num_heads, int head_size, int past_len, int max_input_len, int len_cap,
                         int rotary_dim, float q_scaling, int int8_kv, int nsplit, cudaStream_t stream) {
  if (!block_pointers || tokens_per_block < 16 || (tokens_per_block & (tokens_per_block - 1)) || max_blocks_per_seq < 1) return -1;
  const int max_seq_len = tokens_per_block * max_blocks_per_seq;
  return mmha_launch(out, qkv, nullptr, reinterpret_cast<const long long*>(block_pointers), tokens_per_block,
                     max_blocks_per_seq, seq_lens, input_lengths, masked_tokens, max_input_len_dev, kv_scale_orig_quant,
                     kv_scale_quant_orig, nullptr, nullptr, batch, num_heads, head_size, max_seq_len, past_len,
                     max_input_len, len_cap, rotary_dim, q_scaling, int8_kv, nsplit, stream);
}
}   // extern "C"

template <bool INT8, bool MMA>
static int mmha_launch_t(const cudaLaunchConfig_t& cfg, const MmhaParams& p, size_t smem, bool paged) {
  if constexpr (!MMA) {
    if (p.cache_indir) {
      if (smem > 48 * 1024) TB_CHECK_CUDA(cudaFuncSetAttribute(mmha_decode_kernel<INT8, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      return (int) cudaLaunchKernelEx(&cfg, mmha_decode_kernel<INT8, false, false, true>, p);
    }
  }
  if (paged) {
    if (smem > 48 * 1024) TB_CHECK_CUDA(cudaFuncSetAttribute(mmha_decode_kernel<INT8, MMA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    return (int) cudaLaunchKernelEx(&cfg, mmha_decode_kernel<INT8, MMA, true>, p);
  }
  if (smem > 48 * 1024) TB_CHECK_CUDA(cudaFuncSetAttribute(mmha_decode_kernel<INT8, MMA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  return (int) cudaLaunchKernelEx(&cfg, mmha_decode_kernel<INT8, MMA, false>, p);
}

static int mmha_launch(void* out, const void* qkv, void* kv_cache, const long long* block_ptrs, int tokens_per_block,
                       int max_blocks, const int* seq_lens, const int* input_lengths, const int* masked_tokens,
                       const int* max_input_len_dev, const float* kv_scale_orig_quant, const float* kv_scale_quant_orig,
                       void* workspace, int* counters, int batch, int num_heads, int head_size, int max_seq_len,
                       int past_len, int max_input_len, int len_cap, int rotary_dim, float q_scaling, int int8_kv,
                       int nsplit, cudaStream_t stream, const int* cache_indir, int beam_width) {
  if (head_size != kDh) return -1;
  if (cache_indir && (block_ptrs || beam_width < 1 || batch % beam_width)) return -1;   // beams read the contiguous cache                 // LLaMA-7B head size; other sizes are not built
  if (rotary_dim != 0 && rotary_dim != kDh) return -1;
  if (past_len + 1 > max_seq_len || len_cap + 1 > max_seq_len + 1) return -2;
  if (int8_kv && (!kv_scale_orig_quant || !kv_scale_quant_orig)) return -1;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > kMaxClusterSplits) return -1;
  (void) workspace; (void) counters;   // kept in the ABI: split partials now live in distributed shared memory
  MmhaParams p{};
  p.qkv = (const __half*) qkv; p.kv_cache = kv_cache; p.out = (__half*) out; p.seq_lens = seq_lens;
  p.input_lengths = input_lengths; p.masked_tokens = masked_tokens; p.max_in_dev = max_input_len_dev;
  p.kv_scale_orig_quant = kv_scale_orig_quant; p.kv_scale_quant_orig = kv_scale_quant_orig;
  p.counters = counters;
  p.partial = reinterpret_cast<float*>(workspace);
  p.past_len = past_len; p.max_input_len = max_input_len; p.S_max = max_seq_len; p.H = num_heads;
  p.rotary_dim = rotary_dim; p.inv_sqrt_dh = 1.f / (sqrtf((float) head_size) * q_scaling);
  p.block_ptrs = block_ptrs; p.max_blocks = max_blocks;
  p.cache_indir = cache_indir; p.beam_width = beam_width;
  // The projection that follows is launched with programmatic stream serialisation and requests its first weights before it
  // waits for this kernel: triggering at entry lets its CTAs become resident on the SMs this grid leaves free (same-run A/B,
  // step ms off -> on: int4 B=1 1.692 -> 1.627, W8 B=1 1.878 -> 1.801, cfg3 int8-KV 3.03 -> 2.95, cfg2 2.62 -> 2.60).
  // (Launching THIS kernel with programmatic serialisation too — rotary angles above a griddepcontrol.wait — measured slower
  // again with the ring GEMV in front of it: cfg2 2.50 -> 2.53 ms, SmoothQuant 1.66 -> 1.72, int4 1.63 -> 1.68.)
  static const int pdl_env = getenv("TB_MMHA_PDL") ? atoi(getenv("TB_MMHA_PDL")) : 1;   // A/B switch
  p.pdl_trigger = pdl_env;
  p.tpb_log2 = 0;
  while (block_ptrs && (1 << p.tpb_log2) < tokens_per_block) ++p.tpb_log2;
  const bool paged = block_ptrs != nullptr;
  // int8 caches with long contexts: tensor-core loops, two CTAs per SM, so no more splits than fit one wave
  const int mma_env = g_mma_mode;   // A/B switch (TB_MMHA_MMA / tb_mmha_set_mode): 0 off, 1 always, -1 automatic
  const bool use_mma = int8_kv && !cache_indir && (mma_env == 1 || (mma_env != 0 && len_cap >= 512));
  static const bool nofit = getenv("TB_MMHA_NOFIT") && atoi(getenv("TB_MMHA_NOFIT")) != 0;   // A/B switch
  if (use_mma && !nofit) {
    const int fit = (2 * kNumSMs) / (batch * num_heads);
    if (nsplit > fit) nsplit = fit < 1 ? 1 : fit;
  }
  // shared memory sized for the longest possible split (len_cap = upper bound on any tlength)
  const int kpi = int8_kv ? KvTraits<true>::kKeysPerIter : KvTraits<false>::kKeysPerIter;
  int chunk = (len_cap + nsplit - 1) / nsplit;
  chunk = (chunk + kpi - 1) / kpi * kpi;
  const size_t smem = ((size_t) ((chunk + 1 + 3) & ~3) + (size_t) kpi * kDh) * sizeof(float);
  if (smem > 200 * 1024) return -3;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(num_heads, batch, nsplit);
  cfg.blockDim = dim3(kMmhaThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = nsplit;
  cfg.attrs = attr;
  cfg.numAttrs = nsplit > 1 ? 1 : 0;
  if (use_mma) return mmha_launch_t<true, true>(cfg, p, smem, paged);
  if (int8_kv) return mmha_launch_t<true, false>(cfg, p, smem, paged);
  return mmha_launch_t<false, false>(cfg, p, smem, paged);
}
