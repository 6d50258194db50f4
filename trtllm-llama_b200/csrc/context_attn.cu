// Context-phase (prefill) attention: one fused flash-style causal kernel in place of the reference's
// unfused path, plus the RoPE / KV-cache-write pre-pass.
//
// Replaces (reference, K/ = T/cpp/tensorrt_llm/kernels/, P/ = .../plugins/):
//   P/gptAttentionCommon/gptAttentionCommon.cpp:361-620  enqueueContext (unfused: 7 launches,
//       B*H*S*S fp32 score buffer + fp16 copy materialised in HBM)
//   K/unfusedAttentionKernels.cu:1252-1424  add_fusedQKV_bias_transpose_kernel (RoPE, split)
//   K/unfusedAttentionKernels.cu:1553-1646  transpose4dBatchMajorKVCache (cache write, int8 quant)
//   K/unfusedAttentionKernels.cu:180-257    softmax_kernel ; K/gptKernels.cu:136-199 mask
//
// ctx_prep_kernel   : RoPE(neox, position = index in sequence) on q,k written back in place into
//                     the packed QKV activations (the reference does the same, :1401-1403), K/V
//                     appended to the cache [B,2,H,S_max,Dh] (int8: cvt.rni.sat(x*scale)); padded
//                     rows are stored as zeros.
// flash_ctx_kernel  : S = QK^T (fp32 accum) -> causal+length mask -> online softmax -> P (fp16) . V,
//                     never materialising the score matrix.  Tensor-core mma (m16n8k16 f16, fp32
//                     accumulate).  Round-1 kernel: warp-level MMA; the tcgen05/TMEM port of this
//                     kernel is the next step for the prefill config (DESIGN.md "next").
// Algorithmic traffic per layer: B*S*4*H*Dh*2 bytes (qkv in + out) instead of the reference's
// additional ~6 * B*H*S^2 bytes of score traffic.
#include "common.cuh"
#include "kernels.h"

namespace tb {

constexpr int kD = 128;

// ------------------------------------------------------------------------------------------------
// One CTA per token: the rotary angles depend on (position, pair) only, so cos / sin are computed once (64 threads) and
// reused by all heads; then 256 threads sweep the (head, pair) items of the row.
constexpr int kPrepThreads = 256;
__global__ void __launch_bounds__(kPrepThreads) ctx_prep_kernel(__half* qkv, void* kv_cache, const int* input_lengths,
                                                                const float* kv_scale_orig_quant, int S, int H, int S_max,
                                                                int rotary_dim, int int8_kv, const long long* block_ptrs,
                                                                int tpb_log2, int max_blocks) {
  __shared__ float cs[kD / 2], sn_s[kD / 2];
  const int tok = blockIdx.x;
  const int b = tok / S, s = tok % S;
  const int hidden = H * kD;
  const int len = input_lengths ? input_lengths[b] : S;
  const bool valid = s < len;
  const int half_rot = rotary_dim / 2;
  if (threadIdx.x < kD / 2) {
    float c = 1.f, sn = 0.f;
    if ((int) threadIdx.x < half_rot) {
      const float ang = (float) s / powf(10000.0f, (2 * threadIdx.x) / (float) rotary_dim);
      c = cosf(ang);
      sn = sinf(ang);
    }
    cs[threadIdx.x] = c;
    sn_s[threadIdx.x] = sn;
  }
  __syncthreads();
  const size_t elt = int8_kv ? 1 : 2;
  const float qs = int8_kv ? kv_scale_orig_quant[0] : 1.f;
  // item = (head, group of 8 pairs): pairs (t, t + 64) for t = 8g .. 8g+7 are two 16-byte chunks of the head row.
  // rotary_dim is 0 (cos = 1, sin = 0: identity, any pairing) or 128 (neox halves), checked by the host.
  for (int item = threadIdx.x; item < H * 8; item += kPrepThreads) {
    const int h = item >> 3, g8 = item & 7;
    __half* row = qkv + (size_t) tok * 3 * hidden + (size_t) h * kD;
    uint4 klo = make_uint4(0, 0, 0, 0), khi = klo, vlo = klo, vhi = klo;
    if (valid) {
      auto rotate = [&](__half* base, uint4& lo_out, uint4& hi_out) {
        uint4 lo = *reinterpret_cast<const uint4*>(base + g8 * 8), hi = *reinterpret_cast<const uint4*>(base + 64 + g8 * 8);
        __half* a = reinterpret_cast<__half*>(&lo);
        __half* bq = reinterpret_cast<__half*>(&hi);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float c = cs[g8 * 8 + i], sn = sn_s[g8 * 8 + i];
          const float xa = __half2float(a[i]), xb = __half2float(bq[i]);
          a[i] = __float2half_rn(c * xa - sn * xb);
          bq[i] = __float2half_rn(c * xb + sn * xa);
        }
        *reinterpret_cast<uint4*>(base + g8 * 8) = lo;
        *reinterpret_cast<uint4*>(base + 64 + g8 * 8) = hi;
        lo_out = lo;
        hi_out = hi;
      };
      uint4 qlo, qhi;
      rotate(row, qlo, qhi);
      rotate(row + hidden, klo, khi);
      vlo = *reinterpret_cast<const uint4*>(row + 2 * hidden + g8 * 8);
      vhi = *reinterpret_cast<const uint4*>(row + 2 * hidden + 64 + g8 * 8);
    }
    uint8_t *kc, *vc;
    if (block_ptrs) {   // paged cache (K/kvCacheUtils.h:34-112): block s >> log2(tpb), row h * tpb + (s & (tpb - 1))
      const size_t row = ((size_t) (h << tpb_log2) + (s & ((1 << tpb_log2) - 1))) * kD * elt;
      kc = reinterpret_cast<uint8_t*>(block_ptrs[(size_t) b * 2 * max_blocks + (s >> tpb_log2)]) + row;
      vc = reinterpret_cast<uint8_t*>(block_ptrs[(size_t) b * 2 * max_blocks + max_blocks + (s >> tpb_log2)]) + row;
    } else {
      kc = reinterpret_cast<uint8_t*>(kv_cache) + ((size_t) b * 2 * H + h) * S_max * kD * elt + (size_t) s * kD * elt;
      vc = kc + (size_t) H * S_max * kD * elt;
    }
    if (int8_kv) {
      auto q8 = [&](const uint4& x) {
        const __half* hx = reinterpret_cast<const __half*>(&x);
        uint2 o;
        o.x = pack4_i8(__half2float(hx[0]) * qs, __half2float(hx[1]) * qs, __half2float(hx[2]) * qs, __half2float(hx[3]) * qs);
        o.y = pack4_i8(__half2float(hx[4]) * qs, __half2float(hx[5]) * qs, __half2float(hx[6]) * qs, __half2float(hx[7]) * qs);
        return o;
      };
      *reinterpret_cast<uint2*>(kc + g8 * 8) = q8(klo);
      *reinterpret_cast<uint2*>(kc + 64 + g8 * 8) = q8(khi);
      *reinterpret_cast<uint2*>(vc + g8 * 8) = q8(vlo);
      *reinterpret_cast<uint2*>(vc + 64 + g8 * 8) = q8(vhi);
    } else {
      *reinterpret_cast<uint4*>(kc + (g8 * 8) * 2) = klo;
      *reinterpret_cast<uint4*>(kc + (64 + g8 * 8) * 2) = khi;
      *reinterpret_cast<uint4*>(vc + (g8 * 8) * 2) = vlo;
      *reinterpret_cast<uint4*>(vc + (64 + g8 * 8) * 2) = vhi;
    }
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int kBM = 64;          // query rows per CTA (4 warps x 16)
constexpr int kBN = 64;          // keys per tile
constexpr int kQPad = kD + 8;    // smem row pitch (halfs) for Q / K tiles
constexpr int kVPad = kBN + 8;   // smem row pitch for the transposed V tile [Dh][keys]

__global__ void __launch_bounds__(128) flash_ctx_kernel(const __half* __restrict__ qkv, __half* __restrict__ out,
                                                         const int* __restrict__ input_lengths, int S, int H,
                                                         float qk_scale) {
  extern __shared__ __align__(16) __half sm[];
  __half* Qs = sm;                       // [kBM][kQPad]
  __half* Ks = Qs + kBM * kQPad;         // [kBN][kQPad]
  __half* Vt = Ks + kBN * kQPad;         // [kD][kVPad]

  const int qt = gridDim.x - 1 - blockIdx.x;  // heavy (late) query tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int hidden = H * kD;
  const int len = input_lengths ? min(input_lengths[b], S) : S;
  const int q0 = qt * kBM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const size_t tok_stride = (size_t) 3 * hidden;
  const __half* base = qkv + (size_t) b * S * tok_stride + (size_t) h * kD;
  __half* obase = out + (size_t) b * S * hidden + (size_t) h * kD;

  if (q0 >= len) {  // whole tile is padding: defined output (zeros)
    for (int i = tid; i < kBM * kD / 8; i += 128) {
      const int r = i / (kD / 8), c8 = i % (kD / 8);
      if (q0 + r < S) *reinterpret_cast<uint4*>(obase + (size_t) (q0 + r) * hidden + c8 * 8) = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  // Q tile -> smem (zero beyond S)
  for (int i = tid; i < kBM * kD / 8; i += 128) {
    const int r = i / (kD / 8), c8 = i % (kD / 8);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q0 + r < S) v = *reinterpret_cast<const uint4*>(base + (size_t) (q0 + r) * tok_stride + c8 * 8);
    *reinterpret_cast<uint4*>(Qs + r * kQPad + c8 * 8) = v;
  }
  __syncthreads();
  // A fragments of Q for this warp's 16 rows: 8 k-steps of 16
  uint32_t qa[8][4];
  {
    const __half* qr0 = Qs + (warp * 16 + g) * kQPad;
    const __half* qr1 = qr0 + 8 * kQPad;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(qr0 + ks * 16 + 2 * t4);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(qr1 + ks * 16 + 2 * t4);
      qa[ks][2] = *reinterpret_cast<const uint32_t*>(qr0 + ks * 16 + 8 + 2 * t4);
      qa[ks][3] = *reinterpret_cast<const uint32_t*>(qr1 + ks * 16 + 8 + 2 * t4);
    }
  }

  float o[16][4];  // 16 n8-blocks over Dh = 128
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -3.0e38f, m1 = -3.0e38f, l0 = 0.f, l1 = 0.f;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;  // this thread's two query rows

  const int kv_end = min(len, q0 + kBM);           // causal: keys <= last query of the tile
  for (int k0 = 0; k0 < kv_end; k0 += kBN) {
    __syncthreads();  // previous tile fully consumed
    for (int i = tid; i < kBN * kD / 8; i += 128) {
      const int r = i / (kD / 8), c8 = i % (kD / 8);
      uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (k0 + r < len) {
        const __half* src = base + (size_t) (k0 + r) * tok_stride + c8 * 8;
        kv = *reinterpret_cast<const uint4*>(src + hidden);
        vv = *reinterpret_cast<const uint4*>(src + 2 * hidden);
      }
      *reinterpret_cast<uint4*>(Ks + r * kQPad + c8 * 8) = kv;
      const __half* vh = reinterpret_cast<const __half*>(&vv);
#pragma unroll
      for (int j = 0; j < 8; ++j) Vt[(c8 * 8 + j) * kVPad + r] = vh[j];
    }
    __syncthreads();

    // S = Q K^T : 8 n8-blocks of keys
    float sacc[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      sacc[nb][0] = sacc[nb][1] = sacc[nb][2] = sacc[nb][3] = 0.f;
      const __half* kr = Ks + (nb * 8 + g) * kQPad;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        uint32_t bf[2];
        bf[0] = *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 2 * t4);
        bf[1] = *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8 + 2 * t4);
        mma16816(sacc[nb], qa[ks], bf);
      }
    }
    // scale + mask + online softmax (rows r0: c0,c1 ; r1: c2,c3)
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = k0 + nb * 8 + 2 * t4 + (j & 1);
        const int qr = (j < 2) ? r0 : r1;
        float v = sacc[nb][j] * qk_scale;
        if (key > qr || key >= len) v = -3.0e38f;
        sacc[nb][j] = v;
        if (j < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float corr0 = __expf(m0 - mx0), corr1 = __expf(m1 - mx1);
    m0 = mx0;
    m1 = mx1;
    float ps0 = 0.f, ps1 = 0.f;
    uint32_t pa[4][4];  // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      float p[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v = sacc[nb][j];
        const float e = v <= -1.0e38f ? 0.f : __expf(v - (j < 2 ? m0 : m1));
        // the reference rounds P to fp16 before P.V (softmax_kernel writes T); sum in fp32
        p[j] = e;
      }
      ps0 += p[0] + p[1];
      ps1 += p[2] + p[3];
      const int ks = nb >> 1;
      if ((nb & 1) == 0) {
        pa[ks][0] = pack_h2(p[0], p[1]);
        pa[ks][1] = pack_h2(p[2], p[3]);
      } else {
        pa[ks][2] = pack_h2(p[0], p[1]);
        pa[ks][3] = pack_h2(p[2], p[3]);
      }
    }
    l0 = l0 * corr0 + ps0;
    l1 = l1 * corr1 + ps1;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i][0] *= corr0; o[i][1] *= corr0; o[i][2] *= corr1; o[i][3] *= corr1;
    }
    // O += P V : B fragment from the transposed V tile
#pragma unroll
    for (int nb = 0; nb < 16; ++nb) {
      const __half* vr = Vt + (nb * 8 + g) * kVPad;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bf[2];
        bf[0] = *reinterpret_cast<const uint32_t*>(vr + ks * 16 + 2 * t4);
        bf[1] = *reinterpret_cast<const uint32_t*>(vr + ks * 16 + 8 + 2 * t4);
        mma16816(o[nb], pa[ks], bf);
      }
    }
  }
  // finalise: row sums across the 4 lanes of a quad
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = __fdividef(1.f, l0 + 1.e-6f), inv1 = __fdividef(1.f, l1 + 1.e-6f);
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    const int col = nb * 8 + 2 * t4;
    if (r0 < S) {
      const uint32_t v = r0 < len ? pack_h2(o[nb][0] * inv0, o[nb][1] * inv0) : 0u;
      *reinterpret_cast<uint32_t*>(obase + (size_t) r0 * hidden + col) = v;
    }
    if (r1 < S) {
      const uint32_t v = r1 < len ? pack_h2(o[nb][2] * inv1, o[nb][3] * inv1) : 0u;
      *reinterpret_cast<uint32_t*>(obase + (size_t) r1 * hidden + col) = v;
    }
  }
}

}  // namespace tb

using namespace tb;

extern "C" size_t tb_context_attention_workspace_bytes(int batch, int seq_len, int num_heads) {
  return flash_ctx_tc_workspace_bytes(batch, seq_len, num_heads);
}

static int context_attention_impl(void* out, void* qkv, void* kv_cache, const long long* block_ptrs, int tokens_per_block,
                                  int max_blocks, const int* input_lengths, const float* kv_scale_orig_quant,
                                  void* workspace, int batch, int seq_len, int num_heads, int head_size, int max_seq_len,
                                  int rotary_dim, float q_scaling, int int8_kv, cudaStream_t stream);

extern "C" int tb_context_attention(void* out, void* qkv, void* kv_cache, const int* input_lengths,
                                    const float* kv_scale_orig_quant, void* workspace, int batch, int seq_len,
                                    int num_heads, int head_size, int max_seq_len, int rotary_dim, float q_scaling,
                                    int int8_kv, cudaStream_t stream) {
  return context_attention_impl(out, qkv, kv_cache, nullptr, 0, 0, input_lengths, kv_scale_orig_quant, workspace, batch, seq_len,
                                num_heads, head_size, max_seq_len, rotary_dim, q_scaling, int8_kv, stream);
}

extern "C" int tb_context_attention_paged(void* out, void* qkv, const int64_t* block_pointers, int tokens_per_block,
                                          int max_blocks_per_seq, const int* input_lengths, const float* kv_scale_orig_quant,
                                          void* workspace, int batch, int seq_len, int num_heads, int head_size,
                                          int rotary_dim, float q_scaling, int int8_kv, cudaStream_t stream) {
  if (!block_pointers || tokens_per_block < 16 || (tokens_per_block & (tokens_per_block - 1)) || max_blocks_per_seq < 1) return -1;
  return context_attention_impl(out, qkv, nullptr, reinterpret_cast<const long long*>(block_pointers), tokens_per_block,
                                max_blocks_per_seq, input_lengths, kv_scale_orig_quant, workspace, batch, seq_len, num_heads,
                                head_size, tokens_per_block * max_blocks_per_seq, rotary_dim, q_scaling, int8_kv, stream);
}

static int context_attention_impl(void* out, void* qkv, void* kv_cache, const long long* block_ptrs, int tokens_per_block,
                                  int max_blocks, const int* input_lengths, const float* kv_scale_orig_quant,
                                  void* workspace, int batch, int seq_len, int num_heads, int head_size, int max_seq_len,
                                  int rotary_dim, float q_scaling, int int8_kv, cudaStream_t stream) {
  if (head_size != kD) return -1;
  if (rotary_dim != 0 && rotary_dim != kD) return -1;
  if (seq_len > max_seq_len || batch <= 0 || seq_len <= 0) return -2;
  if (int8_kv && !kv_scale_orig_quant) return -1;
  int tpb_log2 = 0;
  while (block_ptrs && (1 << tpb_log2) < tokens_per_block) ++tpb_log2;
  ctx_prep_kernel<<<dim3(batch * seq_len), kPrepThreads, 0, stream>>>((__half*) qkv, kv_cache, input_lengths,
                                                                      kv_scale_orig_quant, seq_len, num_heads,
                                                                      max_seq_len, rotary_dim, int8_kv, block_ptrs, tpb_log2,
                                                                      max_blocks);
  const float qk_scale_tc = 1.f / (sqrtf((float) head_size) * q_scaling);
  if (workspace)   // tcgen05 path (context_attn_tc.cu); without a workspace the warp-MMA kernel below runs
    return launch_flash_ctx_tc(out, qkv, workspace, input_lengths, batch, seq_len, num_heads, qk_scale_tc, stream);
  const size_t smem = (size_t) (kBM * kQPad + kBN * kQPad + kD * kVPad) * sizeof(__half);
  static bool attr_set = false;
  if (!attr_set) {
    TB_CHECK_CUDA(cudaFuncSetAttribute(flash_ctx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    attr_set = true;
  }
  const float qk_scale = 1.f / (sqrtf((float) head_size) * q_scaling);
  dim3 grid((seq_len + kBM - 1) / kBM, num_heads, batch);
  flash_ctx_kernel<<<grid, 128, smem, stream>>>((const __half*) qkv, (__half*) out, input_lengths, seq_len,
                                                num_heads, qk_scale);
  return (int) cudaGetLastError();
}
