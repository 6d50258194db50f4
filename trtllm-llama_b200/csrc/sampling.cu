// Token sampling beyond greedy: temperature, top-k, top-p (nucleus) and top-k + top-p, one CTA per sequence.
//
// Replaces (reference, K/ = T/cpp/tensorrt_llm/kernels/): the sampling half of DynamicDecodeOp
// (T/cpp/tensorrt_llm/thop/dynamicDecodeOp.cpp:359-363 -> TopKSamplingLayer / TopPSamplingLayer):
//   K/samplingPenaltyKernels.cu:77-93      invokeApplyTemperaturePenalty   logits * 1/(T + 1e-6)
//   K/samplingTopKKernels.cu:118-319       topk_stage1 + topk_stage2_sampling (k iterative arg-max passes, then
//                                          r = u * top_p * sum(exp(l_i - l_max)); walk the k candidates)
//   K/samplingTopPKernels.cu:882-1010,1160-1236  softmax, sort by probability, r = u * top_p, first token whose inclusive
//                                          cumulative probability reaches r
//   K/samplingTopKKernels.cu:38-62         curand state per sequence -> here a counter-based Philox4x32-10 keyed by
//                                          (seed; step, row): no state buffer, CUDA-graph replayable (the step counter is
//                                          read from device memory), every row its own stream (the reference seeds all
//                                          rows identically)
// The row (<= 56 K logits) lives in shared memory.  Top-k: k block-wide arg-max passes as the reference.  Pure top-p:
// no sort — the picked token is the largest probability v with mass{p >= v} >= r, found by a 31-step bisection on the
// float bit pattern (a block-wide masked sum per step); lowest index wins ties, as a stable descending sort would.
// HBM traffic is one read of the logits row (128 KB): latency-bound bookkeeping, not a roofline kernel.
#include "common.cuh"
#include "kernels.h"

namespace tb {

constexpr int kSampThreads = 1024;
constexpr int kSampMaxK = 1024;

struct SampleParams {
  int* out_ids;
  const float* logits;
  int vocab, stride, top_k;
  float top_p, inv_temp;
  unsigned long long seed;
  const int* step_dev;      // generation step counter on the device (NULL: `step`)
  int step;
  const int* finished;      // optional [rows]: finished sequences emit end_id
  int end_id;
  float* uniform_out;       // optional [rows]: the uniform drawn (tests)
};

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

__device__ __forceinline__ ArgMax block_argmax(ArgMax a, ArgMax* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = better(a, b);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = a;
  __syncthreads();
  ArgMax r = red[0];
  for (int w = 1; w < kSampThreads / 32; ++w) r = better(r, red[w]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < kSampThreads / 32; ++w) r += red[w];
  return r;
}

__global__ void __launch_bounds__(kSampThreads) sample_kernel(const SampleParams p) {
  extern __shared__ float row[];                       // [vocab]
  __shared__ ArgMax red_a[kSampThreads / 32];
  __shared__ float red_f[kSampThreads / 32];
  __shared__ float k_val[kSampMaxK];
  __shared__ int k_id[kSampMaxK];
  const int b = blockIdx.x, tid = threadIdx.x, V = p.vocab;
  if (p.finished && p.finished[b]) {
    if (tid == 0) p.out_ids[b] = p.end_id;
    return;
  }
  const float* src = p.logits + (size_t) b * p.stride;
  for (int i = tid; i < V; i += kSampThreads) row[i] = src[i] * p.inv_temp;
  uint32_t c[4] = {(uint32_t) (p.step_dev ? p.step_dev[0] : p.step), 0u, (uint32_t) b, 0u};
  philox4x32_10(c, (uint32_t) p.seed, (uint32_t) (p.seed >> 32));
  const float u = (float) c[0] * 2.3283064365386963e-10f + 1.1641532182693481e-10f;     // curand_uniform's map: (0, 1]
  if (tid == 0 && p.uniform_out) p.uniform_out[b] = u;
  __syncthreads();

  if (p.top_k > 0) {
    const int k = p.top_k < V ? p.top_k : V;
    for (int ite = 0; ite < k; ++ite) {
      ArgMax a{-3.4e38f, 0x7fffffff};
      for (int i = tid; i < V; i += kSampThreads) a = better(a, ArgMax{row[i], i});
      a = block_argmax(a, red_a);
      if (tid == 0) {
        k_val[ite] = a.v;
        k_id[ite] = a.i;
        row[a.i] = -3.4e38f;
      }
      __syncthreads();
    }
    if (tid == 0) {
      const float mx = k_val[0];
      float s = 0.f;
      for (int i = 0; i < k; ++i) {
        k_val[i] = __expf(k_val[i] - mx);
        s += k_val[i];
      }
      float r = u * p.top_p * s;
      int pick = k - 1;
      for (int i = 0; i < k; ++i) {
        r -= k_val[i];
        if (r <= 0.f) { pick = i; break; }
      }
      p.out_ids[b] = k_id[pick];
    }
    return;
  }

  // ---- pure top-p over the whole vocabulary --------------------------------------------------------------------------
  ArgMax a{-3.4e38f, 0x7fffffff};
  for (int i = tid; i < V; i += kSampThreads) a = better(a, ArgMax{row[i], i});
  a = block_argmax(a, red_a);
  float se = 0.f;
  for (int i = tid; i < V; i += kSampThreads) {
    const float e = __expf(row[i] - a.v);
    row[i] = e;
    se += e;
  }
  se = block_sum(se, red_f);
  const float inv = 1.f / se;
  for (int i = tid; i < V; i += kSampThreads) row[i] *= inv;
  __syncthreads();
  const float r = u * p.top_p;
  {
    // r beyond the row's total mass (u = 1, top_p = 1 and a rounded-down sum): the reference's scan never fires and
    // leaves the most probable token (samplingTopPKernels.cu:957)
    float m0 = 0.f;
    for (int i = tid; i < V; i += kSampThreads) m0 += row[i];
    m0 = block_sum(m0, red_f);
    if (m0 < r) {
      if (tid == 0) p.out_ids[b] = a.i;
      return;
    }
  }
  // largest probability value v (as a bit pattern: positive floats order like integers) with mass{p >= v} >= r
  uint32_t lo = 0u, hi = __float_as_uint(1.0f);          // invariant: mass{p >= lo} >= r (lo = 0: the whole row, mass ~1 >= r)
  while (lo < hi) {
    const uint32_t mid = lo + (hi - lo + 1) / 2;
    const float t = __uint_as_float(mid);
    float m = 0.f;
    for (int i = tid; i < V; i += kSampThreads) m += row[i] >= t ? row[i] : 0.f;
    m = block_sum(m, red_f);
    if (m >= r) lo = mid; else hi = mid - 1;
  }
  // the token: the largest probability <= lo that exists in the row ... which is the smallest probability >= lo's
  // predecessor set; equivalently the minimum of {p_i >= v}, lowest index first
  const float v = __uint_as_float(lo);
  ArgMax best{-3.4e38f, 0x7fffffff};
  for (int i = tid; i < V; i += kSampThreads)
    if (row[i] >= v) best = better(best, ArgMax{-row[i], i});          // arg-min of the probabilities at or above v
  best = block_argmax(best, red_a);
  if (tid == 0) p.out_ids[b] = best.i == 0x7fffffff ? a.i : best.i;
}

}  // namespace tb

using namespace tb;

extern "C" int tb_sample(int* out_ids, const float* logits, int rows, int vocab, int vocab_stride, int top_k, float top_p,
                         float temperature, unsigned long long seed, const int* step_dev, int step, const int* finished,
                         int end_id, float* uniform_out, cudaStream_t stream) {
  if (!out_ids || !logits || rows < 1 || vocab < 1 || vocab_stride < vocab) return -1;
  if (top_k < 0 || top_k > kSampMaxK || !(top_p > 0.f) || top_p > 1.f || !(temperature >= 0.f)) return -1;
  const size_t smem = (size_t) vocab * sizeof(float);
  if (smem > 200 * 1024) return -2;
  static bool attr_done = false;
  if (!attr_done) {
    TB_CHECK_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  SampleParams p{};
  p.out_ids = out_ids; p.logits = logits; p.vocab = vocab; p.stride = vocab_stride; p.top_k = top_k; p.top_p = top_p;
  p.inv_temp = 1.f / (temperature + 1e-6f);
  p.seed = seed; p.step_dev = step_dev; p.step = step; p.finished = finished; p.end_id = end_id; p.uniform_out = uniform_out;
  sample_kernel<<<rows, kSampThreads, smem, stream>>>(p);
  return (int) cudaGetLastError();
}
