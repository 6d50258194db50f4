// Beam search (SamplingConfig.num_beams > 1): one decoding step, the cache-indirection update and the final gather_tree.
//
// Replaces (reference, T/ = tensorrt_llm_july-release-v1/, K/ = T/cpp/tensorrt_llm/kernels/), for the path the Python
// runtime drives (DynamicDecodeOp passes no BeamHypotheses, T/cpp/tensorrt_llm/thop/dynamicDecodeOp.cpp):
//   K/onlineSoftmaxBeamsearchKernels.cu:402-592  per (batch, beam) row: log-softmax and its top 2W candidates
//                                                (a finished beam proposes only end_id, with log-probability 0)
//   K/onlineSoftmaxBeamsearchKernels.cu:112-300  batch_topk_kernel: the W best of the 2 W^2 candidates of a batch entry by
//                                                cum_log_prob / length^length_penalty
//   T/cpp/tensorrt_llm/layers/onlineBeamSearchLayer.cu:30-62    update_kernel (parent / token / finished / lengths)
//   T/cpp/tensorrt_llm/layers/baseBeamSearchLayer.cu:29-67      update_indir_cache_kernel
//   K/decodingKernels.cu:31-170                                 gatherTree
// Kept reference behaviour: the length used to normalise candidate j of ANY beam is the length of beam (j mod W)
// (batch_topk_kernel indexes `elem_id % K` when there are no BeamHypotheses) — it only matters once some beams of a batch
// entry have finished; ties go to the candidate with the lower (beam, rank) index.
// One CTA per (batch, beam) row for the candidates, one CTA per batch entry for the selection: bookkeeping kernels (one read
// of the logits, 128 KB per row), latency- not roofline-bound.  The generation step is read from device memory so a captured
// step graph can be replayed.
#include <cfloat>
#include "common.cuh"
#include "kernels.h"

namespace tb {

constexpr int kBeamThreads = 1024;
constexpr int kBeamMaxW = 16;            // beam widths 1..16: 2W <= 32 candidates per row

struct Cand { float v; int i; };
__device__ __forceinline__ Cand cand_better(Cand a, Cand b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

struct BeamCandParams {
  const float* logits;       // [rows or batch][stride]
  int vocab, stride, W, broadcast;    // broadcast: every beam of a batch entry reads the entry's one row (context step)
  const int* finished;       // [rows]
  const float* cum;          // [rows]
  int end_id;
  int* cand_id;              // [rows][2W] token ids, best first
  float* cand_val;           // [rows][2W] cum + log-probability
};

__global__ void __launch_bounds__(kBeamThreads) beam_candidates_kernel(const BeamCandParams p) {
  extern __shared__ float row[];                      // [vocab]
  __shared__ Cand red[kBeamThreads / 32];
  __shared__ float red_f[kBeamThreads / 32];
  __shared__ float top_v[2 * kBeamMaxW];
  __shared__ int top_i[2 * kBeamMaxW];
  const int r = blockIdx.x, tid = threadIdx.x, V = p.vocab, n = 2 * p.W;
  const int lane = tid & 31, warp = tid >> 5;
  if (p.finished[r]) {
    // MAX for end_id, -MAX elsewhere: log-softmax is 0 for end_id and -inf for the rest (Kernels.cu:437-452)
    if (tid < n) {
      p.cand_id[(size_t) r * n + tid] = tid == 0 ? p.end_id : (tid <= p.end_id ? tid - 1 : tid);
      p.cand_val[(size_t) r * n + tid] = tid == 0 ? p.cum[r] : -INFINITY;
    }
    return;
  }
  const float* src = p.logits + (size_t) (p.broadcast ? r / p.W : r) * p.stride;
  float mx = -FLT_MAX;
  for (int i = tid; i < V; i += kBeamThreads) {
    const float v = src[i];
    row[i] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  if (lane == 0) red_f[warp] = mx;
  __syncthreads();
  mx = red_f[0];
  for (int w = 1; w < kBeamThreads / 32; ++w) mx = fmaxf(mx, red_f[w]);
  float s = 0.f;
  for (int i = tid; i < V; i += kBeamThreads) s += __expf(row[i] - mx);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red_f[warp] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < kBeamThreads / 32; ++w) s += red_f[w];
  const float log_d = logf(s);
  for (int ite = 0; ite < n; ++ite) {                 // 2W block-wide arg-max passes
    Cand a{-FLT_MAX, 0x7fffffff};
    for (int i = tid; i < V; i += kBeamThreads) a = cand_better(a, Cand{row[i], i});
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Cand b;
      b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
      b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
      a = cand_better(a, b);
    }
    __syncthreads();
    if (lane == 0) red[warp] = a;
    __syncthreads();
    if (tid == 0) {
      Cand t = red[0];
      for (int w = 1; w < kBeamThreads / 32; ++w) t = cand_better(t, red[w]);
      top_v[ite] = t.v;
      top_i[ite] = t.i;
      row[t.i] = -FLT_MAX;
    }
    __syncthreads();
  }
  if (tid < n) {
    p.cand_id[(size_t) r * n + tid] = top_i[tid];
    p.cand_val[(size_t) r * n + tid] = (top_v[tid] - mx - log_d) + p.cum[r];
  }
}

struct BeamSelectParams {
  const int* cand_id;
  const float* cand_val;
  int W, rows, S_max, out_stride;
  float length_penalty;
  int end_id;
  const int* step_dev;       // generated-token column being produced (device)
  const int* max_in_dev;     // padded prompt length (device): sequence position of column c is max_in + c
  float* cum;                // [rows] in/out
  int* finished;             // [rows] in/out
  int* beam_lens;            // [rows] in/out: the decoder's sequence lengths (length penalty, gather_tree)
  int* out_ids_t;            // [max_out][rows] time-major token ids
  int* parent_t;             // [max_out][rows] time-major parent beams
  int* next_ids;             // [rows] input ids of the next step
  const int* src_indir;      // [batch][W][S_max]
  int* tgt_indir;
};

// one CTA per batch entry, W <= 16: thread 0 walks the 2 W^2 candidates (<= 512) W times — cheaper than any reduction
__global__ void __launch_bounds__(256) beam_select_kernel(const BeamSelectParams p) {
  __shared__ float elem[2 * kBeamMaxW * kBeamMaxW];
  __shared__ int sel[kBeamMaxW];
  __shared__ int new_len[kBeamMaxW], old_len_inc[kBeamMaxW], new_fin[kBeamMaxW], parent[kBeamMaxW], tok[kBeamMaxW];
  __shared__ float new_cum[kBeamMaxW];
  __shared__ unsigned char taken[2 * kBeamMaxW * kBeamMaxW];
  const int b = blockIdx.x, W = p.W, n = 2 * W, tid = threadIdx.x;
  const int col = p.step_dev[0];
  const int base = b * W;
  for (int e = tid; e < W * n; e += blockDim.x) {
    const int j = e % n;
    const int i = j % W;                              // reference: elem_id % K picks the beam whose length normalises
    float v = p.cand_val[(size_t) base * n + e];
    if (p.length_penalty != 0.f) {
      const int len = p.finished[base + i] ? p.beam_lens[base + i] : p.beam_lens[base + i] + 1;
      if (len != 1) v = v / powf((float) len, p.length_penalty);
    }
    elem[e] = v;
    taken[e] = 0;
  }
  if (tid < W) old_len_inc[tid] = p.beam_lens[base + tid] + (p.finished[base + tid] ? 0 : 1);
  __syncthreads();
  if (tid == 0) {
    for (int k = 0; k < W; ++k) {                     // strict > keeps the lowest index among ties (and among -inf)
      int best = -1;
      float bv = 0.f;
      for (int e = 0; e < W * n; ++e)
        if (!taken[e] && (best < 0 || elem[e] > bv)) { best = e; bv = elem[e]; }
      taken[best] = 1;
      sel[k] = best;
    }
  }
  __syncthreads();
  if (tid < W) {
    const int e = sel[tid], pb = e / n;
    tok[tid] = p.cand_id[(size_t) base * n + e];
    new_cum[tid] = p.cand_val[(size_t) base * n + e];
    parent[tid] = pb;
    new_len[tid] = old_len_inc[pb];
    new_fin[tid] = tok[tid] == p.end_id ? 1 : 0;
  }
  __syncthreads();
  if (tid < W) {
    p.out_ids_t[(size_t) col * p.rows + base + tid] = tok[tid];
    p.parent_t[(size_t) col * p.rows + base + tid] = parent[tid];
    p.next_ids[base + tid] = tok[tid];
    p.cum[base + tid] = new_cum[tid];
    p.finished[base + tid] = new_fin[tid];
    p.beam_lens[base + tid] = new_len[tid];
  }
  // cache indirection: positions [0, pos] of every unfinished beam inherit the parent's row; position pos (the token just
  // chosen, whose K/V the next forward pass writes) is the beam's own
  const int pos = p.max_in_dev[0] + col;
  for (int w = 0; w < W; ++w) {
    if (new_fin[w]) continue;
    const int* s = p.src_indir + ((size_t) base + parent[w]) * p.S_max;
    int* t = p.tgt_indir + ((size_t) base + w) * p.S_max;
    for (int i = tid; i <= pos && i < p.S_max; i += blockDim.x) t[i] = i == pos ? w : s[i];
  }
}

struct GatherTreeParams {
  const int* out_ids_t;
  const int* parent_t;
  int rows, W, n, end_id;
  int* out;                  // [rows][n]
};

__global__ void gather_tree_kernel(const GatherTreeParams p) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.rows) return;
  const int base = r / p.W * p.W;
  int beam = r % p.W;
  for (int c = p.n - 1; c >= 0; --c) {
    p.out[(size_t) r * p.n + c] = p.out_ids_t[(size_t) c * p.rows + base + beam];
    beam = p.parent_t[(size_t) c * p.rows + base + beam];
  }
  bool done = false;                                   // everything after the first end_id is end_id
  for (int c = 0; c < p.n; ++c) {
    if (done) p.out[(size_t) r * p.n + c] = p.end_id;
    else if (p.out[(size_t) r * p.n + c] == p.end_id) done = true;
  }
}

__global__ void beam_init_kernel(float* cum, int* finished, int* beam_lens, int* indir_a, int* indir_b, const int* max_in_dev,
                                 int rows, int W, int S_max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) {
    cum[i] = (i % W) == 0 ? 0.f : -1e20f;             // generation.py:392-396
    finished[i] = 0;
    beam_lens[i] = max_in_dev[0];                      // generation.py:808-811 sequence_lengths = max_input_length
  }
  for (size_t j = i; j < (size_t) rows * S_max; j += (size_t) gridDim.x * blockDim.x) { indir_a[j] = 0; indir_b[j] = 0; }
}

}  // namespace tb

using namespace tb;

extern "C" {

size_t tb_beam_workspace_bytes(int rows, int beam_width) { return (size_t) rows * 2 * beam_width * 8 + 256; }

int tb_beam_init(float* cum_log_probs, int* finished, int* beam_lens, int* indir_a, int* indir_b, const int* max_in_dev,
                 int rows, int beam_width, int max_seq_len, cudaStream_t stream) {
  if (rows < 1 || beam_width < 1 || rows % beam_width) return -1;
  beam_init_kernel<<<(rows + 255) / 256 + 64, 256, 0, stream>>>(cum_log_probs, finished, beam_lens, indir_a, indir_b, max_in_dev,
                                                                 rows, beam_width, max_seq_len);
  return (int) cudaGetLastError();
}

int tb_beam_search_step(const float* logits, int vocab, int vocab_stride, int broadcast_rows, int rows, int beam_width,
                        float length_penalty, int end_id, const int* step_dev, const int* max_in_dev, float* cum_log_probs,
                        int* finished, int* beam_lens, int* out_ids_t, int* parent_ids_t, int* next_ids, const int* src_indir,
                        int* tgt_indir, int max_seq_len, void* workspace, cudaStream_t stream) {
  if (beam_width < 1 || beam_width > kBeamMaxW || rows < 1 || rows % beam_width || 2 * beam_width > vocab) return -1;
  if ((size_t) vocab * 4 > 200 * 1024) return -2;
  static bool attr_done = false;
  if (!attr_done) {
    TB_CHECK_CUDA(cudaFuncSetAttribute(beam_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  int* cand_id = reinterpret_cast<int*>(workspace);
  float* cand_val = reinterpret_cast<float*>(cand_id + (size_t) rows * 2 * beam_width);
  BeamCandParams c{logits, vocab, vocab_stride, beam_width, broadcast_rows, finished, cum_log_probs, end_id, cand_id, cand_val};
  beam_candidates_kernel<<<rows, kBeamThreads, (size_t) vocab * 4, stream>>>(c);
  BeamSelectParams s{cand_id, cand_val, beam_width, rows, max_seq_len, 0, length_penalty, end_id, step_dev, max_in_dev,
                     cum_log_probs, finished, beam_lens, out_ids_t, parent_ids_t, next_ids, src_indir, tgt_indir};
  beam_select_kernel<<<rows / beam_width, 256, 0, stream>>>(s);
  return (int) cudaGetLastError();
}

int tb_gather_tree(int* out, const int* out_ids_t, const int* parent_ids_t, int rows, int beam_width, int n_steps, int end_id,
                   cudaStream_t stream) {
  if (rows < 1 || beam_width < 1 || rows % beam_width || n_steps < 1) return -1;
  GatherTreeParams p{out_ids_t, parent_ids_t, rows, beam_width, n_steps, end_id, out};
  gather_tree_kernel<<<(rows + 127) / 128, 128, 0, stream>>>(p);
  return (int) cudaGetLastError();
}
}
