// Small element-wise / gather kernels that TensorRT generated natively in the reference (k14, no
// source in the reference tree; semantics from the Python graph):
//   embedding gather           T/tensorrt_llm/layers/embedding.py, LQ/llama_model.py:159-170
//   silu(fc(x)) * gate(x)      T/tensorrt_llm/layers/mlp.py:68-73, T/tensorrt_llm/functional.py:521-551
//   residual add               LQ/llama_model.py:100-119
//   last-token gather          T/tensorrt_llm/functional.py:3316- (gather_last_token_logits)
//   greedy argmax              DynamicDecodeOp top_k = 1 (T/tensorrt_llm/runtime/generation.py:943-961)
// All HBM-bound, 16-byte vector accesses.
#include "common.cuh"
#include "kernels.h"

namespace tb {

__global__ void embedding_kernel(__half* out, const __half* table, const int* ids, int hidden, int vocab) {
  const int tok = blockIdx.x;
  int id = ids[tok];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const uint4* src = reinterpret_cast<const uint4*>(table + (size_t) id * hidden);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t) tok * hidden);
  for (int i = threadIdx.x; i < hidden / 8; i += blockDim.x) dst[i] = src[i];
}

// in: [rows, 2*inter] (gate | up) or two separate pointers; out [rows, inter]
__global__ void swiglu_kernel(__half* out, const __half* gate, const __half* up, int inter, int in_stride, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n8; i += (int64_t) gridDim.x * blockDim.x) {
    const int64_t row = i / (inter / 8);
    const int c8 = (int) (i % (inter / 8));
    uint4 g4 = *reinterpret_cast<const uint4*>(gate + row * in_stride + c8 * 8);
    uint4 u4 = *reinterpret_cast<const uint4*>(up + row * in_stride + c8 * 8);
    const __half2* g = reinterpret_cast<const __half2*>(&g4);
    const __half2* u = reinterpret_cast<const __half2*>(&u4);
    uint4 o4;
    __half2* o = reinterpret_cast<__half2*>(&o4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 gf = __half22float2(g[j]), uf = __half22float2(u[j]);
      // act output rounded to fp16 (TRT fp16 activation layer), then fp16 multiply
      __half2 a = __floats2half2_rn(silu_fast(gf.x), silu_fast(gf.y));
      float2 af = __half22float2(a);
      o[j] = __floats2half2_rn(af.x * uf.x, af.y * uf.y);
    }
    *reinterpret_cast<uint4*>(out + row * inter + c8 * 8) = o4;
  }
}

__global__ void add_kernel(__half* out, const __half* a, const __half* b, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n8; i += (int64_t) gridDim.x * blockDim.x) {
    uint4 a4 = reinterpret_cast<const uint4*>(a)[i], b4 = reinterpret_cast<const uint4*>(b)[i], o4;
    const __half2* x = reinterpret_cast<const __half2*>(&a4);
    const __half2* y = reinterpret_cast<const __half2*>(&b4);
    __half2* o = reinterpret_cast<__half2*>(&o4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 xf = __half22float2(x[j]), yf = __half22float2(y[j]);
      o[j] = __floats2half2_rn(xf.x + yf.x, xf.y + yf.y);
    }
    reinterpret_cast<uint4*>(out)[i] = o4;
  }
}

// rows of [B, S, hidden] at index last_ids[b] - 1 -> [B, hidden]
__global__ void gather_last_kernel(__half* out, const __half* in, const int* last_ids, int S, int hidden) {
  const int b = blockIdx.x;
  int s = last_ids[b] - 1;
  s = s < 0 ? 0 : (s >= S ? S - 1 : s);
  const uint4* src = reinterpret_cast<const uint4*>(in + ((size_t) b * S + s) * hidden);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t) b * hidden);
  for (int i = threadIdx.x; i < hidden / 8; i += blockDim.x) dst[i] = src[i];
}

// packed rows (remove_input_padding): sequence b ends at packed row sum(lens[:b + 1]) - 1
__global__ void gather_last_packed_kernel(__half* out, const __half* in, const int* lens, int hidden) {
  __shared__ int row_s;
  const int b = blockIdx.x;
  if (threadIdx.x < 32) {
    int a = 0;
    for (int j = threadIdx.x; j <= b; j += 32) a += lens[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (threadIdx.x == 0) row_s = a > 0 ? a - 1 : 0;
  }
  __syncthreads();
  const uint4* src = reinterpret_cast<const uint4*>(in + (size_t) row_s * hidden);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t) b * hidden);
  for (int i = threadIdx.x; i < hidden / 8; i += blockDim.x) dst[i] = src[i];
}

// argmax over fp32 logits [B, V]; lowest index wins ties.  One CTA per row.
__global__ void __launch_bounds__(1024) argmax_kernel(int* out, const float* logits, int vocab, int vocab_stride) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const float* row = logits + (size_t) blockIdx.x * vocab_stride;
  float bv = -3.4e38f;
  int bi = 0x7fffffff;
  if ((vocab & 3) == 0 && (vocab_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) {
    // 8 independent 16-byte loads per thread per pass: one memory round trip for a 32k vocabulary instead of a chain
    // of 32 scalar loads (17 us -> a few us per decode step)
    const float4* row4 = reinterpret_cast<const float4*>(row);
    const int n4 = vocab >> 2;
    for (int base = 0; base < n4; base += 8 * blockDim.x) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i4 = base + u * blockDim.x + threadIdx.x;
        v[u] = i4 < n4 ? row4[i4] : make_float4(-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i4 = base + u * blockDim.x + threadIdx.x, i = i4 * 4;
        const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        if (i4 < n4) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (e[j] > bv || (e[j] == bv && i + j < bi)) { bv = e[j]; bi = i + j; }
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
      const float v = row[i];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    bv = lane < (blockDim.x >> 5) ? sv[lane] : -3.4e38f;
    bi = lane < (blockDim.x >> 5) ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) out[blockIdx.x] = bi;
  }
}

// decode-loop bookkeeping kept on the device so a captured CUDA graph can be replayed:
// appends the new token to output_ids[b, *step_pos], sets next input ids, advances the lengths.
__global__ void advance_step_kernel(const int* new_ids, int* input_ids, int* output_ids, int* seq_lens, int* step_pos,
                                    int batch, int out_stride) {
  const int b = threadIdx.x;
  const int pos = *step_pos;
  if (b < batch) {
    const int id = new_ids[b];
    input_ids[b] = id;
    if (output_ids && pos < out_stride) output_ids[(size_t) b * out_stride + pos] = id;
    if (seq_lens) seq_lens[b] += 1;
  }
  __syncthreads();
  if (b == 0) *step_pos = pos + 1;
}

// Stop criterion of greedy decoding (replaces K/stopCriteriaKernels.cu + the `finished` bookkeeping of the dynamic
// decoder for the greedy case): one thread per sequence scans its generated ids [0, n_done).  *all_done = 1 when every
// sequence has produced end_id; with pad != 0 every position after a sequence's first end_id, up to n_pad, is
// overwritten with end_id (what the reference leaves in output_ids for finished sequences).
__global__ void finished_kernel(int* all_done, int* out_ids, int batch, int out_stride, int n_done, int n_pad, int end_id,
                                int pad) {
  const int b = threadIdx.x;
  int done = 1;
  if (b < batch) {
    int* row = out_ids + (size_t) b * out_stride;
    int first = -1;
    for (int i = 0; i < n_done; ++i)
      if (row[i] == end_id) { first = i; break; }
    done = first >= 0;
    if (pad && done)
      for (int i = first + 1; i < n_pad; ++i) row[i] = end_id;
  }
  const int all = __syncthreads_and(done);
  if (threadIdx.x == 0 && all_done) *all_done = all;
}

__global__ void half_to_float_kernel(float* out, const __half* in, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
    out[i] = __half2float(in[i]);
}

// packed <-> padded token rows (remove_input_padding): sequence b owns packed rows [sum(lens[:b]), +lens[b]) and padded rows
// [b * S, b * S + lens[b]); padded tail rows are zero-filled.  One CTA per (s, b); the prefix sum is a warp loop (B is small).
__global__ void pack_rows_kernel(uint4* dst, const uint4* src, const int* lens, int S, int row16, int to_packed) {
  __shared__ int off_s;
  const int s = blockIdx.x, b = blockIdx.y, len = lens[b];
  if (threadIdx.x < 32) {
    int a = 0;
    for (int j = threadIdx.x; j < b; j += 32) a += lens[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (threadIdx.x == 0) off_s = a;
  }
  __syncthreads();
  const size_t packed = (size_t) (off_s + s) * row16, padded = ((size_t) b * S + s) * row16;
  if (to_packed) {
    if (s < len)
      for (int i = threadIdx.x; i < row16; i += blockDim.x) dst[packed + i] = src[padded + i];
  } else {
    for (int i = threadIdx.x; i < row16; i += blockDim.x) dst[padded + i] = s < len ? src[packed + i] : make_uint4(0, 0, 0, 0);
  }
}
// teacher forcing: the token chosen for the column produced last (step_pos - 1) is replaced by ids[b]
__global__ void force_ids_kernel(const int* ids, int* input_ids, int* output_ids, const int* step_pos, int batch, int out_stride) {
  const int b = threadIdx.x;
  if (b >= batch) return;
  const int col = step_pos[0] - 1;
  input_ids[b] = ids[b];
  if (col >= 0 && col < out_stride) output_ids[(size_t) b * out_stride + col] = ids[b];
}
__global__ void tile_int_kernel(int* p, int n, int w) {
  const int t = threadIdx.x;
  const int v = t < n * w ? p[t / w] : 0;
  __syncthreads();
  if (t < n * w) p[t] = v;
}
__global__ void fill_int_kernel(int* p, int v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// all-gathered vocab-parallel logits [tp, rows, Vl] fp16 -> [rows, tp*Vl] fp32
// (the slice+concat after allgather in T/tensorrt_llm/layers/linear.py:78-97, plus the fp32 cast of LQ/llama_model.py:279)
__global__ void gather_logits_kernel(float* out, const __half* in, int rows, int vl, int tp, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
    const int j = (int) (i % vl);
    const int r = (int) ((i / vl) % rows);
    const int t = (int) (i / ((int64_t) vl * rows));
    out[(size_t) r * vl * tp + (size_t) t * vl + j] = __half2float(in[i]);
  }
}

static inline int grid_for(int64_t work, int threads) {
  int64_t b = (work + threads - 1) / threads;
  const int64_t cap = (int64_t) kNumSMs * 16;
  return (int) (b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace tb

using namespace tb;

extern "C" {

int tb_embedding(void* out, const void* table, const int* ids, int tokens, int hidden, int vocab, cudaStream_t s) {
  if (hidden % 8) return -1;
  embedding_kernel<<<tokens, 128, 0, s>>>((__half*) out, (const __half*) table, ids, hidden, vocab);
  return (int) cudaGetLastError();
}

int tb_swiglu(void* out, const void* gate, const void* up, int rows, int inter, int in_stride, cudaStream_t s) {
  if (inter % 8 || in_stride % 8) return -1;
  const int64_t n8 = (int64_t) rows * (inter / 8);
  swiglu_kernel<<<grid_for(n8, 256), 256, 0, s>>>((__half*) out, (const __half*) gate, (const __half*) up, inter,
                                                  in_stride, n8);
  return (int) cudaGetLastError();
}

int tb_add(void* out, const void* a, const void* b, int64_t n, cudaStream_t s) {
  if (n % 8) return -1;
  add_kernel<<<grid_for(n / 8, 256), 256, 0, s>>>((__half*) out, (const __half*) a, (const __half*) b, n / 8);
  return (int) cudaGetLastError();
}

int tb_gather_last_token(void* out, const void* in, const int* last_ids, int batch, int seq, int hidden,
                         cudaStream_t s) {
  if (hidden % 8) return -1;
  gather_last_kernel<<<batch, 128, 0, s>>>((__half*) out, (const __half*) in, last_ids, seq, hidden);
  return (int) cudaGetLastError();
}

int tb_gather_last_token_packed(void* out, const void* in, const int* lens, int batch, int hidden, cudaStream_t s) {
  if (hidden % 8) return -1;
  gather_last_packed_kernel<<<batch, 128, 0, s>>>((__half*) out, (const __half*) in, lens, hidden);
  return (int) cudaGetLastError();
}

int tb_argmax(int* out, const float* logits, int rows, int vocab, int vocab_stride, cudaStream_t s) {
  argmax_kernel<<<rows, 1024, 0, s>>>(out, logits, vocab, vocab_stride);
  return (int) cudaGetLastError();
}

int tb_advance_step(const int* new_ids, int* input_ids, int* output_ids, int* seq_lens, int* step_pos, int batch,
                    int out_stride, cudaStream_t s) {
  if (batch > 1024) return -1;
  advance_step_kernel<<<1, ((batch + 31) / 32) * 32, 0, s>>>(new_ids, input_ids, output_ids, seq_lens, step_pos, batch,
                                                             out_stride);
  return (int) cudaGetLastError();
}

int tb_finished(int* all_done, int* out_ids, int batch, int out_stride, int n_done, int n_pad, int end_id, int pad,
                cudaStream_t s) {
  if (batch < 1 || batch > 1024 || n_done < 0 || n_pad > out_stride) return -1;
  finished_kernel<<<1, ((batch + 31) / 32) * 32, 0, s>>>(all_done, out_ids, batch, out_stride, n_done, n_pad, end_id, pad);
  return (int) cudaGetLastError();
}

int tb_copy(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  return (int) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s);
}

int tb_fill_int(int* p, int value, int n, cudaStream_t s) {
  fill_int_kernel<<<(n + 255) / 256, 256, 0, s>>>(p, value, n);
  return (int) cudaGetLastError();
}

int tb_unpack_rows(void* padded, const void* packed, const int* lens, int batch, int seq, int row_bytes, cudaStream_t s) {
  if (batch < 1 || seq < 1 || row_bytes < 16 || (row_bytes & 15)) return -1;
  pack_rows_kernel<<<dim3(seq, batch), 128, 0, s>>>((uint4*) padded, (const uint4*) packed, lens, seq, row_bytes / 16, 0);
  return (int) cudaGetLastError();
}
int tb_pack_rows(void* packed, const void* padded, const int* lens, int batch, int seq, int row_bytes, cudaStream_t s) {
  if (batch < 1 || seq < 1 || row_bytes < 16 || (row_bytes & 15)) return -1;
  pack_rows_kernel<<<dim3(seq, batch), 128, 0, s>>>((uint4*) packed, (const uint4*) padded, lens, seq, row_bytes / 16, 1);
  return (int) cudaGetLastError();
}

int tb_force_ids(const int* ids, int* input_ids, int* output_ids, const int* step_pos, int batch, int out_stride, cudaStream_t s) {
  if (batch < 1 || batch > 1024) return -1;
  force_ids_kernel<<<1, ((batch + 31) / 32) * 32, 0, s>>>(ids, input_ids, output_ids, step_pos, batch, out_stride);
  return (int) cudaGetLastError();
}

// in place: p[i * w + j] = p[i] for i < n, j < w (generation.py:30-38 _tile_beam_width on a 1-D tensor); n * w <= 1024
int tb_tile_int(int* p, int n, int w, cudaStream_t s) {
  if (n < 1 || w < 1 || n * w > 1024) return -1;
  tile_int_kernel<<<1, 1024, 0, s>>>(p, n, w);
  return (int) cudaGetLastError();
}

int tb_gather_logits(float* out, const void* in, int rows, int vocab_local, int tp, cudaStream_t s) {
  const int64_t n = (int64_t) rows * vocab_local * tp;
  gather_logits_kernel<<<grid_for(n, 256), 256, 0, s>>>(out, (const __half*) in, rows, vocab_local, tp, n);
  return (int) cudaGetLastError();
}

int tb_half_to_float(float* out, const void* in, int64_t n, cudaStream_t s) {
  half_to_float_kernel<<<grid_for(n, 256), 256, 0, s>>>(out, (const __half*) in, n);
  return (int) cudaGetLastError();
}
}
