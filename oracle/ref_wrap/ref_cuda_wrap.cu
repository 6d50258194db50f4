// oracle/_ref GPU-side wrapper: calls the REFERENCE's own CUDA kernels (compiled for sm_100a from
// /root/reference where they lie, never copied) through a C ABI taking raw device pointers, so the
// -m gpu tests can compare this repo's kernels with the real thing on identical inputs.
// TEST INFRASTRUCTURE ONLY — nothing in the product path links or loads this library.
//
// Reference entry points wrapped:
//   kernels/decoderMaskedMultiheadAttention/decoderMaskedMultiheadAttentionLaunch.h:177-189 mmha_launch_kernel<.., 128>
//   kernels/gptKernels.cu:239-253                     invokeUpdatePaddingCount
//   kernels/layernormKernels.cu:233-264               invokeGeneralLayerNorm<half>
//   kernels/quantization.cu:67-84, 119-130            invokeQuantization<half>, invokePerTokenQuantization<half>
//   kernels/weightOnlyMatrixVectorMultiplication.cu:371-378 weight_only_gemv_launcher
#include <cstdint>
#include <cstring>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "tensorrt_llm/kernels/decoderMaskedMultiheadAttention.h"
#include "tensorrt_llm/kernels/gptKernels.h"
#include "tensorrt_llm/kernels/kvCacheUtils.h"
#include "tensorrt_llm/kernels/layernormKernels.h"
#include "tensorrt_llm/kernels/quantization.h"
#include "tensorrt_llm/kernels/weightOnlyMatrixVectorMultiplication.h"

namespace tensorrt_llm { namespace kernels { namespace mmha {
template <typename T, typename KVCacheBuffer, typename T_PARAMS, int Dh>
void mmha_launch_kernel(const T_PARAMS& params, const KVCacheBuffer& kv_cache_buffer, const cudaStream_t& stream);
}}}

using namespace tensorrt_llm::kernels;

extern "C" {

// Mirrors GPTAttentionPluginCommon::enqueueGeneration + fusedQKV_masked_attention_dispatch
// (plugins/gptAttentionCommon/gptAttentionCommon.cpp:107-207, 649-780) for T = half, Dh = 128,
// KVLinearBuffer, beam 1, multi_block off.
int ref_mmha_decode_half(void* out, const void* qkv, void* kv_cache, int batch, int num_heads, int head_size,
                         int max_seq_len, int past_kv_len, int max_input_len, const int* sequence_lengths,
                         const int* input_lengths, const int* masked_tokens, int* padding_ws,
                         const float* kv_scale_orig_quant, const float* kv_scale_quant_orig, int int8_kv,
                         int rotary_dim, float q_scaling, cudaStream_t stream) {
  if (head_size != 128) return 2;
  invokeUpdatePaddingCount(padding_ws, input_lengths, max_input_len, batch, stream);
  Masked_multihead_attention_params<uint16_t> p;
  memset(&p, 0, sizeof(p));
  const int hidden = num_heads * head_size;
  p.out = reinterpret_cast<uint16_t*>(out);
  p.q = reinterpret_cast<const uint16_t*>(qkv);
  p.k = p.q + hidden;
  p.v = p.q + 2 * hidden;
  p.stride = 3 * hidden;
  p.batch_size = batch;
  p.beam_width = 1;
  p.memory_max_len = max_seq_len;
  p.length_per_sample = sequence_lengths;
  p.timestep = past_kv_len;  // step + 0 - 1 with step = past_kv_len + 1
  p.num_heads = num_heads;
  p.hidden_size_per_head = head_size;
  p.rotary_embedding_dim = rotary_dim;
  p.neox_rotary_style = true;
  p.inv_sqrt_dh = 1.f / (sqrtf((float) head_size) * q_scaling);
  p.total_padding_tokens = padding_ws;
  p.masked_tokens = masked_tokens;
  p.max_input_length = max_input_len;
  p.int8_kv_cache = int8_kv != 0;
  if (int8_kv) { p.kv_scale_orig_quant = kv_scale_orig_quant; p.kv_scale_quant_orig = kv_scale_quant_orig; }
  const int elem = int8_kv ? 1 : 2;
  KVLinearBuffer kv(batch, 1, max_seq_len, num_heads * head_size * elem);
  kv.data = reinterpret_cast<int8_t*>(kv_cache);
  mmha::mmha_launch_kernel<uint16_t, KVLinearBuffer, Masked_multihead_attention_params<uint16_t>, 128>(p, kv, stream);
  return (int) cudaGetLastError();
}

int ref_layernorm_quant_half(void* out, const void* x, const void* gamma, const void* beta, float eps, int tokens,
                             int hidden, int use_diff_of_squares, const float* scale, float* dyn_scale,
                             int8_t* out_quant, cudaStream_t stream) {
  invokeGeneralLayerNorm<half>((half*) out, (const half*) x, (const half*) gamma, (const half*) beta, eps, tokens,
                               hidden, stream, use_diff_of_squares != 0, scale, dyn_scale, out_quant);
  return (int) cudaGetLastError();
}

int ref_per_token_quant_half(int8_t* dst, const void* src, int64_t rows, int64_t cols, float* scales,
                             cudaStream_t stream) {
  invokePerTokenQuantization<half>(dst, (const half*) src, rows, cols, scales, stream);
  return (int) cudaGetLastError();
}

int ref_quantize_half(int8_t* dst, const void* src, int64_t size, const float* scale, cudaStream_t stream) {
  invokeQuantization<half>(dst, (const half*) src, size, scale, stream, 65535);
  return (int) cudaGetLastError();
}

// weight must be in the reference's pre-processed (interleaved, biased) layout -> produce it with
// libref_host.so:ref_symmetric_quantize.
int ref_weight_only_gemv_half(const void* x, const int8_t* w_processed, const void* scales, void* out, int k, int n,
                              int bits, cudaStream_t stream) {
  weight_only_gemv_launcher<int8_t, half>((const half*) x, w_processed, (const half*) scales, nullptr, (half*) out,
                                          k, n, cutlass_kernels::ActivationType::Identity,
                                          bits == 8 ? QuantType::INT8_WEIGHT_ONLY : QuantType::PACKED_INT4_WEIGHT_ONLY,
                                          stream);
  return (int) cudaGetLastError();
}
}
