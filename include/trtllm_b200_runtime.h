/*
 * trtllm_b200_runtime.h — engine-level C ABI: the C++ runtime that plays TensorRT's role for the LLaMA
 * decoder (there is no TensorRT in this build).  It looks the plugin creators up in the
 * registry, creates the operators, and runs the fixed layer schedule of
 * LQ/llama_model.py:78-119,159-287 by calling IPluginV2DynamicExt::enqueue on each, with the
 * TensorRT-native glue ops (RMSNorm, residual, SwiGLU, embedding, last-token gather, lm_head, argmax)
 * fused into plugin epilogues or run as small kernels.  The decode step is captured in a CUDA graph, or — for
 * 1..4 sequences — runs as one persistent kernel over the same buffers (tbrt_set_decode_mode).
 *
 * Replaces: the serialised TensorRT engine + IExecutionContext (T/tensorrt_llm/runtime/generation.py:61-100)
 * and the step loop of GenerationSession.decode (generation.py:782-997) for greedy / sampling / beam search.
 */
#ifndef TRTLLM_B200_RUNTIME_H
#define TRTLLM_B200_RUNTIME_H

#include <stddef.h>
#include <stdint.h>

#include "trtllm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tbrt_engine tbrt_engine;

enum { TBRT_MODE_FP16 = 0, TBRT_MODE_W8 = 1, TBRT_MODE_W4 = 2, TBRT_MODE_SQ = 3 };

typedef struct {
  int32_t hidden, heads, inter, layers, vocab, head_size; /* LQ/build.py:59-71 */
  float rms_eps;
  int32_t mode;          /* TBRT_MODE_*: use_weight_only (+int4) / use_smooth_quant per-token+per-channel (LQ/build.py:276-324) */
  int32_t int8_kv;       /* --int8_kv_cache */
  int32_t max_batch, max_input_len, max_output_len; /* builder limits (LQ/build.py:73-75) */
  int32_t tp_size, tp_rank;   /* Mapping(world_size, rank) (T/tensorrt_llm/mapping.py) */
  int32_t use_cuda_graph;
  int32_t paged_kv_tokens_per_block; /* --paged_kv_cache (LQ/build.py:190-196): 0 = contiguous cache; else a power of two >= 16 */
} tbrt_config;

tbrt_engine* tbrt_create(const tbrt_config* cfg);
void tbrt_destroy(tbrt_engine* e);
const char* tbrt_last_error(void);

/* Bind one weight tensor (device pointer, not copied; the caller keeps it alive).  Names follow the
 * reference's module tree (LQ/weight_quant.py:176-446):
 *   vocab_embedding.weight [V, hidden] fp16     ln_f.weight [hidden] fp16     lm_head.weight [V/tp, hidden] fp16
 *   layers.{i}.input_layernorm.weight, layers.{i}.post_layernorm.weight [hidden] fp16
 *   layers.{i}.attention.qkv.weight    [3*hidden/tp, hidden]   layers.{i}.attention.dense.weight [hidden, hidden/tp]
 *   layers.{i}.mlp.fc_gate.weight      [2*inter/tp, hidden] (fc = gate_proj rows, then gate = up_proj rows)
 *   layers.{i}.mlp.proj.weight         [hidden, inter/tp]
 *   weight dtype by mode: fp16 | int8 [N,K] | packed int4 [N,K/2] | int8 [N,K];
 *   <linear>.per_channel_scale: fp16 [N] (weight-only) or fp32 [N] (SmoothQuant)
 *   layers.{i}.attention.kv_orig_quant_scale / kv_quant_orig_scale fp32 [1] (int8 KV, LQ/weight_quant.py:439-446) */
int tbrt_set_tensor(tbrt_engine* e, const char* name, const void* device_ptr, size_t bytes);
/* Check that every tensor is bound, create plugins, allocate activations / KV caches / workspace. */
int tbrt_finalize(tbrt_engine* e);
size_t tbrt_device_bytes(const tbrt_engine* e);

/* Context phase over a padded batch: ids [B,S] int32 and input_lengths [B] are DEVICE pointers. */
int tbrt_context(tbrt_engine* e, const int32_t* ids, const int32_t* input_lengths, int batch, int seq, tb_stream_t s);
/* One generation step for every sequence (token ids come from the previous step's argmax). */
/* tbrt_context on packed input (build.py --remove_input_padding; generation.py:355-363): device ids [tokens] = the prompts
 * back to back, device input_lengths [batch], seq = the longest prompt.  The projections run on `tokens` rows instead of
 * batch x seq; cache layout and generation steps are those of the padded batch. */
int tbrt_context_packed(tbrt_engine* e, const int32_t* ids, const int32_t* input_lengths, int batch, int tokens, int seq,
                        tb_stream_t s);
int tbrt_step(tbrt_engine* e, tb_stream_t s);
/* Teacher forcing for parity checks: replaces the token the last tbrt_context / tbrt_step chose by ids [batch] (device), so
 * two engines (e.g. tp = 1 and tp = N) can be stepped along the same token path while their logits are compared. */
int tbrt_force_ids(tbrt_engine* e, const int32_t* ids, tb_stream_t s);
/* fp32 logits [B, vocab] of the last context/step call (device pointer). */
const float* tbrt_logits(const tbrt_engine* e);
/* generated ids so far: device int32 [B, max_output_len] (column j = j-th new token). */
const int32_t* tbrt_output_ids(const tbrt_engine* e);
/* device pointer of layer i's KV cache [B,2,H/tp,S_max,Dh] (tests) */
void* tbrt_kv_cache(const tbrt_engine* e, int layer);

/* Whole request with HOST buffers (pinned for async copies): H2D ids/lengths, context, max_new-1
 * steps, D2H of out_ids [B, max_new]; returns after the stream is synchronised.
 * This is GenerationSession.decode (generation.py:782-997) for greedy sampling.  With tbrt_set_end_id(e, id >= 0) the
 * loop checks every 16 steps whether every sequence has produced end_id and stops early, and finished sequences are
 * padded with end_id, as the reference's decoder does; id < 0 (default) always runs max_new - 1 steps. */
int tbrt_generate(tbrt_engine* e, const int32_t* host_ids, const int32_t* host_lengths, int batch, int seq, int max_new,
                  int32_t* host_out_ids, tb_stream_t s);
/* Tensor parallel only: peer-memory all-reduce of the decode path (tb_ar_*).  After tbrt_finalize every rank reads its
 * 64-byte IPC handle, the host all-gathers them (rank order) and hands the table back; without this call the engine
 * uses the NCCL AllReduce plugin for every message size. */
int tbrt_ar_handle(tbrt_engine* e, void* out64);
int tbrt_ar_open(tbrt_engine* e, const void* handles);
int tbrt_set_end_id(tbrt_engine* e, int end_id);
/* Paged KV cache (cfg.paged_kv_tokens_per_block > 0): every layer owns a pool of max_batch * tbrt_kv_max_blocks_per_seq()
 * blocks [2][blocks][H/tp][tokens_per_block][Dh]; the host assigns pool blocks to sequences (KVCacheManager of
 * T/tensorrt_llm/runtime/kv_cache_manager.py) and hands the table over before the context phase and whenever a sequence
 * grows into a new block: block_ids host int32 [batch, blocks_per_seq], -1 = not allocated.  The engine turns it into the
 * per-layer pointer tables the GPTAttention plugin reads (block_pointers input). */
int tbrt_kv_max_blocks_per_seq(const tbrt_engine* e);
int tbrt_set_kv_blocks(tbrt_engine* e, const int32_t* host_block_ids, int batch, int blocks_per_seq, tb_stream_t s);
/* SamplingConfig (T/tensorrt_llm/runtime/generation.py:119-138): top_k = 1 (default) is greedy arg-max; top_k > 1 samples
 * among the k largest logits (top_p > 0 additionally restricts to that share of their mass); top_k = 0 with top_p > 0 is
 * nucleus sampling over the vocabulary; temperature scales the logits first (tb_sample).  The random stream is keyed by
 * (seed, generation step, batch row). */
int tbrt_set_sampling(tbrt_engine* e, int top_k, float top_p, float temperature, unsigned long long seed);
/* Beam search (SamplingConfig.num_beams > 1; generation.py:365-409,823-997 + DynamicDecodeOp's beam layer):
 *   tbrt_context(batch rows) -> tbrt_beam_begin (tiles the KV cache, lengths and logits beam_width times as
 *   generation.py:898-915, redoes the first token as a beam step; the engine then runs batch x beam_width rows, which must
 *   fit max_batch) -> tbrt_step x (n - 1) -> tbrt_beam_finalize (gather_tree: host_out [batch][beam_width][n_steps], best
 *   beam first; cum_log_probs_out [batch][beam_width] or NULL).  length_penalty: score = cum_log_prob / length^penalty
 *   (0 = none; SamplingConfig default 1).  Contiguous KV cache, beam_width in [2, 16]. */
int tbrt_beam_begin(tbrt_engine* e, int beam_width, float length_penalty, int end_id, tb_stream_t s);
int tbrt_beam_finalize(tbrt_engine* e, int32_t* host_out, float* cum_log_probs_out, int n_steps, tb_stream_t s);
/* Generation steps can run as ONE persistent kernel (tb_decode_step_*) whenever the engine's configuration and the batch
 * allow it (mode 1); mode 0 forces the per-operator plugin schedule (IPluginV2DynamicExt::enqueue per operator, CUDA
 * graph) — same weights, caches and step state, so the two can be compared step by step.  Mode -1 (default) picks the
 * path that measured faster on B200: the persistent kernel under tensor parallelism, the plugin schedule on one GPU.
 * tbrt_fused_step_available: largest batch the fused step takes for this engine (0: not available). */
int tbrt_set_decode_mode(tbrt_engine* e, int mode);
int tbrt_fused_step_available(const tbrt_engine* e);
/* the engine's tb_decode_step (NULL when not available), for tb_decode_step_info / tb_decode_step_trace */
void* tbrt_decode_step_handle(tbrt_engine* e);
/* generation steps (incl. the context phase) the last tbrt_generate actually ran */
int tbrt_last_steps(const tbrt_engine* e);
/* kernels launched by the last tbrt_context / tbrt_step / tbrt_generate call */
int64_t tbrt_last_launches(const tbrt_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* TRTLLM_B200_RUNTIME_H */
