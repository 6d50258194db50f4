/*
 * trtllm_b200.h — C ABI of the B200-native LLaMA decoder hot path (libtrtllm_llama_b200.so).
 *
 * Plain pointers and sizes only (device pointers unless a parameter says "host"); every call is
 * asynchronous on `stream` and returns 0 on success, a negative code for a rejected argument, or a
 * positive cudaError_t.  There is no CPU fallback: every entry point launches sm_100a kernels.
 *
 * Each group cites the reference interface it replaces
 * (T/ = tensorrt_llm_july-release-v1, K/ = T/cpp/tensorrt_llm/kernels, P/ = T/cpp/tensorrt_llm/plugins).
 * The plugin-level (IPluginV2DynamicExt) and engine-level entry points are in trtllm_b200_plugin.h /
 * trtllm_b200_runtime.h and sit on top of these.
 */
#ifndef TRTLLM_B200_H
#define TRTLLM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* tb_stream_t; /* == cudaStream_t */

/* ---- library ------------------------------------------------------------------------------- */
const char* tb_version(void);
/* 0 iff a CUDA device of compute capability 10.x is current. */
int tb_check_device(void);

/* ---- RMSNorm / LayerNorm (+ residual add) (+ int8 quantisation) ------------------------------
 * replaces K/layernormKernels.cu:233-264 invokeGeneralLayerNorm (+quant args) and the TRT-native
 * rms_norm T/tensorrt_llm/functional.py:3195-3219.
 * x, residual, sum_out, gamma, beta, out: fp16.  If residual != NULL, h = x + residual is
 * normalised and (if sum_out != NULL) h is written to sum_out.                                  */
int tb_rmsnorm(void* out, const void* x, const void* residual, void* sum_out, const void* gamma, float eps,
               int rows, int hidden, tb_stream_t stream);
/* dynamic != 0: per-token scales written to scale_out[rows]; else static scale *scale_in.
 * layernorm != 0 selects the reference's mean-subtracting LayerNorm (beta may be NULL).         */
int tb_rmsnorm_quant(int8_t* out_q, float* scale_out, const void* x, const void* residual, void* sum_out,
                     const void* gamma, const void* beta, const float* scale_in, float eps, int rows, int hidden,
                     int dynamic, int layernorm, tb_stream_t stream);

/* ---- quantisers: replace K/quantization.cu:67-84 invokeQuantization, :119-130 invokePerTokenQuantization */
int tb_quantize_per_token(int8_t* dst, float* scales, const void* src, int rows, int cols, int src_is_fp32,
                          tb_stream_t stream);
int tb_quantize_tensor(int8_t* dst, const void* src, int64_t size, const float* scale, int src_is_fp32,
                       tb_stream_t stream);

/* ---- decode-shape GEMV (M <= tb_gemv_max_rows: 8 on the tensor-core kernel, 4 otherwise) ----------
 * kind: 0 fp16 weights [N,K]; 1 int8 weight-only [N,K] + fp16 scales[N]; 2 int4 weight-only [N,K/2];
 *       3 W8A8 SmoothQuant (x int8 [M,K], w int8 [N,K], sc per-channel, sr per-token fp32).
 * replaces K/weightOnlyMatrixVectorMultiplication.cu:371-378 weight_only_gemv_launcher and the M<=4
 * calls of CutlassInt8GemmRunner::gemm / cuBLAS GemmPlugin.
 * swiglu != 0: w holds [N = 2*inter, K] (gate rows then up rows), y is [M, inter] = silu(gate)*up.
 * y_f32 != NULL writes fp32 instead of fp16 (lm_head logits).                                   */
int tb_gemv_max_rows(int kind, int K);
/* 1: this (kind, rows, K) runs on the tensor-core GEMV (gemv_mma_kernel), 0: on the FMA GEMV (gemv_kernel) — the reference
 * picks between weight_only_gemv_launcher and the CUTLASS runner the same way (weightOnlyQuantMatmulPlugin.cpp:236-262) */
int tb_gemv_on_tensor_cores(int kind, int M, int K);
int tb_gemv(int kind, void* y, float* y_f32, const void* x, const void* w, const void* w_scale, const float* sc,
            const float* sr, int sc_per_channel, int sr_per_token, const void* residual, int M, int N, int K,
            int swiglu, tb_stream_t stream);
/* tb_gemv with a fused activation prologue (x is fp16 [M,K] whenever prologue != 0):
 *   1  RMSNorm(x, gamma, eps)                        — the TRT-native rms_norm in front of qkv / gate+up
 *   2  RMSNorm + dynamic per-token int8 (kind 3)     — RmsnormQuantization (K/layernormKernels.cu:141-193 semantics)
 *   3  dynamic per-token int8 (kind 3)               — QuantizePerToken (K/quantization.cu:93-117)
 * With 2 and 3 the per-token scales are computed in the kernel and `sr` is ignored.               */
int tb_gemv_fused(int kind, void* y, float* y_f32, const void* x, const void* w, const void* w_scale, const float* sc,
                  const float* sr, int sc_per_channel, int sr_per_token, const void* residual, int M, int N, int K,
                  int swiglu, int prologue, const void* gamma, float eps, tb_stream_t stream);
/* Optional one-shot hint for the calling thread's NEXT tb_gemv / tb_gemv_fused launch: byte ranges (the head of the weight
 * matrix the projection after it will stream; two ranges for a gate|up matrix) that its warps request into L2 as they run
 * out of rows, so HBM keeps streaming through the grid's tail, the launch gap and a short attention kernel in between.
 * No effect on results.  NULL / 0 clears the hint.                                                   */
int tb_gemv_hint_next(const void* a, size_t a_bytes, const void* b, size_t b_bytes);

/* ---- tcgen05 GEMM (any M) ---------------------------------------------------------------------
 * kind as tb_gemv.  out_type: 0 fp16, 1 fp32, 2 int32 (SmoothQuantGemm type_id half/float/int32).
 * replaces CutlassInt8GemmRunner<T>::gemm (K/cutlass_kernels/int8_gemm/int8_gemm.h:108-110),
 * CutlassFpAIntBGemmRunner<T,W>::gemm (K/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm.h:75-76) and
 * the cuBLASLt GemmPlugin (P/gemmPlugin/gemmPlugin.cpp:121-230).
 * workspace: tb_gemm_tc_workspace_bytes(M,N,K) bytes of scratch (need not be initialised: a TensorRT
 * workspace is fine).  counters: tb_gemm_tc_counter_bytes() bytes, zeroed ONCE by the owner; the
 * split-K arrival counters reset themselves.  force_splits / force_nt: 0 = automatic (test hooks). */
size_t tb_gemm_tc_workspace_bytes(int M, int N, int K);
size_t tb_gemm_tc_counter_bytes(void);
int tb_gemm_tc(int kind, void* c, int out_type, const void* x, const void* w, const void* w_scale, const float* sc,
               const float* sr, int sc_per_channel, int sr_per_token, const void* residual, int M, int N, int K,
               void* workspace, size_t workspace_bytes, int* counters, int force_splits, int force_nt,
               tb_stream_t stream);
/* The gate / up projection of the GatedMLP with SwiGLU in the tcgen05 epilogue (prefill shapes; north_star's "fused dequant +
 * SwiGLU epilogue"): w [N = 2*inter, K] holds the gate rows, then the up rows; c fp16 [M, inter] = silu(x.gate^T) * (x.up^T)
 * (T/tensorrt_llm/layers/mlp.py:43-73).  kind 0 fp16 | 3 SmoothQuant int8 (sc per-channel [N] or [1], sr per-token [M] or [1],
 * scale grouping of CE/epilogue/threadblock/epilogue_per_row_per_col_scale.h:279-349).  CTA-pair kernel: each stage holds the
 * pair's gate rows and the matching up rows, two tcgen05.mma per k-step accumulate them side by side in TMEM.  Bit-identical
 * to tb_gemm_tc + tb_swiglu. */
int tb_gemm_tc_swiglu(int kind, void* c, const void* x, const void* w, const float* sc, const float* sr, int sc_per_channel,
                      int sr_per_token, int M, int N, int K, tb_stream_t stream);

/* ---- attention --------------------------------------------------------------------------------
 * decode step: replaces masked_multihead_attention(params, kvbuf, stream)
 * (K/decoderMaskedMultiheadAttention.h:184-199) as driven by
 * GPTAttentionPluginCommon::enqueueGeneration (P/gptAttentionCommon/gptAttentionCommon.cpp:649-780).
 * kv_cache [B,2,H,S_max,Dh] int8|fp16 updated in place at position seq_lens[b] (or past_len).
 * len_cap: host upper bound on any seq_lens[b] (sizes shared memory; == past_len when the host
 * knows it).  workspace: tb_mmha_workspace_bytes() of scratch; counters: tb_mmha_counter_bytes(),
 * zeroed ONCE by the owner (self-resetting split-L arrival counters).                            */
size_t tb_mmha_workspace_bytes(int batch, int num_heads, int max_splits);
size_t tb_mmha_counter_bytes(int batch, int num_heads);
int tb_mmha_num_splits(int batch, int num_heads, int len_hint, int max_splits);
int tb_mmha_decode(void* out, const void* qkv, void* kv_cache, const int* seq_lens, const int* input_lengths,
                   const int* masked_tokens, const float* kv_scale_orig_quant, const float* kv_scale_quant_orig,
                   void* workspace, int* counters, int batch, int num_heads, int head_size, int max_seq_len,
                   int past_len,
                   int max_input_len, int len_cap, int rotary_dim, float q_scaling, int int8_kv, int nsplit,
                   tb_stream_t stream);
/* test / A-B hook: which streaming loops an int8 cache uses: -1 automatic (tensor-core loops from 512 cached positions),
 * 0 FMA loops, 1 tensor-core loops.  Returns the previous mode.  Process-wide; not for concurrent use.            */
int tb_mmha_set_mode(int mode);
/* same, with max_input_len read from a device int when max_input_len_dev != NULL (together with seq_lens this leaves
 * no per-request value in the launch arguments: one captured CUDA graph serves every step of every prompt length). */
int tb_mmha_decode_dev(void* out, const void* qkv, void* kv_cache, const int* seq_lens, const int* input_lengths,
                       const int* masked_tokens, const int* max_input_len_dev, const float* kv_scale_orig_quant,
                       const float* kv_scale_quant_orig, void* workspace, int* counters, int batch, int num_heads,
                       int head_size, int max_seq_len, int past_len, int max_input_len, int len_cap, int rotary_dim,
                       float q_scaling, int int8_kv, int nsplit, tb_stream_t stream);
/* Beam search (GPTAttention input 7, cache_indirection [batch / beam_width, beam_width, max_seq_len] int32): cached position t
 * of row (b, w) is read from the cache of row (b, cache_indirection[b][w][t]); the appended position is the row's own
 * (K/decoderMaskedMultiheadAttention/decoderMaskedMultiheadAttentionTemplate.h:1137-1146,1624-1631).  `batch` counts rows
 * (batch entries x beams).  Contiguous cache only.  Everything else as tb_mmha_decode_dev. */
int tb_mmha_decode_beams(void* out, const void* qkv, void* kv_cache, const int* cache_indirection, int beam_width,
                         const int* seq_lens, const int* input_lengths, const int* masked_tokens, const int* max_input_len_dev,
                         const float* kv_scale_orig_quant, const float* kv_scale_quant_orig, int batch, int num_heads,
                         int head_size, int max_seq_len, int past_len, int max_input_len, int len_cap, int rotary_dim,
                         float q_scaling, int int8_kv, int nsplit, tb_stream_t stream);
/* Paged KV cache (GPTAttention field paged_kv_cache; K/kvCacheUtils.h:34-112 KVBlockArray, beam width 1): instead of one
 * buffer, block_pointers [B, 2, max_blocks_per_seq] (int64 device addresses; K table then V table per sequence, as
 * T/tensorrt_llm/runtime/kv_cache_manager.py:163-184 builds it) name blocks of tokens_per_block positions laid out
 * [H, tokens_per_block, Dh]; position t of head h is row h * tokens_per_block + t % tokens_per_block of block
 * t / tokens_per_block.  tokens_per_block: a power of two >= 16.  Everything else as tb_mmha_decode_dev. */
int tb_mmha_decode_paged(void* out, const void* qkv, const int64_t* block_pointers, int tokens_per_block,
                         int max_blocks_per_seq, const int* seq_lens, const int* input_lengths, const int* masked_tokens,
                         const int* max_input_len_dev, const float* kv_scale_orig_quant, const float* kv_scale_quant_orig,
                         int batch, int num_heads, int head_size, int past_len, int max_input_len, int len_cap,
                         int rotary_dim, float q_scaling, int int8_kv, int nsplit, tb_stream_t stream);
int tb_context_attention_paged(void* out, void* qkv, const int64_t* block_pointers, int tokens_per_block,
                               int max_blocks_per_seq, const int* input_lengths, const float* kv_scale_orig_quant,
                               void* workspace, int batch, int seq_len, int num_heads, int head_size, int rotary_dim,
                               float q_scaling, int int8_kv, tb_stream_t stream);
/* context phase: replaces GPTAttentionPluginCommon::enqueueContext (gptAttentionCommon.cpp:361-620).
 * qkv [B,S,3*H*Dh] fp16 is rotated in place (q,k), out [B,S,H*Dh].
 * workspace: non-NULL (tb_context_attention_workspace_bytes(), a nominal 256 bytes since V is consumed in place as an
 * MN-major tcgen05 operand) selects the tcgen05 kernel; NULL selects the older warp-MMA kernel.       */
size_t tb_context_attention_workspace_bytes(int batch, int seq_len, int num_heads);
int tb_context_attention(void* out, void* qkv, void* kv_cache, const int* input_lengths,
                         const float* kv_scale_orig_quant, void* workspace, int batch, int seq_len, int num_heads,
                         int head_size, int max_seq_len, int rotary_dim, float q_scaling, int int8_kv,
                         tb_stream_t stream);

/* ---- glue (TRT-native ops in the reference, k14) ---------------------------------------------- */
int tb_embedding(void* out, const void* table, const int* ids, int tokens, int hidden, int vocab, tb_stream_t s);
int tb_swiglu(void* out, const void* gate, const void* up, int rows, int inter, int in_stride, tb_stream_t s);
/* SwiGLU + QuantizePerToken in one pass (prefill, SmoothQuant): dst int8 [rows, inter], scales fp32 [rows]; bit-identical
 * to tb_swiglu followed by tb_quantize_per_token (LQ/llama_model.py MLP act + T/quantization/functional.py:135-151). */
int tb_swiglu_quant(int8_t* dst, float* scales, const void* gate, const void* up, int rows, int inter, int in_stride,
                    tb_stream_t s);
int tb_add(void* out, const void* a, const void* b, int64_t n, tb_stream_t s);
int tb_gather_last_token(void* out, const void* in, const int* last_ids, int batch, int seq, int hidden,
                         tb_stream_t s);
/* the same for packed rows (remove_input_padding): in [sum(lens), hidden], sequence b ends at row sum(lens[:b+1]) - 1 */
int tb_gather_last_token_packed(void* out, const void* in, const int* lens, int batch, int hidden, tb_stream_t s);
int tb_argmax(int* out, const float* logits, int rows, int vocab, int vocab_stride, tb_stream_t s);
int tb_advance_step(const int* new_ids, int* input_ids, int* output_ids, int* seq_lens, int* step_pos, int batch,
                    int out_stride, tb_stream_t s);
/* greedy stop criterion (replaces K/stopCriteriaKernels.cu for top_k = 1): *all_done = every sequence of out_ids
 * [batch, out_stride] has end_id among its first n_done ids; pad != 0 also overwrites the positions after a sequence's
 * first end_id, up to n_pad, with end_id (the reference's output for finished sequences). all_done may be NULL. */
int tb_finished(int* all_done, int* out_ids, int batch, int out_stride, int n_done, int n_pad, int end_id, int pad,
                tb_stream_t s);
int tb_half_to_float(float* out, const void* in, int64_t n, tb_stream_t s);
int tb_fill_int(int* p, int value, int n, tb_stream_t s);
int tb_tile_int(int* p, int n, int w, tb_stream_t s);
/* teacher forcing (parity checks): the token of the column produced last (*step_pos - 1) becomes ids[b], in the next step's
 * input ids and in output_ids [batch, out_stride] */
int tb_force_ids(const int* ids, int* input_ids, int* output_ids, const int* step_pos, int batch, int out_stride, tb_stream_t s);
/* packed <-> padded token rows (GPTAttention remove_input_padding, gptAttentionCommon.cpp:467-478): sequence b owns packed
 * rows [sum(lens[:b]), + lens[b]) and padded rows [b * seq, b * seq + lens[b]); lens is a DEVICE array; unpack zero-fills the
 * padded tail rows.  row_bytes: a multiple of 16. */
int tb_unpack_rows(void* padded, const void* packed, const int* lens, int batch, int seq, int row_bytes, tb_stream_t s);
int tb_pack_rows(void* packed, const void* padded, const int* lens, int batch, int seq, int row_bytes, tb_stream_t s); /* in place p[i*w + j] = p[i]; n*w <= 1024 (_tile_beam_width) */
int tb_copy(void* dst, const void* src, size_t bytes, tb_stream_t s); /* device-to-device */
/* in [tp, rows, vocab_local] fp16 (all-gathered vocab-parallel lm_head) -> out [rows, tp*vocab_local] fp32 */
int tb_gather_logits(float* out, const void* in, int rows, int vocab_local, int tp, tb_stream_t s);

/* ---- sampling beyond greedy: temperature, top-k, top-p, top-k + top-p ---------------------------------
 * replaces the sampling half of DynamicDecodeOp (T/cpp/tensorrt_llm/thop/dynamicDecodeOp.cpp:359-363):
 * K/samplingPenaltyKernels.cu:77-93 (temperature), K/samplingTopKKernels.cu:118-319 (top_k > 0: the k largest logits,
 * r = u * top_p * sum exp(l - l_max), walk in descending order), K/samplingTopPKernels.cu:882-1010 (top_k == 0:
 * softmax, r = u * top_p, first token of the descending order whose inclusive cumulative probability reaches r).
 * u in (0, 1] is word 0 of Philox4x32-10(counter = (step, 0, row, 0), key = seed) under curand_uniform's mapping;
 * step is read from *step_dev when step_dev != NULL (CUDA-graph replay).  logits fp32 [rows, vocab_stride];
 * finished (optional, [rows]): finished rows emit end_id; uniform_out (optional, [rows]) receives u.
 * top_k <= 1024, 0 < top_p <= 1, vocab <= 51200.                                                               */
int tb_sample(int* out_ids, const float* logits, int rows, int vocab, int vocab_stride, int top_k, float top_p,
              float temperature, unsigned long long seed, const int* step_dev, int step, const int* finished,
              int end_id, float* uniform_out, tb_stream_t stream);

/* ---- beam search (SamplingConfig.num_beams > 1; replaces the beam half of DynamicDecodeOp:
 * K/onlineSoftmaxBeamsearchKernels.cu:112-300,402-592, layers/onlineBeamSearchLayer.cu:30-62,
 * layers/baseBeamSearchLayer.cu:29-67, K/decodingKernels.cu:31-170 gatherTree) ------------------------------------
 * rows = batch entries x beam_width (beam fastest).  All state is device-resident:
 *   cum_log_probs [rows] fp32, finished [rows] int32, beam_lens [rows] int32 (the decoder's sequence lengths),
 *   out_ids_t / parent_ids_t [max_new][rows] time-major, cache indirections [batch][beam][max_seq_len] int32.
 * tb_beam_init: cum = {0, -1e20, ...} per batch entry, finished = 0, beam_lens = *max_in_dev, both indirections 0.
 * tb_beam_search_step: logits fp32 [rows][vocab_stride] (broadcast_rows != 0: [batch][vocab_stride], every beam reads its
 *   entry's row — the step after the context phase); writes column *step_dev of out_ids_t / parent_ids_t, next_ids [rows],
 *   updates cum / finished / beam_lens and writes tgt_indir from src_indir for positions [0, *max_in_dev + *step_dev].
 *   length_penalty 0 = none.  beam_width <= 16.  workspace: tb_beam_workspace_bytes.
 * tb_gather_tree: out [rows][n_steps] = each final beam's token path through parent_ids, end_id after the first end_id. */
size_t tb_beam_workspace_bytes(int rows, int beam_width);
int tb_beam_init(float* cum_log_probs, int* finished, int* beam_lens, int* indir_a, int* indir_b, const int* max_in_dev,
                 int rows, int beam_width, int max_seq_len, tb_stream_t stream);
int tb_beam_search_step(const float* logits, int vocab, int vocab_stride, int broadcast_rows, int rows, int beam_width,
                        float length_penalty, int end_id, const int* step_dev, const int* max_in_dev, float* cum_log_probs,
                        int* finished, int* beam_lens, int* out_ids_t, int* parent_ids_t, int* next_ids, const int* src_indir,
                        int* tgt_indir, int max_seq_len, void* workspace, tb_stream_t stream);
int tb_gather_tree(int* out, const int* out_ids_t, const int* parent_ids_t, int rows, int beam_width, int n_steps, int end_id,
                   tb_stream_t stream);

/* ---- whole decode step in one persistent kernel (1..tb_decode_step_max_batch() token rows) -----------------------
 * replaces, per generated token, the plugin schedule of GenerationSession.decode's step
 * (T/tensorrt_llm/runtime/generation.py:852-963): every layer's projections (Gemm / WeightOnlyQuantMatmul /
 * SmoothQuantGemm at decode shapes), GPTAttention's generation phase, RmsnormQuantization / QuantizePerToken, the
 * TensorRT-native glue, lm_head and the greedy DynamicDecodeOp — one cooperative launch of one CTA per SM whose
 * producer warps stream the model's weights through shared memory across phase boundaries (csrc/decode_step.cu).
 * kind as tb_gemv; weights in this library's layouts ([N,K] rows; fc_gate = gate rows then up rows); scales fp16
 * (weight-only) or fp32 (SmoothQuant), NULL for fp16 weights.  All buffers are device pointers owned by the caller:
 * h_a/h_b/qkv/att/act fp16 scratch of max_batch rows, logits fp32 [max_batch, vocab]; ids / seq_lens / step_pos /
 * out_ids / next_ids / in_lens / max_in are the device-resident step state (as tb_advance_step / tb_mmha_decode_dev).
 * create returns < 0 when the configuration is not supported (the caller keeps the per-operator path).          */
typedef struct tb_decode_step tb_decode_step;
typedef struct {
  int32_t kind, layers, hidden, heads_local, inter_local, vocab_local, vocab, max_batch, max_seq_len, int8_kv, out_stride;
  float rms_eps;
  int32_t tp_size, tp_rank;
} tb_decode_step_config;
typedef struct {
  const void *w_qkv, *w_dense, *w_fc_gate, *w_proj, *s_qkv, *s_dense, *s_fc_gate, *s_proj, *ln_in, *ln_post;
  void* kv_cache;
  const float *kv_orig_quant, *kv_quant_orig;
} tb_decode_step_layer;
typedef struct {
  const void *emb, *ln_f, *lm_head;
  void *h_a, *h_b, *qkv, *att, *act;
  float* logits;
  int32_t *ids, *seq_lens, *step_pos, *out_ids, *next_ids;
  const int32_t *in_lens, *max_in;
  /* tensor parallel (cfg.tp_size > 1): rank r's peer-mapped scratch of tb_decode_step_tp_bytes() bytes, zero-initialised,
   * as mapped into this process (tp_peers[tp_rank] is the local one) — e.g. tb_ar_extra().  The gathered fp32 logits
   * [max_batch, vocab] of a tensor-parallel step are at tb_decode_step_tp_logits(). */
  void* tp_peers[8];
} tb_decode_step_buffers;
size_t tb_decode_step_tp_bytes(const tb_decode_step_config* cfg);
const float* tb_decode_step_tp_logits(const tb_decode_step* d);
int tb_decode_step_max_batch(void);
int tb_decode_step_create(tb_decode_step** out, const tb_decode_step_config* cfg, const tb_decode_step_layer* layers,
                          const tb_decode_step_buffers* buffers);
void tb_decode_step_destroy(tb_decode_step* d);
int tb_decode_step_launch(tb_decode_step* d, int batch, tb_stream_t stream);
/* ring depth (4 KB stages), dynamic shared memory and grid of the launch for `batch` rows (diagnostics) */
int tb_decode_step_info(const tb_decode_step* d, int batch, int* stages, size_t* smem_bytes, int* grid);
/* diagnostics: record %globaltimer stamps of every CTA's pipeline in the next launches (per projection: after the grid
 * barrier, after activation staging, after the last weight stage, after the epilogue; per attention phase: after the
 * barrier, after the items); out_host (grid x 2048 u64, may be NULL) receives what was recorded so far. */
int tb_decode_step_trace(tb_decode_step* d, int enable, unsigned long long* out_host);

/* ---- measurement support: tensor-pipe ceiling of this GPU at the clock it sustains (SURVEY 8d asks for a measured
 * tcgen05 kind::i8 peak; MEASURED_PEAKS.json has HBM and cuBLAS bf16 only).  Launches `ctas` CTAs, each issuing
 * iters x 4 back-to-back tcgen05.mma (128 x 256 x 32 int8 for kind 0, 128 x 256 x 16 fp16 for kind 1) on resident
 * shared-memory tiles; *ops_out (host) receives the operation count of the launch; the caller times it with events. */
int tb_mma_peak(int kind, int iters, int ctas, int* sink, double* ops_out, tb_stream_t stream);

/* ---- one-shot NVLink all-reduce + residual add for decode-size messages ------------------------------------
 * replaces AllreducePlugin::enqueue -> ncclAllReduce (P/ncclPlugin/allreducePlugin.cpp:80-97) and the residual
 * add that follows it, for messages <= max_bytes.  Setup: every rank creates its context, the host exchanges the
 * 64-byte IPC handles (rank order) and every rank opens its peers.  Per call: the producer writes its fp16 partial
 * into tb_ar_buffer(set); consecutive calls must alternate set = 0, 1, 0, 1, ...; out = residual + sum_r partial_r
 * (rank-ordered fp32 sum: bit-identical on all ranks).                                                        */
typedef struct tb_ar tb_ar;
int tb_ar_create(tb_ar** out, int rank, int world, size_t max_bytes);
/* same, with `extra_bytes` of additional zero-initialised peer-mapped memory behind the all-reduce buffers (the fused decode
 * step keeps its cross-GPU flags, partial sums, arg-max candidates and gathered logits there); tb_ar_extra(a, r) is rank
 * r's copy of that area as mapped into this process (valid for r != own rank after tb_ar_open_peers). */
int tb_ar_create_ex(tb_ar** out, int rank, int world, size_t max_bytes, size_t extra_bytes);
void* tb_ar_extra(tb_ar* a, int r);
size_t tb_ar_extra_bytes(tb_ar* a);
void tb_ar_destroy(tb_ar* a);
int tb_ar_ipc_handle(tb_ar* a, void* out64);
int tb_ar_open_peers(tb_ar* a, const void* handles);
void* tb_ar_buffer(tb_ar* a, int set);
int tb_ar_allreduce(tb_ar* a, int set, void* out, const void* residual, int64_t n_half, tb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TRTLLM_B200_H */
