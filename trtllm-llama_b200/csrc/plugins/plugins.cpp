// The IPluginV2DynamicExt operators of the LLaMA decoder hot path, each a thin shape/type shim over the
// sm_100a kernels behind the C ABI of include/trtllm_b200.h.  Names, versions ("1"), namespace
// ("tensorrt_llm"), field names, input order, output shapes/types and serialisation order follow the
// reference plugin of the same name (cited per class) so an engine builder that looks these creators up
// finds drop-in replacements.  Fields marked [ext] are additions (absent => reference behaviour).
#include <cuda_fp16.h>
#include <cmath>

#include "pluginBase.h"

using namespace nvinfer1;

namespace tb {
namespace plugins {

namespace {
using FT = PluginFieldType;

DimsExprs with_last_dim(const DimsExprs& in, const IDimensionExpr* last) {
  DimsExprs r = in;
  r.d[r.nbDims - 1] = last;
  return r;
}
bool linear(const PluginTensorDesc& d, DataType t) { return d.type == t && d.format == TensorFormat::kLINEAR; }
}  // namespace

// =====================================================================================================
// GPTAttention v1 — P/gptAttentionPlugin/gptAttentionPlugin.{h,cpp}, P/gptAttentionCommon/gptAttentionCommon.{h,cpp}
// inputs  0 qkv [B,S,3*H*Dh]  1 past_key_value [B,2,H,S_max,Dh]  2 sequence_length [B]  3 past_key_value_length [2] (HOST:
//         {past_len, is_context})  4 masked_tokens [B,S_max]  5 input_lengths [B]  6 max_input_length [max_in] (shape only)
//         7 cache_indirection [B,beam,S_max]  (8 kv_orig_quant_scale [1], 9 kv_quant_orig_scale [1] iff int8 KV)
//         (next: block_pointers [B,beam,2,2*max_blocks] int32 view of int64 addresses iff paged_kv_cache — then input 1 is
//          the block pool [blocks,2,H,tokens_per_block,Dh]; P/gptAttentionPlugin/gptAttentionPlugin.cpp:204-235)
//         (last two iff in_flight_batching: host_input_lengths [B] int32 HOST, host_request_types [B] int32 HOST —
//          0 context, 1 generation, 2 none; gptAttentionPlugin.h:106-170)
//         remove_input_padding: input 0 is [1, num_tokens, 3*H*Dh], sequences packed back to back, output 0 likewise
// outputs 0 context [B,S,H*Dh]   1 present_key_value (same buffer as input 1: updated in place)
// =====================================================================================================
class GPTAttentionPlugin : public BasePlugin {
 public:
  static const char* type_name() { return "GPTAttention"; }
  static const std::vector<PluginField>& field_table() {
    static const std::vector<PluginField> t = {
        field_decl("num_heads", FT::kINT32), field_decl("head_size", FT::kINT32), field_decl("unidirectional", FT::kINT32),
        field_decl("q_scaling", FT::kFLOAT32), field_decl("rotary_embedding_dim", FT::kINT32),
        field_decl("neox_rotary_style", FT::kINT8), field_decl("context_fmha_type", FT::kINT8),
        field_decl("multi_block_mode", FT::kINT8), field_decl("multi_query_mode", FT::kINT8),
        field_decl("int8_kv_cache", FT::kINT32), field_decl("fp8_kv_cache", FT::kINT32),
        field_decl("remove_input_padding", FT::kINT8), field_decl("mask_type", FT::kINT32),
        field_decl("paged_kv_cache", FT::kINT32), field_decl("type_id", FT::kINT32),
        field_decl("in_flight_batching", FT::kINT32), field_decl("device_lengths", FT::kINT32) /*[ext]*/};
    return t;
  }
  explicit GPTAttentionPlugin(Fields& f) {
    num_heads_ = f.required<int32_t>("num_heads");
    head_size_ = f.required<int32_t>("head_size");
    unidirectional_ = f.optional<int32_t>("unidirectional", 1);
    q_scaling_ = f.optional<float>("q_scaling", 1.f);
    rotary_dim_ = f.optional<int32_t>("rotary_embedding_dim", 0);
    neox_ = f.optional<int32_t>("neox_rotary_style", 1) != 0;
    context_fmha_ = f.optional<int32_t>("context_fmha_type", 0);
    multi_block_ = f.optional<int32_t>("multi_block_mode", 0) != 0;
    multi_query_ = f.optional<int32_t>("multi_query_mode", 0) != 0;
    int8_kv_ = f.optional<int32_t>("int8_kv_cache", 0) != 0;
    fp8_kv_ = f.optional<int32_t>("fp8_kv_cache", 0) != 0;
    remove_padding_ = f.optional<int32_t>("remove_input_padding", 0) != 0;
    mask_type_ = f.optional<int32_t>("mask_type", 1 /*causal*/);
    paged_kv_ = f.optional<int32_t>("paged_kv_cache", 0) != 0;
    type_ = f.optional<int32_t>("type_id", (int32_t) DataType::kHALF);
    ifb_ = f.optional<int32_t>("in_flight_batching", 0) != 0;
    device_lengths_ = f.optional<int32_t>("device_lengths", 0) != 0;
    validate();
  }
  explicit GPTAttentionPlugin(Reader& r) {
    // same member order as the reference blob (P/gptAttentionCommon/gptAttentionCommon.cpp:862-890,
    // P/gptAttentionPlugin/gptAttentionPlugin.cpp:443-455), then this library's extension word
    num_heads_ = r.get<int32_t>(); head_size_ = r.get<int32_t>(); unidirectional_ = r.get<int32_t>();
    q_scaling_ = r.get<float>(); rotary_dim_ = r.get<int32_t>(); neox_ = r.get<bool>();
    context_fmha_ = r.get<bool>(); (void) r.get<bool>() /*fmha fp32 acc*/; multi_block_ = r.get<bool>();
    multi_query_ = r.get<bool>(); int8_kv_ = r.get<bool>(); fp8_kv_ = r.get<bool>(); remove_padding_ = r.get<bool>();
    mask_type_ = r.get<int32_t>(); paged_kv_ = r.get<bool>(); type_ = r.get<int32_t>(); ifb_ = r.get<bool>();
    device_lengths_ = r.get<bool>();
    validate();
  }
  size_t getSerializationSize() const noexcept override {
    return 4 * sizeof(int32_t) + sizeof(float) + 8 * sizeof(bool) + sizeof(int32_t) + sizeof(bool) + sizeof(int32_t) +
           2 * sizeof(bool);
  }
  void serialize(void* buf) const noexcept override {
    Writer w{static_cast<char*>(buf)};
    w.put(num_heads_); w.put(head_size_); w.put(unidirectional_); w.put(q_scaling_); w.put(rotary_dim_); w.put(neox_);
    w.put((bool) (context_fmha_ != 0)); w.put((bool) (context_fmha_ == 2)); w.put(multi_block_); w.put(multi_query_);
    w.put(int8_kv_); w.put(fp8_kv_); w.put(remove_padding_); w.put(mask_type_); w.put(paged_kv_); w.put(type_);
    w.put(ifb_); w.put(device_lengths_);
  }
  GPTAttentionPlugin* clone() const noexcept override {
    return new GPTAttentionPlugin(*this);
  }
  const char* getPluginType() const noexcept override { return type_name(); }
  int32_t getNbOutputs() const noexcept override { return 2; }
  DimsExprs getOutputDimensions(int32_t idx, const DimsExprs* in, int32_t, IExprBuilder& eb) noexcept override {
    if (idx == 0) return with_last_dim(in[0], eb.constant(num_heads_ * head_size_));
    return in[1];
  }
  DataType getOutputDataType(int32_t idx, const DataType* in, int32_t) const noexcept override {
    return idx == 0 ? in[0] : in[1];
  }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nb_in, int32_t) noexcept override {
    if (pos >= 2 && pos <= 7) return linear(io[pos], DataType::kINT32);
    if (int8_kv_ && (pos == 8 || pos == 9)) return linear(io[pos], DataType::kFLOAT);
    if (paged_kv_ && pos == (int8_kv_ ? 10 : 8)) return linear(io[pos], DataType::kINT32);
    if (ifb_ && (pos == nb_in - 2 || pos == nb_in - 1)) return linear(io[pos], DataType::kINT32);
    if (int8_kv_ && (pos == 1 || pos == nb_in + 1)) return linear(io[pos], DataType::kINT8);
    return linear(io[pos], (DataType) type_);
  }
  size_t getWorkspaceSize(const PluginTensorDesc* in, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
    // generation: split-L partials.  context: V^T for the tcgen05 kernel, 2 * B*S*hidden bytes (the reference sizes
    // ~6.4 GB of score buffers here, gptAttentionCommon.cpp:267-305).
    if (remove_padding_) {   // packed input: padded staging copies of qkv and of the output, then the padded kernels' own
      const size_t nseq = in[5].dims.d[0], max_in = in[6].dims.d[0], hid = (size_t) num_heads_ * head_size_;
      return align128(nseq * max_in * 3 * hid * 2) + align128(nseq * max_in * hid * 2) +
             align128(tb_context_attention_workspace_bytes((int) nseq, (int) max_in, num_heads_)) +
             align128(tb_mmha_workspace_bytes((int) nseq, num_heads_, kMaxSplits));
    }
    const size_t gen = tb_mmha_workspace_bytes(in[0].dims.d[0], num_heads_, kMaxSplits);
    const size_t ctx = tb_context_attention_workspace_bytes(in[0].dims.d[0], in[0].dims.d[1], num_heads_);
    return align128(gen > ctx ? gen : ctx);
  }
  int32_t enqueue(const PluginTensorDesc* id, const PluginTensorDesc*, const void* const* in, void* const* out,
                  void* workspace, cudaStream_t stream) noexcept override {
    if (remove_padding_) return enqueue_packed(id, nb_inputs(), in, out, workspace, stream);
    return guarded("GPTAttention::enqueue", [&]() -> int {
      const int B = id[0].dims.d[0], S = id[0].dims.d[1];
      const int max_in = id[6].dims.d[0];
      // paged cache: input 1 is the pool [blocks,2,H,tokens_per_block,Dh]; the per-sequence block tables come as an
      // int32 view of int64 device addresses [B, beam = 1, 2, 2 * max_blocks] (T/tensorrt_llm/runtime/kv_cache_manager.py:286)
      const int bp_idx = int8_kv_ ? 10 : 8;
      const int tpb = paged_kv_ ? id[1].dims.d[3] : 0;
      const int max_blocks = paged_kv_ ? id[bp_idx].dims.d[id[bp_idx].dims.nbDims - 1] / 2 : 0;
      const int64_t* block_ptrs = paged_kv_ ? static_cast<const int64_t*>(in[bp_idx]) : nullptr;
      if (paged_kv_) {
        TBP_REQUIRE(block_ptrs != nullptr && max_blocks > 0, "paged_kv_cache needs the block_pointers input");
        TBP_REQUIRE(id[bp_idx].dims.nbDims < 2 || id[bp_idx].dims.d[1] == 1, "paged cache: beam width 1 only");
      }
      const int S_max = paged_kv_ ? tpb * max_blocks : id[1].dims.d[3];
      const int* host_len = static_cast<const int*>(in[3]);          // HOST tensor {past_len, is_context}
      TBP_REQUIRE(host_len != nullptr, "past_key_value_length must be a host tensor");
      const int past_len = host_len[0];
      const bool is_context = host_len[1] != 0;
      const float* s_oq = int8_kv_ ? static_cast<const float*>(in[8]) : nullptr;
      const float* s_qo = int8_kv_ ? static_cast<const float*>(in[9]) : nullptr;
      void* cache = out[1] ? out[1] : const_cast<void*>(in[1]);      // in-place: runtime binds one buffer to both
      if (is_context) {
        if (paged_kv_)
          return tb_context_attention_paged(out[0], const_cast<void*>(in[0]), block_ptrs, tpb, max_blocks,
                                            static_cast<const int*>(in[5]), s_oq, workspace, B, S, num_heads_, head_size_,
                                            rotary_dim_, q_scaling_, int8_kv_, stream);
        return tb_context_attention(out[0], const_cast<void*>(in[0]), cache, static_cast<const int*>(in[5]), s_oq, workspace,
                                    B, S, num_heads_, head_size_, S_max, rotary_dim_, q_scaling_, int8_kv_, stream);
      }
      TBP_REQUIRE(S == 1, "generation phase expects one token per sequence");
      // device_lengths [ext]: the step position is read from sequence_length on the device so one
      // captured CUDA graph serves every step; shared memory is then sized for S_max.
      const int cap = device_lengths_ ? S_max - 1 : past_len;
      const int nsplit = tb_mmha_num_splits(B, num_heads_, cap, kMaxSplits);
      // device_lengths [ext]: a non-NULL input 6 then holds max_input_length as ONE device int (the reference only uses
      // its shape), so no per-request value is baked into a captured launch.  Split partials live in distributed
      // shared memory: the kernel needs no counters.
      const int* max_in_dev = device_lengths_ ? static_cast<const int*>(in[6]) : nullptr;
      if (paged_kv_)
        return tb_mmha_decode_paged(out[0], in[0], block_ptrs, tpb, max_blocks, static_cast<const int*>(in[2]),
                                    static_cast<const int*>(in[5]), static_cast<const int*>(in[4]), max_in_dev, s_oq, s_qo, B,
                                    num_heads_, head_size_, device_lengths_ ? 0 : past_len, max_in, cap, rotary_dim_,
                                    q_scaling_, int8_kv_, nsplit, stream);
      // beam search: cache_indirection [B / beam, beam, S_max] with beam > 1 selects, per cached position, which beam's
      // cache row is read (gptAttentionCommon.cpp:700-712 passes it to the kernel whenever beam_width > 1)
      const int beam = (in[7] && id[7].dims.nbDims == 3) ? id[7].dims.d[1] : 1;
      if (beam > 1) {
        TBP_REQUIRE(B % beam == 0 && id[7].dims.d[0] * beam == B && id[7].dims.d[2] == S_max,
                    "cache_indirection must be [batch / beam, beam, max_seq_len]");
        return tb_mmha_decode_beams(out[0], in[0], cache, static_cast<const int*>(in[7]), beam, static_cast<const int*>(in[2]),
                                    static_cast<const int*>(in[5]), static_cast<const int*>(in[4]), max_in_dev, s_oq, s_qo, B,
                                    num_heads_, head_size_, S_max, device_lengths_ ? 0 : past_len, max_in, cap, rotary_dim_,
                                    q_scaling_, int8_kv_, nsplit, stream);
      }
      return tb_mmha_decode_dev(out[0], in[0], cache, static_cast<const int*>(in[2]), static_cast<const int*>(in[5]),
                                static_cast<const int*>(in[4]), max_in_dev, s_oq, s_qo, workspace, nullptr, B, num_heads_,
                                head_size_, S_max, device_lengths_ ? 0 : past_len, max_in, cap, rotary_dim_, q_scaling_,
                                int8_kv_, nsplit, stream);
    });
  }

 private:
  int nb_inputs() const { return 8 + (int8_kv_ ? 2 : 0) + (paged_kv_ ? 1 : 0) + (ifb_ ? 2 : 0); }

  // remove_input_padding (+ in_flight_batching): P/gptAttentionPlugin/gptAttentionPlugin.cpp:150-200 enqueueImpl groups
  // consecutive requests of one type and runs each group (:202-372 enqueueSome) on its slice of the packed tokens, the
  // sequence-indexed inputs and the cache.  A context group is staged into the padded layout the tcgen05 prefill kernel
  // consumes (tb_unpack_rows), attended, and packed back; a generation group is one token per sequence already.
  int32_t enqueue_packed(const PluginTensorDesc* id, int nb_in, const void* const* in, void* const* out, void* workspace,
                         cudaStream_t stream) noexcept {
    return guarded("GPTAttention::enqueue (packed)", [&]() -> int {
      const int nseq = id[5].dims.d[0], max_in = id[6].dims.d[0];
      const int hid = num_heads_ * head_size_;
      const int bp_idx = int8_kv_ ? 10 : 8;
      const int tpb = paged_kv_ ? id[1].dims.d[3] : 0;
      const int max_blocks = paged_kv_ ? id[bp_idx].dims.d[id[bp_idx].dims.nbDims - 1] / 2 : 0;
      const int S_max = paged_kv_ ? tpb * max_blocks : id[1].dims.d[3];
      const int* host_len = static_cast<const int*>(in[3]);
      TBP_REQUIRE(host_len != nullptr, "past_key_value_length must be a host tensor");
      const float* s_oq = int8_kv_ ? static_cast<const float*>(in[8]) : nullptr;
      const float* s_qo = int8_kv_ ? static_cast<const float*>(in[9]) : nullptr;
      char* cache = static_cast<char*>(out[1] ? out[1] : const_cast<void*>(in[1]));
      const size_t seq_stride = paged_kv_ ? 0 : (size_t) 2 * num_heads_ * S_max * head_size_ * (int8_kv_ ? 1 : 2);
      const __half* x = static_cast<const __half*>(in[0]);
      __half* y = static_cast<__half*>(out[0]);
      const int* seq_lens = static_cast<const int*>(in[2]);
      const int* in_lens = static_cast<const int*>(in[5]);
      const int64_t* block_ptrs = paged_kv_ ? static_cast<const int64_t*>(in[bp_idx]) : nullptr;
      char* ws = static_cast<char*>(workspace);
      __half* pad_qkv = reinterpret_cast<__half*>(ws);
      __half* pad_out = reinterpret_cast<__half*>(ws + align128((size_t) nseq * max_in * 3 * hid * 2));
      void* sub_ws = reinterpret_cast<char*>(pad_out) + align128((size_t) nseq * max_in * hid * 2);

      auto some = [&](int seq0, int n, int tok0, bool is_context, int S) -> int {
        const int64_t* bp = block_ptrs ? block_ptrs + (size_t) seq0 * 2 * max_blocks : nullptr;
        if (is_context) {
          if (int rc = tb_unpack_rows(pad_qkv, x + (size_t) tok0 * 3 * hid, in_lens + seq0, n, S, 3 * hid * 2, stream)) return rc;
          int rc = paged_kv_ ? tb_context_attention_paged(pad_out, pad_qkv, bp, tpb, max_blocks, in_lens + seq0, s_oq, sub_ws, n, S,
                                                          num_heads_, head_size_, rotary_dim_, q_scaling_, int8_kv_, stream)
                             : tb_context_attention(pad_out, pad_qkv, cache + seq0 * seq_stride, in_lens + seq0, s_oq, sub_ws, n, S,
                                                    num_heads_, head_size_, S_max, rotary_dim_, q_scaling_, int8_kv_, stream);
          if (rc) return rc;
          return tb_pack_rows(y + (size_t) tok0 * hid, pad_out, in_lens + seq0, n, S, hid * 2, stream);
        }
        // generation: per-request lengths on the device.  In-flight batching has no padded batch: every request sits at its own
        // position and nothing is masked (the reference passes input_seq_length = 1, gptAttentionPlugin.cpp:287-290)
        const bool own_len = ifb_ || device_lengths_;
        const int cap = own_len ? S_max - 1 : host_len[0];
        const int nsplit = tb_mmha_num_splits(n, num_heads_, cap, kMaxSplits);
        const int* lens_arg = ifb_ ? nullptr : in_lens + seq0;
        const int* mask_arg = (ifb_ || !in[4]) ? nullptr : static_cast<const int*>(in[4]) + (size_t) seq0 * S_max;
        const int* max_in_dev = (!ifb_ && device_lengths_) ? static_cast<const int*>(in[6]) : nullptr;
        if (paged_kv_)
          return tb_mmha_decode_paged(y + (size_t) tok0 * hid, x + (size_t) tok0 * 3 * hid, bp, tpb, max_blocks, seq_lens + seq0,
                                      lens_arg, mask_arg, max_in_dev, s_oq, s_qo, n, num_heads_, head_size_, own_len ? 0 : host_len[0],
                                      ifb_ ? 1 : max_in, cap, rotary_dim_, q_scaling_, int8_kv_, nsplit, stream);
        return tb_mmha_decode_dev(y + (size_t) tok0 * hid, x + (size_t) tok0 * 3 * hid, cache + seq0 * seq_stride, seq_lens + seq0,
                                  lens_arg, mask_arg, max_in_dev, s_oq, s_qo, sub_ws, nullptr, n, num_heads_, head_size_, S_max,
                                  own_len ? 0 : host_len[0], ifb_ ? 1 : max_in, cap, rotary_dim_, q_scaling_, int8_kv_, nsplit,
                                  stream);
      };

      if (!ifb_) return some(0, nseq, 0, host_len[1] != 0, max_in);
      const int* host_in_lens = static_cast<const int*>(in[nb_in - 2]);      // HOST
      const int* req_types = static_cast<const int*>(in[nb_in - 1]);         // HOST: 0 context, 1 generation, 2 none
      TBP_REQUIRE(host_in_lens && req_types, "in_flight_batching needs host_input_lengths and host_request_types");
      TBP_REQUIRE(!in[7] || id[7].dims.nbDims < 3 || id[7].dims.d[1] == 1, "in-flight batching: beam width 1 only");
      int seq0 = 0, tok0 = 0, tok1 = 0, ref = req_types[0];
      for (int i = 0; i <= nseq; ++i) {
        if (i < nseq && req_types[i] == ref) {
          tok1 += host_in_lens[i];
          continue;
        }
        if (ref != 2) {
          TBP_REQUIRE(ref == 0 || ref == 1, "host_request_types holds 0 (context), 1 (generation) or 2 (none)");
          int S = 1;
          if (ref == 0)
            for (int j = seq0; j < i; ++j) S = host_in_lens[j] > S ? host_in_lens[j] : S;
          TBP_REQUIRE(S <= max_in, "a context request is longer than max_input_length");
          if (int rc = some(seq0, i - seq0, tok0, ref == 0, S)) return rc;
        }
        if (i < nseq) {
          seq0 = i;
          ref = req_types[i];
          tok0 = tok1;
          tok1 += host_in_lens[i];
        }
      }
      return 0;
    });
  }

  static constexpr int kMaxSplits = 32;
  void validate() const {
    TBP_REQUIRE(head_size_ == 128, "only head_size 128 (LLaMA-7B) is built");
    TBP_REQUIRE(rotary_dim_ == 0 || rotary_dim_ == head_size_, "rotary_embedding_dim must be 0 or head_size");
    TBP_REQUIRE(rotary_dim_ == 0 || neox_, "only neox-style rotary embedding is built");
    TBP_REQUIRE(is_half(type_), "only type_id = half is built");
    TBP_REQUIRE(!multi_query_ && !fp8_kv_, "multi-query / fp8 KV are out of scope (SURVEY 8f)");
    TBP_REQUIRE(!ifb_ || remove_padding_, "in_flight_batching needs remove_input_padding (gptAttentionPlugin.cpp:285)");
    TBP_REQUIRE(unidirectional_ == 1, "causal attention only");
  }
  int32_t num_heads_ = 0, head_size_ = 0, unidirectional_ = 1, rotary_dim_ = 0, context_fmha_ = 0, mask_type_ = 1;
  int32_t type_ = (int32_t) DataType::kHALF;
  float q_scaling_ = 1.f;
  bool neox_ = true, multi_block_ = false, multi_query_ = false, int8_kv_ = false, fp8_kv_ = false;
  bool remove_padding_ = false, paged_kv_ = false, ifb_ = false, device_lengths_ = false;
};

// =====================================================================================================
// SmoothQuantGemm v1 — P/smoothQuantGemmPlugin/smoothQuantGemmPlugin.{h,cpp}
// inputs 0 act int8 [..,K]  1 weight int8 [N,K] (declared fp32 [N,K/4])  2 scale_tokens fp32 [M,1]|[1,1]
//        3 scale_channels fp32 [1,N]|[1,1]  (4 residual fp16 [..,N] iff fused_residual [ext])
// output [..,N] type_id in {half, float, int32}
// =====================================================================================================
class SmoothQuantGemmPlugin : public BasePlugin {
 public:
  static const char* type_name() { return "SmoothQuantGemm"; }
  static const std::vector<PluginField>& field_table() {
    static const std::vector<PluginField> t = {field_decl("has_per_channel_scaling", FT::kINT32),
                                               field_decl("has_per_token_scaling", FT::kINT32),
                                               field_decl("type_id", FT::kINT32), field_decl("fused_swiglu", FT::kINT32),
                                               field_decl("fused_residual", FT::kINT32),
                                               field_decl("fused_prologue", FT::kINT32), field_decl("eps", FT::kFLOAT32)};
    return t;
  }
  explicit SmoothQuantGemmPlugin(Fields& f) {
    per_channel_ = f.optional<int32_t>("has_per_channel_scaling", 0) != 0;
    per_token_ = f.optional<int32_t>("has_per_token_scaling", 0) != 0;
    type_ = f.required<int32_t>("type_id");
    swiglu_ = f.optional<int32_t>("fused_swiglu", 0) != 0;
    residual_ = f.optional<int32_t>("fused_residual", 0) != 0;
    prologue_ = f.optional<int32_t>("fused_prologue", 0);
    eps_ = f.optional<float>("eps", 1e-6f);
    validate();
  }
  explicit SmoothQuantGemmPlugin(Reader& r) {
    // reference order perChannel, perToken, type (smoothQuantGemmPlugin.cpp:253-282); the reference's
    // CUTLASS tactic table that follows is replaced by this library's two extension flags
    per_channel_ = r.get<bool>(); per_token_ = r.get<bool>(); type_ = r.get<int32_t>();
    swiglu_ = r.get<bool>(); residual_ = r.get<bool>(); prologue_ = r.get<int32_t>(); eps_ = r.get<float>();
    validate();
  }
  size_t getSerializationSize() const noexcept override { return 4 * sizeof(bool) + 2 * sizeof(int32_t) + sizeof(float); }
  void serialize(void* buf) const noexcept override {
    Writer w{static_cast<char*>(buf)};
    w.put(per_channel_); w.put(per_token_); w.put(type_); w.put(swiglu_); w.put(residual_); w.put(prologue_); w.put(eps_);
  }
  SmoothQuantGemmPlugin* clone() const noexcept override {
    auto* p = new SmoothQuantGemmPlugin(*this);
    p->counters_ = DeviceCounters();
    return p;
  }
  const char* getPluginType() const noexcept override { return type_name(); }
  int32_t getNbOutputs() const noexcept override { return 1; }
  DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& eb) noexcept override {
    const IDimensionExpr* n = in[1].d[0];
    if (swiglu_) n = eb.operation(DimensionOperation::kFLOOR_DIV, *n, *eb.constant(2));
    return with_last_dim(in[0], n);
  }
  DataType getOutputDataType(int32_t, const DataType*, int32_t) const noexcept override { return (DataType) type_; }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nb_in, int32_t) noexcept override {
    if (pos == 0) return linear(io[pos], prologue_ ? DataType::kHALF : DataType::kINT8);
    if (pos == 1) return io[pos].format == TensorFormat::kLINEAR && (io[pos].type == DataType::kINT8 || io[pos].type == DataType::kFLOAT);
    if (pos == 2 || pos == 3) return linear(io[pos], DataType::kFLOAT);
    if (pos < nb_in) return linear(io[pos], DataType::kHALF);   // residual / gamma [ext]
    return linear(io[pos], (DataType) type_);
  }
  size_t getWorkspaceSize(const PluginTensorDesc* in, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
    return align128(tb_gemm_tc_workspace_bytes((int) rows_of(in[0].dims), in[1].dims.d[0], last_dim(in[0].dims)));
  }
  int32_t enqueue(const PluginTensorDesc* id, const PluginTensorDesc* od, const void* const* in, void* const* out,
                  void* workspace, cudaStream_t stream) noexcept override {
    return guarded("SmoothQuantGemm::enqueue", [&]() -> int {
      const int M = (int) rows_of(id[0].dims), N = id[1].dims.d[0], K = last_dim(id[0].dims);
      int next = 4;
      const void* res = residual_ ? in[next++] : nullptr;
      const void* gamma = prologue_ == 2 ? in[next++] : nullptr;
      const float* st = static_cast<const float*>(in[2]);
      const float* sc = static_cast<const float*>(in[3]);
      if (M <= tb_gemv_max_rows(3, K) && is_half(type_))
        return tb_gemv_fused(3, out[0], nullptr, in[0], in[1], nullptr, sc, st, per_channel_, per_token_, res, M, N, K,
                             swiglu_, prologue_, gamma, eps_, stream);
      if (swiglu_ && !prologue_ && !res && is_half(type_))   // prefill shapes: SwiGLU in the tcgen05 epilogue (gemm_tc2.cu)
        return tb_gemm_tc_swiglu(3, out[0], in[0], in[1], sc, st, per_channel_, per_token_, M, N, K, stream);
      TBP_REQUIRE(!swiglu_ && !prologue_, "fused_prologue is only available on the decode (M <= tb_gemv_max_rows) path");
      const int ot = is_half(type_) ? 0 : (type_ == (int32_t) DataType::kFLOAT ? 1 : 2);
      (void) od;
      return tb_gemm_tc(3, out[0], ot, in[0], in[1], nullptr, sc, st, per_channel_, per_token_, res, M, N, K, workspace,
                        tb_gemm_tc_workspace_bytes(M, N, K), counters_.get(tb_gemm_tc_counter_bytes()), 0, 0, stream);
    });
  }

 private:
  void validate() const {
    TBP_REQUIRE(is_half(type_) || type_ == (int32_t) DataType::kFLOAT || type_ == (int32_t) DataType::kINT32,
                "type_id must be half, float or int32");
    TBP_REQUIRE(!(swiglu_ || residual_) || is_half(type_), "fused epilogues need a half output");
    // [ext] fused_prologue: 2 = RmsnormQuantization (extra input gamma, field eps), 3 = QuantizePerToken; input 0 is
    // then the fp16 activation and scale_tokens (input 2) is ignored
    TBP_REQUIRE(prologue_ == 0 || prologue_ == 2 || prologue_ == 3, "fused_prologue must be 0, 2 or 3");
  }
  bool per_channel_ = false, per_token_ = false, swiglu_ = false, residual_ = false;
  int32_t type_ = (int32_t) DataType::kHALF, prologue_ = 0;
  float eps_ = 1e-6f;
  DeviceCounters counters_;
};

// =====================================================================================================
// WeightOnlyQuantMatmul v1 — P/weightOnlyQuantMatmulPlugin/weightOnlyQuantMatmulPlugin.{h,cpp}
// inputs 0 act fp16 [..,K]  1 weight (declared fp32 [K, N/4] int8 | [K, N/8] int4; bytes hold this library's
//        processed layout [N,K] int8 / [N,K/2] packed int4, see tb_preprocess_weights)  2 scales fp16 [N]
//        (3 residual fp16 iff fused_residual [ext])
// =====================================================================================================
class WeightOnlyQuantMatmulPlugin : public BasePlugin {
 public:
  static const char* type_name() { return "WeightOnlyQuantMatmul"; }
  static const std::vector<PluginField>& field_table() {
    static const std::vector<PluginField> t = {field_decl("type_id", FT::kINT32), field_decl("weight_type_id", FT::kINT32),
                                               field_decl("fused_swiglu", FT::kINT32),
                                               field_decl("fused_residual", FT::kINT32),
                                               field_decl("fused_prologue", FT::kINT32), field_decl("eps", FT::kFLOAT32)};
    return t;
  }
  explicit WeightOnlyQuantMatmulPlugin(Fields& f) {
    type_ = f.required<int32_t>("type_id");
    weight_type_ = f.required<int32_t>("weight_type_id");
    swiglu_ = f.optional<int32_t>("fused_swiglu", 0) != 0;
    residual_ = f.optional<int32_t>("fused_residual", 0) != 0;
    prologue_ = f.optional<int32_t>("fused_prologue", 0);
    eps_ = f.optional<float>("eps", 1e-6f);
    validate();
  }
  explicit WeightOnlyQuantMatmulPlugin(Reader& r) {
    type_ = r.get<int32_t>(); weight_type_ = r.get<int32_t>();   // weightOnlyQuantMatmulPlugin.cpp:256-267
    swiglu_ = r.get<bool>(); residual_ = r.get<bool>(); prologue_ = r.get<int32_t>(); eps_ = r.get<float>();
    validate();
  }
  size_t getSerializationSize() const noexcept override { return 3 * sizeof(int32_t) + 2 * sizeof(bool) + sizeof(float); }
  void serialize(void* buf) const noexcept override {
    Writer w{static_cast<char*>(buf)};
    w.put(type_); w.put(weight_type_); w.put(swiglu_); w.put(residual_); w.put(prologue_); w.put(eps_);
  }
  WeightOnlyQuantMatmulPlugin* clone() const noexcept override {
    auto* p = new WeightOnlyQuantMatmulPlugin(*this);
    p->counters_ = DeviceCounters();
    return p;
  }
  const char* getPluginType() const noexcept override { return type_name(); }
  int32_t getNbOutputs() const noexcept override { return 1; }
  int pack() const { return weight_type_ == 1 ? 4 : 8; }   // output channels per declared fp32 element
  DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& eb) noexcept override {
    const IDimensionExpr* n = eb.operation(DimensionOperation::kPROD, *in[1].d[1], *eb.constant(pack()));
    if (swiglu_) n = eb.operation(DimensionOperation::kFLOOR_DIV, *n, *eb.constant(2));
    return with_last_dim(in[0], n);
  }
  DataType getOutputDataType(int32_t, const DataType*, int32_t) const noexcept override { return (DataType) type_; }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t, int32_t) noexcept override {
    if (pos == 1) return linear(io[pos], DataType::kFLOAT);
    return linear(io[pos], (DataType) type_);
  }
  size_t getWorkspaceSize(const PluginTensorDesc* in, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
    return align128(tb_gemm_tc_workspace_bytes((int) rows_of(in[0].dims), in[1].dims.d[1] * pack(), last_dim(in[0].dims)));
  }
  int32_t enqueue(const PluginTensorDesc* id, const PluginTensorDesc*, const void* const* in, void* const* out,
                  void* workspace, cudaStream_t stream) noexcept override {
    return guarded("WeightOnlyQuantMatmul::enqueue", [&]() -> int {
      const int M = (int) rows_of(id[0].dims), N = id[1].dims.d[1] * pack(), K = last_dim(id[0].dims);
      TBP_REQUIRE(id[1].dims.d[0] == K, "weight rows must equal the activation's last dim");
      const int kind = weight_type_ == 1 ? 1 : 2;
      int next = 3;
      const void* res = residual_ ? in[next++] : nullptr;
      const void* gamma = prologue_ ? in[next++] : nullptr;
      if (M <= tb_gemv_max_rows(kind, K))
        return tb_gemv_fused(kind, out[0], nullptr, in[0], in[1], in[2], nullptr, nullptr, 0, 0, res, M, N, K, swiglu_,
                             prologue_, gamma, eps_, stream);
      TBP_REQUIRE(!swiglu_ && !prologue_, "fused_swiglu / fused_prologue are only available on the decode (M <= tb_gemv_max_rows) path");
      return tb_gemm_tc(kind, out[0], 0, in[0], in[1], in[2], nullptr, nullptr, 0, 0, res, M, N, K, workspace,
                        tb_gemm_tc_workspace_bytes(M, N, K), counters_.get(tb_gemm_tc_counter_bytes()), 0, 0, stream);
    });
  }

 private:
  void validate() const {
    TBP_REQUIRE(is_half(type_), "only type_id = half is supported (as in the reference, weightOnlyQuantMatmulPlugin.cpp:47-63)");
    TBP_REQUIRE(weight_type_ == 1 || weight_type_ == 2, "weight_type_id must be 1 (int8) or 2 (int4)");
    // [ext] fused_prologue 1: RMSNorm(x, gamma, eps) applied while staging the activations (extra last input gamma)
    TBP_REQUIRE(prologue_ == 0 || prologue_ == 1, "fused_prologue must be 0 or 1");
  }
  int32_t type_ = (int32_t) DataType::kHALF, weight_type_ = 1, prologue_ = 0;
  float eps_ = 1e-6f;
  bool swiglu_ = false, residual_ = false;
  DeviceCounters counters_;
};

// =====================================================================================================
// Gemm v1 — P/gemmPlugin/gemmPlugin.{h,cpp}: fp16 C = A . B^T, the only form the LLaMA graph uses
// (T/tensorrt_llm/layers/linear.py:13-35: transa = 0, transb = 1).  inputs 0 A [..,K]  1 B [N,K]
// (2 residual iff fused_residual [ext]).  out_fp32 [ext]: fp32 logits for lm_head.
// =====================================================================================================
class GemmPlugin : public BasePlugin {
 public:
  static const char* type_name() { return "Gemm"; }
  static const std::vector<PluginField>& field_table() {
    static const std::vector<PluginField> t = {field_decl("transa", FT::kINT32), field_decl("transb", FT::kINT32),
                                               field_decl("type_id", FT::kINT32), field_decl("fused_swiglu", FT::kINT32),
                                               field_decl("fused_residual", FT::kINT32), field_decl("out_fp32", FT::kINT32),
                                               field_decl("fused_prologue", FT::kINT32), field_decl("eps", FT::kFLOAT32)};
    return t;
  }
  explicit GemmPlugin(Fields& f) {
    transa_ = f.optional<int32_t>("transa", 0);
    transb_ = f.optional<int32_t>("transb", 1);
    type_ = f.required<int32_t>("type_id");
    swiglu_ = f.optional<int32_t>("fused_swiglu", 0) != 0;
    residual_ = f.optional<int32_t>("fused_residual", 0) != 0;
    out_fp32_ = f.optional<int32_t>("out_fp32", 0) != 0;
    prologue_ = f.optional<int32_t>("fused_prologue", 0);
    eps_ = f.optional<float>("eps", 1e-6f);
    validate();
  }
  explicit GemmPlugin(Reader& r) {
    transa_ = r.get<int32_t>(); transb_ = r.get<int32_t>(); type_ = r.get<int32_t>();
    swiglu_ = r.get<bool>(); residual_ = r.get<bool>(); out_fp32_ = r.get<bool>(); prologue_ = r.get<int32_t>();
    eps_ = r.get<float>();
    validate();
  }
  size_t getSerializationSize() const noexcept override { return 4 * sizeof(int32_t) + 3 * sizeof(bool) + sizeof(float); }
  void serialize(void* buf) const noexcept override {
    Writer w{static_cast<char*>(buf)};
    w.put(transa_); w.put(transb_); w.put(type_); w.put(swiglu_); w.put(residual_); w.put(out_fp32_); w.put(prologue_);
    w.put(eps_);
  }
  GemmPlugin* clone() const noexcept override {
    auto* p = new GemmPlugin(*this);
    p->counters_ = DeviceCounters();
    return p;
  }
  const char* getPluginType() const noexcept override { return type_name(); }
  int32_t getNbOutputs() const noexcept override { return 1; }
  DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& eb) noexcept override {
    const IDimensionExpr* n = in[1].d[0];
    if (swiglu_) n = eb.operation(DimensionOperation::kFLOOR_DIV, *n, *eb.constant(2));
    return with_last_dim(in[0], n);
  }
  DataType getOutputDataType(int32_t, const DataType*, int32_t) const noexcept override {
    return out_fp32_ ? DataType::kFLOAT : (DataType) type_;
  }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nb_in, int32_t) noexcept override {
    if (pos == nb_in && out_fp32_) return linear(io[pos], DataType::kFLOAT);
    return linear(io[pos], (DataType) type_);
  }
  size_t getWorkspaceSize(const PluginTensorDesc* in, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
    return align128(tb_gemm_tc_workspace_bytes((int) rows_of(in[0].dims), in[1].dims.d[0], last_dim(in[0].dims)));
  }
  int32_t enqueue(const PluginTensorDesc* id, const PluginTensorDesc*, const void* const* in, void* const* out,
                  void* workspace, cudaStream_t stream) noexcept override {
    return guarded("Gemm::enqueue", [&]() -> int {
      const int M = (int) rows_of(id[0].dims), N = id[1].dims.d[0], K = last_dim(id[0].dims);
      TBP_REQUIRE(id[1].dims.d[1] == K, "B must be [N, K]");
      int next = 2;
      const void* res = residual_ ? in[next++] : nullptr;
      const void* gamma = prologue_ ? in[next++] : nullptr;
      if (M <= tb_gemv_max_rows(0, K))
        return tb_gemv_fused(0, out_fp32_ ? nullptr : out[0], out_fp32_ ? static_cast<float*>(out[0]) : nullptr, in[0],
                             in[1], nullptr, nullptr, nullptr, 0, 0, res, M, N, K, swiglu_, prologue_, gamma, eps_, stream);
      if (swiglu_ && !prologue_ && !res && !out_fp32_)       // prefill shapes: SwiGLU in the tcgen05 epilogue (gemm_tc2.cu)
        return tb_gemm_tc_swiglu(0, out[0], in[0], in[1], nullptr, nullptr, 0, 0, M, N, K, stream);
      TBP_REQUIRE(!swiglu_ && !prologue_, "fused_prologue is only available on the decode (M <= tb_gemv_max_rows) path");
      return tb_gemm_tc(0, out[0], out_fp32_ ? 1 : 0, in[0], in[1], nullptr, nullptr, nullptr, 0, 0, res, M, N, K,
                        workspace, tb_gemm_tc_workspace_bytes(M, N, K), counters_.get(tb_gemm_tc_counter_bytes()), 0, 0,
                        stream);
    });
  }

 private:
  void validate() const {
    TBP_REQUIRE(is_half(type_), "only type_id = half is built");
    TBP_REQUIRE(transa_ == 0 && transb_ == 1, "only C = A . B^T (transa=0, transb=1) is built");
    TBP_REQUIRE(!(out_fp32_ && (residual_ || swiglu_)), "out_fp32 excludes the fused epilogues");
    TBP_REQUIRE(prologue_ == 0 || prologue_ == 1, "fused_prologue must be 0 or 1 (RMSNorm; extra last input gamma)");
  }
  int32_t transa_ = 0, transb_ = 1, type_ = (int32_t) DataType::kHALF, prologue_ = 0;
  float eps_ = 1e-6f;
  bool swiglu_ = false, residual_ = false, out_fp32_ = false;
  DeviceCounters counters_;
};

// =====================================================================================================
// RmsnormQuantization v1 (new) and LayernormQuantization v1 (the reference's,
// P/layernormQuantizationPlugin/layernormQuantizationPlugin.{h,cpp}) share one implementation.
// inputs 0 x [..,H]  1 weight [H]  2 bias [H]  3 scale_to_int fp32 [1]   (4 residual iff fused_residual [ext])
// outputs 0 int8 [..,H]  (1 fp32 [..,1] iff dyn_act_scaling)  (+ fp16 [..,H] = x + residual iff fused_residual)
// Field semantics follow the header (eps, use_diff_of_squares, dyn_act_scaling, type_id); the reference's
// definition crosses the two flags — not reproduced.
// =====================================================================================================
template <bool RMS>
class NormQuantizationPlugin : public BasePlugin {
 public:
  static const char* type_name() { return RMS ? "RmsnormQuantization" : "LayernormQuantization"; }
  static const std::vector<PluginField>& field_table() {
    static const std::vector<PluginField> t = {field_decl("eps", FT::kFLOAT32), field_decl("use_diff_of_squares", FT::kINT32),
                                               field_decl("dyn_act_scaling", FT::kINT32), field_decl("type_id", FT::kINT32),
                                               field_decl("fused_residual", FT::kINT32)};
    return t;
  }
  explicit NormQuantizationPlugin(Fields& f) {
    eps_ = f.optional<float>("eps", RMS ? 1e-6f : 1e-5f);
    diff_of_squares_ = f.optional<int32_t>("use_diff_of_squares", 0) != 0;
    dynamic_ = f.optional<int32_t>("dyn_act_scaling", 0) != 0;
    type_ = f.required<int32_t>("type_id");
    residual_ = f.optional<int32_t>("fused_residual", 0) != 0;
    TBP_REQUIRE(is_half(type_), "only type_id = half is built");
  }
  explicit NormQuantizationPlugin(Reader& r) {
    eps_ = r.get<float>(); diff_of_squares_ = r.get<bool>(); dynamic_ = r.get<bool>(); type_ = r.get<int32_t>();
    residual_ = r.get<bool>();   // layernormQuantizationPlugin.cpp:206-219 order + extension
    TBP_REQUIRE(is_half(type_), "only type_id = half is built");
  }
  size_t getSerializationSize() const noexcept override { return sizeof(float) + 3 * sizeof(bool) + sizeof(int32_t); }
  void serialize(void* buf) const noexcept override {
    Writer w{static_cast<char*>(buf)};
    w.put(eps_); w.put(diff_of_squares_); w.put(dynamic_); w.put(type_); w.put(residual_);
  }
  NormQuantizationPlugin* clone() const noexcept override { return new NormQuantizationPlugin(*this); }
  const char* getPluginType() const noexcept override { return type_name(); }
  int32_t getNbOutputs() const noexcept override { return 1 + (dynamic_ ? 1 : 0) + (residual_ ? 1 : 0); }
  DimsExprs getOutputDimensions(int32_t idx, const DimsExprs* in, int32_t, IExprBuilder& eb) noexcept override {
    if (dynamic_ && idx == 1) return with_last_dim(in[0], eb.constant(1));
    return in[0];
  }
  DataType getOutputDataType(int32_t idx, const DataType* in, int32_t) const noexcept override {
    if (idx == 0) return DataType::kINT8;
    if (dynamic_ && idx == 1) return DataType::kFLOAT;
    return in[0];
  }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nb_in, int32_t) noexcept override {
    if (pos == 3) return linear(io[pos], DataType::kFLOAT);
    if (pos == nb_in) return linear(io[pos], DataType::kINT8);
    if (dynamic_ && pos == nb_in + 1) return linear(io[pos], DataType::kFLOAT);
    return linear(io[pos], (DataType) type_);
  }
  int32_t enqueue(const PluginTensorDesc* id, const PluginTensorDesc*, const void* const* in, void* const* out, void*,
                  cudaStream_t stream) noexcept override {
    return guarded(type_name(), [&]() -> int {
      const int rows = (int) rows_of(id[0].dims), hidden = last_dim(id[0].dims);
      float* dyn = dynamic_ ? static_cast<float*>(out[1]) : nullptr;
      void* sum = residual_ ? out[1 + (dynamic_ ? 1 : 0)] : nullptr;
      return tb_rmsnorm_quant(static_cast<int8_t*>(out[0]), dyn, in[0], residual_ ? in[4] : nullptr, sum, in[1], in[2],
                              static_cast<const float*>(in[3]), eps_, rows, hidden, dynamic_, RMS ? 0 : 1, stream);
    });
  }

 private:
  float eps_ = 1e-6f;
  bool diff_of_squares_ = false, dynamic_ = false, residual_ = false;
  int32_t type_ = (int32_t) DataType::kHALF;
};

// =====================================================================================================
// QuantizePerToken v1 / QuantizeTensor v1 — P/quantizePerTokenPlugin, P/quantizeTensorPlugin (no fields)
// =====================================================================================================
class QuantizePerTokenPlugin : public BasePlugin {
 public:
  static const char* type_name() { return "QuantizePerToken"; }
  static const std::vector<PluginField>& field_table() { static const std::vector<PluginField> t; return t; }
  explicit QuantizePerTokenPlugin(Fields&) {}
  explicit QuantizePerTokenPlugin(Reader&) {}
  size_t getSerializationSize() const noexcept override { return 0; }
  void serialize(void*) const noexcept override {}
  QuantizePerTokenPlugin* clone() const noexcept override { return new QuantizePerTokenPlugin(*this); }
  const char* getPluginType() const noexcept override { return type_name(); }
  int32_t getNbOutputs() const noexcept override { return 2; }
  DimsExprs getOutputDimensions(int32_t idx, const DimsExprs* in, int32_t, IExprBuilder& eb) noexcept override {
    return idx == 0 ? in[0] : with_last_dim(in[0], eb.constant(1));
  }
  DataType getOutputDataType(int32_t idx, const DataType*, int32_t) const noexcept override {
    return idx == 0 ? DataType::kINT8 : DataType::kFLOAT;
  }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t, int32_t) noexcept override {
    if (pos == 0) return io[0].format == TensorFormat::kLINEAR && (io[0].type == DataType::kHALF || io[0].type == DataType::kFLOAT);
    return linear(io[pos], pos == 1 ? DataType::kINT8 : DataType::kFLOAT);
  }
  int32_t enqueue(const PluginTensorDesc* id, const PluginTensorDesc*, const void* const* in, void* const* out, void*,
                  cudaStream_t stream) noexcept override {
    return guarded(type_name(), [&]() -> int {
      return tb_quantize_per_token(static_cast<int8_t*>(out[0]), static_cast<float*>(out[1]), in[0],
                                   (int) rows_of(id[0].dims), last_dim(id[0].dims), id[0].type == DataType::kFLOAT, stream);
    });
  }
};

class QuantizeTensorPlugin : public BasePlugin {
 public:
  static const char* type_name() { return "QuantizeTensor"; }
  static const std::vector<PluginField>& field_table() { static const std::vector<PluginField> t; return t; }
  explicit QuantizeTensorPlugin(Fields&) {}
  explicit QuantizeTensorPlugin(Reader&) {}
  size_t getSerializationSize() const noexcept override { return 0; }
  void serialize(void*) const noexcept override {}
  QuantizeTensorPlugin* clone() const noexcept override { return new QuantizeTensorPlugin(*this); }
  const char* getPluginType() const noexcept override { return type_name(); }
  int32_t getNbOutputs() const noexcept override { return 1; }
  DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder&) noexcept override { return in[0]; }
  DataType getOutputDataType(int32_t, const DataType*, int32_t) const noexcept override { return DataType::kINT8; }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t, int32_t) noexcept override {
    if (pos == 0) return io[0].format == TensorFormat::kLINEAR && (io[0].type == DataType::kHALF || io[0].type == DataType::kFLOAT);
    return linear(io[pos], pos == 1 ? DataType::kFLOAT : DataType::kINT8);
  }
  int32_t enqueue(const PluginTensorDesc* id, const PluginTensorDesc*, const void* const* in, void* const* out, void*,
                  cudaStream_t stream) noexcept override {
    return guarded(type_name(), [&]() -> int {
      return tb_quantize_tensor(static_cast<int8_t*>(out[0]), in[0], volume(id[0].dims), static_cast<const float*>(in[1]),
                                id[0].type == DataType::kFLOAT, stream);
    });
  }
};

// =====================================================================================================
// AllReduce v1 / AllGather v1 — P/ncclPlugin/allreducePlugin.{h,cpp}, allgatherPlugin.{h,cpp}
// fields group:i32[n], type_id:i32.  The communicator for `group` must have been registered with
// tb_comm_init (the reference bootstraps over MPI, allreducePlugin.cpp:128-167; there is no MPI here).
// Both are no-ops while IS_BUILDING=1 (P/common/plugin.h:145-157).
// AllReduce [ext] fused_residual: inputs (x, residual) -> out = allreduce(x) + residual.
// =====================================================================================================
template <bool GATHER>
class NcclPlugin : public BasePlugin {
 public:
  static const char* type_name() { return GATHER ? "AllGather" : "AllReduce"; }
  static const std::vector<PluginField>& field_table() {
    static const std::vector<PluginField> t = {PluginField("group", nullptr, FT::kINT32, 1), field_decl("type_id", FT::kINT32)};
    return t;
  }
  explicit NcclPlugin(Fields& f) {
    group_ = f.int_list("group");
    type_ = f.required<int32_t>("type_id");
    TBP_REQUIRE(!group_.empty(), "group must list at least one rank");
    TBP_REQUIRE(is_half(type_), "only type_id = half is built");
  }
  explicit NcclPlugin(Reader& r) {
    // allreducePlugin.cpp:170-183 writes type, then the group "until the end of the blob"; a blob that was padded by its
    // container would then yield phantom ranks, so trailing ranks are only accepted while they are distinct and in
    // [0, 4096) (a zero-padded tail repeats rank 0 and stops the scan)
    type_ = r.get<int32_t>();
    while (r.p + sizeof(int32_t) <= r.end) {
      const int32_t g = r.get<int32_t>();
      bool dup = g < 0 || g >= 4096;
      for (int32_t h : group_) dup = dup || h == g;
      if (dup) break;
      group_.push_back(g);
    }
    TBP_REQUIRE(!group_.empty() && is_half(type_), "bad serialised NCCL plugin");
  }
  size_t getSerializationSize() const noexcept override { return sizeof(int32_t) * (1 + group_.size()); }
  void serialize(void* buf) const noexcept override {
    Writer w{static_cast<char*>(buf)};
    w.put(type_);
    for (int32_t g : group_) w.put(g);
  }
  NcclPlugin* clone() const noexcept override { return new NcclPlugin(*this); }
  const char* getPluginType() const noexcept override { return type_name(); }
  int32_t getNbOutputs() const noexcept override { return 1; }
  DimsExprs getOutputDimensions(int32_t, const DimsExprs* in, int32_t, IExprBuilder& eb) noexcept override {
    DimsExprs r = in[0];
    if (GATHER) r.d[0] = eb.operation(DimensionOperation::kPROD, *in[0].d[0], *eb.constant((int32_t) group_.size()));
    return r;
  }
  DataType getOutputDataType(int32_t, const DataType* in, int32_t) const noexcept override { return in[0]; }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t, int32_t) noexcept override {
    return linear(io[pos], (DataType) type_);
  }
  int32_t initialize() noexcept override {
    const char* building = std::getenv("IS_BUILDING");
    if (building && building[0] == '1') return 0;
    comm_ = find_comm(group_);
    if (!comm_) {
      log_msg(ILogger::Severity::kERROR, "%s: no communicator registered for this group (call tb_comm_init first)", type_name());
      return -1;
    }
    return 0;
  }
  int32_t enqueue(const PluginTensorDesc* id, const PluginTensorDesc*, const void* const* in, void* const* out, void*,
                  cudaStream_t stream) noexcept override {
    return guarded(type_name(), [&]() -> int {
      const char* building = std::getenv("IS_BUILDING");
      if (building && building[0] == '1') return 0;
      if (!comm_) comm_ = find_comm(group_);
      TBP_REQUIRE(comm_ != nullptr, "communicator not initialised");
      const size_t n = (size_t) volume(id[0].dims);
      return GATHER ? comm_allgather_half(comm_, in[0], out[0], n, stream) : comm_allreduce_half(comm_, in[0], out[0], n, stream);
    });
  }

 private:
  std::vector<int32_t> group_;
  int32_t type_ = (int32_t) DataType::kHALF;
  CommHandle* comm_ = nullptr;
};

// ---- creator instances (registered by initLibNvInferPlugins, registry.cpp) ------------------------------------
std::vector<IPluginCreator*>& all_creators() {
  static Creator<GPTAttentionPlugin> c0;
  static Creator<SmoothQuantGemmPlugin> c1;
  static Creator<WeightOnlyQuantMatmulPlugin> c2;
  static Creator<NormQuantizationPlugin<true>> c3;
  static Creator<NormQuantizationPlugin<false>> c4;
  static Creator<QuantizePerTokenPlugin> c5;
  static Creator<QuantizeTensorPlugin> c6;
  static Creator<NcclPlugin<false>> c7;
  static Creator<NcclPlugin<true>> c8;
  static Creator<GemmPlugin> c9;
  static std::vector<IPluginCreator*> v = {&c0, &c1, &c2, &c3, &c4, &c5, &c6, &c7, &c8, &c9};
  return v;
}

}  // namespace plugins
}  // namespace tb
