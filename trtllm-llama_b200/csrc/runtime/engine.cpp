// The C++ runtime that stands where TensorRT stands in the reference: it owns no math.  It looks plugin
// creators up in the registry by (name, "1", "tensorrt_llm"), creates the operators from
// PluginFieldCollections exactly as T/tensorrt_llm/functional.py:2828-2928 and
// T/tensorrt_llm/quantization/functional.py:12-212 do, and per step calls
// IPluginV2DynamicExt::enqueue(inputDesc, outputDesc, inputs, outputs, workspace, stream) in the order
// of LQ/llama_model.py:78-119,159-287.  Glue ops that TensorRT generates natively in the reference
//  are either fused into a plugin epilogue ([ext] fields) or one small kernel each.
//
// Device-resident step state (token ids, sequence lengths, output position) lets one captured CUDA
// graph serve every decode step; the reference's loop rebuilds host shape buffers and calls .item()
// every token (T/tensorrt_llm/runtime/generation.py:852-963).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/trtllm_b200_runtime.h"
#include "../plugins/pluginBase.h"

using namespace nvinfer1;
using tb::plugins::kNamespace;
using tb::plugins::kVersion;

extern "C" bool initLibNvInferPlugins(void* logger, const char* libNamespace);

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }

#define RT_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)
#define RT_CALL(expr)                                                         \
  do {                                                                        \
    const int rc_ = (expr);                                                   \
    if (rc_ != 0) return fail(std::string(#expr) + " failed with code " + std::to_string(rc_)); \
  } while (0)

struct Tensor { const void* ptr = nullptr; size_t bytes = 0; };

PluginTensorDesc desc(std::initializer_list<int> dims, DataType t) {
  PluginTensorDesc d{};
  d.dims.nbDims = (int32_t) dims.size();
  int i = 0;
  for (int v : dims) d.dims.d[i++] = v;
  d.type = t;
  d.format = TensorFormat::kLINEAR;
  d.scale = 1.f;
  return d;
}

struct PluginDeleter { void operator()(IPluginV2DynamicExt* p) const { if (p) p->destroy(); } };
using PluginPtr = std::unique_ptr<IPluginV2DynamicExt, PluginDeleter>;

struct FieldList {
  std::vector<PluginField> f;
  std::vector<std::unique_ptr<char[]>> store;
  template <class T> void add(const char* name, PluginFieldType t, T v) {
    store.emplace_back(new char[sizeof(T)]);
    std::memcpy(store.back().get(), &v, sizeof(T));
    f.emplace_back(name, store.back().get(), t, 1);
  }
  void add_list(const char* name, const std::vector<int32_t>& v) {
    store.emplace_back(new char[sizeof(int32_t) * v.size()]);
    std::memcpy(store.back().get(), v.data(), sizeof(int32_t) * v.size());
    f.emplace_back(name, store.back().get(), PluginFieldType::kINT32, (int32_t) v.size());
  }
};

PluginPtr make_plugin(const char* name, FieldList& fl) {
  IPluginCreator* c = getPluginRegistry()->getPluginCreator(name, kVersion, kNamespace);
  if (!c) { g_err = std::string("plugin creator not registered: ") + name; return nullptr; }
  PluginFieldCollection fc{(int32_t) fl.f.size(), fl.f.data()};
  auto* p = static_cast<IPluginV2DynamicExt*>(c->createPlugin(name, &fc));
  if (!p) { g_err = std::string("createPlugin failed: ") + name; return nullptr; }
  if (p->initialize() != 0) { p->destroy(); g_err = std::string("plugin initialize failed: ") + name; return nullptr; }
  return PluginPtr(p);
}

struct LinearW { const void* w = nullptr; const void* scale = nullptr; int N = 0, K = 0; };
struct LayerW {
  const void *ln_in = nullptr, *ln_post = nullptr;
  LinearW qkv, dense, fc_gate, proj;
  const float *kv_oq = nullptr, *kv_qo = nullptr;
};

}  // namespace

struct tbrt_engine {
  tbrt_config c{};
  std::map<std::string, Tensor> tensors;
  std::vector<LayerW> L;
  const void *emb = nullptr, *ln_f = nullptr, *lm_head = nullptr;
  int Hl = 0, hid_l = 0, inter_l = 0, vocab_l = 0, S_max = 0;
  bool finalized = false;

  // plugins (one instance per distinct configuration, shared by all layers)
  PluginPtr lin, lin_res, lin_swiglu, lm, attn, attn_packed, normq, qpt, allreduce, allgather;
  int packed_B = 0;            // > 0 while a packed (remove_input_padding) context phase runs: its sequence count
  // paged KV cache: per-layer pools live in kv[]; one device table [layers][max_batch][2][max_blocks] of block addresses
  int tpb = 0, max_blocks = 0, pool_blocks = 0;
  long long* d_block_tables = nullptr;
  std::vector<long long> h_block_tables;
  // decode-shape (M <= 4) variants with the norm / quantiser fused into the projection's prologue ([ext] fields)
  PluginPtr lin_n, lin_n_swiglu, lin_q_res, lm_n, normq_res;

  // device memory
  std::vector<void*> allocs;
  size_t dev_bytes = 0;
  __half *h = nullptr, *h2 = nullptr, *x = nullptr, *qkv = nullptr, *att = nullptr, *gu = nullptr, *act = nullptr,
         *o = nullptr, *hl = nullptr;
  int8_t* xq = nullptr;
  float *xs = nullptr, *logits = nullptr;
  __half* logits_h = nullptr;
  std::vector<void*> kv;
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
  int *d_ids = nullptr, *d_in_lens = nullptr, *d_seq_lens = nullptr, *d_step_pos = nullptr, *d_next = nullptr,
      *d_out_ids = nullptr, *d_prompt = nullptr, *d_flag = nullptr, *d_max_in = nullptr;
  int* h_flag = nullptr;        // pinned
  int end_id = -1;              // >= 0: greedy stop criterion + end_id padding in tbrt_generate
  int last_steps = 0;
  float* d_dummy_scale = nullptr;

  // session state
  int B = 0, S_in = 0, steps_done = 0;
  int64_t launches = 0;
  std::map<int, cudaGraphExec_t> graphs;
  std::map<int, int64_t> graph_nodes;
  std::map<int, int> eager_steps;
  cudaStream_t cap_stream = nullptr;
  tb_decode_step* ds = nullptr; // whole-step persistent kernel (csrc/decode_step.cu); NULL when the configuration is not taken
  bool last_step_fused = false;
  // -1 (default): the faster path as measured on B200 — the fused step under tensor parallelism (tp2 cfg2 1.91 vs 2.23
  // ms, cfg5 1.64 vs 2.15 ms per step), the per-operator schedule on one GPU (cfg2 2.66 vs 2.78 ms);
  // 1: fused step kernel whenever available; 0: per-operator plugin schedule (CUDA graph)
  int decode_mode = -1;
  // sampling (SamplingConfig, generation.py:119-138): top_k = 1 / top_p = 0 is greedy arg-max
  int top_k = 1;
  float top_p = 0.f, temperature = 1.f;
  unsigned long long seed = 0;
  bool sampling() const { return top_k != 1 && !(top_k == 0 && top_p <= 0.f); }
  // beam search (SamplingConfig.num_beams > 1, generation.py:365-409,823-997): rows = batch entries x beams after
  // tbrt_beam_begin; all decoder state is device-resident so the step stays one replayable graph
  int beam_W = 1, beam_end_id = -1;
  float beam_length_penalty = 0.f;
  float* d_cum = nullptr;
  int *d_fin = nullptr, *d_beam_lens = nullptr, *d_ids_t = nullptr, *d_parent_t = nullptr, *d_indir[2] = {nullptr, nullptr},
      *d_beam_out = nullptr;
  void* d_beam_ws = nullptr;
  int beam_step(bool broadcast, cudaStream_t s);
  tb_ar* ar = nullptr;          // peer-memory all-reduce of the decode path (tensor parallel)
  bool ar_open = false;
  int ar_site = 0;              // call-site parity, reset per step (two calls per layer: even per step)

  template <class T> int alloc(T*& p, size_t bytes) {
    void* q = nullptr;
    RT_CUDA(cudaMalloc(&q, bytes ? bytes : 16));
    allocs.push_back(q);
    dev_bytes += bytes;
    p = static_cast<T*>(q);
    return 0;
  }
  ~tbrt_engine() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.second);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (ar) tb_ar_destroy(ar);
    if (ds) tb_decode_step_destroy(ds);
    for (void* p : allocs) cudaFree(p);
    if (h_flag) cudaFreeHost(h_flag);
  }

  int bind();
  int build_plugins();
  int linear(IPluginV2DynamicExt* p, const LinearW& w, const void* in, const float* in_scales, void* out,
             const void* residual, int M, DataType out_t, cudaStream_t s, const void* gamma = nullptr,
             bool half_in = false);
  int layers_forward(int M, int S, bool context, cudaStream_t s);
  int head(int rows, const __half* src, cudaStream_t s);
  int step_body(cudaStream_t s);
  void build_decode_step();
  bool fused_step() const {
    const bool want = decode_mode < 0 ? c.tp_size > 1 : decode_mode != 0;
    return ds && want && !sampling() && beam_W == 1 && B <= tb_decode_step_max_batch();
  }
};

// ---------------------------------------------------------------------------------------------------------
int tbrt_engine::bind() {
  auto get = [&](const std::string& n, size_t want, const void*& out) -> int {
    auto it = tensors.find(n);
    if (it == tensors.end()) return fail("tensor not bound: " + n);
    if (want && it->second.bytes != want)
      return fail("tensor " + n + " has " + std::to_string(it->second.bytes) + " bytes, expected " + std::to_string(want));
    out = it->second.ptr;
    return 0;
  };
  const int hid = c.hidden, tp = c.tp_size;
  Hl = c.heads / tp; hid_l = Hl * c.head_size; inter_l = c.inter / tp; vocab_l = c.vocab / tp;
  S_max = c.max_input_len + c.max_output_len;
  if (c.heads % tp || c.inter % tp || c.vocab % tp) return fail("heads, inter and vocab must divide by tp_size");
  if (c.hidden != c.heads * c.head_size) return fail("hidden must equal heads * head_size");
  auto wbytes = [&](int N, int K) -> size_t {
    switch (c.mode) {
      case TBRT_MODE_FP16: return (size_t) N * K * 2;
      case TBRT_MODE_W4: return (size_t) N * K / 2;
      default: return (size_t) N * K;
    }
  };
  auto bind_linear = [&](const std::string& base, int N, int K, LinearW& w) -> int {
    w.N = N; w.K = K;
    if (get(base + ".weight", wbytes(N, K), w.w)) return -1;
    if (c.mode == TBRT_MODE_W8 || c.mode == TBRT_MODE_W4) return get(base + ".per_channel_scale", (size_t) N * 2, w.scale);
    if (c.mode == TBRT_MODE_SQ) return get(base + ".per_channel_scale", (size_t) N * 4, w.scale);
    return 0;
  };
  if (get("vocab_embedding.weight", (size_t) c.vocab * hid * 2, emb)) return -1;
  if (get("ln_f.weight", (size_t) hid * 2, ln_f)) return -1;
  if (get("lm_head.weight", (size_t) vocab_l * hid * 2, lm_head)) return -1;
  L.resize(c.layers);
  for (int i = 0; i < c.layers; ++i) {
    const std::string p = "layers." + std::to_string(i);
    LayerW& l = L[i];
    if (get(p + ".input_layernorm.weight", (size_t) hid * 2, l.ln_in)) return -1;
    if (get(p + ".post_layernorm.weight", (size_t) hid * 2, l.ln_post)) return -1;
    if (bind_linear(p + ".attention.qkv", 3 * hid_l, hid, l.qkv)) return -1;
    if (bind_linear(p + ".attention.dense", hid, hid_l, l.dense)) return -1;
    if (bind_linear(p + ".mlp.fc_gate", 2 * inter_l, hid, l.fc_gate)) return -1;
    if (bind_linear(p + ".mlp.proj", hid, inter_l, l.proj)) return -1;
    if (c.int8_kv) {
      const void *a = nullptr, *b = nullptr;
      if (get(p + ".attention.kv_orig_quant_scale", 4, a)) return -1;
      if (get(p + ".attention.kv_quant_orig_scale", 4, b)) return -1;
      l.kv_oq = static_cast<const float*>(a);
      l.kv_qo = static_cast<const float*>(b);
    }
  }
  return 0;
}

int tbrt_engine::build_plugins() {
  initLibNvInferPlugins(nullptr, kNamespace);
  const int32_t half_t = (int32_t) DataType::kHALF;
  auto make_linear = [&](bool swiglu, bool residual, int prologue = 0) -> PluginPtr {
    FieldList fl;
    const char* name = nullptr;
    if (c.mode == TBRT_MODE_FP16) {
      name = "Gemm";
      fl.add<int32_t>("transa", PluginFieldType::kINT32, 0);
      fl.add<int32_t>("transb", PluginFieldType::kINT32, 1);
      fl.add<int32_t>("type_id", PluginFieldType::kINT32, half_t);
    } else if (c.mode == TBRT_MODE_SQ) {
      name = "SmoothQuantGemm";
      fl.add<int32_t>("has_per_channel_scaling", PluginFieldType::kINT32, 1);
      fl.add<int32_t>("has_per_token_scaling", PluginFieldType::kINT32, 1);
      fl.add<int32_t>("type_id", PluginFieldType::kINT32, half_t);
    } else {
      name = "WeightOnlyQuantMatmul";
      fl.add<int32_t>("type_id", PluginFieldType::kINT32, half_t);
      fl.add<int32_t>("weight_type_id", PluginFieldType::kINT32, c.mode == TBRT_MODE_W8 ? 1 : 2);
    }
    if (swiglu) fl.add<int32_t>("fused_swiglu", PluginFieldType::kINT32, 1);
    if (residual) fl.add<int32_t>("fused_residual", PluginFieldType::kINT32, 1);
    if (prologue) {
      fl.add<int32_t>("fused_prologue", PluginFieldType::kINT32, prologue);
      fl.add<float>("eps", PluginFieldType::kFLOAT32, c.rms_eps);
    }
    return make_plugin(name, fl);
  };
  const bool sq_mode = c.mode == TBRT_MODE_SQ;
  if (!(lin = make_linear(false, false))) return -1;
  if (!(lin_res = make_linear(false, true))) return -1;
  if (!(lin_swiglu = make_linear(true, false))) return -1;
  if (!(lin_n = make_linear(false, false, sq_mode ? 2 : 1))) return -1;
  if (!(lin_n_swiglu = make_linear(true, false, sq_mode ? 2 : 1))) return -1;
  if (sq_mode && !(lin_q_res = make_linear(false, true, 3))) return -1;
  {
    FieldList fl;
    fl.add<int32_t>("transa", PluginFieldType::kINT32, 0);
    fl.add<int32_t>("transb", PluginFieldType::kINT32, 1);
    fl.add<int32_t>("type_id", PluginFieldType::kINT32, half_t);
    fl.add<int32_t>("out_fp32", PluginFieldType::kINT32, c.tp_size == 1 ? 1 : 0);
    if (!(lm = make_plugin("Gemm", fl))) return -1;
    fl.add<int32_t>("fused_prologue", PluginFieldType::kINT32, 1);     // ln_f rides in the lm_head GEMV (rows <= 4)
    fl.add<float>("eps", PluginFieldType::kFLOAT32, c.rms_eps);
    if (!(lm_n = make_plugin("Gemm", fl))) return -1;
  }
  {
    // the fields T/tensorrt_llm/functional.py:2833-2891 passes for a LLaMA layer
    FieldList fl;
    fl.add<int32_t>("num_heads", PluginFieldType::kINT32, Hl);
    fl.add<int32_t>("head_size", PluginFieldType::kINT32, c.head_size);
    fl.add<int32_t>("unidirectional", PluginFieldType::kINT32, 1);
    fl.add<float>("q_scaling", PluginFieldType::kFLOAT32, 1.f);
    fl.add<int32_t>("rotary_embedding_dim", PluginFieldType::kINT32, c.head_size);
    fl.add<int8_t>("neox_rotary_style", PluginFieldType::kINT8, 1);
    fl.add<int8_t>("context_fmha_type", PluginFieldType::kINT8, 1);
    fl.add<int8_t>("multi_block_mode", PluginFieldType::kINT8, 1);
    fl.add<int8_t>("multi_query_mode", PluginFieldType::kINT8, 0);
    fl.add<int32_t>("int8_kv_cache", PluginFieldType::kINT32, c.int8_kv);
    fl.add<int32_t>("fp8_kv_cache", PluginFieldType::kINT32, 0);
    fl.add<int8_t>("remove_input_padding", PluginFieldType::kINT8, 0);
    fl.add<int32_t>("mask_type", PluginFieldType::kINT32, 1);
    fl.add<int32_t>("paged_kv_cache", PluginFieldType::kINT32, c.paged_kv_tokens_per_block > 0 ? 1 : 0);
    fl.add<int32_t>("type_id", PluginFieldType::kINT32, half_t);
    fl.add<int32_t>("in_flight_batching", PluginFieldType::kINT32, 0);
    fl.add<int32_t>("device_lengths", PluginFieldType::kINT32, 1);
    if (!(attn = make_plugin("GPTAttention", fl))) return -1;
  }
  {
    // the same operator for packed input (build.py --remove_input_padding): input 0 is [1, num_tokens, 3 * hidden]
    FieldList fl;
    fl.add<int32_t>("num_heads", PluginFieldType::kINT32, Hl);
    fl.add<int32_t>("head_size", PluginFieldType::kINT32, c.head_size);
    fl.add<int32_t>("unidirectional", PluginFieldType::kINT32, 1);
    fl.add<float>("q_scaling", PluginFieldType::kFLOAT32, 1.f);
    fl.add<int32_t>("rotary_embedding_dim", PluginFieldType::kINT32, c.head_size);
    fl.add<int8_t>("neox_rotary_style", PluginFieldType::kINT8, 1);
    fl.add<int8_t>("context_fmha_type", PluginFieldType::kINT8, 1);
    fl.add<int8_t>("multi_block_mode", PluginFieldType::kINT8, 1);
    fl.add<int8_t>("multi_query_mode", PluginFieldType::kINT8, 0);
    fl.add<int32_t>("int8_kv_cache", PluginFieldType::kINT32, c.int8_kv);
    fl.add<int32_t>("fp8_kv_cache", PluginFieldType::kINT32, 0);
    fl.add<int8_t>("remove_input_padding", PluginFieldType::kINT8, 1);
    fl.add<int32_t>("mask_type", PluginFieldType::kINT32, 1);
    fl.add<int32_t>("paged_kv_cache", PluginFieldType::kINT32, c.paged_kv_tokens_per_block > 0 ? 1 : 0);
    fl.add<int32_t>("type_id", PluginFieldType::kINT32, half_t);
    fl.add<int32_t>("in_flight_batching", PluginFieldType::kINT32, 0);
    fl.add<int32_t>("device_lengths", PluginFieldType::kINT32, 1);
    if (!(attn_packed = make_plugin("GPTAttention", fl))) return -1;
  }
  if (c.mode == TBRT_MODE_SQ) {
    FieldList fl;
    fl.add<float>("eps", PluginFieldType::kFLOAT32, c.rms_eps);
    fl.add<int32_t>("use_diff_of_squares", PluginFieldType::kINT32, 0);
    fl.add<int32_t>("dyn_act_scaling", PluginFieldType::kINT32, 1);
    fl.add<int32_t>("type_id", PluginFieldType::kINT32, half_t);
    if (!(normq = make_plugin("RmsnormQuantization", fl))) return -1;
    fl.add<int32_t>("fused_residual", PluginFieldType::kINT32, 1);      // (x, w, b, scale, residual) -> (q, scales, x + residual)
    if (!(normq_res = make_plugin("RmsnormQuantization", fl))) return -1;
    FieldList none;
    if (!(qpt = make_plugin("QuantizePerToken", none))) return -1;
  }
  if (c.tp_size > 1) {
    std::vector<int32_t> group;
    for (int r = 0; r < c.tp_size; ++r) group.push_back(r);
    FieldList a, g;
    a.add_list("group", group);
    a.add<int32_t>("type_id", PluginFieldType::kINT32, half_t);
    g.add_list("group", group);
    g.add<int32_t>("type_id", PluginFieldType::kINT32, half_t);
    if (!(allreduce = make_plugin("AllReduce", a))) return -1;
    if (!(allgather = make_plugin("AllGather", g))) return -1;
  }
  return 0;
}

// one projection through its plugin: fp16 / weight-only take (x, w[, scales][, residual]); SmoothQuant
// takes (x int8, w, scale_tokens, scale_channels[, residual]) — same input order as the reference plugins
int tbrt_engine::linear(IPluginV2DynamicExt* p, const LinearW& w, const void* in, const float* in_scales, void* out,
                        const void* residual, int M, DataType out_t, cudaStream_t s, const void* gamma, bool half_in) {
  PluginTensorDesc id[6], od[1];
  const void* inputs[6];
  void* outputs[1] = {out};
  int n = 0;
  const bool is_lm = (p == lm.get() || p == lm_n.get());   // lm_head stays fp16 (LQ/quant.py:58-59)
  if (c.mode == TBRT_MODE_SQ && !is_lm) {
    id[n] = desc({M, w.K}, half_in ? DataType::kHALF : DataType::kINT8); inputs[n++] = in;
    id[n] = desc({w.N, w.K / 4}, DataType::kFLOAT); inputs[n++] = w.w;
    id[n] = desc({M, 1}, DataType::kFLOAT); inputs[n++] = in_scales;
    id[n] = desc({1, w.N}, DataType::kFLOAT); inputs[n++] = w.scale;
  } else if ((c.mode == TBRT_MODE_W8 || c.mode == TBRT_MODE_W4) && !is_lm) {
    const int pack = c.mode == TBRT_MODE_W8 ? 4 : 8;
    id[n] = desc({M, w.K}, DataType::kHALF); inputs[n++] = in;
    id[n] = desc({w.K, w.N / pack}, DataType::kFLOAT); inputs[n++] = w.w;
    id[n] = desc({w.N}, DataType::kHALF); inputs[n++] = w.scale;
  } else {
    id[n] = desc({M, w.K}, DataType::kHALF); inputs[n++] = in;
    id[n] = desc({w.N, w.K}, DataType::kHALF); inputs[n++] = w.w;
  }
  if (residual) { id[n] = desc({M, w.N}, DataType::kHALF); inputs[n++] = residual; }
  if (gamma) { id[n] = desc({w.K}, DataType::kHALF); inputs[n++] = gamma; }
  od[0] = desc({M, w.N}, out_t);
  ++launches;
  return p->enqueue(id, od, inputs, outputs, workspace, s);
}

int tbrt_engine::layers_forward(int M, int S, bool context, cudaStream_t s) {
  const bool packed = context && packed_B > 0;          // M = number of real tokens, S = the longest prompt
  const int hid = c.hidden, Bq = context ? (packed ? packed_B : M / S) : M;
  const bool sq = c.mode == TBRT_MODE_SQ, tp = c.tp_size > 1;
  const int kind = c.mode;   // TBRT_MODE_* == tb_gemv kind
  int gemv_rows = tb_gemv_max_rows(kind, c.hidden);
  if (tb_gemv_max_rows(kind, hid_l) < gemv_rows) gemv_rows = tb_gemv_max_rows(kind, hid_l);
  if (tb_gemv_max_rows(kind, inter_l) < gemv_rows) gemv_rows = tb_gemv_max_rows(kind, inter_l);
  const bool fuse_swiglu = M <= gemv_rows;
  // SwiGLU in the tcgen05 epilogue (tb_gemm_tc_swiglu) is built, bit-identical and reachable through the plugins' fused_swiglu
  // field at any M, but NOT the default here: measured at M = 16384 (tools/swiglu_tc_bench.py) the fused GEMM runs 1.46 ms
  // (2.0 POPS) against 0.98 ms (3.0 POPS) for the plain one — two 128-token accumulators per stage move 40 KB of operands
  // through shared memory per 512 MMA cycles instead of 32 KB — so GEMM + SwiGLU/quantise in two kernels (1.22 ms) beats
  // fused GEMM + quantise (1.59 ms).  TB_SWIGLU_TC = 1 opts in.
  static const int swiglu_tc_env = getenv("TB_SWIGLU_TC") ? atoi(getenv("TB_SWIGLU_TC")) : 0;
  const bool fuse_swiglu_tc = swiglu_tc_env != 0 && M >= 2048 && inter_l % 128 == 0;
  int host_len[2] = {context ? 0 : c.max_input_len, context ? 1 : 0};   // device_lengths [ext]: step position is on the device

  auto norm = [&](const __half* src, const void* gamma, const __half* residual, __half* sum_out) -> int {
    launches += 1;
    if (sq) {
      PluginTensorDesc id[4] = {desc({M, hid}, DataType::kHALF), desc({hid}, DataType::kHALF), desc({hid}, DataType::kHALF),
                                desc({1}, DataType::kFLOAT)};
      PluginTensorDesc od[2] = {desc({M, hid}, DataType::kINT8), desc({M, 1}, DataType::kFLOAT)};
      if (residual) {   // one kernel: h = x + residual (written to sum_out), RMSNorm(h), per-token int8
        PluginTensorDesc id5[5] = {id[0], id[1], id[2], id[3], desc({M, hid}, DataType::kHALF)};
        PluginTensorDesc od3[3] = {od[0], od[1], desc({M, hid}, DataType::kHALF)};
        const void* in5[5] = {src, gamma, nullptr, d_dummy_scale, residual};
        void* out3[3] = {xq, xs, sum_out};
        return normq_res->enqueue(id5, od3, in5, out3, workspace, s);
      }
      const void* in[4] = {src, gamma, nullptr, d_dummy_scale};
      void* out[2] = {xq, xs};
      return normq->enqueue(id, od, in, out, workspace, s);
    }
    return tb_rmsnorm(x, src, residual, sum_out, gamma, c.rms_eps, M, hid, s);
  };
  auto quant = [&](const __half* src, int cols) -> int {
    PluginTensorDesc id[1] = {desc({M, cols}, DataType::kHALF)};
    PluginTensorDesc od[2] = {desc({M, cols}, DataType::kINT8), desc({M, 1}, DataType::kFLOAT)};
    const void* in[1] = {src};
    void* out[2] = {xq, xs};
    launches += 1;
    return qpt->enqueue(id, od, in, out, workspace, s);
  };
  const void* lin_in = sq ? static_cast<const void*>(xq) : static_cast<const void*>(x);
  // decode shapes: RMSNorm (+ per-token quantisation) rides in the projection's prologue -> 5 kernels per layer
  const bool fused = M <= gemv_rows;
  // A/B switch TB_FUSE_NORM_ROWS=4: from 5 rows on run RMSNorm once as its own PDL-chained kernel instead of in the prologue of
  // every CTA of the tensor-core GEMV (48 % of the QKV launch at 8 rows before the CTA pairs shared it).  Measured equal at
  // step level (cfg3 int8-KV 2.934 vs 2.941 ms, int4 B=8 2.042 vs 2.025 ms): the extra launches cost what the prologue did.
  static const int fuse_norm_rows = getenv("TB_FUSE_NORM_ROWS") ? atoi(getenv("TB_FUSE_NORM_ROWS")) : 8;
  const bool fuse_norm = fused && (sq || M <= fuse_norm_rows);
  // decode shapes: every projection asks L2 for the head of the weights the NEXT projection streams (tb_gemv_hint_next);
  // the whole dense matrix when the attention kernel runs in between.  TB_PF_MB / TB_PF_ATTN_MB = 0 switch it off.
  static const size_t pf_mb = getenv("TB_PF_MB") ? (size_t) atoi(getenv("TB_PF_MB")) : 12;
  static const size_t pf_attn_mb = getenv("TB_PF_ATTN_MB") ? (size_t) atoi(getenv("TB_PF_ATTN_MB")) : 12;
  auto wbytes = [&](const LinearW& w) -> size_t {
    return c.mode == TBRT_MODE_FP16 ? (size_t) w.N * w.K * 2 : (c.mode == TBRT_MODE_W4 ? (size_t) w.N * w.K / 2 : (size_t) w.N * w.K);
  };
  auto hint = [&](const LinearW* nx, bool swiglu_next, size_t cap_mb) {
    if (!fused || !nx || !cap_mb) { tb_gemv_hint_next(nullptr, 0, nullptr, 0); return; }
    size_t total = wbytes(*nx), cap = cap_mb << 20;
    if (total > cap) total = cap;
    if (swiglu_next) {   // gate rows [0, inter), up rows [inter, 2 inter): the kernel walks both fronts together
      const size_t half = (total / 2) & ~(size_t) 127;
      tb_gemv_hint_next(nx->w, half, static_cast<const uint8_t*>(nx->w) + wbytes(*nx) / 2, half);
    } else {
      tb_gemv_hint_next(nx->w, total & ~(size_t) 127, nullptr, 0);
    }
  };

  __half* cur = h;   // residual stream
  __half* nxt = h2;
  if (!fused) RT_CALL(norm(cur, L[0].ln_in, nullptr, nullptr));
  for (int li = 0; li < c.layers; ++li) {
    const LayerW& l = L[li];
    hint(&l.dense, false, pf_attn_mb);
    if (fuse_norm) {
      RT_CALL(linear(lin_n.get(), l.qkv, cur, xs, qkv, nullptr, M, DataType::kHALF, s, l.ln_in, true));
    } else {
      if (fused) RT_CALL(norm(cur, l.ln_in, nullptr, nullptr));
      RT_CALL(linear(lin.get(), l.qkv, lin_in, xs, qkv, nullptr, M, DataType::kHALF, s));
    }
    {
      const DataType kvt = c.int8_kv ? DataType::kINT8 : DataType::kHALF;
      PluginTensorDesc id[11] = {packed ? desc({1, M, 3 * hid_l}, DataType::kHALF) : desc({Bq, context ? S : 1, 3 * hid_l}, DataType::kHALF),
                                 tpb ? desc({pool_blocks, 2, Hl, tpb, c.head_size}, kvt) : desc({Bq, 2, Hl, S_max, c.head_size}, kvt),
                                 desc({Bq}, DataType::kINT32), desc({2}, DataType::kINT32),
                                 desc({Bq, S_max}, DataType::kINT32), desc({Bq}, DataType::kINT32),
                                 desc({S_in}, DataType::kINT32), desc({Bq, 1, S_max}, DataType::kINT32),
                                 desc({1}, DataType::kFLOAT), desc({1}, DataType::kFLOAT)};
      PluginTensorDesc od[2] = {packed ? desc({1, M, hid_l}, DataType::kHALF) : desc({Bq, context ? S : 1, hid_l}, DataType::kHALF), id[1]};
      // masked_tokens = NULL: derived from input_lengths / max_input_length on the device ([ext])
      // input 6: max_input_length as one device int (device_lengths [ext]) — a replayed step graph must not bake S_in
      const void* in[11] = {qkv, kv[li], d_seq_lens, host_len, nullptr, d_in_lens, context ? nullptr : d_max_in, nullptr,
                            l.kv_oq, l.kv_qo, nullptr};
      if (tpb) {   // block_pointers [B, 1, 2, 2 * max_blocks] (int32 view of int64) right after the optional KV scales
        const int bp = c.int8_kv ? 10 : 8;
        id[bp] = desc({Bq, 1, 2, 2 * max_blocks}, DataType::kINT32);
        in[bp] = d_block_tables + (size_t) li * c.max_batch * 2 * max_blocks;
      }
      if (beam_W > 1 && !context) {   // cache_indirection [batch, beam, S_max]: which beam's cache row holds position t
        id[7] = desc({Bq / beam_W, beam_W, S_max}, DataType::kINT32);
        in[7] = d_indir[0];
      }
      void* out[2] = {att, kv[li]};
      launches += context ? 2 : 1;
      RT_CALL((packed ? attn_packed : attn)->enqueue(id, od, in, out, workspace, s));
    }
    IPluginV2DynamicExt* row_lin = (fused && sq) ? lin_q_res.get() : lin_res.get();   // QuantizePerToken fused in
    IPluginV2DynamicExt* row_lin_nores = (fused && sq) ? nullptr : lin.get();
    const void* dense_in = att;
    if (sq && !(fused && !tp)) { RT_CALL(quant(att, hid_l)); dense_in = xq; }
    (void) row_lin_nores;
    hint(&l.fc_gate, true, pf_mb);
    if (!tp && fused) {
      RT_CALL(linear(row_lin, l.dense, dense_in, xs, nxt, cur, M, DataType::kHALF, s, nullptr, sq));   // nxt = cur + dense(att)
    } else if (!tp) {
      // prefill shapes: the residual add rides in the norm kernel that follows (one fused add + RMSNorm (+ quantise)
      // pass) instead of the GEMM epilogue, whose per-element residual loads stall the TMEM drain (ncu: 7 % tensor pipe)
      RT_CALL(linear(lin.get(), l.dense, dense_in, xs, o, nullptr, M, DataType::kHALF, s));
      RT_CALL(norm(o, l.ln_post, cur, nxt));                                                     // nxt = o + cur, x = norm(nxt)
    } else if (fused && ar_open) {
      // row-parallel partial straight into the peer-mapped buffer, then one kernel: all-reduce + residual add
      const int set = ar_site++ & 1;
      RT_CALL(linear(lin.get(), l.dense, dense_in, xs, tb_ar_buffer(ar, set), nullptr, M, DataType::kHALF, s));
      launches += 1;
      RT_CALL(tb_ar_allreduce(ar, set, nxt, cur, (int64_t) M * hid, s));
    } else {
      RT_CALL(linear(lin.get(), l.dense, dense_in, xs, o, nullptr, M, DataType::kHALF, s));
      PluginTensorDesc d1[1] = {desc({M, hid}, DataType::kHALF)};
      const void* in[1] = {o};
      void* out[1] = {o};
      launches += 1;
      RT_CALL(allreduce->enqueue(d1, d1, in, out, workspace, s));
      if (fused) {
        launches += 1;
        RT_CALL(tb_add(nxt, o, cur, (int64_t) M * hid, s));
      } else {
        RT_CALL(norm(o, l.ln_post, cur, nxt));                                                  // nxt = o + cur, x = norm(nxt)
      }
    }
    std::swap(cur, nxt);
    bool act_quantised = false;
    hint(&l.proj, false, pf_mb);
    if (fuse_norm) {
      RT_CALL(linear(lin_n_swiglu.get(), l.fc_gate, cur, xs, act, nullptr, M, DataType::kHALF, s, l.ln_post, true));
    } else if (fused) {
      RT_CALL(norm(cur, l.ln_post, nullptr, nullptr));
      RT_CALL(linear(lin_swiglu.get(), l.fc_gate, lin_in, xs, act, nullptr, M, DataType::kHALF, s));
    } else if (fuse_swiglu) {
      RT_CALL(linear(lin_swiglu.get(), l.fc_gate, lin_in, xs, act, nullptr, M, DataType::kHALF, s));
    } else if (fuse_swiglu_tc && (sq || c.mode == TBRT_MODE_FP16)) {
      // prefill shapes: silu(gate) * up computed in the tcgen05 epilogue (two accumulators side by side in TMEM); the
      // SmoothQuant per-token quantisation then reads the fp16 activation once (QuantizePerToken below)
      RT_CALL(linear(lin_swiglu.get(), l.fc_gate, lin_in, xs, act, nullptr, M, DataType::kHALF, s));
    } else if (sq && inter_l <= 16384) {
      // SmoothQuant prefill: SwiGLU and the per-token quantisation of its output in one pass over the GEMM output
      RT_CALL(linear(lin.get(), l.fc_gate, lin_in, xs, gu, nullptr, M, DataType::kHALF, s));
      launches += 1;
      RT_CALL(tb_swiglu_quant(xq, xs, gu, gu + inter_l, M, inter_l, 2 * inter_l, s));
      act_quantised = true;
    } else {
      RT_CALL(linear(lin.get(), l.fc_gate, lin_in, xs, gu, nullptr, M, DataType::kHALF, s));
      launches += 1;
      RT_CALL(tb_swiglu(act, gu, gu + inter_l, M, inter_l, 2 * inter_l, s));
    }
    const void* proj_in = act;
    if (act_quantised) proj_in = xq;
    else if (sq && !(fused && !tp)) { RT_CALL(quant(act, inter_l)); proj_in = xq; }
    const void* next_gamma = li + 1 < c.layers ? L[li + 1].ln_in : nullptr;
    if (li + 1 < c.layers) {
      hint(&L[li + 1].qkv, false, pf_mb);
    } else if (fused && pf_mb) {   // lm_head stays fp16 in every mode
      size_t b = (size_t) vocab_l * c.hidden * 2;
      if (b > (pf_mb << 20)) b = pf_mb << 20;
      tb_gemv_hint_next(lm_head, b & ~(size_t) 127, nullptr, 0);
    }
    if (!tp && fused) {
      RT_CALL(linear(row_lin, l.proj, proj_in, xs, nxt, cur, M, DataType::kHALF, s, nullptr, sq));
      std::swap(cur, nxt);
    } else if (!tp) {
      RT_CALL(linear(lin.get(), l.proj, proj_in, xs, o, nullptr, M, DataType::kHALF, s));
      if (next_gamma) {
        RT_CALL(norm(o, next_gamma, cur, nxt));
      } else {
        launches += 1;
        RT_CALL(tb_add(nxt, o, cur, (int64_t) M * hid, s));
      }
      std::swap(cur, nxt);
    } else if (fused && ar_open) {
      const int set = ar_site++ & 1;
      RT_CALL(linear(lin.get(), l.proj, proj_in, xs, tb_ar_buffer(ar, set), nullptr, M, DataType::kHALF, s));
      launches += 1;
      RT_CALL(tb_ar_allreduce(ar, set, nxt, cur, (int64_t) M * hid, s));
      std::swap(cur, nxt);
    } else {
      RT_CALL(linear(lin.get(), l.proj, proj_in, xs, o, nullptr, M, DataType::kHALF, s));
      PluginTensorDesc d1[1] = {desc({M, hid}, DataType::kHALF)};
      const void* in[1] = {o};
      void* out[1] = {o};
      launches += 1;
      RT_CALL(allreduce->enqueue(d1, d1, in, out, workspace, s));
      if (next_gamma && !fused) {
        RT_CALL(norm(o, next_gamma, cur, nxt));
      } else {
        launches += 1;
        RT_CALL(tb_add(nxt, o, cur, (int64_t) M * hid, s));
      }
      std::swap(cur, nxt);
    }
  }
  if (cur != h) {   // keep the final residual stream in `h` (layer count is even in practice; copy otherwise)
    RT_CUDA(cudaMemcpyAsync(h, cur, (size_t) M * hid * 2, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

// ln_f -> lm_head (fp32 logits) -> greedy argmax -> device-side bookkeeping
int tbrt_engine::head(int rows, const __half* src, cudaStream_t s) {
  LinearW w;
  w.w = lm_head; w.N = vocab_l; w.K = c.hidden;
  launches += 2;
  if (c.tp_size == 1 && rows <= tb_gemv_max_rows(0, c.hidden)) {
    RT_CALL(linear(lm_n.get(), w, src, nullptr, logits, nullptr, rows, DataType::kFLOAT, s, ln_f));
  } else if (c.tp_size == 1) {
    launches += 1;
    RT_CALL(tb_rmsnorm(x, src, nullptr, nullptr, ln_f, c.rms_eps, rows, c.hidden, s));
    RT_CALL(linear(lm.get(), w, x, nullptr, logits, nullptr, rows, DataType::kFLOAT, s));
  } else {
    launches += 1;
    RT_CALL(tb_rmsnorm(x, src, nullptr, nullptr, ln_f, c.rms_eps, rows, c.hidden, s));
    // vocab-parallel lm_head + all-gather (T/tensorrt_llm/layers/linear.py:78-97 gather_output)
    RT_CALL(linear(lm.get(), w, x, nullptr, logits_h, nullptr, rows, DataType::kHALF, s));
    PluginTensorDesc id[1] = {desc({rows, vocab_l}, DataType::kHALF)};
    PluginTensorDesc od[1] = {desc({rows * c.tp_size, vocab_l}, DataType::kHALF)};
    const void* in[1] = {logits_h};
    void* out[1] = {logits_h + (size_t) rows * vocab_l};
    launches += 2;
    RT_CALL(allgather->enqueue(id, od, in, out, workspace, s));
    RT_CALL(tb_gather_logits(logits, logits_h + (size_t) rows * vocab_l, rows, vocab_l, c.tp_size, s));
  }
  if (beam_W > 1) {
    if (beam_step(false, s)) return -1;
  } else if (sampling()) {
    // the generation step (column of output_ids being produced) keys the random stream: read on the device, so the
    // captured step graph draws fresh numbers on every replay
    RT_CALL(tb_sample(d_next, logits, rows, c.vocab, c.vocab, top_k, top_p > 0.f ? top_p : 1.f, temperature, seed, d_step_pos, 0,
                      nullptr, end_id, nullptr, s));
  } else {
    RT_CALL(tb_argmax(d_next, logits, rows, c.vocab, c.vocab, s));
  }
  RT_CALL(tb_advance_step(d_next, d_ids, d_out_ids, d_seq_lens, d_step_pos, rows, c.max_output_len, s));
  return 0;
}

// one beam-search decoding step on the logits of all rows: candidates, selection, bookkeeping, cache indirection
// (tgt written from src, then copied back so the attention plugin always reads d_indir[0])
int tbrt_engine::beam_step(bool broadcast, cudaStream_t s) {
  const int rows = broadcast ? B * beam_W : B;
  launches += 2;
  RT_CALL(tb_beam_search_step(logits, c.vocab, c.vocab, broadcast ? 1 : 0, rows, beam_W, beam_length_penalty, beam_end_id,
                              d_step_pos, d_max_in, d_cum, d_fin, d_beam_lens, d_ids_t, d_parent_t, d_next, d_indir[0],
                              d_indir[1], S_max, d_beam_ws, s));
  RT_CUDA(cudaMemcpyAsync(d_indir[0], d_indir[1], (size_t) rows * S_max * 4, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int tbrt_engine::step_body(cudaStream_t s) {
  ar_site = 0;
  launches += 1;
  RT_CALL(tb_embedding(h, emb, d_ids, B, c.hidden, c.vocab, s));
  if (layers_forward(B, 1, false, s)) return -1;
  return head(B, h, s);
}

// The whole decode step as one persistent kernel over the same weights, KV caches, activation arena and device-resident
// step state the plugin schedule uses (either path can run any step).  Not an error when the configuration is not taken.
void tbrt_engine::build_decode_step() {
  static const bool off = getenv("TB_DECODE_STEP") && atoi(getenv("TB_DECODE_STEP")) == 0;
  if (off || ds || tpb) return;     // the fused step reads a contiguous cache
  if (c.tp_size > 1 && !(ar && ar_open && tb_ar_extra_bytes(ar) > 0)) return;   // needs the peers' scratch mapped
  tb_decode_step_config dc{};
  dc.kind = c.mode; dc.layers = c.layers; dc.hidden = c.hidden; dc.heads_local = Hl; dc.inter_local = inter_l;
  dc.vocab_local = vocab_l; dc.vocab = c.vocab; dc.max_batch = c.max_batch; dc.max_seq_len = S_max; dc.int8_kv = c.int8_kv;
  dc.out_stride = c.max_output_len; dc.rms_eps = c.rms_eps; dc.tp_size = c.tp_size; dc.tp_rank = c.tp_rank;
  std::vector<tb_decode_step_layer> dl(c.layers);
  for (int i = 0; i < c.layers; ++i) {
    const LayerW& l = L[i];
    dl[i] = tb_decode_step_layer{l.qkv.w, l.dense.w, l.fc_gate.w, l.proj.w, l.qkv.scale, l.dense.scale, l.fc_gate.scale,
                                 l.proj.scale, l.ln_in, l.ln_post, kv[i], l.kv_oq, l.kv_qo};
  }
  tb_decode_step_buffers db{};
  db.emb = emb; db.ln_f = ln_f; db.lm_head = lm_head; db.h_a = h; db.h_b = h2; db.qkv = qkv; db.att = att; db.act = act;
  db.logits = logits; db.ids = d_ids; db.seq_lens = d_seq_lens; db.step_pos = d_step_pos; db.out_ids = d_out_ids;
  db.next_ids = d_next; db.in_lens = d_in_lens; db.max_in = d_max_in;
  for (int r = 0; r < 8; ++r) db.tp_peers[r] = (c.tp_size > 1 && r < c.tp_size) ? tb_ar_extra(ar, r) : nullptr;
  if (tb_decode_step_create(&ds, &dc, dl.data(), &db) != 0) ds = nullptr;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" {

const char* tbrt_last_error(void) { return g_err.c_str(); }

tbrt_engine* tbrt_create(const tbrt_config* cfg) {
  if (!cfg || cfg->head_size != 128 || cfg->tp_size < 1 || cfg->mode < 0 || cfg->mode > 3) {
    g_err = "bad tbrt_config";
    return nullptr;
  }
  if (tb_check_device() != 0) {
    g_err = "no sm_100 device is current: this runtime has no CPU or other-architecture fallback";
    return nullptr;
  }
  auto* e = new tbrt_engine();
  e->c = *cfg;
  return e;
}
void tbrt_destroy(tbrt_engine* e) { delete e; }

int tbrt_set_tensor(tbrt_engine* e, const char* name, const void* ptr, size_t bytes) {
  if (!e || !name || !ptr) return fail("tbrt_set_tensor: null argument");
  if (e->finalized) return fail("engine already finalized");
  e->tensors[name] = Tensor{ptr, bytes};
  return 0;
}

int tbrt_finalize(tbrt_engine* e) {
  if (e->finalized) return 0;
  if (e->bind()) return -1;
  if (e->build_plugins()) return -1;
  const tbrt_config& c = e->c;
  const size_t Mmax = (size_t) c.max_batch * c.max_input_len, hid = c.hidden;
  if (e->alloc(e->h, Mmax * hid * 2) || e->alloc(e->h2, Mmax * hid * 2) || e->alloc(e->x, Mmax * hid * 2) ||
      e->alloc(e->qkv, Mmax * 3 * e->hid_l * 2) || e->alloc(e->att, Mmax * e->hid_l * 2) ||
      e->alloc(e->gu, Mmax * 2 * e->inter_l * 2) || e->alloc(e->act, Mmax * e->inter_l * 2) ||
      e->alloc(e->o, Mmax * hid * 2) || e->alloc(e->hl, (size_t) c.max_batch * hid * 2))
    return -1;
  if (c.mode == TBRT_MODE_SQ) {
    const size_t widest = (size_t) (e->inter_l > c.hidden ? e->inter_l : c.hidden);
    if (e->alloc(e->xq, Mmax * widest) || e->alloc(e->xs, Mmax * 4)) return -1;
  }
  if (e->alloc(e->logits, (size_t) c.max_batch * c.vocab * 4)) return -1;
  if (c.tp_size > 1 && e->alloc(e->logits_h, (size_t) c.max_batch * e->vocab_l * 2 * (1 + c.tp_size))) return -1;
  e->kv.resize(c.layers);
  if (c.paged_kv_tokens_per_block > 0) {
    const int t = c.paged_kv_tokens_per_block;
    if (t < 16 || (t & (t - 1))) return fail("paged_kv_tokens_per_block must be a power of two >= 16");
    e->tpb = t;
    e->max_blocks = (e->S_max + t - 1) / t;
    e->pool_blocks = c.max_batch * e->max_blocks;
    e->h_block_tables.assign((size_t) c.layers * c.max_batch * 2 * e->max_blocks, 0);
    if (e->alloc(e->d_block_tables, e->h_block_tables.size() * sizeof(long long))) return -1;
    RT_CUDA(cudaMemset(e->d_block_tables, 0, e->h_block_tables.size() * sizeof(long long)));
  }
  const size_t kv_bytes = e->tpb ? (size_t) 2 * e->pool_blocks * e->Hl * e->tpb * c.head_size * (c.int8_kv ? 1 : 2)
                                 : (size_t) c.max_batch * 2 * e->Hl * e->S_max * c.head_size * (c.int8_kv ? 1 : 2);
  for (int i = 0; i < c.layers; ++i) {
    if (e->alloc(e->kv[i], kv_bytes)) return -1;
    RT_CUDA(cudaMemset(e->kv[i], 0, kv_bytes));
  }
  // workspace: the maximum any plugin asks for over the shapes this engine runs (TensorRT does the same)
  size_t ws = tb_mmha_workspace_bytes(c.max_batch, e->Hl, 32);
  {
    const size_t w = tb_context_attention_workspace_bytes(c.max_batch, c.max_input_len, e->Hl);
    if (w > ws) ws = w;
  }
  const int Ns[4] = {3 * e->hid_l, c.hidden, 2 * e->inter_l, e->vocab_l};
  const int Ks[4] = {c.hidden, e->inter_l, c.hidden, c.hidden};
  const int Ms[2] = {c.max_batch, (int) Mmax};
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 2; ++b) {
      const size_t w = tb_gemm_tc_workspace_bytes(Ms[b], Ns[a], Ks[a]);
      if (w > ws) ws = w;
    }
  {
    const size_t w = tb_gemm_tc_workspace_bytes(c.max_batch, c.hidden, e->hid_l);
    if (w > ws) ws = w;
  }
  {
    // packed input: the attention plugin stages padded copies of qkv and of its output in front of the kernels' own space
    const size_t a = ((Mmax * 3 * e->hid_l * 2 + 127) & ~(size_t) 127) + ((Mmax * e->hid_l * 2 + 127) & ~(size_t) 127);
    const size_t w = a + tb_context_attention_workspace_bytes(c.max_batch, c.max_input_len, e->Hl) + 256 +
                     tb_mmha_workspace_bytes(c.max_batch, e->Hl, 32) + 256;
    if (w > ws) ws = w;
  }
  e->workspace_bytes = ws + 1024;
  if (e->alloc(e->workspace, e->workspace_bytes)) return -1;
  if (e->alloc(e->d_ids, (size_t) c.max_batch * 4) || e->alloc(e->d_in_lens, (size_t) c.max_batch * 4) ||
      e->alloc(e->d_seq_lens, (size_t) c.max_batch * 4) || e->alloc(e->d_step_pos, 4) ||
      e->alloc(e->d_next, (size_t) c.max_batch * 4) || e->alloc(e->d_out_ids, (size_t) c.max_batch * c.max_output_len * 4) ||
      e->alloc(e->d_flag, 4) || e->alloc(e->d_max_in, 4) || cudaMallocHost(reinterpret_cast<void**>(&e->h_flag), 4) != cudaSuccess ||
      e->alloc(e->d_prompt, Mmax * 4) || e->alloc(e->d_dummy_scale, 4))
    return -1;
  RT_CUDA(cudaMemset(e->d_dummy_scale, 0, 4));
  if (c.tp_size > 1 && c.tp_size <= 8) {
    tb_decode_step_config dc{};
    dc.hidden = c.hidden; dc.vocab = c.vocab; dc.tp_size = c.tp_size;
    RT_CALL(tb_ar_create_ex(&e->ar, c.tp_rank, c.tp_size, (size_t) 8 * c.hidden * 2, tb_decode_step_tp_bytes(&dc)));
  }
  e->build_decode_step();
  RT_CUDA(cudaDeviceSynchronize());
  e->finalized = true;
  return 0;
}

size_t tbrt_device_bytes(const tbrt_engine* e) { return e->dev_bytes; }
const float* tbrt_logits(const tbrt_engine* e) {
  // a tensor-parallel fused step gathers the vocabulary shards into its peer-mapped scratch
  if (e->last_step_fused && e->c.tp_size > 1) return tb_decode_step_tp_logits(e->ds);
  return e->logits;
}
const int32_t* tbrt_output_ids(const tbrt_engine* e) { return e->d_out_ids; }
int tbrt_set_end_id(tbrt_engine* e, int end_id) { e->end_id = end_id; return 0; }
int tbrt_kv_max_blocks_per_seq(const tbrt_engine* e) { return e->max_blocks; }
int tbrt_set_kv_blocks(tbrt_engine* e, const int32_t* ids, int batch, int blocks_per_seq, tb_stream_t st) {
  if (!e->tpb) return fail("the engine was not built with a paged KV cache");
  if (!ids || batch < 1 || batch > e->c.max_batch || blocks_per_seq < 1 || blocks_per_seq > e->max_blocks)
    return fail("tbrt_set_kv_blocks: bad table shape");
  const size_t elt = e->c.int8_kv ? 1 : 2;
  const size_t block_bytes = (size_t) e->Hl * e->tpb * e->c.head_size * elt;
  const size_t per_layer = (size_t) e->c.max_batch * 2 * e->max_blocks;
  for (int li = 0; li < e->c.layers; ++li) {
    // pool layout [2][blocks][H][tokens_per_block][Dh]: K blocks, then V blocks (kv_cache_manager.py:84-93)
    const long long base = reinterpret_cast<long long>(e->kv[li]);
    long long* t = e->h_block_tables.data() + (size_t) li * per_layer;
    for (int b = 0; b < batch; ++b)
      for (int j = 0; j < e->max_blocks; ++j) {
        const int id = j < blocks_per_seq ? ids[(size_t) b * blocks_per_seq + j] : -1;
        if (id >= e->pool_blocks) return fail("tbrt_set_kv_blocks: block id outside the pool");
        t[((size_t) b * 2 + 0) * e->max_blocks + j] = id < 0 ? 0 : base + (long long) ((size_t) id * block_bytes);
        t[((size_t) b * 2 + 1) * e->max_blocks + j] = id < 0 ? 0 : base + (long long) (((size_t) e->pool_blocks + id) * block_bytes);
      }
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(st);
  RT_CUDA(cudaMemcpyAsync(e->d_block_tables, e->h_block_tables.data(), e->h_block_tables.size() * sizeof(long long),
                          cudaMemcpyHostToDevice, s));
  RT_CUDA(cudaStreamSynchronize(s));      // the staging vector is reused by the next call
  return 0;
}
int tbrt_set_sampling(tbrt_engine* e, int top_k, float top_p, float temperature, unsigned long long seed) {
  if (top_k < 0 || top_k > 1024 || top_p < 0.f || top_p > 1.f || !(temperature >= 0.f)) return fail("bad sampling parameters");
  const bool changed = e->top_k != top_k || e->top_p != top_p || e->temperature != temperature || e->seed != seed;
  e->top_k = top_k; e->top_p = top_p; e->temperature = temperature; e->seed = seed;
  if (changed) {                       // captured step graphs bake the sampling kernel and its arguments
    for (auto& g : e->graphs) cudaGraphExecDestroy(g.second);
    e->graphs.clear();
    e->graph_nodes.clear();
    e->eager_steps.clear();
  }
  return 0;
}
int tbrt_set_decode_mode(tbrt_engine* e, int mode) { e->decode_mode = mode < 0 ? -1 : (mode ? 1 : 0); return 0; }
void* tbrt_decode_step_handle(tbrt_engine* e) { return e->ds; }
int tbrt_fused_step_available(const tbrt_engine* e) { return e->ds ? tb_decode_step_max_batch() : 0; }
int tbrt_last_steps(const tbrt_engine* e) { return e->last_steps; }
void* tbrt_kv_cache(const tbrt_engine* e, int layer) { return (layer >= 0 && layer < (int) e->kv.size()) ? e->kv[layer] : nullptr; }
int64_t tbrt_last_launches(const tbrt_engine* e) { return e->launches; }
int tbrt_ar_handle(tbrt_engine* e, void* out64) {
  if (!e->ar) return fail("no peer all-reduce context (tp_size == 1 or engine not finalized)");
  RT_CALL(tb_ar_ipc_handle(e->ar, out64));
  return 0;
}
int tbrt_ar_open(tbrt_engine* e, const void* handles) {
  if (!e->ar) return fail("no peer all-reduce context (tp_size == 1 or engine not finalized)");
  RT_CALL(tb_ar_open_peers(e->ar, handles));
  e->ar_open = true;
  e->build_decode_step();       // tensor parallel: the fused step pushes partial sums / flags into the peers' scratch
  return 0;
}

int tbrt_context(tbrt_engine* e, const int32_t* ids, const int32_t* input_lengths, int batch, int seq, tb_stream_t st) {
  if (!e->finalized) return fail("engine not finalized");
  if (batch < 1 || batch > e->c.max_batch || seq < 1 || seq > e->c.max_input_len) return fail("batch / seq outside the engine limits");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(st);
  e->B = batch; e->S_in = seq; e->steps_done = 0; e->launches = 0; e->ar_site = 0; e->last_step_fused = false;
  e->beam_W = 1;
  const int M = batch * seq;
  // step state: the first generated token lands in column 0; every sequence of the padded batch sits at
  // position seq afterwards (sequence_length = max_input_len + step, generation.py:686-687)
  RT_CUDA(cudaMemsetAsync(e->d_step_pos, 0, 4, s));
  RT_CALL(tb_fill_int(e->d_seq_lens, seq - 1, batch, s));
  RT_CALL(tb_fill_int(e->d_max_in, seq, 1, s));
  RT_CUDA(cudaMemcpyAsync(e->d_in_lens, input_lengths, (size_t) batch * 4, cudaMemcpyDeviceToDevice, s));
  e->launches += 1;
  RT_CALL(tb_embedding(e->h, e->emb, ids, M, e->c.hidden, e->c.vocab, s));
  if (e->layers_forward(M, seq, true, s)) return -1;
  e->launches += 1;
  RT_CALL(tb_gather_last_token(e->hl, e->h, e->d_in_lens, batch, seq, e->c.hidden, s));
  return e->head(batch, e->hl, s);
}

// Beam search over the request tbrt_context just ran (generation.py:898-915: the context phase runs once per batch entry,
// then the KV cache, lengths and logits are tiled beam_width times).  Afterwards the engine's rows are batch x beam_width:
// tbrt_step advances all beams, tbrt_beam_finalize walks the parent pointers (gather_tree).
int tbrt_beam_begin(tbrt_engine* e, int beam_width, float length_penalty, int end_id, tb_stream_t st) {
  if (!e->finalized || e->B == 0 || e->steps_done != 0 || e->beam_W != 1) return fail("tbrt_beam_begin must follow tbrt_context");
  if (beam_width < 2 || beam_width > 16) return fail("beam_width must be in [2, 16]");
  const int Bq = e->B, W = beam_width, rows = Bq * W;
  if (rows > e->c.max_batch) return fail("batch x beam_width exceeds the engine's max_batch (build with max_batch_size x max_beam_width rows)");
  if (e->tpb) return fail("beam search reads the contiguous KV cache (paged_kv_cache engines: beam width 1)");
  if (e->c.tp_size > 1 && rows > 8) return fail("tensor parallel decode handles at most 8 rows");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(st);
  if (!e->d_cum) {
    const size_t mb = (size_t) e->c.max_batch;
    if (e->alloc(e->d_cum, mb * 4) || e->alloc(e->d_fin, mb * 4) || e->alloc(e->d_beam_lens, mb * 4) ||
        e->alloc(e->d_ids_t, mb * e->c.max_output_len * 4) || e->alloc(e->d_parent_t, mb * e->c.max_output_len * 4) ||
        e->alloc(e->d_indir[0], mb * e->S_max * 4) || e->alloc(e->d_indir[1], mb * e->S_max * 4) ||
        e->alloc(e->d_beam_out, mb * e->c.max_output_len * 4) || e->alloc(e->d_beam_ws, tb_beam_workspace_bytes((int) mb, 16)))
      return -1;
  }
  // tile the cache rows: row b -> rows [b W, b W + W), highest entry first so no source row is overwritten before it is read
  const size_t row_bytes = (size_t) 2 * e->Hl * e->S_max * e->c.head_size * (e->c.int8_kv ? 1 : 2);
  for (int li = 0; li < e->c.layers; ++li) {
    char* base = static_cast<char*>(e->kv[li]);
    for (int b = Bq - 1; b >= 0; --b)
      for (int w = W - 1; w >= 0; --w) {
        const int dst = b * W + w;
        if (dst == b) continue;
        RT_CUDA(cudaMemcpyAsync(base + (size_t) dst * row_bytes, base + (size_t) b * row_bytes, row_bytes, cudaMemcpyDeviceToDevice, s));
      }
  }
  RT_CALL(tb_tile_int(e->d_in_lens, Bq, W, s));
  e->beam_W = W; e->beam_end_id = end_id; e->beam_length_penalty = length_penalty;
  RT_CALL(tb_beam_init(e->d_cum, e->d_fin, e->d_beam_lens, e->d_indir[0], e->d_indir[1], e->d_max_in, rows, W, e->S_max, s));
  // the context phase chose column 0 greedily; redo it as the first beam step on the same logits (every beam of an entry
  // reads the entry's one logits row; cum_log_probs {0, -1e20, ...} make the W candidates come from that one distribution)
  RT_CUDA(cudaMemsetAsync(e->d_step_pos, 0, 4, s));
  RT_CALL(tb_fill_int(e->d_seq_lens, e->S_in - 1, rows, s));
  if (e->beam_step(true, s)) return -1;
  e->B = rows;
  RT_CALL(tb_advance_step(e->d_next, e->d_ids, e->d_out_ids, e->d_seq_lens, e->d_step_pos, rows, e->c.max_output_len, s));
  return 0;
}

// host_out [batch][beam_width][n_steps] (best beam first), cum_log_probs_out [batch][beam_width] or NULL
int tbrt_beam_finalize(tbrt_engine* e, int32_t* host_out, float* cum_log_probs_out, int n_steps, tb_stream_t st) {
  if (e->beam_W < 2) return fail("tbrt_beam_finalize without tbrt_beam_begin");
  if (n_steps < 1 || n_steps > e->steps_done + 1) return fail("n_steps exceeds the steps run");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(st);
  RT_CALL(tb_gather_tree(e->d_beam_out, e->d_ids_t, e->d_parent_t, e->B, e->beam_W, n_steps, e->beam_end_id, s));
  RT_CUDA(cudaMemcpyAsync(host_out, e->d_beam_out, (size_t) e->B * n_steps * 4, cudaMemcpyDeviceToHost, s));
  if (cum_log_probs_out) RT_CUDA(cudaMemcpyAsync(cum_log_probs_out, e->d_cum, (size_t) e->B * 4, cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  return 0;
}

// Context phase on packed input (build.py --remove_input_padding; generation.py:355-363): ids [tokens] holds the prompts back
// to back, input_lengths [batch] on the device, seq = the longest prompt.  Every projection runs on `tokens` rows instead of
// batch x seq; the KV cache and the generation steps are those of the padded batch.
int tbrt_context_packed(tbrt_engine* e, const int32_t* ids, const int32_t* input_lengths, int batch, int tokens, int seq,
                        tb_stream_t st) {
  if (!e->finalized) return fail("engine not finalized");
  if (batch < 1 || batch > e->c.max_batch || seq < 1 || seq > e->c.max_input_len || tokens < batch || tokens > batch * seq)
    return fail("batch / seq / tokens outside the engine limits");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(st);
  e->B = batch; e->S_in = seq; e->steps_done = 0; e->launches = 0; e->ar_site = 0; e->last_step_fused = false;
  e->beam_W = 1;
  RT_CUDA(cudaMemsetAsync(e->d_step_pos, 0, 4, s));
  RT_CALL(tb_fill_int(e->d_seq_lens, seq - 1, batch, s));
  RT_CALL(tb_fill_int(e->d_max_in, seq, 1, s));
  RT_CUDA(cudaMemcpyAsync(e->d_in_lens, input_lengths, (size_t) batch * 4, cudaMemcpyDeviceToDevice, s));
  e->launches += 1;
  RT_CALL(tb_embedding(e->h, e->emb, ids, tokens, e->c.hidden, e->c.vocab, s));
  e->packed_B = batch;
  const int rc = e->layers_forward(tokens, seq, true, s);
  e->packed_B = 0;
  if (rc) return -1;
  e->launches += 1;
  RT_CALL(tb_gather_last_token_packed(e->hl, e->h, e->d_in_lens, batch, e->c.hidden, s));
  return e->head(batch, e->hl, s);
}

int tbrt_step(tbrt_engine* e, tb_stream_t st) {
  if (!e->finalized || e->B == 0) return fail("tbrt_step before tbrt_context");
  if (e->S_in + e->steps_done + 1 >= e->S_max) return fail("KV cache is full");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(st);
  e->steps_done += 1;
  e->last_step_fused = e->fused_step();
  if (e->fused_step()) {
    e->launches = 1;
    RT_CALL(tb_decode_step_launch(e->ds, e->B, st));
    return 0;
  }
  if (!e->c.use_cuda_graph) { e->launches = 0; return e->step_body(s); }
  const int gkey = e->B | (e->beam_W << 16);   // a beam-search step graph holds different kernels than a greedy one
  auto it = e->graphs.find(gkey);
  if (it == e->graphs.end()) {
    // the first step at a batch size runs eagerly (plugins allocate their counters, kernels set their
    // attributes); the second is captured; every later one replays the graph
    if (e->eager_steps[gkey]++ == 0) { e->launches = 0; return e->step_body(s); }
    // capture on a private stream (the caller's may be the legacy default stream, which cannot capture);
    // nothing executes during capture, the instantiated graph is then launched on the caller's stream
    cudaGraph_t g = nullptr;
    if (!e->cap_stream) RT_CUDA(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    RT_CUDA(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    e->launches = 0;
    const int rc = e->step_body(e->cap_stream);
    cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &g);
    if (rc != 0) { if (g) cudaGraphDestroy(g); return -1; }
    RT_CUDA(ce);
    size_t nodes = 0;
    RT_CUDA(cudaGraphGetNodes(g, nullptr, &nodes));
    cudaGraphExec_t ge = nullptr;
    RT_CUDA(cudaGraphInstantiate(&ge, g, 0));
    RT_CUDA(cudaGraphDestroy(g));
    e->graphs[gkey] = ge;
    e->graph_nodes[gkey] = (int64_t) nodes;
    it = e->graphs.find(gkey);
  }
  e->launches = e->graph_nodes[gkey];
  RT_CUDA(cudaGraphLaunch(it->second, s));
  return 0;
}

int tbrt_force_ids(tbrt_engine* e, const int32_t* ids, tb_stream_t st) {
  if (!e->finalized || e->B == 0 || !ids) return fail("tbrt_force_ids before tbrt_context");
  if (e->beam_W != 1) return fail("tbrt_force_ids: not during beam search");
  RT_CALL(tb_force_ids(ids, e->d_ids, e->d_out_ids, e->d_step_pos, e->B, e->c.max_output_len, reinterpret_cast<cudaStream_t>(st)));
  return 0;
}

int tbrt_generate(tbrt_engine* e, const int32_t* host_ids, const int32_t* host_lengths, int batch, int seq, int max_new,
                  int32_t* host_out_ids, tb_stream_t st) {
  if (!e->finalized) return fail("engine not finalized");
  if (max_new < 1 || max_new > e->c.max_output_len) return fail("max_new outside the engine limits");
  if (batch < 1 || batch > e->c.max_batch || seq < 1 || seq > e->c.max_input_len) return fail("batch / seq outside the engine limits");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(st);
  RT_CUDA(cudaMemcpyAsync(e->d_prompt, host_ids, (size_t) batch * seq * 4, cudaMemcpyHostToDevice, s));
  RT_CUDA(cudaMemcpyAsync(e->d_next, host_lengths, (size_t) batch * 4, cudaMemcpyHostToDevice, s));
  if (tbrt_context(e, e->d_prompt, e->d_next, batch, seq, st)) return -1;
  int64_t total = e->launches;
  int steps = 1;
  for (int i = 1; i < max_new; ++i) {
    if (e->end_id >= 0 && (i % 16) == 0) {
      // stop criterion: every sequence has produced end_id (checked every 16 steps: one tiny kernel + a 4-byte read)
      if (tb_finished(e->d_flag, e->d_out_ids, batch, e->c.max_output_len, i, i, e->end_id, 0, s)) return fail("tb_finished");
      RT_CUDA(cudaMemcpyAsync(e->h_flag, e->d_flag, 4, cudaMemcpyDeviceToHost, s));
      RT_CUDA(cudaStreamSynchronize(s));
      total += 1;
      if (*e->h_flag) break;
    }
    if (tbrt_step(e, st)) return -1;
    total += e->launches;
    ++steps;
  }
  e->last_steps = steps;
  if (e->end_id >= 0) {
    if (tb_finished(nullptr, e->d_out_ids, batch, e->c.max_output_len, steps, max_new, e->end_id, 1, s)) return fail("tb_finished");
    total += 1;
  }
  RT_CUDA(cudaMemcpy2DAsync(host_out_ids, (size_t) max_new * 4, e->d_out_ids, (size_t) e->c.max_output_len * 4,
                            (size_t) max_new * 4, batch, cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  e->launches = total;
  return 0;
}
}
