"""Torch-tensor front end of the C ABI (include/trtllm_b200.h).

PyTorch is used for device memory and streams only; every function below launches the hand-written
sm_100a kernels in ``libtrtllm_llama_b200.so`` through ctypes and raises if the call is rejected.
Names follow the reference's functional layer (T/tensorrt_llm/quantization/functional.py:12-212,
T/tensorrt_llm/functional.py:2695-2928) so call sites read like the reference's.
"""
from __future__ import annotations

import torch

from ._lib import check, lib

KIND_F16, KIND_W8, KIND_W4, KIND_A8W8 = 0, 1, 2, 3
_OUT_TYPES = {torch.float16: 0, torch.float32: 1, torch.int32: 2}


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise ValueError("expected contiguous CUDA tensors")


_ws_cache = {}
_cnt_cache = {}


def _counters(device) -> torch.Tensor:
    """zeroed once; split-K / split-L arrival counters reset themselves after every launch."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    cur = _cnt_cache.get(key)
    if cur is None:
        cur = torch.zeros(1 << 16, dtype=torch.int32, device=device)
        _cnt_cache[key] = cur
    return cur


def _workspace(nbytes: int, device) -> torch.Tensor:
    """scratch for split-K / split-L partials (contents need not be initialised)."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    cur = _ws_cache.get(key)
    if cur is None or cur.numel() < nbytes:
        cur = torch.empty(max(nbytes, 1 << 22), dtype=torch.uint8, device=device)
        _ws_cache[key] = cur
    return cur


# ------------------------------------------------------------------------------------------------
def rms_norm(x, weight, eps=1e-6, residual=None, return_sum=False):
    """T/tensorrt_llm/functional.py:3195-3219 rms_norm (+ fused residual add)."""
    _chk_cuda(x, weight, residual)
    rows, hidden = x.numel() // x.shape[-1], x.shape[-1]
    out = torch.empty_like(x)
    s = torch.empty_like(x) if (residual is not None and return_sum) else None
    check(lib.tb_rmsnorm(_p(out), _p(x), _p(residual), _p(s), _p(weight), eps, rows, hidden, _stream()), "tb_rmsnorm")
    return (out, s) if return_sum else out


def smooth_quant_rms_norm(x, weight, scale=None, eps=1e-6, dynamic_act_scaling=True, residual=None,
                          return_sum=False, bias=None, layernorm=False):
    """RmsnormQuantization plugin op (modelled on smooth_quant_layer_norm,
    T/tensorrt_llm/quantization/functional.py:77-129): returns int8 (and per-token scales if dynamic)."""
    _chk_cuda(x, weight, residual, scale, bias)
    rows, hidden = x.numel() // x.shape[-1], x.shape[-1]
    q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    ds = torch.empty(x.shape[:-1] + (1,), dtype=torch.float32, device=x.device) if dynamic_act_scaling else None
    s = torch.empty_like(x) if (residual is not None and return_sum) else None
    check(lib.tb_rmsnorm_quant(_p(q), _p(ds), _p(x), _p(residual), _p(s), _p(weight), _p(bias), _p(scale), eps, rows,
                               hidden, int(dynamic_act_scaling), int(layernorm), _stream()), "tb_rmsnorm_quant")
    res = (q, ds) if dynamic_act_scaling else (q,)
    return res + ((s,) if return_sum else ())


def quantize_per_token(x):
    """T/tensorrt_llm/quantization/functional.py:160-186."""
    _chk_cuda(x)
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    s = torch.empty(x.shape[:-1] + (1,), dtype=torch.float32, device=x.device)
    check(lib.tb_quantize_per_token(_p(q), _p(s), _p(x), rows, cols, int(x.dtype == torch.float32), _stream()),
          "tb_quantize_per_token")
    return q, s


def quantize_tensor(x, scale):
    """T/tensorrt_llm/quantization/functional.py:188-212."""
    _chk_cuda(x, scale)
    q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    check(lib.tb_quantize_tensor(_p(q), _p(x), x.numel(), _p(scale), int(x.dtype == torch.float32), _stream()),
          "tb_quantize_tensor")
    return q


# ------------------------------------------------------------------------------------------------
PRO_NONE, PRO_RMS, PRO_RMS_QUANT, PRO_QUANT = 0, 1, 2, 3


def gemv(kind, x, w, *, w_scale=None, sc=None, sr=None, residual=None, swiglu=False, out_fp32=False, prologue=0,
         gamma=None, eps=1e-6):
    _chk_cuda(x, w, w_scale, sc, sr, residual, gamma)
    M, K = x.numel() // x.shape[-1], x.shape[-1]
    N = w.shape[0]
    n_out = N // 2 if swiglu else N
    if out_fp32:
        y32 = torch.empty(x.shape[:-1] + (n_out,), dtype=torch.float32, device=x.device)
        y = None
    else:
        y = torch.empty(x.shape[:-1] + (n_out,), dtype=torch.float16, device=x.device)
        y32 = None
    check(lib.tb_gemv_fused(kind, _p(y), _p(y32), _p(x), _p(w), _p(w_scale), _p(sc), _p(sr),
                            int(sc is not None and sc.numel() > 1), int(sr is not None and sr.numel() > 1), _p(residual),
                            M, N, K, int(swiglu), int(prologue), _p(gamma), float(eps), _stream()), "tb_gemv_fused")
    return y32 if out_fp32 else y


def gemm_tc(kind, x, w, *, w_scale=None, sc=None, sr=None, residual=None, out_dtype=torch.float16, force_splits=0,
            force_nt=0):
    _chk_cuda(x, w, w_scale, sc, sr, residual)
    M, K = x.numel() // x.shape[-1], x.shape[-1]
    N = w.shape[0]
    c = torch.empty(x.shape[:-1] + (N,), dtype=out_dtype, device=x.device)
    need = lib.tb_gemm_tc_workspace_bytes(M, N, K) * (4 if force_splits else 1) + (1 << 20)
    if force_splits:
        need = max(need, 4096 + ((N + 127) // 128) * ((M + 15) // 16) * force_splits * 256 * 128 * 4)
    ws = _workspace(need, x.device)
    check(lib.tb_gemm_tc(kind, _p(c), _OUT_TYPES[out_dtype], _p(x), _p(w), _p(w_scale), _p(sc), _p(sr),
                         int(sc is not None and sc.numel() > 1), int(sr is not None and sr.numel() > 1), _p(residual),
                         M, N, K, _p(ws), ws.numel(), _p(_counters(x.device)), force_splits, force_nt, _stream()),
          "tb_gemm_tc")
    return c


def smooth_quant_gemm(x_i8, w_i8, scale_tokens, scale_channels, per_token_scaling, per_channel_scaling,
                      out_dtype=torch.float16, use_gemv=None):
    """SmoothQuantGemm plugin op — T/tensorrt_llm/quantization/functional.py:12-53.
    x int8 [..., K], w int8 [N, K]; scales fp32 ([M,1]|[1,1], [1,N]|[1,1])."""
    M = x_i8.numel() // x_i8.shape[-1]
    if use_gemv is None:
        use_gemv = M <= lib.tb_gemv_max_rows(KIND_A8W8, x_i8.shape[-1]) and out_dtype == torch.float16
    if use_gemv:
        return gemv(KIND_A8W8, x_i8, w_i8, sc=scale_channels, sr=scale_tokens)
    return gemm_tc(KIND_A8W8, x_i8, w_i8, sc=scale_channels, sr=scale_tokens, out_dtype=out_dtype)


def weight_only_quant_matmul(x, w_processed, scales, weight_type_id, use_gemv=None):
    """WeightOnlyQuantMatmul plugin op — T/tensorrt_llm/quantization/functional.py:56-74.
    weight_type_id 1 = int8 ([N,K] processed), 2 = int4 ([N,K/2] processed)."""
    kind = KIND_W8 if weight_type_id == 1 else KIND_W4
    M = x.numel() // x.shape[-1]
    if use_gemv is None:
        use_gemv = M <= lib.tb_gemv_max_rows(kind, x.shape[-1])
    if use_gemv:
        return gemv(kind, x, w_processed, w_scale=scales)
    return gemm_tc(kind, x, w_processed, w_scale=scales)


def matmul_f16(x, w, residual=None, out_fp32=False, use_gemv=None):
    """Gemm plugin / TRT-native MatMul(x, W^T), W [N,K] (T/tensorrt_llm/layers/linear.py:13-35)."""
    M = x.numel() // x.shape[-1]
    if use_gemv is None:
        use_gemv = M <= lib.tb_gemv_max_rows(KIND_F16, x.shape[-1])
    if use_gemv:
        return gemv(KIND_F16, x, w, residual=residual, out_fp32=out_fp32)
    return gemm_tc(KIND_F16, x, w, residual=residual, out_dtype=torch.float32 if out_fp32 else torch.float16)


# ------------------------------------------------------------------------------------------------
def mmha_decode(qkv, kv_cache, past_len, *, num_heads, head_size, max_input_len, seq_lens=None, input_lengths=None,
                masked_tokens=None, kv_scale_orig_quant=None, kv_scale_quant_orig=None, q_scaling=1.0,
                rotary_dim=None, nsplit=0, len_cap=None, max_splits=32, max_input_len_dev=None):
    """GPTAttention plugin, generation phase (T/tensorrt_llm/functional.py:2695-2928 with
    past_key_value_length = [past_len, 0]).  kv_cache [B,2,H,S_max,Dh] is updated in place."""
    _chk_cuda(qkv, kv_cache, seq_lens, input_lengths, masked_tokens, kv_scale_orig_quant, kv_scale_quant_orig)
    B = qkv.shape[0]
    S_max = kv_cache.shape[3]
    int8_kv = kv_cache.dtype == torch.int8
    rot = head_size if rotary_dim is None else rotary_dim
    cap = past_len if len_cap is None else len_cap
    if nsplit <= 0:
        nsplit = lib.tb_mmha_num_splits(B, num_heads, cap, max_splits)
    ws = _workspace(lib.tb_mmha_workspace_bytes(B, num_heads, max(nsplit, max_splits)), qkv.device)
    out = torch.empty((B, num_heads * head_size), dtype=torch.float16, device=qkv.device)
    # max_input_len_dev: one device int that overrides max_input_len (what the engine's replayed step graph passes)
    check(lib.tb_mmha_decode_dev(_p(out), _p(qkv), _p(kv_cache), _p(seq_lens), _p(input_lengths), _p(masked_tokens),
                                 _p(max_input_len_dev), _p(kv_scale_orig_quant), _p(kv_scale_quant_orig), _p(ws), None, B,
                                 num_heads, head_size, S_max, int(past_len), int(max_input_len), int(cap), rot,
                                 float(q_scaling), int(int8_kv), nsplit, _stream()), "tb_mmha_decode")
    return out


def context_attention(qkv, kv_cache, input_lengths, *, num_heads, head_size, kv_scale_orig_quant=None, q_scaling=1.0,
                      rotary_dim=None, use_tc=True):
    """GPTAttention plugin, context phase (past_key_value_length = [0, 1]).  qkv [B,S,3*H*Dh] is
    rotated in place; kv_cache written for positions [0, S)."""
    _chk_cuda(qkv, kv_cache, input_lengths, kv_scale_orig_quant)
    B, S = qkv.shape[0], qkv.shape[1]
    rot = head_size if rotary_dim is None else rotary_dim
    out = torch.empty((B, S, num_heads * head_size), dtype=torch.float16, device=qkv.device)
    ws = _workspace(lib.tb_context_attention_workspace_bytes(B, S, num_heads), qkv.device) if use_tc else None
    check(lib.tb_context_attention(_p(out), _p(qkv), _p(kv_cache), _p(input_lengths), _p(kv_scale_orig_quant), _p(ws), B, S,
                                   num_heads, head_size, kv_cache.shape[3], rot, float(q_scaling),
                                   int(kv_cache.dtype == torch.int8), _stream()), "tb_context_attention")
    return out


def sample(logits, top_k=0, top_p=1.0, temperature=1.0, seed=0, step=0, finished=None, end_id=2, return_uniform=False):
    """DynamicDecodeOp's sampling for one step (T/cpp/tensorrt_llm/thop/dynamicDecodeOp.cpp:359-363): logits fp32 [B, V]
    -> ids int32 [B].  ``step`` may be a device int32 tensor (graph-replayable counter)."""
    _chk_cuda(logits, finished)
    B, V = logits.shape
    out = torch.empty((B,), dtype=torch.int32, device=logits.device)
    u = torch.empty((B,), dtype=torch.float32, device=logits.device) if return_uniform else None
    step_dev = step if isinstance(step, torch.Tensor) else None
    check(lib.tb_sample(_p(out), _p(logits), B, V, logits.stride(0), int(top_k), float(top_p), float(temperature), int(seed),
                        _p(step_dev), 0 if step_dev is not None else int(step), _p(finished), int(end_id), _p(u), _stream()),
          "tb_sample")
    return (out, u) if return_uniform else out


# ------------------------------------------------------------------------------------------------
def embedding(ids, table):
    _chk_cuda(ids, table)
    out = torch.empty(ids.shape + (table.shape[1],), dtype=torch.float16, device=table.device)
    check(lib.tb_embedding(_p(out), _p(table), _p(ids), ids.numel(), table.shape[1], table.shape[0], _stream()),
          "tb_embedding")
    return out


def swiglu(gate_up):
    """gate_up [..., 2*inter] (gate | up) -> silu(gate) * up."""
    _chk_cuda(gate_up)
    inter = gate_up.shape[-1] // 2
    rows = gate_up.numel() // gate_up.shape[-1]
    out = torch.empty(gate_up.shape[:-1] + (inter,), dtype=torch.float16, device=gate_up.device)
    g = gate_up.view(rows, 2 * inter)
    check(lib.tb_swiglu(_p(out), g.data_ptr(), g.data_ptr() + inter * 2, rows, inter, 2 * inter, _stream()), "tb_swiglu")
    return out


def swiglu_quant(gate_up):
    """SwiGLU + per-token int8 quantisation in one pass: gate_up [..., 2*inter] -> (int8 [..., inter], fp32 scales [..., 1])."""
    _chk_cuda(gate_up)
    inter = gate_up.shape[-1] // 2
    rows = gate_up.numel() // gate_up.shape[-1]
    q = torch.empty(gate_up.shape[:-1] + (inter,), dtype=torch.int8, device=gate_up.device)
    sc = torch.empty(gate_up.shape[:-1] + (1,), dtype=torch.float32, device=gate_up.device)
    g = gate_up.view(rows, 2 * inter)
    check(lib.tb_swiglu_quant(_p(q), _p(sc), g.data_ptr(), g.data_ptr() + inter * 2, rows, inter, 2 * inter, _stream()),
          "tb_swiglu_quant")
    return q, sc


def add(a, b):
    _chk_cuda(a, b)
    out = torch.empty_like(a)
    check(lib.tb_add(_p(out), _p(a), _p(b), a.numel(), _stream()), "tb_add")
    return out


def argmax(logits):
    _chk_cuda(logits)
    rows, vocab = logits.numel() // logits.shape[-1], logits.shape[-1]
    out = torch.empty(logits.shape[:-1], dtype=torch.int32, device=logits.device)
    check(lib.tb_argmax(_p(out), _p(logits), rows, vocab, vocab, _stream()), "tb_argmax")
    return out


def mmha_decode_beams(qkv, kv_cache, cache_indirection, past_len, *, num_heads, head_size, max_input_len, input_lengths=None,
                      kv_scale_orig_quant=None, kv_scale_quant_orig=None, q_scaling=1.0, nsplit=0):
    """GPTAttention plugin, generation phase with beam search: cache_indirection [batch, beam, S_max] int32 names, per cached
    position, the beam whose cache row is read (T/tensorrt_llm/functional.py:2695-2928, input 7)."""
    _chk_cuda(qkv, kv_cache, cache_indirection, input_lengths, kv_scale_orig_quant, kv_scale_quant_orig)
    rows, S_max = qkv.shape[0], kv_cache.shape[3]
    beam = cache_indirection.shape[1]
    if nsplit <= 0:
        nsplit = lib.tb_mmha_num_splits(rows, num_heads, past_len, 32)
    out = torch.empty((rows, num_heads * head_size), dtype=torch.float16, device=qkv.device)
    check(lib.tb_mmha_decode_beams(_p(out), _p(qkv), _p(kv_cache), _p(cache_indirection), beam, None, _p(input_lengths), None,
                                   None, _p(kv_scale_orig_quant), _p(kv_scale_quant_orig), rows, num_heads, head_size, S_max,
                                   int(past_len), int(max_input_len), int(past_len), head_size, float(q_scaling),
                                   int(kv_cache.dtype == torch.int8), nsplit, _stream()), "tb_mmha_decode_beams")
    return out


class BeamSearchState:
    """Device-resident decoder state of tb_beam_search_step (the beam half of DynamicDecodeOp)."""

    def __init__(self, batch, beam_width, max_new, max_seq_len, max_input_len, device="cuda"):
        rows = batch * beam_width
        i32 = dict(dtype=torch.int32, device=device)
        self.rows, self.W, self.S_max = rows, beam_width, max_seq_len
        self.cum = torch.empty(rows, dtype=torch.float32, device=device)
        self.finished, self.lens, self.next_ids = (torch.empty(rows, **i32) for _ in range(3))
        self.ids_t, self.parent_t = (torch.zeros((max_new, rows), **i32) for _ in range(2))
        self.indir = [torch.empty((batch, beam_width, max_seq_len), **i32) for _ in range(2)]
        self.step = torch.zeros(1, **i32)
        self.max_in = torch.full((1,), max_input_len, **i32)
        self.ws = _workspace(lib.tb_beam_workspace_bytes(rows, beam_width), device)
        check(lib.tb_beam_init(_p(self.cum), _p(self.finished), _p(self.lens), _p(self.indir[0]), _p(self.indir[1]),
                               _p(self.max_in), rows, beam_width, max_seq_len, _stream()), "tb_beam_init")

    def advance(self, logits, end_id, length_penalty=1.0, broadcast=False):
        """logits fp32 [rows, V] ([batch, V] with broadcast) -> writes column ``step`` and the target indirection."""
        check(lib.tb_beam_search_step(_p(logits), logits.shape[1], logits.stride(0), int(broadcast), self.rows, self.W,
                                      float(length_penalty), int(end_id), _p(self.step), _p(self.max_in), _p(self.cum),
                                      _p(self.finished), _p(self.lens), _p(self.ids_t), _p(self.parent_t), _p(self.next_ids),
                                      _p(self.indir[0]), _p(self.indir[1]), self.S_max, _p(self.ws), _stream()),
              "tb_beam_search_step")
        self.indir[0].copy_(self.indir[1])
        self.step += 1

    def gather_tree(self, n, end_id):
        out = torch.empty((self.rows, n), dtype=torch.int32, device=self.cum.device)
        check(lib.tb_gather_tree(_p(out), _p(self.ids_t), _p(self.parent_t), self.rows, self.W, n, int(end_id), _stream()),
              "tb_gather_tree")
        return out


def gemm_tc_swiglu(kind, x, w_gate_up, *, sc=None, sr=None):
    """The gate / up projection with SwiGLU in the tcgen05 epilogue: w_gate_up [2*inter, K] (gate rows, then up rows) ->
    fp16 [..., inter] = silu(x.gate^T) * (x.up^T).  kind KIND_F16 or KIND_A8W8.  Bit-identical to gemm_tc + swiglu."""
    _chk_cuda(x, w_gate_up, sc, sr)
    M, K = x.numel() // x.shape[-1], x.shape[-1]
    N = w_gate_up.shape[0]
    c = torch.empty(x.shape[:-1] + (N // 2,), dtype=torch.float16, device=x.device)
    check(lib.tb_gemm_tc_swiglu(kind, _p(c), _p(x), _p(w_gate_up), _p(sc), _p(sr), int(sc is not None and sc.numel() > 1),
                                int(sr is not None and sr.numel() > 1), M, N, K, _stream()), "tb_gemm_tc_swiglu")
    return c
