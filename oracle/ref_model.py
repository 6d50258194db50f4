"""CPU oracle (numpy) of the whole LLaMA decoder as the reference wires it through its plugins.

TEST INFRASTRUCTURE ONLY (same rule as ref_ops.py): imported by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs only.

Wiring followed (LQ/ = tensorrt_llm_july-release-v1/examples/llama_quant/, T/ = tensorrt_llm_july-release-v1/):
  LQ/llama_model.py:78-119   LLaMADecoderLayer.forward: input_layernorm -> attention -> +residual ->
                              post_layernorm -> GatedMLP -> +residual
  LQ/llama_model.py:159-287  embedding -> layers -> ln_f -> gather_last_token_logits -> lm_head -> fp32
  T/tensorrt_llm/layers/attention.py:128-184   qkv ColumnLinear -> gpt_attention plugin -> dense RowLinear
  T/tensorrt_llm/layers/mlp.py:43-73           GatedMLP: proj(act(fc(x)) * gate(x))
  T/tensorrt_llm/quantization/layer.py:120-153,204-220,306-382,685-852  SmoothQuant linear / MLP / attention:
        per-token dynamic activation scales, per-channel weight scales, quantize_per_token before the
        row-parallel GEMMs (dense, proj)
  T/tensorrt_llm/runtime/generation.py:782-997 padded-batch greedy decode loop
        (context: past_key_value_length=[0,1]; step s: [max_input_len+s, 0])
The reference's own SmoothQuant-LLaMA driver never built and is semantically wrong in four ways
(SURVEY.md F2); the SQ mode below is the *intended* arithmetic (RMSNorm, RoPE on, q_scaling 1,
quantised GatedMLP), built from the per-kernel semantics of ref_ops.py.

Modes (``mode``): "fp16" | "w8" | "w4" (weight-only) | "sq" (W8A8 SmoothQuant, per-token+per-channel).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import ref_ops as R

F16, F32 = np.float16, np.float32


@dataclass
class LlamaCfg:
    hidden: int = 4096
    heads: int = 32
    inter: int = 11008
    layers: int = 32
    vocab: int = 32000
    eps: float = 1e-6
    head_size: int = 128

    @staticmethod
    def tiny(layers=2, hidden=256, inter=384, vocab=512):
        return LlamaCfg(hidden=hidden, heads=hidden // 128, inter=inter, layers=layers, vocab=vocab)


def random_weights(cfg: LlamaCfg, seed=0, std=0.02):
    """fp16 weights in torch-Linear layout [out, in]; qkv rows are q | k | v
    (LQ/hf_llama_convert.py:364-384 stacks q,k,v)."""
    rng = np.random.default_rng(seed)
    n = lambda *s: (rng.standard_normal(s, dtype=F32) * std).astype(F16)  # noqa: E731
    g = lambda k: (1.0 + 0.1 * rng.standard_normal(k, dtype=F32)).astype(F16)  # noqa: E731
    w = {"vocab_embedding": n(cfg.vocab, cfg.hidden), "ln_f": g(cfg.hidden), "lm_head": n(cfg.vocab, cfg.hidden),
         "layers": []}
    for _ in range(cfg.layers):
        w["layers"].append({
            "input_layernorm": g(cfg.hidden), "qkv": n(3 * cfg.hidden, cfg.hidden), "dense": n(cfg.hidden, cfg.hidden),
            "post_layernorm": g(cfg.hidden), "gate": n(cfg.inter, cfg.hidden), "up": n(cfg.inter, cfg.hidden),
            "down": n(cfg.hidden, cfg.inter)})
    return w


# ------------------------------------------------------------------------------------------------
# build-time quantisation of one Linear weight [N, K]
# ------------------------------------------------------------------------------------------------
def quantize_linear(w_nk, mode):
    """-> dict consumed by ``linear`` below and (same tensors) by the engine under test."""
    w_nk = np.asarray(w_nk, F16)
    if mode == "fp16":
        return {"w": w_nk}
    if mode in ("w8", "w4"):
        # LQ/weight_quant.py:264-271: symmetric_quantize_last_axis_of_batched_matrix on W^T [K, N]
        q, s = R.symmetric_quantize(np.ascontiguousarray(w_nk.T), 8 if mode == "w8" else 4)
        return {"q": np.ascontiguousarray(q.T), "scales": s}          # q [N, K] unprocessed ints
    if mode == "sq":
        # per-channel symmetric int8: LQ/convert.py:27-103 generate_int8 (scale_w_orig_quant_c = 127/amax_c)
        wf = w_nk.astype(F32)
        amax = np.maximum(np.abs(wf).max(axis=1), F32(1e-8)).astype(F32)
        q = np.clip(np.rint(wf * (F32(127.0) / amax)[:, None]), -127, 127).astype(np.int8)
        return {"q": q, "scale_ch": (amax / F32(127.0)).astype(F32)}
    raise ValueError(mode)


def quantize_model(weights, mode):
    qw = {k: weights[k] for k in ("vocab_embedding", "ln_f", "lm_head")}   # lm_head stays fp16 (LQ/quant.py:58-59)
    qw["layers"] = []
    for lw in weights["layers"]:
        e = {"input_layernorm": lw["input_layernorm"], "post_layernorm": lw["post_layernorm"]}
        for name in ("qkv", "dense", "gate", "up", "down"):
            e[name] = quantize_linear(lw[name], mode)
        qw["layers"].append(e)
    return qw


def linear(x, lw, mode, act_q=None):
    """x [M, K] fp16 (or, for sq, act_q = (int8 [M,K], scale_tokens [M,1])) -> fp16 [M, N]."""
    if mode == "fp16":
        return R.gemm_f16(x, lw["w"])
    if mode in ("w8", "w4"):
        return R.weight_only_matmul(x, np.ascontiguousarray(lw["q"].T), lw["scales"])
    q, st = act_q
    return R.sq_gemm(q, lw["q"], st, lw["scale_ch"], F16)


# ------------------------------------------------------------------------------------------------
class OracleLlama:
    """Padded-batch greedy generation, KV cache [L][B,2,H,S_max,Dh] (int8 or fp16)."""

    def __init__(self, cfg: LlamaCfg, qweights, mode="fp16", int8_kv=False, kv_scale=None, max_seq_len=256):
        self.cfg, self.w, self.mode, self.int8_kv, self.S_max = cfg, qweights, mode, int8_kv, max_seq_len
        # LQ/weight_quant.py:439-446: kv_orig_quant_scale = 1/t, kv_quant_orig_scale = t
        t = F32(kv_scale if kv_scale is not None else 4.0 / 127.0)
        self.kv_oq, self.kv_qo = (F32(1.0) / t, t) if int8_kv else (None, None)
        self.cache = None

    def _norm_in(self, h, gamma):
        if self.mode == "sq":
            return None, R.rmsnorm_quant(h, gamma, self.cfg.eps, dynamic=True)
        return R.rmsnorm(h, gamma, self.cfg.eps), None

    def _layer(self, li, h, attn_fn):
        lw, m = self.w["layers"][li], self.mode
        x, xq = self._norm_in(h, lw["input_layernorm"])
        qkv = linear(x, lw["qkv"], m, xq)
        a = attn_fn(li, qkv)
        aq = R.quantize_per_token(a) if m == "sq" else None
        h = R.residual_add(linear(a, lw["dense"], m, aq), h)
        x, xq = self._norm_in(h, lw["post_layernorm"])
        act = R.swiglu(linear(x, lw["gate"], m, xq), linear(x, lw["up"], m, xq))
        actq = R.quantize_per_token(act) if m == "sq" else None
        return R.residual_add(linear(act, lw["down"], m, actq), h)

    def _logits(self, h_last):
        x = R.rmsnorm(h_last, self.w["ln_f"], self.cfg.eps)
        # lm_head fp16 GEMM, logits cast to fp32 (LQ/llama_model.py:272-279)
        return R.gemm_f16(x, self.w["lm_head"]).astype(F32)

    def context(self, input_ids, input_lengths):
        """input_ids [B, S] (padded), returns fp32 logits [B, V] at each sequence's last real token."""
        c = self.cfg
        B, S = input_ids.shape
        self.B, self.max_in, self.in_lens = B, S, np.asarray(input_lengths, np.int32)
        dt = np.int8 if self.int8_kv else F16
        self.cache = [np.zeros((B, 2, c.heads, self.S_max, c.head_size), dt) for _ in range(c.layers)]
        h = self.w["vocab_embedding"][input_ids.reshape(-1)]

        def attn(li, qkv):
            o = R.context_attention(qkv.reshape(B, S, -1), self.cache[li], self.in_lens, num_heads=c.heads,
                                    head_size=c.head_size, kv_scale_orig_quant=self.kv_oq)
            return o.reshape(B * S, -1)

        for li in range(c.layers):
            h = self._layer(li, h, attn)
        h = h.reshape(B, S, -1)[np.arange(B), self.in_lens - 1]
        self.past = S
        return self._logits(h)

    def step(self, token_ids):
        """token_ids [B] -> fp32 logits [B, V]; appends to the cache at position ``past``."""
        c = self.cfg
        h = self.w["vocab_embedding"][np.asarray(token_ids).reshape(-1)]

        def attn(li, qkv):
            return R.mmha_decode(qkv, self.cache[li], self.past, self.in_lens, self.max_in, num_heads=c.heads,
                                 head_size=c.head_size, kv_scale_orig_quant=self.kv_oq, kv_scale_quant_orig=self.kv_qo)

        for li in range(c.layers):
            h = self._layer(li, h, attn)
        self.past += 1
        return self._logits(h)

    def generate_beams(self, input_ids, input_lengths, max_new_tokens, beam_width, end_id, length_penalty=1.0):
        """Beam search as GenerationSession.decode drives it (generation.py:365-409,823-997): context once per batch entry,
        cache / lengths / logits tiled beam_width times, cum_log_probs {0, -1e20, ...}, one beam step per token, gather_tree.
        Returns (ids [B, W, max_new_tokens], cum_log_probs [B, W], min selection margin over all steps)."""
        c, W = self.cfg, int(beam_width)
        logits = self.context(input_ids, input_lengths)
        B = self.B
        rows = B * W
        self.cache = [np.repeat(k, W, axis=0) for k in self.cache]
        self.in_lens = np.repeat(self.in_lens, W)
        self.B = rows
        logits = np.repeat(logits, W, axis=0)
        cum = np.tile(np.array([0.0] + [-1e20] * (W - 1), dtype=F32), B)
        fin, lens = np.zeros(rows, bool), np.full(rows, self.max_in, np.int64)
        indir = np.zeros((B, W, self.S_max), np.int32)
        ids_t, par_t, margin = [], [], np.inf
        for s in range(max_new_tokens):
            tok, par, cum, fin, lens, indir, m = R.beam_search_step(logits, cum, fin, lens, indir, self.max_in + s, beam_width=W,
                                                                    end_id=end_id, length_penalty=length_penalty)
            margin = min(margin, m)
            ids_t.append(tok)
            par_t.append(par)
            if s + 1 < max_new_tokens:
                logits = self.step_beams(tok, indir, W)
        out = R.gather_tree(np.stack(ids_t), np.stack(par_t), W, end_id)
        return out.reshape(B, W, max_new_tokens), cum.reshape(B, W), margin

    def step_beams(self, token_ids, cache_indir, beam_width):
        c = self.cfg
        h = self.w["vocab_embedding"][np.asarray(token_ids).reshape(-1)]

        def attn(li, qkv):
            return R.mmha_decode_beams(qkv, self.cache[li], cache_indir, beam_width, self.past, self.in_lens, self.max_in,
                                       num_heads=c.heads, head_size=c.head_size, kv_scale_orig_quant=self.kv_oq,
                                       kv_scale_quant_orig=self.kv_qo)

        for li in range(c.layers):
            h = self._layer(li, h, attn)
        self.past += 1
        return self._logits(h)

    def generate(self, input_ids, input_lengths, max_new_tokens, return_logits=False):
        logits = self.context(input_ids, input_lengths)
        ids, all_logits = [], []
        for s in range(max_new_tokens):
            tok = R.greedy_argmax(logits)
            ids.append(tok)
            all_logits.append(logits)
            if s + 1 < max_new_tokens:
                logits = self.step(tok)
        out = np.stack(ids, axis=1)
        return (out, np.stack(all_logits, axis=1)) if return_logits else out
