"""Generates the golden fixtures under tests/golden/ — run ONCE in the authoring container, where
/root/reference exists; the fixtures (not the reference) travel with the repo.

Sources of truth, in the reference's own code:
  * the reference's C++ weight-only quantiser compiled from its sources (oracle/_ref/libref_host.so:
    K/cutlass_kernels/cutlass_preprocessors.cpp:615-721 symmetric_quantize)
  * the reference's Python test oracles, IMPORTED from /root/reference/.../tests/quantization/_utils.py
    (gt_matmul_smooth_quant :92-121, gt_quantize_per_token :124-129, woq_gt_matmul :36-62, woq_gen_weights :17-25)
    with `tensorrt_llm` stubbed (it only supplies a dtype map there) and `.cuda()` neutralised
  * HF transformers' LlamaForCausalLM (the oracle of T/tests/model/test_llama.py:153-354) with seeded weights.

    python tests/golden/make_golden.py
"""
import ctypes as C
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/tensorrt_llm_july-release-v1"
sys.path.insert(0, ROOT)


def load_ref_utils():
    stub = types.ModuleType("tensorrt_llm")
    stub._utils = types.SimpleNamespace(str_dtype_to_torch=lambda s: {"float16": torch.float16, "float32": torch.float32,
                                                                     "int32": torch.int32}[s])
    sys.modules["tensorrt_llm"] = stub
    torch.Tensor.cuda = lambda self, *a, **k: self       # the oracles only use the GPU as a scratch pad
    spec = importlib.util.spec_from_file_location("ref_quant_utils", os.path.join(REF, "tests/quantization/_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_host():
    return C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_host.so"))


def main():
    U = load_ref_utils()
    out = {}
    # ---- (1) SmoothQuant GEMM: T/tests/quantization/test_smooth_quant_gemm.py:23-109 at reduced m, n ----------
    torch.manual_seed(1234)
    m, n, k = 8, 256, 768
    a = torch.randint(-128, 128, (m, k), dtype=torch.int8)
    b = torch.randint(-128, 128, (n, k), dtype=torch.int8)
    for per_token in (False, True):
        for per_channel in (False, True):
            sa = torch.randint(1, 10, (m if per_token else 1, 1)).float() * 1e-2
            sb = torch.randint(1, 10, (1, n if per_channel else 1)).float() * 1e-2
            for dt in ("float16", "float32", "int32"):
                ref = U.gt_matmul_smooth_quant(a, b, sa, sb, dt)
                key = f"sq_{int(per_token)}{int(per_channel)}_{dt}"
                out[key] = ref.numpy()
                out[key + "_sa"] = sa.numpy()
                out[key + "_sb"] = sb.numpy()
    out["sq_a"], out["sq_b"] = a.numpy(), b.numpy()
    # ---- (2) per-token quantiser: T/tests/quantization/test_functional.py:110-155 --------------------------------
    torch.manual_seed(7)
    x = torch.randn(6, 4, 512, dtype=torch.float32).half()
    q, s = U.gt_quantize_per_token(x)
    out["qpt_x"], out["qpt_q"], out["qpt_s"] = x.numpy(), q.numpy(), s.numpy()
    # ---- (3)+(4) weight-only: reference C++ quantiser + reference Python matmul oracle ----------------------------
    lib = ref_host()
    for bits in (8, 4):
        n_, k_ = 128, 256
        w_nk = U.woq_gen_weights(n_, k_, "float16")           # seed 0, rand*2-1 (test_weight_only_quant_matmul.py:87)
        w_kn = w_nk.t().contiguous()                          # the op quantises [K, N] along K per column N
        wk = w_kn.numpy()
        processed = np.zeros(k_ * n_ * bits // 8, np.int8)
        unprocessed = np.zeros(k_ * n_ * bits // 8, np.int8)
        scales = np.zeros(n_, np.float16)
        rc = lib.ref_symmetric_quantize(wk.view(np.uint16).ctypes.data_as(C.c_void_p), C.c_int64(k_), C.c_int64(n_), bits,
                                        processed.ctypes.data_as(C.c_void_p), unprocessed.ctypes.data_as(C.c_void_p),
                                        scales.view(np.uint16).ctypes.data_as(C.c_void_p))
        assert rc == 0
        out[f"woq{bits}_w_kn"] = wk
        out[f"woq{bits}_unprocessed"] = unprocessed.reshape(k_, n_ * bits // 8)
        out[f"woq{bits}_scales"] = scales
        torch.manual_seed(0)
        act = (torch.rand(4, k_, dtype=torch.float16) * 2 - 1.0)
        from oracle import ref_ops as R
        q_int = unprocessed.reshape(k_, n_) if bits == 8 else R.unpack_int4(unprocessed.reshape(k_, n_ // 2))
        ref = U.woq_gt_matmul(4, act, torch.from_numpy(q_int.astype(np.float32)), torch.from_numpy(scales.astype(np.float32)),
                              "float16")
        out[f"woq{bits}_act"], out[f"woq{bits}_ref"] = act.numpy(), ref.numpy()
    # ---- (5) whole model: HF LlamaForCausalLM fp32, tiny config with head_dim 128 ---------------------------------------
    from transformers import LlamaConfig, LlamaForCausalLM
    from oracle import ref_model as RM
    cfg = RM.LlamaCfg.tiny(layers=2, hidden=256, inter=384, vocab=512)
    w = RM.random_weights(cfg, seed=11, std=0.05)
    hf = LlamaForCausalLM(LlamaConfig(hidden_size=cfg.hidden, intermediate_size=cfg.inter, num_hidden_layers=cfg.layers,
                                      num_attention_heads=cfg.heads, num_key_value_heads=cfg.heads, vocab_size=cfg.vocab,
                                      rms_norm_eps=cfg.eps, max_position_embeddings=64, tie_word_embeddings=False,
                                      attention_bias=False, mlp_bias=False))
    sd = {"model.embed_tokens.weight": w["vocab_embedding"], "model.norm.weight": w["ln_f"], "lm_head.weight": w["lm_head"]}
    for i, lw in enumerate(w["layers"]):
        p = f"model.layers.{i}."
        q_, k_w, v_ = np.split(lw["qkv"], 3, axis=0)
        sd.update({p + "input_layernorm.weight": lw["input_layernorm"], p + "post_attention_layernorm.weight": lw["post_layernorm"],
                   p + "self_attn.q_proj.weight": q_, p + "self_attn.k_proj.weight": k_w, p + "self_attn.v_proj.weight": v_,
                   p + "self_attn.o_proj.weight": lw["dense"], p + "mlp.gate_proj.weight": lw["gate"],
                   p + "mlp.up_proj.weight": lw["up"], p + "mlp.down_proj.weight": lw["down"]})
    missing = hf.load_state_dict({k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in sd.items()}, strict=False)
    assert not [m_ for m_ in missing.missing_keys if "inv_freq" not in m_], missing
    hf.eval()
    rng = np.random.default_rng(12)
    ids = rng.integers(3, cfg.vocab, (2, 9)).astype(np.int64)
    with torch.no_grad():
        o = hf(input_ids=torch.from_numpy(ids), use_cache=True)
        logits0 = o.logits[:, -1].numpy()
        tok = o.logits[:, -1].argmax(-1, keepdim=True)
        o2 = hf(input_ids=tok, past_key_values=o.past_key_values, use_cache=True)
        logits1 = o2.logits[:, -1].numpy()
    out["hf_ids"], out["hf_logits0"], out["hf_logits1"], out["hf_tok0"] = ids.astype(np.int32), logits0, logits1, tok.numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), sum(v.nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    main()
