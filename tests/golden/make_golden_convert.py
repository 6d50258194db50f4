"""Golden vectors for the HF -> FT converter pieces — run ONCE in the authoring container, where
/root/reference exists; only the fixture (tests/golden/convert_vectors.npz) travels with the repo.

The reference's own ``examples/llama_quant/convert.py`` and ``smoothquant.py`` are IMPORTED (``tensorrt_llm`` stubbed: they
only take ``torch_to_numpy`` from it) and run on seeded inputs:

  * generate_int8 (LQ/convert.py:27-103): plain matrix with fp32 ranges, plain matrix with fp16 ranges, fused QKV
  * smooth_gemm (LQ/smoothquant.py:37-67): one matrix, and two matrices sharing an input with a layernorm to fold into
  * split_and_save_weight (LQ/convert.py:160-325): file names and bytes for every tensor kind at tp = 1, and for the
    row-parallel / QKV kinds at tp = 2, with int8_outputs in {None, "kv_cache_only", "all"}

    python tests/golden/make_golden_convert.py
"""
import importlib.util
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/tensorrt_llm_july-release-v1/examples/llama_quant"


def load_ref(name):
    stub = types.ModuleType("tensorrt_llm")
    stub._utils = types.ModuleType("tensorrt_llm._utils")
    stub._utils.torch_to_numpy = lambda t: t.detach().cpu().numpy()
    sys.modules["tensorrt_llm"] = stub
    sys.modules["tensorrt_llm._utils"] = stub._utils
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def seeded_case(seed, k, n, qkv=False, rng_dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    shape = (k, 3, n) if qkv else (k, n)
    w = (torch.randn(shape, generator=g) * 0.05).to(torch.float16)
    cols = 3 * n if qkv else n
    rng = {"x": torch.rand(k if not qkv else cols, generator=g) * 4 + 0.1, "y": torch.rand(cols, generator=g) * 9 + 0.1,
           "w": w.abs().reshape(k, -1).max(dim=0)[0].to(rng_dtype).clip(1e-8, None)}
    return w.numpy(), rng


def main():
    C = load_ref("convert")
    S = load_ref("smoothquant")
    out = {}
    # ---- generate_int8 -------------------------------------------------------------------------------------------------
    for tag, seed, k, n, qkv, dt in (("plain32", 1, 48, 40, False, torch.float32), ("plain16", 2, 48, 40, False, torch.float16),
                                     ("qkv32", 3, 32, 32, True, torch.float32)):
        w, rng = seeded_case(seed, k, n, qkv, dt)
        res = C.generate_int8(w, rng, is_qkv=qkv)
        for key, v in res.items():
            out[f"gi8_{tag}_{key}"] = np.asarray(v)
    # ---- smooth_gemm ---------------------------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(7)
    w1 = (torch.randn(24, 16, generator=g) * 0.1).to(torch.float16)
    act = torch.rand(16, generator=g) * 3 + 0.05
    w1c = w1.clone()
    s1 = S.smooth_gemm(w1c, act, None, None, 0.5)
    out["sg_one_scales"], out["sg_one_w"] = s1.numpy(), w1c.numpy()
    wa = (torch.randn(24, 16, generator=g) * 0.1).float()
    wb = (torch.randn(8, 16, generator=g) * 0.2).float()
    ln_w, ln_b = torch.rand(16, generator=g) + 0.5, torch.rand(16, generator=g) - 0.5
    wac, wbc, lwc, lbc = wa.clone(), wb.clone(), ln_w.clone(), ln_b.clone()
    s2 = S.smooth_gemm([wac, wbc], act, lwc, lbc, 0.8)
    out["sg_two_scales"], out["sg_two_wa"], out["sg_two_wb"] = s2.numpy(), wac.numpy(), wbc.numpy()
    out["sg_two_lnw"], out["sg_two_lnb"] = lwc.numpy(), lbc.numpy()
    # ---- split_and_save_weight: directory listings + bytes --------------------------------------------------------------
    def run_dir(key, vals, rng, tp, int8_outputs):
        with tempfile.TemporaryDirectory() as d:
            C.split_and_save_weight(0, Path(d), tp, key, vals, "fp16", rng,
                                    {"int8_outputs": int8_outputs, "multi_query_mode": False, "local_dim": None})
            return {f: np.fromfile(os.path.join(d, f), dtype=np.uint8) for f in sorted(os.listdir(d))}
    cases = []
    wq, rq = seeded_case(11, 32, 32, True)
    wd, rd = seeded_case(12, 32, 48)
    for tp in (1, 2):
        for io in (None, "kv_cache_only", "all"):
            cases.append((f"qkv_tp{tp}_{io}", "model.layers.0.attention.query_key_value.weight", wq, rq, tp, io))
            cases.append((f"dense_tp{tp}_{io}", "model.layers.0.attention.dense.weight", wd, rd, tp, io))
            cases.append((f"down_tp{tp}_{io}", "model.layers.0.mlp.down_proj.weight", wd, rd, tp, io))
    for io in (None, "all"):
        cases.append((f"gate_tp1_{io}", "model.layers.0.mlp.gate_proj.weight", wd, rd, 1, io))
        cases.append((f"up_tp1_{io}", "model.layers.0.mlp.up_proj.weight", wd, rd, 1, io))
    cases.append(("ln_tp1_None", "model.layers.0.input_layernorm.weight", np.arange(32, dtype=np.float16), None, 1, None))
    names = []
    for tag, key, vals, rng, tp, io in cases:
        files = run_dir(key, vals, rng, tp, io)
        names.append(tag)
        out[f"ssw_{tag}__files"] = np.array(list(files.keys()))
        for f, b in files.items():
            out[f"ssw_{tag}__{f}"] = b
    out["ssw_cases"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "convert_vectors.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "convert_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
