"""examples/llama_quant entry points: build.py writes the reference's artefacts (engine per rank + config.json) on
CPU; on a GPU run.py reproduces the ids the runtime API generates from the same engine file."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples", "llama_quant")
TINY = ["--n_layer", "2", "--n_embd", "256", "--n_head", "2", "--inter_size", "384", "--vocab_size", "512",
        "--max_batch_size", "2", "--max_input_len", "16", "--max_output_len", "8"]


def _build(out_dir, *extra):
    r = subprocess.run([sys.executable, os.path.join(EX, "build.py"), "--output_dir", str(out_dir), *TINY, *extra],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r


def test_build_writes_reference_artefacts(tmp_path):
    _build(tmp_path, "--use_weight_only", "--weight_only_precision", "int4", "--int8_kv_cache", "--world_size", "2",
           "--use_gpt_attention_plugin", "float16")
    # LQ/build.py:26-27,387-389: one engine per rank + config.json
    for rank in range(2):
        assert (tmp_path / f"llama_float16_tp2_rank{rank}.engine").exists()
    cfg = json.load(open(tmp_path / "config.json"))
    bc = cfg["builder_config"]
    assert bc["tensor_parallel"] == 2 and bc["num_layers"] == 2 and bc["hidden_size"] == 256 and bc["int8"] is True
    assert bc["max_batch_size"] == 2 and bc["max_input_len"] == 16 and bc["max_output_len"] == 8
    assert cfg["plugin_config"]["weight_only_quant_matmul_plugin"] == "float16" and cfg["plugin_config"]["nccl_plugin"] == "float16"
    from trtllm_llama_b200 import builder as B
    t = B.deserialize_engine(str(tmp_path / "llama_float16_tp2_rank1.engine"), device="cpu")
    assert t["layers.0.attention.qkv.weight"].shape == (3 * 128, 256 // 2) and t["layers.0.attention.qkv.weight"].dtype == torch.int8
    assert t["layers.1.mlp.proj.weight"].shape == (256, 192 // 2)           # row-parallel: K = inter / tp, int4 packed
    assert t["layers.0.mlp.fc_gate.per_channel_scale"].shape == (2 * 192,)
    assert t["lm_head.weight"].shape == (256, 256) and t["layers.0.attention.kv_quant_orig_scale"].numel() == 1
    mc = B.model_config_from_json(str(tmp_path / "config.json"), rank=1)
    assert mc.tp_rank == 1 and mc.quant_mode.is_int4_weight_only() and mc.quant_mode.has_int8_kv_cache()


def test_build_rejects_out_of_scope_flags(tmp_path):
    for bad in (["--paged_kv_cache"], ["--n_kv_head", "1"], ["--dtype", "bfloat16"], ["--use_smooth_quant"]):
        r = subprocess.run([sys.executable, os.path.join(EX, "build.py"), "--output_dir", str(tmp_path), *TINY, *bad],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode != 0


@pytest.mark.gpu
def test_run_py_matches_runtime_api(tmp_path):
    _build(tmp_path, "--use_weight_only", "--int8_kv_cache")
    ids = np.random.default_rng(3).integers(3, 512, (2, 11)).astype(np.int32)
    np.save(tmp_path / "in.npy", ids)
    r = subprocess.run([sys.executable, os.path.join(EX, "run.py"), "--max_output_len", "8", "--engine_dir", str(tmp_path),
                        "--input_tokens", str(tmp_path / "in.npy"), "--output_npy", str(tmp_path / "out.npy"),
                        "--iterations", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mean latency" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    out = np.load(tmp_path / "out.npy")
    from trtllm_llama_b200 import builder as B
    from trtllm_llama_b200 import runtime as rt
    mc = B.model_config_from_json(str(tmp_path / "config.json"))
    sess = rt.GenerationSession(mc, B.deserialize_engine(str(tmp_path / "llama_float16_tp1_rank0.engine")))
    sess.setup(2, 11, 8)
    ref = sess.decode(torch.from_numpy(ids).pin_memory(), torch.full((2,), 11, dtype=torch.int32).pin_memory()).numpy()
    assert out.shape == (2, 8) and np.array_equal(out, ref)
