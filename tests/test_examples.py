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
    for bad in (["--n_kv_head", "1"], ["--dtype", "bfloat16"], ["--use_smooth_quant"],
                ["--max_beam_width", "17"], ["--max_beam_width", "2", "--paged_kv_cache"]):
        r = subprocess.run([sys.executable, os.path.join(EX, "build.py"), "--output_dir", str(tmp_path), *TINY, *bad],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode != 0


def test_build_paged_kv_cache_flag(tmp_path):
    """--paged_kv_cache / --tokens_per_block (LQ/build.py:190-196) reach config.json and the runtime's ModelConfig."""
    _build(tmp_path, "--int8_kv_cache", "--paged_kv_cache", "--tokens_per_block", "32")
    cfg = json.load(open(tmp_path / "config.json"))
    assert cfg["plugin_config"]["paged_kv_cache"] is True and cfg["plugin_config"]["tokens_per_block"] == 32
    from trtllm_llama_b200 import builder as B
    mc = B.model_config_from_json(str(tmp_path / "config.json"))
    assert mc.paged_kv_cache and mc.tokens_per_block == 32


def test_summarize_rejects_flags_it_does_not_honour(tmp_path):
    for bad in (["--num_beams", "4", "--top_k", "8"], ["--test_hf"]):
        r = subprocess.run([sys.executable, os.path.join(EX, "summarize.py"), "--engine_dir", str(tmp_path), *bad],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode != 0 and ("does not combine" in r.stderr or "offline" in r.stderr), r.stderr[-500:]


def test_build_max_beam_width_sizes_the_engine_rows(tmp_path):
    """--max_beam_width (LQ/build.py:35): the engine holds max_batch_size x max_beam_width rows."""
    _build(tmp_path, "--int8_kv_cache", "--max_batch_size", "2", "--max_beam_width", "3")
    cfg = json.load(open(tmp_path / "config.json"))
    assert cfg["builder_config"]["max_batch_size"] == 6


@pytest.mark.gpu
def test_run_py_beam_search(tmp_path):
    """run.py --num_beams 3 (LQ/run.py:56-59): [batch, beams, output_len] ids, best beam first, equal to the runtime API."""
    _build(tmp_path, "--use_weight_only", "--int8_kv_cache", "--max_batch_size", "2", "--max_beam_width", "3")
    ids = np.random.default_rng(4).integers(3, 512, (2, 9)).astype(np.int32)
    np.save(tmp_path / "in.npy", ids)
    r = subprocess.run([sys.executable, os.path.join(EX, "run.py"), "--max_output_len", "8", "--engine_dir", str(tmp_path),
                        "--input_tokens", str(tmp_path / "in.npy"), "--output_npy", str(tmp_path / "out.npy"), "--num_beams", "3",
                        "--iterations", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mean latency" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    out = np.load(tmp_path / "out.npy")
    from trtllm_llama_b200 import builder as B
    from trtllm_llama_b200 import runtime as rt
    mc = B.model_config_from_json(str(tmp_path / "config.json"))
    sess = rt.GenerationSession(mc, B.deserialize_engine(str(tmp_path / "llama_float16_tp1_rank0.engine")))
    sess.setup(2, 9, 8, beam_width=3)
    sc = rt.SamplingConfig(end_id=2, pad_id=2, num_beams=3)
    ref = sess.decode(torch.from_numpy(ids).pin_memory(), torch.full((2,), 9, dtype=torch.int32).pin_memory(), sc).numpy()
    assert out.shape == (2, 3, 8) and np.array_equal(out, ref)


@pytest.mark.gpu
def test_summarize_py_runs_and_checks_agreement(tmp_path):
    """summarize.py on a tiny engine: greedy run with the token-agreement check against the oracle, and a sampled run
    (--top_k 8) that is reproducible for a seed."""
    from oracle import ref_model as RM
    _build(tmp_path, "--use_weight_only", "--int8_kv_cache", "--random_seed", "5", "--max_batch_size", "2")
    cfg = RM.LlamaCfg.tiny(layers=2, hidden=256, inter=384, vocab=512)
    from trtllm_llama_b200 import builder as B
    from trtllm_llama_b200.runtime import ModelConfig
    mc = ModelConfig(vocab_size=512, num_layers=2, num_heads=2, hidden_size=256, inter_size=384)
    w = B.random_llama_weights(mc, seed=5, device="cuda")        # build.py draws on the GPU's generator when there is one
    wn = {k: w[k].cpu().numpy() for k in ("vocab_embedding", "ln_f", "lm_head")}
    wn["layers"] = [{k: v.cpu().numpy() for k, v in lw.items()} for lw in w["layers"]]
    np.save(tmp_path / "w.npy", wn, allow_pickle=True)
    base = [sys.executable, os.path.join(EX, "summarize.py"), "--engine_dir", str(tmp_path), "--batch_size", "2", "--max_ite",
            "2", "--max_input_len", "16", "--output_len", "8"]
    r = subprocess.run(base + ["--check_accuracy", "--oracle_weights", str(tmp_path / "w.npy"), "--agreement_threshold", "80"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "token agreement" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    a = subprocess.run(base + ["--top_k", "8", "--random_seed", "3"], capture_output=True, text=True, timeout=600)
    assert a.returncode == 0 and "tokens/s" in a.stdout, a.stdout[-2000:] + a.stderr[-2000:]


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
