#!/usr/bin/env python
"""examples/llama_quant/hf_llama_convert.py of the reference: HF LLaMA checkpoint -> FT-format directory
(``<out-dir>/<tp>-gpu/``: ``config.ini`` + raw ``model.*.bin`` files), optionally with SmoothQuant int8 weights / scales
and the int8 KV-cache scale.  SURVEY.md §8f-2; file layout and scale algebra: ``trtllm_llama_b200.ft_format``.

Flags of the reference (LQ/hf_llama_convert.py:27-98) are kept.  Differences, all deliberate:

  * calibration data: the reference downloads ``lambada`` through ``datasets``.  Offline images cannot: pass
    ``--calib-ids file.npy`` (int token ids [samples, seq]) or let the script fall back to seeded random ids (a warning is
    printed: ranges from random ids are only good for plumbing tests).
  * the reference calibrates unconditionally (twice without flags); here only when ``-sq`` or ``-kv`` ask for ranges.
  * QKV ranges: the reference's fused ``attention.query_key_value`` entry repeats q_proj's x / y / w ranges three
    times (LQ/hf_llama_convert.py:311-323), so its int8 KV scale is q_proj's output range (``--kv-range q``, kept for
    parity checks).  The default ``--kv-range kv`` takes the K and V thirds from k_proj / v_proj and leaves q_proj's
    output OUT of the y range, so ``scale_y_quant_orig`` — the int8 KV-cache scale — is max(|K|, |V|) / 127.
  * ``-sq``: the reference computes its smoothers on temporary copies, rescales the recorded ranges and writes the
    UNsmoothed weights (LQ/hf_llama_convert.py:106-227): the written int8 columns are clipped wherever the smoothed
    range is smaller than the real one and ``scale_x_orig_quant`` does not match the runtime activations.
    ``--smooth-mode reference`` reproduces exactly that bookkeeping (parity / debugging only); the default
    ``--smooth-mode folded`` writes smoothed q/k/v and gate/up weights and folds 1/s into the two norms
    (the mathematically consistent SmoothQuant for the inputs that follow a norm).
  * tensor parallel splits of gate/up are along the output axis (the reference flattens the matrix before splitting,
    which only coincides for ``-tp 1``).
"""
import argparse
import configparser
import dataclasses
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import trtllm_llama_b200  # noqa: E402,F401
from trtllm_llama_b200.calibration import capture_activation_range, smooth_gemm  # noqa: E402
from trtllm_llama_b200.ft_format import split_and_save_weight  # noqa: E402


@dataclasses.dataclass(frozen=True)
class ProgArgs:
    out_dir: str
    in_file: str
    tensor_parallelism: int = 1
    processes: int = 2
    calibrate_kv_cache: bool = False
    smoothquant: float = None
    model: str = "llama"
    storage_type: str = "fp16"
    dataset_cache_dir: str = None
    calib_ids: str = None
    calib_samples: int = 512
    calib_seq_len: int = 512
    smooth_mode: str = "folded"
    kv_range: str = "kv"
    device: str = None

    @staticmethod
    def parse(args=None) -> 'ProgArgs':
        p = argparse.ArgumentParser(formatter_class=argparse.RawTextHelpFormatter)
        p.add_argument('--out-dir', '-o', type=str, required=True, help='file name of output directory')
        p.add_argument('--in-file', '-i', type=str, required=True, help='file name of input checkpoint file')
        p.add_argument('--tensor-parallelism', '-tp', type=int, default=1, help='Requested tensor parallelism for inference')
        p.add_argument("--processes", "-p", type=int, default=2, help="accepted for CLI parity (conversion is single-process)")
        p.add_argument("--calibrate-kv-cache", "-kv", action="store_true",
                       help="Generate scaling factors for KV cache. Used for storing KV cache in int8.")
        p.add_argument("--smoothquant", "-sq", type=float, default=None,
                       help="Set the alpha parameter (https://arxiv.org/pdf/2211.10438.pdf) to Smoothquant the model, and "
                            "output int8 weights. A good first try is 0.5. Must be in [0, 1]")
        p.add_argument("--model", default="llama", type=str)
        p.add_argument("--storage-type", "-t", type=str, default="float32", choices=["float32", "float16", "bfloat16"])
        p.add_argument("--dataset-cache-dir", type=str, default=None, help="cache dir to load the hugging face dataset")
        p.add_argument("--calib-ids", type=str, default=None, help=".npy of int token ids [samples, seq] (offline calibration)")
        p.add_argument("--calib-samples", type=int, default=512)
        p.add_argument("--calib-seq-len", type=int, default=512)
        p.add_argument("--smooth-mode", choices=["reference", "folded"], default="folded",
                       help="folded (default): write smoothed weights, fold 1/s into the norms; reference: the upstream "
                            "script's bookkeeping (lossy; parity/debug only)")
        p.add_argument("--kv-range", choices=["q", "kv"], default="kv",
                       help="kv (default): int8 KV scale from the k_proj / v_proj output ranges; q: q_proj's (upstream)")
        p.add_argument("--device", type=str, default=None, help="cuda / cpu (default: cuda when available)")
        ns = p.parse_args(args)
        if ns.smoothquant is not None and not 0.0 <= ns.smoothquant <= 1.0:
            p.error("--smoothquant must be in [0, 1]")
        return ProgArgs(**vars(ns))


def _calibration_set(args: ProgArgs, vocab_size: int, in_file: str):
    """(dataset, tokenizer) for capture_activation_range"""
    if args.calib_ids:
        ids = np.load(args.calib_ids)
        return [torch.from_numpy(r.astype(np.int64)) for r in ids], None
    try:
        from datasets import load_dataset
        from transformers import LlamaTokenizer
        ds = load_dataset("lambada", split="validation", cache_dir=args.dataset_cache_dir)
        return ds, LlamaTokenizer.from_pretrained(in_file)
    except Exception as e:   # no datasets package / no network / no tokenizer files
        print(f"[WARNING] lambada is not available ({type(e).__name__}); calibrating on seeded random token ids — "
              f"the ranges are placeholders, pass --calib-ids for real ones")
        g = torch.Generator().manual_seed(0)
        n = min(args.calib_samples, 32)
        return [torch.randint(3, vocab_size, (min(args.calib_seq_len, 128),), generator=g) for _ in range(n)], None


def _fused_ranges(act_range, num_layers, kv_range):
    """entries under the FT names: attention.query_key_value (q_proj's ranges x 3, LQ/hf_llama_convert.py:311-323) and
    attention.dense (= o_proj)"""
    for l in range(num_layers):
        q = act_range[f'model.layers.{l}.self_attn.q_proj']
        parts = {"x": [q["x"]] * 3, "y": [q["y"]] * 3, "w": [q["w"]] * 3}
        if kv_range == "kv":
            k, v = act_range[f'model.layers.{l}.self_attn.k_proj'], act_range[f'model.layers.{l}.self_attn.v_proj']
            # generate_int8 reduces y with one max() over Q|K|V: keep q_proj's output out of it
            parts["y"] = [torch.zeros_like(q["y"]), k["y"], v["y"]]
        act_range[f'model.layers.{l}.attention.query_key_value'] = {n: torch.cat(p, dim=-1) for n, p in parts.items()}
        o = act_range[f'model.layers.{l}.self_attn.o_proj']
        act_range[f'model.layers.{l}.attention.dense'] = {"x": o["x"], "y": o["y"], "w": o["w"]}


@torch.no_grad()
def smooth_llama_model(model, scales, alpha, mode="reference"):
    """Per layer and per projection: smoother from the calibrated input range, recorded x range divided by it, "w" range
    replaced by the per-output-column range of the smoothed matrix (LQ/hf_llama_convert.py:106-227).  mode "reference":
    the model's weights are left untouched, as the reference does; mode "folded": q/k/v and gate/up are smoothed in the
    model and 1/s is folded into input_layernorm / post_attention_layernorm (o_proj and down_proj, which follow no norm,
    keep their weights)."""
    sd = model.state_dict()
    dev = next(model.parameters()).device
    num_layers = model.config.num_hidden_layers

    def f16(name):
        return sd[name].detach().to(torch.float16).to(dev).clone()

    for l in range(num_layers):
        pre = f'model.layers.{l}.'
        # q/k/v share the input: one smoother from the largest of the three weights, on the [out*3, in] stack
        qkv = torch.stack([f16(pre + f'self_attn.{n}_proj.weight') for n in "qkv"], dim=-1).permute(1, 2, 0)   # [in, 3, out]
        key = pre + 'attention.query_key_value'
        flat = qkv.reshape(qkv.shape[0], -1)
        # the reference passes this [in, 3*out] view where smooth_gemm expects [out, in]; it only type-checks because
        # in == out for LLaMA attention, and it makes the smoother follow the OUTPUT index.  Reproduced in "reference" mode.
        if mode == "reference":
            s = smooth_gemm(flat, scales[key]["x"], None, None, alpha)
            scales[key]["x"] = scales[key]["x"] / s
            scales[key]["w"] = qkv.abs().max(dim=0)[0]
        else:
            ws = [sd[pre + f'self_attn.{n}_proj.weight'] for n in "qkv"]
            s = smooth_gemm(ws, scales[key]["x"][:ws[0].shape[1]], sd[pre + 'input_layernorm.weight'], None, alpha)
            scales[key]["x"] = scales[key]["x"] / torch.cat([s, s, s]).to(scales[key]["x"].device)
            scales[key]["w"] = torch.stack([w.abs().amax(dim=1) for w in ws])
        for hf, ft, foldable in (("mlp.down_proj", "mlp.down_proj", False), ("mlp.gate_proj", "mlp.gate_proj", True),
                                 ("mlp.up_proj", "mlp.up_proj", True), ("self_attn.o_proj", "self_attn.o_proj", False)):
            k2 = pre + ft
            if mode == "reference" or not foldable:
                w = f16(pre + hf + '.weight')
                if mode == "reference":
                    s = smooth_gemm(w, scales[k2]["x"], None, None, alpha)
                    scales[k2]["x"] = scales[k2]["x"] / s
                scales[k2]["w"] = w.T.abs().max(dim=0)[0]
        if mode == "folded":
            gu = [sd[pre + 'mlp.gate_proj.weight'], sd[pre + 'mlp.up_proj.weight']]
            xk = scales[pre + 'mlp.gate_proj']["x"]
            s = smooth_gemm(gu, torch.maximum(xk, scales[pre + 'mlp.up_proj']["x"]), sd[pre + 'post_attention_layernorm.weight'],
                            None, alpha)
            for n, w in (("mlp.gate_proj", gu[0]), ("mlp.up_proj", gu[1])):
                scales[pre + n]["x"] = scales[pre + n]["x"] / s.to(scales[pre + n]["x"].device)
                scales[pre + n]["w"] = w.abs().amax(dim=1)
        # attention.dense mirrors o_proj after the update
        o = scales[pre + 'self_attn.o_proj']
        scales[pre + 'attention.dense'] = {"x": o["x"], "y": o["y"], "w": o["w"]}


@torch.no_grad()
def hf_llama_converter(args: ProgArgs, model=None):
    """Convert ``args.in_file`` (or an already loaded ``model``) and write the FT directory; returns its path."""
    infer_tp = args.tensor_parallelism
    multi_query_mode = False
    saved_dir = Path(args.out_dir) / f"{infer_tp}-gpu"
    saved_dir.mkdir(parents=True, exist_ok=True)
    device = args.device or ("cuda" if torch.cuda.is_available() else "cpu")
    if model is None:
        from transformers import LlamaForCausalLM
        model = LlamaForCausalLM.from_pretrained(args.in_file, torch_dtype="auto").to(device)
    num_layers = model.config.num_hidden_layers

    act_range = {}
    if args.smoothquant is not None or args.calibrate_kv_cache:
        os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
        dataset, tokenizer = _calibration_set(args, model.config.vocab_size, args.in_file)
        act_range = capture_activation_range(model, tokenizer, dataset, num_samples=args.calib_samples,
                                             seq_len=args.calib_seq_len)
        _fused_ranges(act_range, num_layers, args.kv_range)
        if args.smoothquant is not None:
            smooth_llama_model(model, act_range, args.smoothquant, args.smooth_mode)

    config = configparser.ConfigParser()
    config["llama"] = {k: f"{v}" for k, v in vars(args).items()}
    for k, v in vars(model.config).items():
        config["llama"][k] = f"{v}".replace("%", "%%")
    config["llama"]["storage_dtype"] = args.storage_type
    config["llama"]["multi_query_mode"] = str(multi_query_mode)
    with open(saved_dir / "config.ini", 'w') as f:
        config.write(f)

    int8_outputs = "all" if args.smoothquant is not None else ("kv_cache_only" if args.calibrate_kv_cache else None)
    conv_cfg = {"int8_outputs": int8_outputs, "multi_query_mode": multi_query_mode, "local_dim": None}
    sd = model.state_dict()

    def f16(name):
        return sd[name].detach().cpu().numpy().astype(np.float16)

    def emit(key, val):
        split_and_save_weight(0, saved_dir, infer_tp, key, val, args.storage_type, act_range.get(key.replace(".weight", "")),
                              conv_cfg)

    for l in range(num_layers):
        pre = f'model.layers.{l}.'
        qkv = np.stack([f16(pre + f'self_attn.{n}_proj.weight') for n in "qkv"], axis=-1)     # [out, in, 3]
        emit(pre + 'attention.query_key_value.weight', np.transpose(qkv, (1, 2, 0)))          # [in, 3, out]
        emit(pre + 'attention.dense.weight', f16(pre + 'self_attn.o_proj.weight').T)
        for n in ("down", "gate", "up"):
            emit(pre + f'mlp.{n}_proj.weight', f16(pre + f'mlp.{n}_proj.weight').T)
        emit(pre + 'input_layernorm.weight', f16(pre + 'input_layernorm.weight'))
        emit(pre + 'post_attention_layernorm.weight', f16(pre + 'post_attention_layernorm.weight'))
    for hf, ft in (('model.embed_tokens.weight', 'model.wte.weight.bin'), ('model.norm.weight', 'model.final_layernorm.weight.bin'),
                   ('lm_head.weight', 'model.lm_head.weight.bin')):
        f16(hf).tofile(saved_dir / ft)
    return saved_dir


def run_conversion(args: ProgArgs):
    print("\n=============== Arguments ===============")
    for key, value in vars(args).items():
        print(f"{key}: {value}")
    print("========================================")
    print(f"written: {hf_llama_converter(args)}")


if __name__ == "__main__":
    run_conversion(ProgArgs.parse())
