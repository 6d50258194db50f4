"""Engine build / serialisation: the part of the reference's ``Builder`` that survives without TensorRT.

Mirrors T/tensorrt_llm/builder.py:57-267 (``Builder.create_builder_config`` / ``build_engine`` / ``save_config``) and
LQ/build.py:276-444: one engine file per rank named ``llama_{dtype}_tp{N}_rank{r}.engine`` plus ``config.json`` with
the ``builder_config`` / ``plugin_config`` sections ``LQ/run.py:76-89`` reads.  The engine file holds this library's
processed tensors (a JSON index + raw bytes) instead of a TensorRT plan: the plugins are re-created from
``config.json`` by the C++ runtime at load time."""
from __future__ import annotations

import json
import os
import struct

import numpy as np
import torch

from .quantization import QuantMode
from .runtime import ModelConfig, build_engine_tensors, shard_weights

MAGIC = b"TB200ENG"
_DT = {torch.float16: "float16", torch.float32: "float32", torch.int8: "int8", torch.int32: "int32"}
_NP = {"float16": np.float16, "float32": np.float32, "int8": np.int8, "int32": np.int32}


def get_engine_name(model, dtype, tp_size, rank):
    """LQ/build.py:26-27."""
    return "{}_{}_tp{}_rank{}.engine".format(model, dtype, tp_size, rank)


def serialize_engine(tensors: dict, path: str):
    index, off = {}, 0
    for name, t in tensors.items():
        nbytes = t.numel() * t.element_size()
        index[name] = {"dtype": _DT[t.dtype], "shape": list(t.shape), "offset": off, "nbytes": nbytes}
        off += (nbytes + 255) & ~255
    hdr = json.dumps(index).encode()
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<Q", len(hdr)) + hdr)
        base = f.tell()
        for name, t in tensors.items():
            f.seek(base + index[name]["offset"])
            f.write(t.detach().cpu().contiguous().numpy().tobytes())


def deserialize_engine(path: str, device="cuda") -> dict:
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path} is not a trtllm_llama_b200 engine file")
        (n,) = struct.unpack("<Q", f.read(8))
        index = json.loads(f.read(n))
        base = f.tell()
        out = {}
        for name, e in index.items():
            f.seek(base + e["offset"])
            a = np.frombuffer(f.read(e["nbytes"]), dtype=_NP[e["dtype"]]).reshape(e["shape"])
            out[name] = torch.from_numpy(a.copy()).to(device)
    return out


def quant_mode_from_args(args) -> QuantMode:
    """LQ/build.py:276-324 flag algebra."""
    if getattr(args, "use_smooth_quant", False):
        qm = QuantMode.use_smooth_quant(args.per_token, args.per_channel)
    elif getattr(args, "use_weight_only", False):
        qm = QuantMode.use_weight_only(args.weight_only_precision == "int4")
    else:
        qm = QuantMode(0)
    if getattr(args, "int8_kv_cache", False):
        qm |= QuantMode.INT8_KV_CACHE
    return qm


def save_config(path, *, precision, world_size, mc: ModelConfig, plugin_config: dict):
    """T/tensorrt_llm/builder.py:259-267 + LQ/build.py:408-426: the keys LQ/run.py:76-89 reads."""
    cfg = {"builder_config": {"name": "llama", "precision": precision, "tensor_parallel": world_size,
                              "num_layers": mc.num_layers, "num_heads": mc.num_heads, "hidden_size": mc.hidden_size,
                              "inter_size": mc.inter_size, "vocab_size": mc.vocab_size, "hidden_act": "silu",
                              "max_position_embeddings": 2048, "max_batch_size": mc.max_batch_size,
                              "max_input_len": mc.max_input_len, "max_output_len": mc.max_output_len,
                              "int8": bool(mc.quant_mode.has_act_and_weight_quant() or mc.quant_mode.has_int8_kv_cache()),
                              "multi_query_mode": False, "quant_mode": int(mc.quant_mode), "rms_eps": mc.rms_eps},
           "plugin_config": plugin_config}
    with open(path, "w") as f:
        json.dump(cfg, f, indent=2)


def model_config_from_json(path, rank=0) -> ModelConfig:
    with open(path) as f:
        c = json.load(f)["builder_config"]
    return ModelConfig(vocab_size=c["vocab_size"], num_layers=c["num_layers"], num_heads=c["num_heads"],
                       hidden_size=c["hidden_size"], inter_size=c["inter_size"], rms_eps=c.get("rms_eps", 1e-6),
                       quant_mode=QuantMode(c.get("quant_mode", 0)), max_batch_size=c["max_batch_size"],
                       max_input_len=c["max_input_len"], max_output_len=c["max_output_len"],
                       tp_size=c["tensor_parallel"], tp_rank=rank)


def random_llama_weights(mc: ModelConfig, seed=0, device="cuda"):
    """seeded random-init fp16 weights (LQ/build.py builds with random weights when --model_dir is not given)."""
    g = torch.Generator(device=device).manual_seed(seed)
    n = lambda *s: (torch.randn(*s, generator=g, device=device, dtype=torch.float32) * 0.02).half()  # noqa: E731
    ones = lambda k: torch.ones(k, device=device, dtype=torch.float16)  # noqa: E731
    hid, inter = mc.hidden_size, mc.inter_size
    w = {"vocab_embedding": n(mc.vocab_size, hid), "ln_f": ones(hid), "lm_head": n(mc.vocab_size, hid), "layers": []}
    for _ in range(mc.num_layers):
        w["layers"].append({"input_layernorm": ones(hid), "qkv": n(3 * hid, hid), "dense": n(hid, hid),
                            "post_layernorm": ones(hid), "gate": n(inter, hid), "up": n(inter, hid), "down": n(hid, inter)})
    return w


def load_from_ft_llama(model_dir: str, mc: ModelConfig, device="cuda"):
    """FT-format fp16 checkpoint written by LQ/hf_llama_convert.py (file names of LQ/weight_quant.py:172-437, the
    unsharded ``.0.bin`` / ``.bin`` variants; [in, out] matrices are transposed to this library's [out, in])."""
    def ff(name, shape=None):
        p = os.path.join(model_dir, name)
        if not os.path.exists(p):
            raise FileNotFoundError(p)
        a = np.fromfile(p, dtype=np.float16)
        return torch.from_numpy(a.reshape(shape) if shape else a).to(device)
    hid, inter = mc.hidden_size, mc.inter_size
    w = {"vocab_embedding": ff("model.wte.weight.bin", [mc.vocab_size, hid]), "ln_f": ff("model.final_layernorm.weight.bin"),
         "lm_head": ff("model.lm_head.weight.bin", [mc.vocab_size, hid]), "layers": []}
    for i in range(mc.num_layers):
        p = f"model.model.layers.{i}."
        w["layers"].append({
            "input_layernorm": ff(p + "input_layernorm.weight.bin"),
            "qkv": ff(p + "attention.query_key_value.weight.0.bin", [hid, 3 * hid]).t().contiguous(),
            "dense": ff(p + "attention.dense.weight.0.bin", [hid, hid]).t().contiguous(),
            "post_layernorm": ff(p + "post_attention_layernorm.weight.bin"),
            "gate": ff(p + "mlp.gate_proj.weight.0.bin", [hid, inter]).t().contiguous(),
            "up": ff(p + "mlp.up_proj.weight.0.bin", [hid, inter]).t().contiguous(),
            "down": ff(p + "mlp.down_proj.weight.0.bin", [inter, hid]).t().contiguous()})
    return w


def build_rank_engine(weights, mc: ModelConfig, rank: int, kv_scale=4.0 / 127.0):
    """LQ/build.py:276-390: shard -> quantise -> processed tensors for one rank."""
    import dataclasses
    mcr = dataclasses.replace(mc, tp_rank=rank)
    return build_engine_tensors(shard_weights(weights, mc.tp_size, rank, mc.num_heads), mcr, kv_scale=kv_scale)
