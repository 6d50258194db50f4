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
        full = json.load(f)
    c, pc = full["builder_config"], full.get("plugin_config", {})
    return ModelConfig(remove_input_padding=bool(pc.get("remove_input_padding", False)),
                       paged_kv_cache=bool(pc.get("paged_kv_cache", False)), tokens_per_block=int(pc.get("tokens_per_block", 64)),
                       vocab_size=c["vocab_size"], num_layers=c["num_layers"], num_heads=c["num_heads"],
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
    """FT-format checkpoint written by ``examples/llama_quant/hf_llama_convert.py`` (file names of
    LQ/weight_quant.py:172-446; unsharded ``.bin`` / ``.0.bin`` variants; [in, out] matrices are transposed to this
    library's [out, in]).  Besides the fp16 weights it picks up, when present,

      * ``attention.query_key_value.scale_y_quant_orig.bin`` -> per-layer ``kv_scale`` (int8 KV cache,
        LQ/weight_quant.py:439-446: kv_quant_orig_scale = t, kv_orig_quant_scale = 1 / t)
      * for a SmoothQuant per-channel ModelConfig the converter's ``.weight.int8.col.0.bin`` + ``scale_w_quant_orig.col``
        (LQ/weight_quant.py:116-147,239-262) -> per-layer ``sq`` = {name: (int8 [out, in], fp32 scale [out])}, which
        ``build_engine_tensors`` uses instead of re-quantising the fp16 weights."""
    def raw(name, dtype, shape=None, required=True):
        p = os.path.join(model_dir, name)
        if not os.path.exists(p):
            if required:
                raise FileNotFoundError(p)
            return None
        a = np.fromfile(p, dtype=dtype)
        return torch.from_numpy(a.reshape(shape) if shape else a).to(device)

    def ff(name, shape=None):
        return raw(name, np.float16, shape)

    def first(names, shape):
        for n in names:
            t = raw(n, np.float16, shape, required=False)
            if t is not None:
                return t
        raise FileNotFoundError(os.path.join(model_dir, names[0]))

    def pieces(stem, rows, cols, axis):
        """a matrix the converter wrote as ``<stem>.<r>.bin`` pieces of a <tp>-gpu tree ([in, out], split along ``axis``):
        every piece present is read and they are joined again, so any tensor-parallel tree loads (sharding for the
        engine's own tp is redone by ``shard_weights``)"""
        n = 0
        while os.path.exists(os.path.join(model_dir, f"{stem}.{n}.bin")):
            n += 1
        if n == 0:
            raise FileNotFoundError(os.path.join(model_dir, f"{stem}.0.bin"))
        shape = [rows // n, cols] if axis == 0 else [rows, cols // n]
        parts = [ff(f"{stem}.{r}.bin", shape) for r in range(n)]
        return parts[0] if n == 1 else torch.cat(parts, dim=axis)

    hid, inter = mc.hidden_size, mc.inter_size
    want_sq = mc.quant_mode.has_act_and_weight_quant() and mc.quant_mode.has_per_channel_scaling()
    w = {"vocab_embedding": ff("model.wte.weight.bin", [mc.vocab_size, hid]), "ln_f": ff("model.final_layernorm.weight.bin"),
         "lm_head": ff("model.lm_head.weight.bin", [mc.vocab_size, hid]), "layers": []}
    for i in range(mc.num_layers):
        p = f"model.model.layers.{i}."
        qkv_base = p + "attention.query_key_value."
        lw = {
            "input_layernorm": ff(p + "input_layernorm.weight.bin"),
            # the converter writes QKV whole as ``.weight.bin`` ([in, 3, out/3]); older trees have ``.weight.0.bin``
            "qkv": first([qkv_base + "weight.bin", qkv_base + "weight.0.bin"], [hid, 3 * hid]).t().contiguous(),
            "dense": pieces(p + "attention.dense.weight", hid, hid, 0).t().contiguous(),
            "post_layernorm": ff(p + "post_attention_layernorm.weight.bin"),
            "gate": pieces(p + "mlp.gate_proj.weight", hid, inter, 1).t().contiguous(),
            "up": pieces(p + "mlp.up_proj.weight", hid, inter, 1).t().contiguous(),
            "down": pieces(p + "mlp.down_proj.weight", inter, hid, 0).t().contiguous()}
        kv = raw(qkv_base + "scale_y_quant_orig.bin", np.float32, required=False)
        if kv is not None:
            lw["kv_scale"] = float(kv.reshape(-1)[0])
        if want_sq:
            sq = {}

            def joined(stem, dtype, full_shape, axis):
                """``<stem>.<r>.bin`` pieces of a <tp>-gpu tree (split along ``axis`` of ``full_shape``), or the unsplit
                ``<stem>.bin``; None when neither exists"""
                n = 0
                while os.path.exists(os.path.join(model_dir, f"{stem}.{n}.bin")):
                    n += 1
                if n == 0:
                    return raw(f"{stem}.bin", dtype, list(full_shape), required=False)
                shp = list(full_shape)
                shp[axis] //= n
                parts = [raw(f"{stem}.{r}.bin", dtype, shp) for r in range(n)]
                return parts[0] if n == 1 else torch.cat(parts, dim=axis)

            # (name, file stem, [in, out] view the converter split, split axis, scale shape, scale split axis or None)
            for name, base, wshape, waxis, sshape, saxis in (
                    ("qkv", qkv_base, (hid, 3, hid), 2, (3, hid), 1),
                    ("dense", p + "attention.dense.", (hid, hid), 0, (hid,), None),
                    ("gate", p + "mlp.gate_proj.", (hid, inter), 1, (inter,), 0),
                    ("up", p + "mlp.up_proj.", (hid, inter), 1, (inter,), 0),
                    ("down", p + "mlp.down_proj.", (inter, hid), 0, (hid,), None)):
                q = joined(base + "weight.int8.col", np.int8, wshape, waxis)
                if saxis is None:
                    sc = raw(base + "scale_w_quant_orig.col.bin", np.float32, list(sshape), required=False)
                else:
                    sc = joined(base + "scale_w_quant_orig.col", np.float32, sshape, saxis)
                if q is not None and sc is not None:
                    q = q.reshape(wshape[0], -1)
                    sq[name] = (q.t().contiguous(), sc.reshape(-1).contiguous())
            if len(sq) == 5:
                lw["sq"] = sq
        w["layers"].append(lw)
    return w


def build_rank_engine(weights, mc: ModelConfig, rank: int, kv_scale=4.0 / 127.0, require_kv_scale=False):
    """LQ/build.py:276-390: shard -> quantise -> processed tensors for one rank.  ``require_kv_scale``: the weights come
    from a checkpoint, so an int8 KV cache needs its calibrated per-layer scale (no placeholder)."""
    import dataclasses
    mcr = dataclasses.replace(mc, tp_rank=rank)
    return build_engine_tensors(shard_weights(weights, mc.tp_size, rank, mc.num_heads), mcr, kv_scale=kv_scale,
                                require_kv_scale=require_kv_scale)
