"""HF -> FT-format checkpoint pieces: int8 scale algebra and the on-disk layout.

Keeps the interface of the reference's ``examples/llama_quant/convert.py`` (function names, arguments, dictionary keys,
file names) so a converter script and ``weight_quant.py``-style loaders read the same:

  * ``generate_int8``          LQ/convert.py:27-103   quantised weights + the four families of scaling factors
  * ``write_int8``             LQ/convert.py:106-146  which of them are written once and which per rank
  * ``split_and_save_weight``  LQ/convert.py:160-325  which tensor is split along which axis, file names
  * ``split / save_val / save_split / str_to_np_dtype``  LQ/convert.py:9-25,149-158

Everything is numpy on the host (conversion is offline, not on the timed path).  ``act_range`` values may be numpy arrays
or torch tensors.  Files are ``model.<key>.bin`` (unsplit) or ``model.<key>.<rank>.bin`` raw little-endian arrays.
"""
from pathlib import Path

import numpy as np

F32 = np.float32


def _np(a):
    """torch tensor or array-like -> numpy (no copy for arrays)"""
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a)


def split(v, tp_size, idx, dim=0):
    """Rank ``idx``'s share of ``v`` (LQ/convert.py:9-15): vectors are cut in ``tp_size`` equal pieces, matrices along ``dim``."""
    if tp_size == 1:
        return v
    if v.ndim == 1:
        return np.ascontiguousarray(np.split(v, tp_size)[idx])
    return np.ascontiguousarray(np.split(v, tp_size, axis=dim)[idx])


def save_val(val, dir, key, tp_num=None):
    """``model.<key>.bin`` or ``model.<key>.<tp_num>.bin`` (LQ/convert.py:17-19)."""
    suffix = "bin" if tp_num is None else f"{tp_num}.bin"
    np.ascontiguousarray(val).tofile(Path(dir) / f"model.{key}.{suffix}")


def save_split(split_vals, dir, key, i, factor):
    """Pieces ``j`` of rank ``i`` get the global index ``i * factor + j`` (LQ/convert.py:22-24)."""
    for j, val in enumerate(split_vals):
        save_val(val, dir, key, i * factor + j)


def generate_int8(weights, act_range, is_qkv=False, multi_query_mode=False):
    """Int8 weights (per-tensor and per-column) and scaling factors of one linear layer — LQ/convert.py:27-103.

    ``weights`` [in, out] (QKV: [in, 3, out/3]); ``act_range`` = {"x": |input| max per channel, "y": |output| max per
    channel, "w": |weight| max per output column}.  Returned keys (all fp32 except the int8 weights):

      weight.int8 / weight.int8.col        round(w * 127 / max|w|) clipped to [-127, 127], max per tensor / per column
      scale_x_orig_quant                   127 / max|x|                      (fp activation -> int8, per tensor)
      scale_w_quant_orig[.col]             max|w| / 127                      (int8 weight -> fp)
      scale_y_accum_quant[.col]            (127 / max|y|) / (scale_x_orig_quant * 127 / max|w|)   (int32 accum -> int8)
      scale_y_quant_orig                   max|y| / 127                      (int8 output -> fp; the int8 KV-cache scale)

    For QKV "per tensor" means one factor for each of Q, K and V (shape [3, 1]), and the per-tensor factors are
    broadcast to the per-column shape [3, out/3].
    """
    if is_qkv and multi_query_mode:
        raise ValueError("Multi-query w/ int8 quant has not been supported yet")
    # dtype flow as in the reference: the ranges keep the dtype they were measured in (fp16 ranges give fp16 factors),
    # Python-float constants do not widen them
    w_rng = _np(act_range["w"])
    weights = _np(weights)
    if is_qkv:
        w3 = w_rng.reshape(3, -1)
        scale_w_orig_quant_t = 127. / w3.max(axis=-1, keepdims=True)
        scale_w_orig_quant_c = 127. / w3
    else:
        scale_w_orig_quant_t = 127. / np.asarray(w_rng.max())
        scale_w_orig_quant_c = 127. / w_rng
    scale_w_quant_orig_t = 1.0 / scale_w_orig_quant_t
    scale_w_quant_orig_c = 1.0 / scale_w_orig_quant_c

    x_max = float(_np(act_range["x"]).max())
    y_max = float(_np(act_range["y"]).max())
    scale_x_orig_quant_t = np.array(127. / x_max)
    scale_y_orig_quant_t = np.array(127. / y_max)
    scale_y_quant_orig_t = np.array(y_max / 127.)
    scale_y_accum_quant_t = scale_y_orig_quant_t / (scale_x_orig_quant_t * scale_w_orig_quant_t)
    scale_y_accum_quant_c = scale_y_orig_quant_t / (scale_x_orig_quant_t * scale_w_orig_quant_c)
    if is_qkv:
        scale_y_accum_quant_t = np.broadcast_to(scale_y_accum_quant_t, scale_w_orig_quant_c.shape)
        scale_w_quant_orig_t = np.broadcast_to(scale_w_quant_orig_t, scale_w_orig_quant_c.shape)

    def to_i8(x):
        return x.round().clip(-127, 127).astype(np.int8)

    return {
        "weight.int8": to_i8(weights * scale_w_orig_quant_t),
        "weight.int8.col": to_i8(weights * scale_w_orig_quant_c),
        "scale_x_orig_quant": scale_x_orig_quant_t.astype(F32),
        "scale_w_quant_orig": np.asarray(scale_w_quant_orig_t).astype(F32),
        "scale_w_quant_orig.col": scale_w_quant_orig_c.astype(F32),
        "scale_y_accum_quant": np.asarray(scale_y_accum_quant_t).astype(F32),
        "scale_y_accum_quant.col": scale_y_accum_quant_c.astype(F32),
        "scale_y_quant_orig": scale_y_quant_orig_t.astype(F32),
    }


def write_int8(vals, dir, base_key, split_dim, tp_rank, split_factor, kv_cache_only=False):
    """Write what ``generate_int8`` produced — LQ/convert.py:106-146.  Int8 weights are split along ``split_dim`` per rank;
    per-column factors are per rank only for column-parallel layers (``split_dim == -1``: QKV, gate, up); per-tensor
    factors are written once by rank 0; with ``kv_cache_only`` only ``scale_y_quant_orig`` is written."""
    if not kv_cache_only:
        save_split(np.split(vals["weight.int8"], split_factor, axis=split_dim), dir, f"{base_key}.weight.int8", tp_rank,
                   split_factor)
        save_split(np.split(vals["weight.int8.col"], split_factor, axis=split_dim), dir, f"{base_key}.weight.int8.col",
                   tp_rank, split_factor)
    saved_keys_once = ["scale_y_quant_orig"]
    if not kv_cache_only:
        saved_keys_once += ["scale_x_orig_quant", "scale_w_quant_orig", "scale_y_accum_quant"]
        if split_dim == -1:
            for k in ("scale_w_quant_orig.col", "scale_y_accum_quant.col"):
                save_split(np.split(vals[k], split_factor, axis=split_dim), dir, f"{base_key}.{k}", tp_rank, split_factor)
        else:
            saved_keys_once += ["scale_w_quant_orig.col", "scale_y_accum_quant.col"]
    if tp_rank == 0:
        for k in saved_keys_once:
            save_val(vals[k], dir, f"{base_key}.{k}")


def str_to_np_dtype(type_str):
    try:
        return {"fp32": np.float32, "fp16": np.float16}[type_str]
    except KeyError:
        raise ValueError(f"{type_str} is an invalid storage type")


_REPLICATED = ("input_layernorm.weight", "input_layernorm.bias", "attention.dense.bias", "post_attention_layernorm.weight",
               "post_attention_layernorm.bias", "mlp.gate_proj.bias", "mlp.up_proj.bias", "mlp.down_proj.bias",
               "final_layernorm.weight", "final_layernorm.bias")


def split_and_save_weight(tp_rank, saved_dir, split_factor, key, vals, storage_type, act_range, config):
    """One tensor of the checkpoint -> its FT files — LQ/convert.py:160-325.

      norm weights / biases            written once (rank 0), unsplit
      attention.dense, mlp.down_proj   [in, out] split along the input axis (row parallel)         -> ``.<rank>.bin``
      mlp.gate_proj, mlp.up_proj       [in, out] split along the output axis (column parallel)      -> ``.<rank>.bin``
      attention.query_key_value        [in, 3, out/3] written whole as ``.weight.bin`` (the loader shards it per head group)

    plus, when ``config["int8_outputs"]`` is ``"all"`` (SmoothQuant) the int8 variants and scales of every matrix, or when
    it is ``"kv_cache_only"`` just the QKV output scale used for the int8 KV cache.  ``storage_type`` is accepted for
    interface parity (the reference stores what it is given)."""
    del storage_type
    int8_outputs = config.get("int8_outputs", None)
    multi_query_mode = config.get("multi_query_mode", False)
    save_int8 = int8_outputs in ("all", "kv_cache_only")
    vals = _np(vals)

    if any(k in key for k in _REPLICATED):
        if tp_rank == 0:
            save_val(vals, saved_dir, key)
    elif "attention.dense.weight" in key or "mlp.down_proj.weight" in key:
        save_split(np.split(vals, split_factor, axis=0), saved_dir, key, tp_rank, split_factor)
        if act_range is not None and int8_outputs == "all":
            write_int8(generate_int8(vals, act_range, multi_query_mode=multi_query_mode), saved_dir,
                       key.replace(".weight", ""), 0, tp_rank, split_factor)
    elif "mlp.gate_proj.weight" in key or "mlp.up_proj.weight" in key:
        save_split(np.split(vals, split_factor, axis=-1), saved_dir, key, tp_rank, split_factor)
        if act_range is not None and int8_outputs == "all":
            write_int8(generate_int8(vals, act_range, multi_query_mode=multi_query_mode), saved_dir,
                       key.replace(".weight", ""), -1, tp_rank, split_factor)
    elif "attention.query_key_value.weight" in key:
        save_val(vals, saved_dir, key, tp_num=None)
        if save_int8:
            write_int8(generate_int8(vals, act_range, is_qkv=True, multi_query_mode=False), saved_dir,
                       key.replace(".weight", ""), -1, tp_rank, split_factor, kv_cache_only=int8_outputs == "kv_cache_only")
    else:
        print(f"[WARNING] {key} not handled by converter")
