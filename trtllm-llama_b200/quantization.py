"""Build-time weight quantisation and layout for the B200 plugins (product code, torch only).

Mirrors the reference's build-time operators — same names, argument meaning and error behaviour:
  * ``QuantMode``                      T/tensorrt_llm/quantization/mode.py:4-137
  * ``symmetric_quantize_last_axis_of_batched_matrix`` and friends
                                       T/cpp/tensorrt_llm/thop/weightOnlyQuantOp.cpp:143-231,343-371
                                       (arithmetic: K/cutlass_kernels/cutlass_preprocessors.cpp:615-721)
  * per-channel int8 for SmoothQuant   LQ/convert.py:27-103 (generate_int8)

The *processed* layout is this library's own (the reference's is an Ampere ldmatrix interleave,
cutlass_preprocessors.cpp:537-578, meaningless on sm_100): weights are stored [N, K] with K contiguous
so each output channel is one TMA row / one coalesced 16-byte stream; int4 is packed two per byte with
the nibbles of every 8-element group interleaved for one-LOP3 extraction (``pack_processed_int4``).  The byte count equals the reference's, so the plugin's declared weight shape
(fp32 [K, N/4] or [K, N/8]) is unchanged.
"""
from __future__ import annotations

from enum import IntFlag, auto

import torch


class QuantMode(IntFlag):
    """T/tensorrt_llm/quantization/mode.py:4-137 (flags used by examples/llama_quant)."""
    INT4_WEIGHTS = auto()
    INT8_WEIGHTS = auto()
    ACTIVATIONS = auto()
    PER_CHANNEL = auto()
    PER_TOKEN = auto()
    PER_GROUP = auto()
    INT8_KV_CACHE = auto()
    FP8_KV_CACHE = auto()
    FP8_QDQ = auto()

    def _all(self, bits, mask=None):
        mask = bits if mask is None else mask
        return (self & mask) == bits

    def is_int8_weight_only(self):
        return self._all(self.INT8_WEIGHTS, self.INT8_WEIGHTS | self.ACTIVATIONS)

    def is_int4_weight_only(self):
        return self._all(self.INT4_WEIGHTS, self.INT4_WEIGHTS | self.ACTIVATIONS)

    def is_weight_only(self):
        return self.is_int4_weight_only() or self.is_int8_weight_only()

    def has_act_and_weight_quant(self):
        return self._all(self.INT8_WEIGHTS | self.ACTIVATIONS)

    def has_per_token_dynamic_scaling(self):
        return self._all(self.PER_TOKEN)

    def has_per_channel_scaling(self):
        return self._all(self.PER_CHANNEL)

    def has_act_static_scaling(self):
        return not self.has_per_token_dynamic_scaling()

    def has_int8_kv_cache(self):
        return self._all(self.INT8_KV_CACHE)

    def has_fp8_kv_cache(self):
        return self._all(self.FP8_KV_CACHE)

    def has_any_quant(self):
        return bool(self & (self.INT4_WEIGHTS | self.INT8_WEIGHTS | self.ACTIVATIONS | self.INT8_KV_CACHE))

    @staticmethod
    def from_description(quantize_weights=False, quantize_activations=False, per_token=False, per_channel=False,
                         use_int4_weights=False, use_int8_kv_cache=False):
        if quantize_activations and not quantize_weights:
            raise ValueError("We do not support activation-only quantization.")   # mode.py wording
        mode = QuantMode(0)
        if quantize_weights and use_int4_weights:
            mode |= QuantMode.INT4_WEIGHTS
        elif quantize_weights:
            mode |= QuantMode.INT8_WEIGHTS
        if quantize_activations:
            mode |= QuantMode.ACTIVATIONS
        if per_channel:
            mode |= QuantMode.PER_CHANNEL
        if per_token:
            mode |= QuantMode.PER_TOKEN
        if use_int8_kv_cache:
            mode |= QuantMode.INT8_KV_CACHE
        return mode

    @staticmethod
    def use_smooth_quant(per_token=False, per_channel=False):
        return QuantMode.from_description(True, True, per_token, per_channel)

    @staticmethod
    def use_weight_only(use_int4_weights=False):
        return QuantMode.from_description(True, False, False, False, use_int4_weights)


def _bits_of(quant_type) -> int:
    if quant_type == torch.int8:
        return 8
    if quant_type == torch.quint4x2:
        return 4
    raise ValueError("Unsupported quantization type. Must be int8 or quint4x2.")   # thop wording


def pack_int8_tensor_to_packed_int4(t: torch.Tensor) -> torch.Tensor:
    """[..., n] int8 in [-8, 7] -> [..., n/2] int8, low nibble = even index (weightOnlyQuantOp.cpp:353-356)."""
    if t.dtype != torch.int8 or t.shape[-1] % 2:
        raise ValueError("expected an int8 tensor with an even last dim")
    u = t.to(torch.uint8) & 0x0F
    return (u[..., 0::2] | (u[..., 1::2] << 4)).view(torch.int8)


def unpack_int4_packed_tensor_to_int8(t: torch.Tensor) -> torch.Tensor:
    """inverse of ``pack_int8_tensor_to_packed_int4`` (weightOnlyQuantOp.cpp:349-352)."""
    u = t.view(torch.uint8)
    lo = (u & 0x0F).to(torch.int16)
    hi = (u >> 4).to(torch.int16)
    lo = torch.where(lo > 7, lo - 16, lo)
    hi = torch.where(hi > 7, hi - 16, hi)
    out = torch.empty(u.shape[:-1] + (u.shape[-1] * 2,), dtype=torch.int8, device=t.device)
    out[..., 0::2] = lo.to(torch.int8)
    out[..., 1::2] = hi.to(torch.int8)
    return out


_I4_ORDER = (0, 2, 4, 6, 1, 3, 5, 7)     # element held by nibble position p of each 32-bit word (8 consecutive k)


def pack_processed_int4(q_nk: torch.Tensor) -> torch.Tensor:
    """[N, K] int8 in [-8, 7] -> this library's processed int4 layout [N, K/2]: every 8 consecutive k share one
    32-bit word whose nibble positions 0..7 hold elements (0,2,4,6,1,3,5,7), so that the masks 0x000F000F << 4i pull
    out the element PAIRS (2i, 2i+1) as two fp16 lanes with one LOP3 each (common.cuh i4x8_to_h2x4).  Same idea as the
    reference's register relayout for its Ampere kernels (cutlass_preprocessors.cpp:383-460)."""
    if q_nk.dtype != torch.int8 or q_nk.shape[-1] % 8:
        raise ValueError("expected an int8 tensor whose last dim is a multiple of 8")
    g = q_nk.reshape(*q_nk.shape[:-1], -1, 8)[..., list(_I4_ORDER)]
    return pack_int8_tensor_to_packed_int4(g.reshape(q_nk.shape).contiguous())


def unpack_processed_int4(p: torch.Tensor) -> torch.Tensor:
    """inverse of ``pack_processed_int4``."""
    u = unpack_int4_packed_tensor_to_int8(p)
    g = u.reshape(*u.shape[:-1], -1, 8)
    out = torch.empty_like(g)
    out[..., list(_I4_ORDER)] = g
    return out.reshape(u.shape)


def preprocess_weights_for_mixed_gemm(q_kn: torch.Tensor, quant_type) -> torch.Tensor:
    """Unprocessed quantised weights ([K, N] int8, or [K, N/2] packed int4) -> this library's processed
    layout ([N, K] int8 / [N, K/2] packed int4).  Same role and signature as
    thop/weightOnlyQuantOp.cpp:347-348 preprocess_weights_for_mixed_gemm."""
    bits = _bits_of(quant_type)
    if bits == 8:
        return q_kn.t().contiguous()
    return pack_processed_int4(unpack_int4_packed_tensor_to_int8(q_kn).t().contiguous())


def _symmetric_quantize(weight: torch.Tensor, bits: int):
    """[K, N] float -> (q int8 [K, N], scales fp16 [N]).  cutlass_preprocessors.cpp:650-701:
    scale = max_k|w| / 2^(bits-1) in fp32 (stored as fp16), q = clip(round_half_away(w / scale))."""
    if weight.dim() != 2:
        raise ValueError("Invalid dim. The dim of weight should be 2 (batched [B, K, N] is handled by the caller)")
    w = weight.to(torch.float32)
    col_max = w.abs().amax(dim=0)
    scale = col_max * (1.0 / (1 << (bits - 1)))
    t = w / scale
    r = torch.sign(t) * torch.floor(t.abs() + 0.5)
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    # all-zero column: 0/0.  The reference's clamp of NaN gives 127 (int8) / INT_MIN -> -8 (int4); the scale is 0.
    r = torch.where(torch.isnan(r), torch.full_like(r, float(hi if bits == 8 else lo)), r)
    q = r.clamp(lo, hi).to(torch.int8)
    return q, scale.to(torch.float16)


def _symmetric_quantize_last_axis_of_batched_matrix(weight: torch.Tensor, quant_type):
    """-> (unprocessed ints, processed weights, scales): weightOnlyQuantOp.cpp:357-359."""
    bits = _bits_of(quant_type)
    if weight.dim() == 3:
        parts = [_symmetric_quantize_last_axis_of_batched_matrix(w, quant_type) for w in weight]
        return tuple(torch.stack(p) for p in zip(*parts))
    q, scales = _symmetric_quantize(weight, bits)
    unprocessed = q if bits == 8 else pack_int8_tensor_to_packed_int4(q)
    q_nk = q.t().contiguous()
    processed = q_nk if bits == 8 else pack_processed_int4(q_nk)
    return unprocessed, processed, scales


def symmetric_quantize_last_axis_of_batched_matrix(weight: torch.Tensor, quant_type):
    """(processed int8 weights, fp16 scales [N]) for a [K, N] weight — the op LQ/weight_quant.py:264-271 calls."""
    _, processed, scales = _symmetric_quantize_last_axis_of_batched_matrix(weight, quant_type)
    return processed, scales


def quantize_per_channel_int8(w_nk: torch.Tensor):
    """SmoothQuant weight quantisation, per output channel: [N, K] -> (int8 [N, K], fp32 scale [N] = amax/127).
    LQ/convert.py:27-103 (scale_w_orig_quant_c = 127 / amax_c; weight.int8.col = round(w * scale).clip(-127,127))."""
    wf = w_nk.to(torch.float32)
    amax = wf.abs().amax(dim=1).clamp_min(1e-8)
    # tensor / tensor: torch evaluates `scalar / tensor` as scalar * reciprocal(tensor), which is 1 ulp off IEEE division
    c127 = torch.full_like(amax, 127.0)
    q = torch.round(wf * (c127 / amax)[:, None]).clamp(-127, 127).to(torch.int8)
    return q, (amax / c127).to(torch.float32)
