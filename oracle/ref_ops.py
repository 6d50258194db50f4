"""CPU oracle (numpy) for the LLaMA decoder hot path of TRT2022/trtllm-llama.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``trtllm-llama_b200/``,
``examples/``) may import this module: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do, and only as
the checker.

Every function restates the arithmetic of one reference kernel, citing the
file:line it follows.  Path aliases (same as SURVEY.md):
  K/ = tensorrt_llm_july-release-v1/cpp/tensorrt_llm/kernels/
  P/ = tensorrt_llm_july-release-v1/cpp/tensorrt_llm/plugins/
  CE/ = .../cutlass_extensions/include/cutlass_extensions/
  T/ = tensorrt_llm_july-release-v1/

Parity pinning (see DESIGN.md "Oracle"):
  * symmetric_quantize / int4 packing: pinned bit-exactly against the reference's
    own ``cutlass_preprocessors.cpp`` compiled by ``oracle/Makefile`` into
    ``oracle/_ref/libref_host.so`` (tests/test_cpu_oracle_and_abi.py, live and through tests/golden/).
  * sq_gemm / quantize_per_token / rmsnorm-quant / weight-only / attention:
    pinned against the reference tests' procedural fixtures restated in
    ``tests/golden/make_golden.py`` (input distributions, seeds, tolerances of
    T/tests/quantization/*.py and T/tests/attention/test_gpt_attention.py) and,
    on the GPU box, against the reference's own CUDA kernels recompiled for
    sm_100a (``oracle/_ref/libref_cuda.so``, tests/test_ref_cuda_parity.py).
"""
from __future__ import annotations

import numpy as np

F16 = np.float16
F32 = np.float32


# --------------------------------------------------------------------------- #
# rounding primitives
# --------------------------------------------------------------------------- #
def cvt_rni_sat_s8(x):
    """``cvt.rni.sat.s8.f32``: round-to-nearest-even, saturate to [-128, 127].

    T/cpp/tensorrt_llm/common/cudaTypeUtils.cuh:327-371 (cuda_cast<int8_t>(float)),
    K/decoderMaskedMultiheadAttentionUtils.h:2276-2286 (cast_to_int8).
    NaN converts to 0 (PTX cvt.sat semantics)."""
    x = np.asarray(x, dtype=F32)
    r = np.rint(x)  # numpy rint == round half to even
    r = np.where(np.isnan(r), 0.0, r)
    return np.clip(r, -128, 127).astype(np.int8)


def f16(x):
    return np.asarray(x, dtype=F32).astype(F16)


def f16r(x):
    """round an fp32 value through fp16 and return fp32."""
    return np.asarray(x, dtype=F32).astype(F16).astype(F32)


# --------------------------------------------------------------------------- #
# a7: weight-only symmetric quantiser  (K/cutlass_kernels/cutlass_preprocessors.cpp:615-721)
# --------------------------------------------------------------------------- #
def symmetric_quantize(weight, bits):
    """Per-output-column symmetric quantisation of a [K, N] weight matrix.

    Returns (q int8 [K, N] unpacked values, scales float16 [N]).
    cutlass_preprocessors.cpp:650-668: per_col_max = max_k |w| ; scale = per_col_max / 2^(bits-1)
    (fp32; the *fp32* value divides the weights, the stored scale is its fp16 cast);
    :682-701: q = clip(round_half_away(w / scale), -2^(bits-1), 2^(bits-1)-1).
    """
    w = np.asarray(weight, dtype=F32)
    assert w.ndim == 2 and bits in (4, 8)
    col_max = np.abs(w).max(axis=0).astype(F32)
    scale32 = (col_max * F32(1.0 / (1 << (bits - 1)))).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (w / scale32[None, :]).astype(F32)
    # C `round()` : half away from zero
    r = np.sign(t) * np.floor(np.abs(t) + F32(0.5))
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    # all-zero column: 0/0 = NaN.  The reference's std::max(-128.f, std::min(127.f, NaN)) yields 127
    # (:685-687) and its int(NaN) yields INT_MIN -> -8 for int4 (:699-701).  The scale is 0, so the
    # dequantised weight is 0 either way; followed here only to stay bit-identical with libref_host.
    r = np.where(np.isnan(r), float(hi if bits == 8 else lo), r)
    q = np.clip(r, lo, hi).astype(np.int8)
    return q, scale32.astype(F16)


def pack_int4(q):
    """[K, N] int8 values in [-8, 7] -> [K, N/2] int8, low nibble = even column.
    cutlass_preprocessors.cpp:684-704 ("(clipped & 0x0F) << (4 * packed_idx)")."""
    q = np.asarray(q, dtype=np.int8)
    assert q.shape[-1] % 2 == 0
    lo = q[..., 0::2].astype(np.uint8) & 0x0F
    hi = q[..., 1::2].astype(np.uint8) & 0x0F
    return (lo | (hi << 4)).astype(np.uint8).view(np.int8)


def unpack_int4(p):
    """inverse of pack_int4 (thop/weightOnlyQuantOp.cpp unpack_int4_packed_tensor_to_int8)."""
    u = np.asarray(p).view(np.uint8)
    lo = (u & 0x0F).astype(np.int8)
    hi = (u >> 4).astype(np.int8)
    lo = np.where(lo > 7, lo - 16, lo).astype(np.int8)
    hi = np.where(hi > 7, hi - 16, hi).astype(np.int8)
    out = np.empty(u.shape[:-1] + (u.shape[-1] * 2,), dtype=np.int8)
    out[..., 0::2] = lo
    out[..., 1::2] = hi
    return out


# --------------------------------------------------------------------------- #
# a6: weight-only matmul  (k10 / k11)
# --------------------------------------------------------------------------- #
def weight_only_dequant(q, scales):
    """fp16(fp16(q) * scale[n])  --  CE/gemm/warp/mma_tensorop_dequantizer.h:255-271 (half2 mul of the
    converted weight fragment by the fp16 scale fragment)."""
    return (np.asarray(q).astype(F16) * np.asarray(scales, dtype=F16)[None, :]).astype(F16)


def weight_only_matmul(act, q, scales):
    """C[m,n] = fp16( sum_k fp32(A[m,k]) * fp32(dequant(W)[k,n]) )   (fp32 accumulate, k10).

    act [M, K] fp16, q [K, N] int8 (unprocessed ints), scales [N] fp16.
    K/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:49-175 (ElementAccumulator = float).
    The M==1 GEMV (K/weightOnlyMatrixVectorMultiplication.cu:136-205) additionally rounds each
    product to fp16 before the fp32 add; both fall inside the reference test's column tolerance
    (T/tests/quantization/test_weight_only_quant_matmul.py:85-122)."""
    w = weight_only_dequant(q, scales).astype(F32)
    a = np.asarray(act, dtype=F16).astype(F32)
    return (a.astype(np.float64) @ w.astype(np.float64)).astype(F32).astype(F16)


# --------------------------------------------------------------------------- #
# a5: SmoothQuant int8 GEMM (k9)
# --------------------------------------------------------------------------- #
def sq_gemm(a_i8, b_i8, scale_tokens, scale_channels, out_dtype=F16):
    """C[m,n] = T( float(sum_k a[m,k]*b[n,k]) * (sc[n] * sr[m]) ).

    a [M,K] int8, b [N,K] int8; scale_tokens [M] or [1]; scale_channels [N] or [1] (fp32).
    CE/epilogue/threadblock/epilogue_per_row_per_col_scale.h:279-349 — note the grouping
    ``accum * (scale_col * scale_row)`` (:325, :341), all in fp32, then a round-to-nearest convert.
    int32 output: float->int32 round-to-nearest (NumericArrayConverter default)."""
    a = np.asarray(a_i8, dtype=np.int32)
    b = np.asarray(b_i8, dtype=np.int32)
    return sq_gemm_epilogue((a @ b.T).astype(np.int32), scale_tokens, scale_channels, out_dtype)


def sq_gemm_epilogue(acc_i32, scale_tokens, scale_channels, out_dtype=F16):
    """The epilogue of ``sq_gemm`` on given exact int32 accumulators [M,N] (epilogue_per_row_per_col_scale.h:279-349);
    lets a test feed accumulators computed elsewhere (exactly) for shapes where an integer matmul on the CPU is too slow."""
    acc = np.asarray(acc_i32, dtype=np.int32)
    sr = np.asarray(scale_tokens, dtype=F32).reshape(-1)
    sc = np.asarray(scale_channels, dtype=F32).reshape(-1)
    sr = np.broadcast_to(sr, (acc.shape[0],)) if sr.size == 1 else sr
    sc = np.broadcast_to(sc, (acc.shape[1],)) if sc.size == 1 else sc
    s = (sc[None, :] * sr[:, None]).astype(F32)
    res = (acc.astype(F32) * s).astype(F32)
    if out_dtype == np.int32:
        return np.rint(res).astype(np.int32)
    return res.astype(out_dtype)


# --------------------------------------------------------------------------- #
# a10: quantisers (k13)
# --------------------------------------------------------------------------- #
def quantize_per_token(x):
    """K/quantization.cu:93-117.  x [M, K] fp16 (or fp32).
    amax = max(1e-6 (cast to T), |x|) in T; scale_out = amax/127; q = cvt.rni.sat(float(x) * (127/amax))."""
    x = np.asarray(x)
    T = x.dtype.type
    amax = np.maximum(np.abs(x).max(axis=-1), T(1e-6)).astype(F32)
    scale_out = (amax / F32(127.0)).astype(F32)
    s = (F32(127.0) / amax).astype(F32)
    q = cvt_rni_sat_s8(x.astype(F32) * s[..., None])
    return q, scale_out[..., None]


def quantize_tensor(x, scale_orig_quant):
    """K/quantization.cu:31-65: q = cvt.rni.sat(float(x) * scale)."""
    return cvt_rni_sat_s8(np.asarray(x).astype(F32) * F32(scale_orig_quant))


# --------------------------------------------------------------------------- #
# a9/a12: RMSNorm (+ quant)   (new plugin modelled on LayernormQuantization)
# --------------------------------------------------------------------------- #
def rmsnorm(x, gamma, eps=1e-6):
    """y = fp16( x * rsqrt(mean(x^2) + eps) * gamma ), body in fp32.
    T/tensorrt_llm/functional.py:3195-3219 (rms_norm: pow/mean/+eps/sqrt/div in fp32 then *weight),
    kernel structure of K/layernormKernels.cu:60-194 with the mean term removed."""
    xf = np.asarray(x, dtype=F16).astype(F32)
    var = (xf.astype(np.float64) ** 2).mean(axis=-1, keepdims=True).astype(F32)
    inv = (F32(1.0) / np.sqrt(var + F32(eps))).astype(F32)
    y = (xf * inv).astype(F32) * np.asarray(gamma, dtype=F16).astype(F32)
    return y.astype(F16)


def rmsnorm_quant(x, gamma, eps=1e-6, scale_orig_quant=None, dynamic=True):
    """RMSNorm followed by int8 quantisation, K/layernormKernels.cu:141-193 semantics:
      static : q = cvt.rni.sat(float(y_fp16) * scale)                       (:162-166)
      dynamic: amax = max(|y_fp16|, 1e-6) (fp16) ; q = cvt.rni.sat(float(y_fp16) * (127/amax));
               scale_out = amax/127                                       (:173-193, smem path)
    """
    y = rmsnorm(x, gamma, eps)
    if not dynamic:
        return cvt_rni_sat_s8(y.astype(F32) * F32(scale_orig_quant)), None
    amax = np.maximum(np.abs(y).max(axis=-1), F16(1e-6)).astype(F32)
    s = (F32(127.0) / amax).astype(F32)
    q = cvt_rni_sat_s8(y.astype(F32) * s[..., None])
    return q, (amax / F32(127.0)).astype(F32)[..., None]


def layernorm_quant(x, gamma, beta, eps=1e-5, scale_orig_quant=None, dynamic=True):
    """The reference's LayernormQuantization proper (K/layernormKernels.cu:60-194), kept so the
    oracle can be pinned on T/tests/quantization/test_smooth_quant_layer_norm.py."""
    xf = np.asarray(x, dtype=F16).astype(F32)
    mean = xf.astype(np.float64).mean(axis=-1, keepdims=True).astype(F32)
    var = ((xf - mean).astype(np.float64) ** 2).mean(axis=-1, keepdims=True).astype(F32)
    inv = (F32(1.0) / np.sqrt(var + F32(eps))).astype(F32)
    y = ((xf - mean) * inv * np.asarray(gamma, F16).astype(F32) + np.asarray(beta, F16).astype(F32)).astype(F16)
    if not dynamic:
        return cvt_rni_sat_s8(y.astype(F32) * F32(scale_orig_quant)), None
    amax = np.maximum(np.abs(y).max(axis=-1), F16(1e-6)).astype(F32)
    q = cvt_rni_sat_s8(y.astype(F32) * (F32(127.0) / amax)[..., None])
    return q, (amax / F32(127.0)).astype(F32)[..., None]


# --------------------------------------------------------------------------- #
# RoPE (neox) and int8 KV helpers
# --------------------------------------------------------------------------- #
def rope_neox(x, pos, rot_dim=None, base=10000.0):
    """x [..., Dh] fp16, pos broadcastable to x.shape[:-1].  Pairs (j, j + rot/2); fp32 math,
    rounded back to fp16.  K/decoderMaskedMultiheadAttentionUtils.h:1511-1531
    (inv_freq = t / pow(10000, 2j/rot); x' = c*x - s*y ; y' = c*y + s*x), neox pairing
    K/decoderMaskedMultiheadAttention/decoderMaskedMultiheadAttentionTemplate.h:1431-1476."""
    x = np.asarray(x, dtype=F16)
    dh = x.shape[-1]
    rot = dh if rot_dim is None else rot_dim
    half = rot // 2
    j = np.arange(half, dtype=F32)
    denom = np.power(F32(base), (2.0 * j / F32(rot)).astype(F32)).astype(F32)
    t = np.asarray(pos, dtype=F32)[..., None]
    ang = (t / denom).astype(F32)
    c, s = np.cos(ang).astype(F32), np.sin(ang).astype(F32)
    a = x[..., :half].astype(F32)
    b = x[..., half:rot].astype(F32)
    out = x.copy()
    out[..., :half] = (c * a - s * b).astype(F16)
    out[..., half:rot] = (c * b + s * a).astype(F16)
    return out


def kv_quant(x, scale_orig_quant):
    """store_8bits_kv_cache_vec, K/decoderMaskedMultiheadAttentionUtils.h:2383-2390:
    cvt.rni.sat.s8(float(x) * scale)."""
    return cvt_rni_sat_s8(np.asarray(x, dtype=F16).astype(F32) * F32(scale_orig_quant))


def kv_dequant(q, scale_quant_orig):
    """load_8bits_kv_cache_vec, K/decoderMaskedMultiheadAttentionUtils.h:2358-2365:
    fp16(float(int8) * scale)."""
    return (np.asarray(q, dtype=np.int8).astype(F32) * F32(scale_quant_orig)).astype(F16)


# --------------------------------------------------------------------------- #
# a1/a2: decode-step masked multi-head attention (k1)
# --------------------------------------------------------------------------- #
def mmha_decode(qkv, kv_cache, past_len, input_lengths, max_input_len, *, num_heads, head_size,
                q_scaling=1.0, rotary_dim=None, kv_scale_orig_quant=None, kv_scale_quant_orig=None):
    """One generation step for every sequence of a padded batch.

    qkv       [B, 3*H*Dh] fp16   (q | k | v, P/gptAttentionCommon/gptAttentionCommon.cpp:134-146)
    kv_cache  [B, 2, H, S_max, Dh] int8 or fp16 — UPDATED IN PLACE at position ``past_len``
              (K/kvCacheUtils.h:114-170 KVLinearBuffer)
    past_len  int: timestep (= sequence_length of every sample in a padded batch,
              gptAttentionCommon.cpp:157; T/tensorrt_llm/runtime/generation.py:686-689)
    input_lengths [B] : real prompt lengths; positions [input_lengths[b], max_input_len) are padding
              -> masked_tokens (generation.py) and total_padding_tokens (K/gptKernels.cu:239-253).
    Returns out [B, H*Dh] fp16.

    Arithmetic (decoderMaskedMultiheadAttentionTemplate.h):
      :1425-1476 RoPE at position past_len - pad on q,k (fp16 result)
      :1493-1509 K[t] <- k (int8: cvt.rni.sat(k*s))   :1913-1926 V[t] <- v
      :1511-1549 qk_t = dot(q,k) * inv_sqrt_dh using the *unquantised* current k
      :1601-1682 qk_i = dot(q, dequant(K_i)) * inv_sqrt_dh, masked tokens excluded from the max
      :1719-1779 p_i = exp(qk_i - max) (0 if masked); p_i *= 1/(sum + 1e-6); p rounded to fp16 (:1765)
      :1829-1950 out = sum_i p_i * dequant(V_i) + p_t * v   (fp32 accumulate)
      :1985-2017 fp16 store.
    """
    qkv = np.asarray(qkv, dtype=F16)
    B = qkv.shape[0]
    H, Dh = num_heads, head_size
    hid = H * Dh
    int8_kv = kv_cache.dtype == np.int8
    inv_sqrt_dh = F32(1.0) / (np.sqrt(F32(Dh)) * F32(q_scaling))
    out = np.zeros((B, hid), dtype=F16)
    t = int(past_len)
    for b in range(B):
        pad = int(max_input_len) - int(input_lengths[b])
        pos = t - pad
        q = qkv[b, 0:hid].reshape(H, Dh)
        k = qkv[b, hid:2 * hid].reshape(H, Dh)
        v = qkv[b, 2 * hid:3 * hid].reshape(H, Dh)
        q = rope_neox(q, np.full((H,), pos), rotary_dim)
        k = rope_neox(k, np.full((H,), pos), rotary_dim)
        if int8_kv:
            kv_cache[b, 0, :, t, :] = kv_quant(k, kv_scale_orig_quant)
            kv_cache[b, 1, :, t, :] = kv_quant(v, kv_scale_orig_quant)
            Kc = kv_dequant(kv_cache[b, 0, :, :t, :], kv_scale_quant_orig)
            Vc = kv_dequant(kv_cache[b, 1, :, :t, :], kv_scale_quant_orig)
        else:
            kv_cache[b, 0, :, t, :] = k
            kv_cache[b, 1, :, t, :] = v
            Kc = kv_cache[b, 0, :, :t, :]
            Vc = kv_cache[b, 1, :, :t, :]
        masked = np.zeros(t + 1, dtype=bool)
        masked[int(input_lengths[b]):int(max_input_len)] = True
        qf = q.astype(F32)
        s = np.empty((H, t + 1), dtype=F32)
        s[:, :t] = np.einsum("hd,htd->ht", qf.astype(np.float64), Kc.astype(np.float64)).astype(F32) * inv_sqrt_dh
        s[:, t] = (qf.astype(np.float64) * k.astype(np.float64)).sum(-1).astype(F32) * inv_sqrt_dh
        smax = np.where(masked[None, :], -np.inf, s).max(axis=1, keepdims=True)
        p = np.where(masked[None, :], 0.0, np.exp((s - smax).astype(F32))).astype(F32)
        inv_sum = F32(1.0) / (p.sum(axis=1, keepdims=True, dtype=F32) + F32(1e-6))
        p16 = (p * inv_sum).astype(F16).astype(np.float64)
        o = np.einsum("ht,htd->hd", p16[:, :t], Vc.astype(np.float64)) + p16[:, t:t + 1] * v.astype(np.float64)
        out[b] = o.astype(F32).astype(F16).reshape(hid)
    return out


# --------------------------------------------------------------------------- #
# a3: context (prefill) attention — unfused reference path k2..k7
# --------------------------------------------------------------------------- #
def context_attention(qkv, kv_cache, input_lengths, *, num_heads, head_size, q_scaling=1.0,
                      rotary_dim=None, kv_scale_orig_quant=None):
    """qkv [B, S, 3*H*Dh] fp16 (padded to S = max_input_len); kv_cache [B,2,H,S_max,Dh] written for
    positions [0, S) (padding rows hold zeros, K/unfusedAttentionKernels.cu:1252-1424 zeroes them).
    Returns out [B, S, H*Dh] fp16; rows at padded positions are unspecified (the reference leaves a
    uniform-softmax artefact there, nothing reads them) and are returned as zeros.

      k2  RoPE position = index in sequence (unfusedAttentionKernels.cu:1252-1424)
      k3  cache write, int8: cvt.rni.sat(x * kvScaleOrigQuant[0]) (:1553-1646)
      k7  QK^T fp16 x fp16 -> fp32 (P/gptAttentionCommon/gptAttentionCommon.cpp:533-547)
      k4  s = qk_scale*qk + (1-mask)*(-10000); p = exp(s-max); p *= 1/(sum+1e-6); p -> fp16 (:180-257)
          mask = causal & both < len (K/gptKernels.cu:136-199)
      k7  P.V fp16 out, fp32 accumulate (:602-605)
    Attention uses the *unquantised* k,v of this call (the cache copy is only for later steps)."""
    qkv = np.asarray(qkv, dtype=F16)
    B, S, _ = qkv.shape
    H, Dh = num_heads, head_size
    hid = H * Dh
    qk_scale = F32(1.0) / (np.sqrt(F32(Dh)) * F32(q_scaling))
    out = np.zeros((B, S, hid), dtype=F16)
    for b in range(B):
        L = int(input_lengths[b])
        q = qkv[b, :, 0:hid].reshape(S, H, Dh).transpose(1, 0, 2).copy()
        k = qkv[b, :, hid:2 * hid].reshape(S, H, Dh).transpose(1, 0, 2).copy()
        v = qkv[b, :, 2 * hid:].reshape(S, H, Dh).transpose(1, 0, 2).copy()
        pos = np.broadcast_to(np.arange(S)[None, :], (H, S))
        q = rope_neox(q, pos, rotary_dim)
        k = rope_neox(k, pos, rotary_dim)
        q[:, L:] = 0
        k[:, L:] = 0
        v[:, L:] = 0
        if kv_cache.dtype == np.int8:
            kv_cache[b, 0, :, :S] = kv_quant(k, kv_scale_orig_quant)
            kv_cache[b, 1, :, :S] = kv_quant(v, kv_scale_orig_quant)
        else:
            kv_cache[b, 0, :, :S] = k
            kv_cache[b, 1, :, :S] = v
        s = np.einsum("hqd,hkd->hqk", q[:, :L].astype(np.float64), k[:, :L].astype(np.float64)).astype(F32)
        mask = np.tril(np.ones((L, L), dtype=bool))
        s = (qk_scale * s + np.where(mask, F32(0), F32(-10000.0))[None]).astype(F32)
        p = np.exp(s - s.max(axis=-1, keepdims=True)).astype(F32)
        p = (p * (F32(1.0) / (p.sum(axis=-1, keepdims=True, dtype=F32) + F32(1e-6)))).astype(F16)
        o = np.einsum("hqk,hkd->hqd", p.astype(np.float64), v[:, :L].astype(np.float64)).astype(F32).astype(F16)
        out[b, :L] = o.transpose(1, 0, 2).reshape(L, hid)
    return out


# --------------------------------------------------------------------------- #
# a12: glue ops (TRT-native in the reference, semantics from the Python graph)
# --------------------------------------------------------------------------- #
def silu(x):
    xf = np.asarray(x).astype(F32)
    return (xf / (F32(1.0) + np.exp(-xf))).astype(F32)


def swiglu(fc_out, gate_out):
    """GatedMLP.forward: inter = act(fc(x)) * gate(x)   (T/tensorrt_llm/layers/mlp.py:68-73);
    fc = gate_proj, gate = up_proj (LQ/weight_quant.py:343-404).  fp16 in / fp16 out, fp32 inside."""
    a = np.asarray(fc_out, dtype=F16)
    g = np.asarray(gate_out, dtype=F16)
    return (silu(a).astype(F16).astype(F32) * g.astype(F32)).astype(F16)


def residual_add(x, r):
    return (np.asarray(x, F16).astype(F32) + np.asarray(r, F16).astype(F32)).astype(F16)


def gemm_f16(a, w_nk):
    """plain fp16 GEMM with fp32 accumulate, w stored [N, K] (torch Linear layout;
    T/tensorrt_llm/layers/linear.py:13-35 matmul(x, W^T) / P/gemmPlugin transb=1)."""
    return (np.asarray(a, F16).astype(np.float64) @ np.asarray(w_nk, F16).astype(np.float64).T).astype(F32).astype(F16)


def greedy_argmax(logits):
    """top_k=1 sampling == argmax with lowest index on ties (DynamicDecodeOp top-k=1;
    T/tensorrt_llm/runtime/generation.py:119-131 SamplingConfig defaults)."""
    return np.argmax(np.asarray(logits, F32), axis=-1).astype(np.int32)


# --------------------------------------------------------------------------- #
# f4: sampling (top-k / top-p), K/samplingTopKKernels.cu, K/samplingTopPKernels.cu, K/samplingPenaltyKernels.cu
# --------------------------------------------------------------------------- #
def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., the generator behind curand's Philox): counter 4 x u32, key 2 x u32 -> 4 x u32."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [int(x) & 0xFFFFFFFF for x in counter]
    k = [int(x) & 0xFFFFFFFF for x in key]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def sampling_uniform(seed, step, row):
    """The uniform in (0, 1] a sampling kernel draws for (seed, generation step, batch row): word 0 of
    Philox4x32-10(counter = (step, 0, row, 0), key = seed) mapped as curand_uniform does (x * 2^-32 + 2^-33).
    The reference draws from curand's XORWOW state initialised with curand_init(seed, 0, 0) for EVERY row
    (samplingTopKKernels.cu:38-62: all rows share one stream); a counter-based generator keyed by (step, row) needs no
    state buffer, survives CUDA-graph replay and gives every row its own stream — same distribution, different numbers."""
    x = philox4x32_10((step, 0, row, 0), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))[0]
    return F32(F32(x) * F32(2.3283064365386963e-10) + F32(1.1641532182693481e-10))


def sample_top_k_top_p(logits, top_k, top_p, temperature, uniforms):
    """logits [B, V] fp32 -> ids [B] (and the per-row candidate tables, for tolerance-aware tests).

    temperature: logits * (1 / (T + 1e-6))                                   (samplingPenaltyKernels.cu:77-93)
    top_k > 0:   the k largest logits (lowest index wins ties), e_i = exp(l_i - l_max), r = u * top_p * sum(e);
                 walk the candidates in descending order subtracting e_i, take the first with r <= 0 (or the last)
                                                                              (samplingTopKKernels.cu:197-319)
    top_k == 0:  probabilities softmax(l), sorted descending (stable), r = u * top_p; first token whose inclusive
                 cumulative probability reaches r                             (samplingTopPKernels.cu:882-1010, :1160-1236)"""
    logits = np.asarray(logits, dtype=F32)
    B, V = logits.shape
    inv_t = F32(1.0) / (F32(temperature) + F32(1e-6))
    ids, tables = np.zeros(B, np.int32), []
    for b in range(B):
        l = (logits[b] * inv_t).astype(F32)
        u = F32(uniforms[b])
        if top_k > 0:
            k = min(int(top_k), V)
            order = np.lexsort((np.arange(V), -l.astype(np.float64)))[:k]      # descending value, ascending index
            e = np.exp((l[order] - l[order[0]]).astype(F32)).astype(F32)
            s = F32(0)
            for x in e:
                s = F32(s + x)
            r = F32(F32(u * F32(top_p)) * s)
            pick = k - 1
            for i in range(k):
                r = F32(r - e[i])
                if r <= 0:
                    pick = i
                    break
            ids[b] = order[pick]
            tables.append((order, e, s))
        else:
            m = l.max()
            e = np.exp((l - m).astype(F32)).astype(F32)
            p = (e / e.sum(dtype=F32)).astype(F32)
            order = np.lexsort((np.arange(V), -p.astype(np.float64)))
            c = np.cumsum(p[order], dtype=F32)
            r = F32(u * F32(top_p))
            j = int(np.searchsorted(c, r, side="left"))
            ids[b] = order[j] if j < V else order[0]      # never reached: the scan leaves the top token (:957)
            tables.append((order, p, c))
    return ids, tables


# --------------------------------------------------------------------------- #
# f4: beam search (num_beams > 1) — the beam half of DynamicDecodeOp as the Python runtime drives it (no BeamHypotheses)
# --------------------------------------------------------------------------- #
def beam_search_step(logits, cum_log_probs, finished, beam_lens, src_indir, pos, *, beam_width, end_id, length_penalty=1.0):
    """One decoding step.  logits [rows, V] fp32 (rows = batch x beam, beam fastest); cum_log_probs [rows] fp32;
    finished [rows] bool; beam_lens [rows] int (the decoder's sequence lengths); src_indir [batch, beam, S_max] int;
    pos = sequence position of the token being chosen (max_input_length + generation step).

    Returns (tokens [rows], parents [rows], cum' [rows], finished' [rows], beam_lens' [rows], tgt_indir, margin) where margin
    is the smallest gap between a selected and the best rejected normalised score (tolerance-aware tests).

      K/onlineSoftmaxBeamsearchKernels.cu:402-592  per row: the 2W largest logits, log-softmax = l - max - log(sum exp(l - max)),
          candidate value = cum_log_probs[row] + log-softmax; a finished row proposes end_id with log-probability 0 and
          nothing else (every other entry is -MAX, i.e. -inf after the subtraction)
      :112-300 batch_topk_kernel: per batch entry the W best of the 2 W^2 candidates by value / len^length_penalty, where —
          with no BeamHypotheses — candidate j (0..2W-1) of ANY beam is normalised with the length of beam (j mod W)
          (`elem_id % K`), len = beam_lens (+1 unless that beam is finished), and len == 1 or length_penalty == 0 skip it
          (:44-52); the new cum_log_probs is the un-normalised value
      layers/onlineBeamSearchLayer.cu:30-62 update_kernel: parent = candidate's beam, token, finished' = token == end_id,
          beam_lens' = the parent's length after its own increment
      layers/baseBeamSearchLayer.cu:29-67 update_indir_cache: unfinished beams: tgt[beam][t] = src[parent][t] for t < pos,
          tgt[beam][pos] = beam
    """
    logits = np.asarray(logits, dtype=F32)
    rows, V = logits.shape
    W, n = int(beam_width), 2 * int(beam_width)
    batch = rows // W
    cum = np.asarray(cum_log_probs, dtype=F32)
    fin = np.asarray(finished).astype(bool)
    lens = np.asarray(beam_lens, dtype=np.int64)
    cand_id = np.zeros((rows, n), np.int64)
    cand_val = np.full((rows, n), -np.inf, dtype=F32)
    for r in range(rows):
        if fin[r]:
            cand_id[r, 0], cand_val[r, 0] = end_id, cum[r]
            continue
        l = logits[r]
        order = np.lexsort((np.arange(V), -l.astype(np.float64)))[:n]
        m = l.max()
        d = np.exp((l - m).astype(F32)).astype(F32).sum(dtype=np.float64)
        lp = ((l[order] - m).astype(F32) - F32(np.log(d))).astype(F32)
        cand_id[r], cand_val[r] = order, (lp + cum[r]).astype(F32)
    tokens, parents = np.zeros(rows, np.int32), np.zeros(rows, np.int32)
    new_cum, new_fin, new_lens = cum.copy(), fin.copy(), lens.copy()
    tgt = np.array(src_indir, dtype=np.int32, copy=True)
    margin = np.inf
    for b in range(batch):
        base = b * W
        inc = lens[base:base + W] + (~fin[base:base + W]).astype(np.int64)
        elem = np.empty(W * n, dtype=F32)
        for e in range(W * n):
            i = (e % n) % W
            v = cand_val[base + e // n, e % n]
            ln = inc[i]
            if length_penalty != 0.0 and ln != 1 and np.isfinite(v):
                v = F32(v / F32(np.power(F32(ln), F32(length_penalty))))
            elem[e] = v
        order = np.lexsort((np.arange(W * n), -elem.astype(np.float64)))
        sel = order[:W]
        if W * n > W and np.isfinite(elem[order[W]]):
            margin = min(margin, float(elem[sel[-1]] - elem[order[W]]))
        for k in range(W - 1):
            margin = min(margin, float(elem[sel[k]] - elem[sel[k + 1]])) if np.isfinite(elem[sel[k + 1]]) else margin
        for k, e in enumerate(sel):
            pb, j = e // n, e % n
            tokens[base + k] = cand_id[base + pb, j]
            parents[base + k] = pb
            new_cum[base + k] = cand_val[base + pb, j]
            new_fin[base + k] = tokens[base + k] == end_id
            new_lens[base + k] = inc[pb]
        for k in range(W):
            if new_fin[base + k]:
                continue
            tgt[b, k, :pos] = np.asarray(src_indir)[b, parents[base + k], :pos]
            tgt[b, k, pos] = k
    return tokens, parents, new_cum, new_fin, new_lens, tgt, margin


def gather_tree(out_ids_t, parent_ids_t, beam_width, end_id):
    """out_ids_t / parent_ids_t [n_steps, rows] -> [rows, n_steps]: each final beam's path through the parent pointers, then
    end_id after the first end_id (K/decodingKernels.cu:31-170 for the generated part)."""
    out_ids_t, parent_ids_t = np.asarray(out_ids_t), np.asarray(parent_ids_t)
    n, rows = out_ids_t.shape
    W = int(beam_width)
    out = np.zeros((rows, n), np.int32)
    for r in range(rows):
        base, beam = r // W * W, r % W
        for c in range(n - 1, -1, -1):
            out[r, c] = out_ids_t[c, base + beam]
            beam = parent_ids_t[c, base + beam]
        hit = np.nonzero(out[r] == end_id)[0]
        if hit.size:
            out[r, hit[0]:] = end_id
    return out


def mmha_decode_beams(qkv, kv_cache, cache_indir, beam_width, past_len, input_lengths, max_input_len, **kw):
    """mmha_decode with beam search: cached position t of row (b, w) is read from row (b, cache_indir[b][w][t]); the
    position appended now is the row's own (decoderMaskedMultiheadAttentionTemplate.h:1137-1146,1624-1631).  kv_cache
    [rows, 2, H, S_max, Dh] is updated in place at ``past_len`` only."""
    rows, W, t = kv_cache.shape[0], int(beam_width), int(past_len)
    ind = np.asarray(cache_indir).reshape(rows, -1)
    gathered = kv_cache.copy()
    for r in range(rows):
        src = r // W * W + ind[r, :t]
        gathered[r, :, :, :t, :] = kv_cache[src, :, :, np.arange(t), :].transpose(1, 2, 0, 3)
    out = mmha_decode(qkv, gathered, past_len, input_lengths, max_input_len, **kw)
    kv_cache[:, :, :, t, :] = gathered[:, :, :, t, :]
    return out
