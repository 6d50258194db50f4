#!/usr/bin/env python
"""The reference's own CUDA kernels, recompiled for this GPU (oracle/_ref/libref_cuda.so: sm_100a; libref_cutlass.so: the
CUTLASS 2.10 GEMMs as compute_90 PTX, JIT-compiled here), timed beside this library's kernel for the same shape in the
SAME harness: CUDA events on the launching stream, 3 warm-up launches, operands rotated through enough distinct copies
to exceed the 126 MB L2 between repeats.  This is the "reference engine on the same box" comparator of BASELINE.md §3 /
SURVEY F6 at kernel level (TensorRT itself cannot be installed here).  MEASUREMENT INFRASTRUCTURE: loads oracle/_ref.

    python tools/ref_kernel_bench.py            # prints one JSON object
`bench.py` embeds the same dict as `reference_kernels`.  Reference entry points:
  K/decoderMaskedMultiheadAttention.h:184-199 (masked_multihead_attention), K/weightOnlyMatrixVectorMultiplication.cu:371-378,
  K/layernormKernels.cu:233-264, K/quantization.cu:119-130, K/cutlass_kernels/int8_gemm/int8_gemm_template.h:56-172,
  K/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:49-175."""
from __future__ import annotations

import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_CUDA = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
REF_CUTLASS = os.path.join(ROOT, "oracle", "_ref", "libref_cutlass.so")
L2_BYTES = 126 << 20
H, DH, HID, INTER = 32, 128, 4096, 11008


def _P(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


def _time(torch, fns, reps=None):
    """fns: closures over distinct operand copies (rotated so that no launch finds its operands in L2); returns the mean
    microseconds per launch.  The `reps` launches (default: 3 passes over the rotation, >= 12) are captured into ONE CUDA
    graph and the replay is timed with CUDA events, so neither side pays Python / ctypes launch overhead; if a kernel
    cannot be captured the launches are timed eagerly (noted by the caller through `_time.eager`)."""
    for f in (fns[:3] if len(fns) >= 3 else fns * 3):
        f()
    torch.cuda.synchronize()
    reps = reps or max(12, 3 * len(fns))

    def run():
        for i in range(reps):
            fns[i % len(fns)]()
    replay = run
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        replay = g.replay
        replay()
    except Exception:      # noqa: BLE001
        _time.eager += 1
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


_time.eager = 0


def _copies(nbytes):
    return max(2, min(12, (2 * L2_BYTES + nbytes - 1) // max(nbytes, 1)))


def reference_kernels(full=True):
    import torch
    import trtllm_llama_b200  # noqa: F401
    from trtllm_llama_b200 import ops
    if not os.path.exists(REF_CUDA):
        return {"unavailable": "oracle/_ref/libref_cuda.so not built"}
    ref = C.CDLL(REF_CUDA)
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    g = torch.Generator(device="cuda").manual_seed(0)
    out = {"harness": "CUDA events around one CUDA-graph replay of >= 12 launches per kernel (no host launch overhead on either "
                      "side), 3 warm-up launches, operands rotated over copies totalling > 2 x L2; us per launch",
           "rows": []}

    def row(kernel, shape, ref_us, ours_us, nbytes, note=None):
        r = {"kernel": kernel, "shape": shape, "reference_us": round(ref_us, 2) if ref_us else None,
             "ours_us": round(ours_us, 2), "speedup": round(ref_us / ours_us, 2) if ref_us else None,
             "ours_gbs": round(nbytes / ours_us / 1e3, 1) if nbytes else None,
             "reference_gbs": round(nbytes / ref_us / 1e3, 1) if (nbytes and ref_us) else None}
        if note:
            r["note"] = note
        out["rows"].append(r)

    # ---- decode attention: cfg2 (B = 1, ~192 cached, int8) and cfg3 (B = 8, 2047 cached, fp16 / int8) -----------------
    for B, past, int8_kv, tag in ((1, 191, True, "cfg2"), (8, 2047, False, "cfg3 fp16 KV"), (8, 2047, True, "cfg3 int8 KV")):
        S_max = past + 129
        elt = 1 if int8_kv else 2
        nb = B * 2 * H * S_max * DH * elt
        n = _copies(nb)
        caches = [(torch.randint(-127, 128, (B, 2, H, S_max, DH), device="cuda", dtype=torch.int8, generator=g) if int8_kv
                   else torch.randn(B, 2, H, S_max, DH, device="cuda", generator=g).half()) for _ in range(n)]
        qkv = torch.randn(B, 3 * HID, device="cuda", generator=g).half()
        lens = torch.full((B,), past, dtype=torch.int32, device="cuda")
        in_lens = torch.full((B,), min(past, 128), dtype=torch.int32, device="cuda")
        masked = torch.zeros((B, S_max), dtype=torch.int32, device="cuda")
        pad_ws = torch.zeros((B,), dtype=torch.int32, device="cuda")
        sq, sdq = torch.tensor([127.0 / 4.0], device="cuda"), torch.tensor([4.0 / 127.0], device="cuda")
        o_ref = torch.zeros((B, HID), dtype=torch.float16, device="cuda")
        max_in = min(past, 128)

        def f_ref(c):
            return lambda: ref.ref_mmha_decode_half(_P(o_ref), _P(qkv), _P(c), B, H, DH, S_max, past, max_in, _P(lens), _P(in_lens),
                                                    _P(masked), _P(pad_ws), _P(sq), _P(sdq), int(int8_kv), DH, C.c_float(1.0), st())
        kw = dict(kv_scale_orig_quant=sq, kv_scale_quant_orig=sdq) if int8_kv else {}

        def f_my(c):
            return lambda: ops.mmha_decode(qkv, c, past, num_heads=H, head_size=DH, max_input_len=max_in, seq_lens=lens,
                                           input_lengths=in_lens, nsplit=0, **kw)
        bytes_read = 2 * H * B * past * DH * elt
        row("masked_multihead_attention (decode)", f"{tag}: B={B} H=32 L={past} {'int8' if int8_kv else 'fp16'} KV",
            _time(torch, [f_ref(c) for c in caches]), _time(torch, [f_my(c) for c in caches]), bytes_read)
        del caches

    # ---- weight-only GEMV, batch 1 (the reference's M = 1 path), the LLaMA-7B projections ---------------------------------
    x1 = (torch.randn(1, INTER, device="cuda", generator=g) * 0.1).half()
    for bits in (8, 4):
        for nm, N, K in (("qkv", 3 * HID, HID), ("dense", HID, HID), ("gate", INTER, HID), ("down", HID, INTER)):
            nb = N * K * bits // 8
            n = _copies(nb)
            ws_ref = [torch.randint(-128, 128, (nb,), device="cuda", dtype=torch.int8, generator=g) for _ in range(n)]
            sc = (torch.rand(N, device="cuda", generator=g) * 0.01).half()
            y = torch.zeros((1, N), dtype=torch.float16, device="cuda")
            xk = x1[:, :K].contiguous()
            kind = ops.KIND_W8 if bits == 8 else ops.KIND_W4
            t_ref = _time(torch, [(lambda w=w: ref.ref_weight_only_gemv_half(_P(xk), _P(w), _P(sc), _P(y), K, N, bits, st()))
                                  for w in ws_ref])
            wv = [w.view(N, K * bits // 8) for w in ws_ref]
            t_my = _time(torch, [(lambda w=w: ops.gemv(kind, xk, w, w_scale=sc)) for w in wv])
            row(f"weight_only_gemv int{bits} (M=1)", f"{nm}: N={N} K={K}", t_ref, t_my, nb)
            del ws_ref, wv

    # ---- norm + quantise, per-token quantise (M = 8 decode rows, M = 16384 prefill rows) ---------------------------------
    for M in (8, 16384):
        x = torch.randn(M, HID, device="cuda", generator=g).half()
        gam = torch.ones(HID, device="cuda", dtype=torch.float16)
        bet = torch.zeros(HID, device="cuda", dtype=torch.float16)
        q = torch.zeros((M, HID), dtype=torch.int8, device="cuda")
        s = torch.zeros((M,), dtype=torch.float32, device="cuda")
        unused = torch.zeros((M, HID), dtype=torch.float16, device="cuda")
        nb = M * HID * 3
        t_ref = _time(torch, [lambda: ref.ref_layernorm_quant_half(_P(unused), _P(x), _P(gam), _P(bet), C.c_float(1e-6), M, HID, 0,
                                                                   _P(None), _P(s), _P(q), st())])
        t_my = _time(torch, [lambda: ops.smooth_quant_rms_norm(x, gam, None, 1e-6, True)])
        row("LayernormQuantization (ref) vs RmsnormQuantization (ours), dynamic", f"M={M} hidden=4096", t_ref, t_my, nb,
            "the reference has no RMS variant (SURVEY F1): its LayerNorm + quantise kernel is the nearest comparator")
        xi = torch.randn(M, INTER, device="cuda", generator=g).half()
        qi = torch.zeros((M, INTER), dtype=torch.int8, device="cuda")
        t_ref = _time(torch, [lambda: ref.ref_per_token_quant_half(_P(qi), _P(xi), C.c_int64(M), C.c_int64(INTER), _P(s), st())])
        t_my = _time(torch, [lambda: ops.quantize_per_token(xi)])
        row("invokePerTokenQuantization", f"M={M} cols=11008", t_ref, t_my, M * INTER * 3)

    # ---- CUTLASS GEMMs of the reference (compute_90 PTX, JIT) ----------------------------------------------------------
    if full and os.path.exists(REF_CUTLASS):
        cut = C.CDLL(REF_CUTLASS)
        ws = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda")
        M = 16384
        a = torch.randint(-127, 128, (M, INTER), device="cuda", dtype=torch.int8, generator=g)
        sr = torch.rand(M, device="cuda", generator=g) * 0.01 + 1e-3
        for nm, N, K in (("qkv", 3 * HID, HID), ("dense", HID, HID), ("gate+up", 2 * INTER, HID), ("down", HID, INTER)):
            b = torch.randint(-127, 128, (N, K), device="cuda", dtype=torch.int8, generator=g)
            sc = torch.rand(N, device="cuda", generator=g) * 0.01 + 1e-3
            ak = a[:, :K].contiguous()
            c = torch.zeros((M, N), dtype=torch.float16, device="cuda")
            best, best_t = None, None
            for t in range(cut.ref_int8_gemm_num_tactics()):
                rc = cut.ref_int8_gemm_half(_P(ak), _P(b), _P(sc), _P(sr), _P(c), M, N, K, 1, 1, t, _P(ws), C.c_size_t(ws.numel()), st())
                torch.cuda.synchronize()
                if rc != 0:
                    continue
                us = _time(torch, [lambda t=t: cut.ref_int8_gemm_half(_P(ak), _P(b), _P(sc), _P(sr), _P(c), M, N, K, 1, 1, t, _P(ws),
                                                                      C.c_size_t(ws.numel()), st())], reps=5)
                if best is None or us < best:
                    best, best_t = us, t
            ours = ops.gemm_tc(ops.KIND_A8W8, ak, b, sc=sc.view(1, -1), sr=sr.view(-1, 1))
            same = bool(torch.equal(ours, c)) if best is not None else None
            t_my = _time(torch, [lambda: ops.gemm_tc(ops.KIND_A8W8, ak, b, sc=sc.view(1, -1), sr=sr.view(-1, 1))], reps=5)
            r_ops = 2.0 * M * N * K
            row("CutlassInt8GemmRunner (SmoothQuant GEMM, per-token x per-channel)", f"{nm}: M={M} N={N} K={K}", best, t_my, 0,
                f"best of {cut.ref_int8_gemm_num_tactics()} reference tactics (#{best_t}); reference {r_ops / best / 1e6:.0f} TOP/s, "
                f"ours {r_ops / t_my / 1e6:.0f} TOP/s; outputs bit-identical: {same}" if best else "reference kernel failed to run")
            del b, c, ours
        # weight-only CUTLASS path the reference takes for batch 8 (M != 1): cfg3
        x8 = (torch.randn(8, INTER, device="cuda", generator=g) * 0.1).half()
        for bits in (8, 4):
            for nm, N, K in (("qkv", 3 * HID, HID), ("dense", HID, HID), ("gate", INTER, HID), ("down", HID, INTER)):
                nb = N * K * bits // 8
                n = _copies(nb)
                wsr = [torch.randint(-128, 128, (nb,), device="cuda", dtype=torch.int8, generator=g) for _ in range(n)]
                sc = (torch.rand(N, device="cuda", generator=g) * 0.01).half()
                y = torch.zeros((8, N), dtype=torch.float16, device="cuda")
                xk = x8[:, :K].contiguous()
                best = None
                for t in range(cut.ref_fpA_intB_gemm_num_tactics()):
                    rc = cut.ref_fpA_intB_gemm_half(_P(xk), _P(wsr[0]), _P(sc), _P(y), 8, N, K, bits, t, _P(ws), C.c_size_t(ws.numel()), st())
                    torch.cuda.synchronize()
                    if rc != 0:
                        continue
                    us = _time(torch, [(lambda w=w, t=t: cut.ref_fpA_intB_gemm_half(_P(xk), _P(w), _P(sc), _P(y), 8, N, K, bits, t, _P(ws),
                                                                                   C.c_size_t(ws.numel()), st())) for w in wsr])
                    best = us if best is None else min(best, us)
                kind = ops.KIND_W8 if bits == 8 else ops.KIND_W4
                wv = [w.view(N, K * bits // 8) for w in wsr]
                t_my = _time(torch, [(lambda w=w: ops.gemv(kind, xk, w, w_scale=sc)) for w in wv])
                row(f"CutlassFpAIntBGemmRunner int{bits} (M=8, the reference's batch-8 decode path)", f"{nm}: N={N} K={K}", best, t_my, nb,
                    f"best of {cut.ref_fpA_intB_gemm_num_tactics()} reference tactics")
                del wsr, wv
    elif full:
        out["cutlass"] = "oracle/_ref/libref_cutlass.so not built"
    sp = [r["speedup"] for r in out["rows"] if r["speedup"]]
    out["min_speedup"], out["rows_faster"], out["rows_total"] = (min(sp) if sp else None), sum(s > 1 for s in sp), len(sp)
    out["eager_fallbacks"] = _time.eager
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    print(json.dumps(reference_kernels(), indent=1))
