"""The reference's OWN CUDA kernels (oracle/_ref/libref_cuda.so, compiled for sm_100a from
/root/reference by oracle/Makefile) against (a) the numpy oracle — this pins the oracle — and
(b) this repo's kernels, on identical seeded inputs.  Skipped when the checker library is absent."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_ops as R  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CUDA = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
REF_HOST = os.path.join(ROOT, "oracle", "_ref", "libref_host.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_CUDA):
        pytest.skip("oracle/_ref/libref_cuda.so not built")
    return C.CDLL(REF_CUDA)


@pytest.fixture(scope="module")
def ops():
    import trtllm_llama_b200  # noqa: F401
    from trtllm_llama_b200 import ops as o
    return o


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def P(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


@pytest.mark.parametrize("int8_kv", [True, False])
@pytest.mark.parametrize("past", [40, 333, 1500])
def test_mmha_reference_kernel(ref, ops, int8_kv, past):
    rng = np.random.default_rng(21)
    B, H, Dh, S_max, max_in = 3, 4, 128, 2048, 32
    in_lens = np.array([32, 20, 1], dtype=np.int32)
    qkv = rng.standard_normal((B, 3 * H * Dh)).astype(np.float16)
    if int8_kv:
        cache = rng.integers(-127, 128, (B, 2, H, S_max, Dh), dtype=np.int8)
        s_q, s_dq = np.float32(127.0 / 4.0), np.float32(4.0 / 127.0)
    else:
        cache = rng.standard_normal((B, 2, H, S_max, Dh)).astype(np.float16)
        s_q = s_dq = None
    masked = np.zeros((B, S_max), dtype=np.int32)
    for b in range(B):
        masked[b, in_lens[b]:max_in] = 1
    seq_lens = np.full((B,), past, dtype=np.int32)

    # reference kernel
    c_ref = dev(cache)
    out_ref = torch.zeros((B, H * Dh), dtype=torch.float16, device="cuda")
    pad_ws = torch.zeros((B,), dtype=torch.int32, device="cuda")
    d_sq = dev(np.array([s_q or 1.0], np.float32))
    d_sdq = dev(np.array([s_dq or 1.0], np.float32))
    d_qkv, d_seq, d_in, d_mask = dev(qkv), dev(seq_lens), dev(in_lens), dev(masked)
    rc = ref.ref_mmha_decode_half(P(out_ref), P(d_qkv), P(c_ref), B, H, Dh, S_max, past, max_in, P(d_seq), P(d_in),
                                  P(d_mask), P(pad_ws), P(d_sq), P(d_sdq), int(int8_kv), Dh, C.c_float(1.0),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    out_ref = host(out_ref).astype(np.float32)
    cache_after_ref = host(c_ref)

    # oracle (pins the restatement against the real kernel)
    c_or = cache.copy()
    out_or = R.mmha_decode(qkv, c_or, past, in_lens, max_in, num_heads=H, head_size=Dh, kv_scale_orig_quant=s_q,
                           kv_scale_quant_orig=s_dq).astype(np.float32)
    # T/tests/attention/test_gpt_attention.py:828-831: atol 2e-3 on outputs of magnitude <= 1; outputs here
    # reach ~4, so the same bound is applied relative to the output scale (2 fp16 ulps at that scale)
    tol = 2e-3 * max(1.0, float(np.abs(out_ref).max()))
    np.testing.assert_allclose(out_or, out_ref, atol=tol)
    if int8_kv:
        assert np.abs(c_or.astype(np.int32) - cache_after_ref.astype(np.int32)).max() <= 1
    else:
        np.testing.assert_allclose(c_or.astype(np.float32), cache_after_ref.astype(np.float32), atol=2e-3)

    # this repo's kernel, both single-CTA and split-L
    for nsplit in (1, 0):
        c_my = dev(cache)
        kw = dict(kv_scale_orig_quant=d_sq, kv_scale_quant_orig=d_sdq) if int8_kv else {}
        out_my = ops.mmha_decode(d_qkv, c_my, past, num_heads=H, head_size=Dh, max_input_len=max_in, seq_lens=d_seq,
                                 input_lengths=d_in, masked_tokens=d_mask, nsplit=nsplit, **kw)
        np.testing.assert_allclose(host(out_my).astype(np.float32), out_ref, atol=tol)
        if int8_kv:
            assert np.abs(host(c_my).astype(np.int32) - cache_after_ref.astype(np.int32)).max() <= 1
        else:
            np.testing.assert_allclose(host(c_my).astype(np.float32), cache_after_ref.astype(np.float32), atol=2e-3)


def test_per_token_quant_reference_kernel(ref, ops):
    rng = np.random.default_rng(22)
    x = rng.standard_normal((64, 4096)).astype(np.float16)
    d_x = dev(x)
    q_ref = torch.zeros((64, 4096), dtype=torch.int8, device="cuda")
    s_ref = torch.zeros((64,), dtype=torch.float32, device="cuda")
    assert ref.ref_per_token_quant_half(P(q_ref), P(d_x), C.c_int64(64), C.c_int64(4096), P(s_ref),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    q_or, s_or = R.quantize_per_token(x)
    assert np.array_equal(host(q_ref), q_or)
    np.testing.assert_array_equal(host(s_ref), s_or[:, 0])
    q_my, s_my = ops.quantize_per_token(d_x)
    assert np.array_equal(host(q_my), host(q_ref))
    np.testing.assert_array_equal(host(s_my)[:, 0], host(s_ref))


@pytest.mark.parametrize("dynamic", [True, False])
def test_layernorm_quant_reference_kernel(ref, ops, dynamic):
    rng = np.random.default_rng(1997)
    rows, hidden = 32, 1024
    x = rng.standard_normal((rows, hidden)).astype(np.float16)
    g = rng.standard_normal(hidden).astype(np.float16)
    b = rng.standard_normal(hidden).astype(np.float16)
    scale = np.array([25.0], np.float32)
    d_x, d_g, d_b, d_s = dev(x), dev(g), dev(b), dev(scale)
    q_ref = torch.zeros((rows, hidden), dtype=torch.int8, device="cuda")
    ds_ref = torch.zeros((rows,), dtype=torch.float32, device="cuda")
    out_unused = torch.zeros((rows, hidden), dtype=torch.float16, device="cuda")
    rc = ref.ref_layernorm_quant_half(P(out_unused), P(d_x), P(d_g), P(d_b), C.c_float(1e-5), rows, hidden, 0,
                                      P(None if dynamic else d_s), P(ds_ref if dynamic else None), P(q_ref),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    q_or, s_or = R.layernorm_quant(x, g, b, 1e-5, scale[0], dynamic)
    assert np.abs(host(q_ref).astype(np.int32) - q_or.astype(np.int32)).max() <= 1
    res = ops.smooth_quant_rms_norm(d_x, d_g, d_s, 1e-5, dynamic, bias=d_b, layernorm=True)
    assert np.abs(host(res[0]).astype(np.int32) - host(q_ref).astype(np.int32)).max() <= 1
    if dynamic:
        np.testing.assert_allclose(host(ds_ref), s_or[:, 0], rtol=1e-3)
        np.testing.assert_allclose(host(res[1])[:, 0], host(ds_ref), rtol=1e-3)


@pytest.mark.parametrize("bits", [8, 4])
def test_weight_only_gemv_reference_kernel(ref, ops, bits):
    if not os.path.exists(REF_HOST):
        pytest.skip("oracle/_ref/libref_host.so not built")
    hostlib = C.CDLL(REF_HOST)
    rng = np.random.default_rng(0)
    K, N = 4096, 1024                                   # test_weight_only_quant_matmul.py:90 (1, 1024, 4096)
    w = (rng.random((K, N), dtype=np.float32) * 2 - 1).astype(np.float16)
    x = (rng.random((1, K), dtype=np.float32) * 2 - 1).astype(np.float16)
    nb = K * N * bits // 8
    proc = np.zeros(nb, np.int8); unproc = np.zeros(nb, np.int8); sc = np.zeros(N, np.uint16)
    assert hostlib.ref_symmetric_quantize(w.ctypes.data_as(C.c_void_p), C.c_int64(K), C.c_int64(N), bits,
                                          proc.ctypes.data_as(C.c_void_p), unproc.ctypes.data_as(C.c_void_p),
                                          sc.ctypes.data_as(C.c_void_p)) == 0
    scales = sc.view(np.float16)
    q, s_or = R.symmetric_quantize(w, bits)
    assert np.array_equal(scales, s_or)
    d_x, d_w, d_s = dev(x), dev(proc), dev(scales)
    y_ref = torch.zeros((1, N), dtype=torch.float16, device="cuda")
    assert ref.ref_weight_only_gemv_half(P(d_x), P(d_w), P(d_s), P(y_ref), K, N, bits,
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    y_ref = host(y_ref).astype(np.float32)
    y_or = R.weight_only_matmul(x, q, scales).astype(np.float32)
    tol = 1.5 * np.abs(y_or).max() / (1 << (bits - 1))   # reference's own tolerance (_utils.py:62-89)
    np.testing.assert_allclose(y_or, y_ref, atol=tol)
    qt = np.ascontiguousarray(q.T)
    from trtllm_llama_b200.quantization import pack_processed_int4
    wp = qt if bits == 8 else pack_processed_int4(torch.from_numpy(np.ascontiguousarray(qt))).numpy()
    for use_gemv in (True, False):
        y_my = host(ops.weight_only_quant_matmul(d_x, dev(wp), d_s, 1 if bits == 8 else 2, use_gemv=use_gemv))
        np.testing.assert_allclose(y_my.astype(np.float32), y_ref, atol=tol)
        np.testing.assert_allclose(y_my.astype(np.float32), y_or, atol=2e-3 * np.abs(y_or).max() + 1e-3)


# ------------------------------------------------------------------------------------------------
# The reference's CUTLASS GEMMs (oracle/_ref/libref_cutlass.so: compute_90 PTX, JIT-compiled on this GPU) as GPU oracle
# ------------------------------------------------------------------------------------------------
REF_CUTLASS = os.path.join(ROOT, "oracle", "_ref", "libref_cutlass.so")


@pytest.fixture(scope="module")
def cutlass_ref():
    if not os.path.exists(REF_CUTLASS):
        pytest.skip("oracle/_ref/libref_cutlass.so not built")
    return C.CDLL(REF_CUTLASS)


@pytest.mark.parametrize("M,N,K", [(8, 256, 512), (300, 384, 1024), (2048, 4096, 4096)])
@pytest.mark.parametrize("per_token,per_channel", [(True, True), (False, False)])
def test_sq_gemm_reference_cutlass_kernel(cutlass_ref, ops, M, N, K, per_token, per_channel):
    """CutlassInt8GemmRunner's kernel (int8_gemm_template.h:56-172) == the oracle == this repo's tcgen05 GEMM, bit for bit
    (test_smooth_quant_gemm.py:20-127 input distributions)."""
    g = torch.Generator(device="cuda").manual_seed(31)
    a = torch.randint(-128, 128, (M, K), device="cuda", dtype=torch.int8, generator=g)
    b = torch.randint(-128, 128, (N, K), device="cuda", dtype=torch.int8, generator=g)
    sr = (torch.randint(1, 10, (M if per_token else 1,), device="cuda", generator=g).float() * 1e-2).contiguous()
    sc = (torch.randint(1, 10, (N if per_channel else 1,), device="cuda", generator=g).float() * 1e-2).contiguous()
    ws = torch.zeros(16 << 20, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ran = 0
    mine = host(ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc.view(1, -1), sr=sr.view(-1, 1)))
    for tactic in range(cutlass_ref.ref_int8_gemm_num_tactics()):
        c = torch.zeros((M, N), dtype=torch.float16, device="cuda")
        rc = cutlass_ref.ref_int8_gemm_half(P(a), P(b), P(sc), P(sr), P(c), M, N, K, int(per_channel), int(per_token), tactic,
                                            P(ws), C.c_size_t(ws.numel()), st)
        if rc != 0:
            continue
        ran += 1
        assert np.array_equal(host(c), mine), f"reference CUTLASS tactic {tactic} differs from the tcgen05 GEMM"
    assert ran > 0, "no reference tactic ran"
    if M <= 300:
        assert np.array_equal(mine, R.sq_gemm(host(a), host(b), host(sr), host(sc), np.float16))


@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("M", [8, 130])
def test_weight_only_gemm_reference_cutlass_kernel(cutlass_ref, ops, bits, M):
    """CutlassFpAIntBGemmRunner's kernel on the reference's own pre-processed weights (the path the reference takes for
    batch > 1, fpA_intB_gemm_template.h:49-175) against the oracle and this repo's kernels on their own layout."""
    if not os.path.exists(REF_HOST):
        pytest.skip("oracle/_ref/libref_host.so not built")
    hostlib = C.CDLL(REF_HOST)
    rng = np.random.default_rng(0)
    K, N = 4096, 1024
    w = (rng.random((K, N), dtype=np.float32) * 2 - 1).astype(np.float16)
    x = (rng.random((M, K), dtype=np.float32) * 0.2 - 0.1).astype(np.float16)
    nb = K * N * bits // 8
    proc = np.zeros(nb, np.int8); unproc = np.zeros(nb, np.int8); sc = np.zeros(N, np.uint16)
    assert hostlib.ref_symmetric_quantize(w.ctypes.data_as(C.c_void_p), C.c_int64(K), C.c_int64(N), bits,
                                          proc.ctypes.data_as(C.c_void_p), unproc.ctypes.data_as(C.c_void_p),
                                          sc.ctypes.data_as(C.c_void_p)) == 0
    scales = sc.view(np.float16)
    q, _ = R.symmetric_quantize(w, bits)
    y_or = R.weight_only_matmul(x, q, scales).astype(np.float32)
    d_x, d_w, d_s = dev(x), dev(proc), dev(scales)
    ws = torch.zeros(16 << 20, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    tol = 1.5 * np.abs(y_or).max() / (1 << (bits - 1))   # the reference's own tolerance (_utils.py:62-89)
    ran = 0
    for tactic in range(cutlass_ref.ref_fpA_intB_gemm_num_tactics()):
        y = torch.zeros((M, N), dtype=torch.float16, device="cuda")
        rc = cutlass_ref.ref_fpA_intB_gemm_half(P(d_x), P(d_w), P(d_s), P(y), M, N, K, bits, tactic, P(ws),
                                                C.c_size_t(ws.numel()), st)
        if rc != 0:
            continue
        ran += 1
        y_ref = host(y).astype(np.float32)
        np.testing.assert_allclose(y_or, y_ref, atol=tol)
    assert ran > 0, "no reference tactic ran"
    from trtllm_llama_b200.quantization import pack_processed_int4
    qt = np.ascontiguousarray(q.T)
    wp = qt if bits == 8 else pack_processed_int4(torch.from_numpy(qt)).numpy()
    y_my = host(ops.weight_only_quant_matmul(d_x, dev(wp), d_s, 1 if bits == 8 else 2, use_gemv=M <= 8)).astype(np.float32)
    np.testing.assert_allclose(y_my, y_ref, atol=tol)
    np.testing.assert_allclose(y_my, y_or, atol=2e-3 * np.abs(y_or).max() + 1e-3)
