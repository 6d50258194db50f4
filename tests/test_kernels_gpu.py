"""Parity of every sm_100a kernel (through the C ABI) against the numpy oracle on seeded inputs.

Tolerances: integer / int8 outputs bit-exact unless stated; fp16 outputs within the stated atol,
chosen at or below the reference tests' own (T/tests/quantization/*, T/tests/attention/*)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_ops as R  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    import trtllm_llama_b200  # noqa: F401
    from trtllm_llama_b200 import ops as o
    assert trtllm_llama_b200.lib.tb_check_device() == 0
    return o


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,hidden", [(1, 4096), (8, 4096), (33, 1024), (5, 11008)])
def test_rmsnorm(ops, rows, hidden):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((rows, hidden)).astype(np.float16)
    g = (1 + 0.1 * rng.standard_normal(hidden)).astype(np.float16)
    y = host(ops.rms_norm(dev(x), dev(g), 1e-6))
    ref = R.rmsnorm(x, g, 1e-6)
    # fp32 reduction order differs -> at most 1 fp16 ulp
    np.testing.assert_allclose(y.astype(np.float32), ref.astype(np.float32), rtol=2e-3, atol=1e-3)
    # fused residual
    r = rng.standard_normal((rows, hidden)).astype(np.float16)
    y2, s2 = ops.rms_norm(dev(x), dev(g), 1e-6, residual=dev(r), return_sum=True)
    hsum = R.residual_add(x, r)
    assert np.array_equal(host(s2), hsum)
    np.testing.assert_allclose(host(y2).astype(np.float32), R.rmsnorm(hsum, g).astype(np.float32), rtol=2e-3, atol=1e-3)


@pytest.mark.parametrize("rows,hidden", [(8, 4096), (64, 1024)])
@pytest.mark.parametrize("dynamic", [True, False])
def test_rmsnorm_quant(ops, rows, hidden, dynamic):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((rows, hidden)).astype(np.float16)
    g = (1 + 0.1 * rng.standard_normal(hidden)).astype(np.float16)
    scale = np.array([31.5], dtype=np.float32)
    res = ops.smooth_quant_rms_norm(dev(x), dev(g), dev(scale), 1e-6, dynamic)
    q_ref, s_ref = R.rmsnorm_quant(x, g, 1e-6, scale[0], dynamic)
    q = host(res[0])
    # T/tests/quantization/test_smooth_quant_layer_norm.py:22-112: int8 atol 1, scales atol 1e-2
    assert np.abs(q.astype(np.int32) - q_ref.astype(np.int32)).max() <= 1
    assert (q != q_ref).mean() < 2e-3
    if dynamic:
        np.testing.assert_allclose(host(res[1]), s_ref, rtol=1e-6)


def test_layernorm_quant_mode(ops):
    rng = np.random.default_rng(1997)
    x = rng.standard_normal((16, 1024)).astype(np.float16)
    g = rng.standard_normal(1024).astype(np.float16)
    b = rng.standard_normal(1024).astype(np.float16)
    q, s = ops.smooth_quant_rms_norm(dev(x), dev(g), None, 1e-5, True, bias=dev(b), layernorm=True)
    q_ref, s_ref = R.layernorm_quant(x, g, b, 1e-5)
    assert np.abs(host(q).astype(np.int32) - q_ref.astype(np.int32)).max() <= 1
    np.testing.assert_allclose(host(s), s_ref, rtol=1e-3)


@pytest.mark.parametrize("dtype", [np.float16, np.float32])
def test_quantize_per_token(ops, dtype):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((4, 2, 4, 4096)).astype(dtype)
    x[0, 0, 0] = 0  # all-zero row: amax floor 1e-6
    q, s = ops.quantize_per_token(dev(x))
    q_ref, s_ref = R.quantize_per_token(x)
    assert np.array_equal(host(q), q_ref)          # bit-exact (test_functional.py:110-155)
    np.testing.assert_array_equal(host(s), s_ref)


def test_quantize_tensor(ops):
    rng = np.random.default_rng(4)
    x = (rng.standard_normal((8, 512)) * 3).astype(np.float16)
    sc = np.array([17.25], dtype=np.float32)
    assert np.array_equal(host(ops.quantize_tensor(dev(x), dev(sc))), R.quantize_tensor(x, sc[0]))


# ------------------------------------------------------------------------------------------------
def _pack_processed_int4(q_nk):
    from trtllm_llama_b200.quantization import pack_processed_int4
    return pack_processed_int4(torch.from_numpy(np.ascontiguousarray(q_nk))).numpy()


def _wo_inputs(rng, M, N, K, bits):
    w = (rng.random((K, N), dtype=np.float32) * 2 - 1).astype(np.float16)       # test_weight_only_quant_matmul.py:87
    q, scales = R.symmetric_quantize(w, bits)
    x = (rng.random((M, K), dtype=np.float32) * 0.2 - 0.1).astype(np.float16)
    ref = R.weight_only_matmul(x, q, scales)
    qt = np.ascontiguousarray(q.T)                                               # this repo's processed layout [N, K]
    wp = qt if bits == 8 else _pack_processed_int4(qt)
    return x, wp, scales, ref


def _wo_tol(ref, bits):
    # column-wise tolerance far below the reference's 1.5*max/2^(bits-1) (_utils.py:62-89)
    return 2e-3 * np.abs(ref.astype(np.float32)).max() + 1e-3


@pytest.mark.parametrize("M", [1, 2, 4, 7, 8])
@pytest.mark.parametrize("N,K", [(512, 4096), (256, 11008), (100, 4096), (64, 136)])
def test_gemv_f16(ops, M, N, K):
    if M > ops.lib.tb_gemv_max_rows(ops.KIND_F16, K):
        pytest.skip("M > 4 needs the tensor-core GEMV (K multiple of 32)")
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
    w = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
    y = host(ops.gemv(ops.KIND_F16, dev(x), dev(w)))
    ref = R.gemm_f16(x, w)
    np.testing.assert_allclose(y.astype(np.float32), ref.astype(np.float32), rtol=2e-3, atol=2e-3)
    y32 = host(ops.gemv(ops.KIND_F16, dev(x), dev(w), out_fp32=True))
    np.testing.assert_allclose(y32, x.astype(np.float64) @ w.astype(np.float64).T, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("M", [1, 3, 8])
@pytest.mark.parametrize("N,K", [(384, 4096), (200, 11008), (96, 320)])
def test_gemv_weight_only(ops, bits, M, N, K):
    if M > ops.lib.tb_gemv_max_rows(ops.KIND_W8 if bits == 8 else ops.KIND_W4, K):
        pytest.skip("M > 4 needs the tensor-core GEMV")
    rng = np.random.default_rng(6)
    x, wp, scales, ref = _wo_inputs(rng, M, N, K, bits)
    y = host(ops.weight_only_quant_matmul(dev(x), dev(wp), dev(scales), 1 if bits == 8 else 2, use_gemv=True))
    np.testing.assert_allclose(y.astype(np.float32), ref.astype(np.float32), atol=_wo_tol(ref, bits))


def _sq_inputs(rng, M, N, K, per_token, per_channel):
    a = rng.integers(-128, 128, (M, K), dtype=np.int8)
    b = rng.integers(-128, 128, (N, K), dtype=np.int8)
    sa = (rng.integers(1, 10, (M, 1) if per_token else (1, 1)) * 1e-2).astype(np.float32)
    sb = (rng.integers(1, 10, (1, N) if per_channel else (1, 1)) * 1e-2).astype(np.float32)
    return a, b, sa, sb


@pytest.mark.parametrize("per_token,per_channel", [(True, True), (False, True), (True, False), (False, False)])
@pytest.mark.parametrize("M,N,K", [(4, 768, 768), (8, 520, 4096), (3, 64, 11008), (2, 100, 208)])
def test_gemv_sq(ops, per_token, per_channel, M, N, K):
    rng = np.random.default_rng(7)
    a, b, sa, sb = _sq_inputs(rng, M, N, K, per_token, per_channel)
    y = host(ops.smooth_quant_gemm(dev(a), dev(b), dev(sa), dev(sb), per_token, per_channel, use_gemv=True))
    assert np.array_equal(y, R.sq_gemm(a, b, sa, sb, np.float16))   # bit-exact (test_smooth_quant_gemm.py:109)


@pytest.mark.parametrize("M,K,inter", [(2, 1024, 384), (8, 4096, 1000), (1, 136, 24)])
def test_gemv_swiglu(ops, M, K, inter):
    rng = np.random.default_rng(8)
    x = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
    w = (rng.standard_normal((2 * inter, K)) * 0.05).astype(np.float16)
    y = host(ops.gemv(ops.KIND_F16, dev(x), dev(w), swiglu=True))
    gu = R.gemm_f16(x, w)
    ref = R.swiglu(gu[:, :inter], gu[:, inter:])
    np.testing.assert_allclose(y.astype(np.float32), ref.astype(np.float32), rtol=4e-3, atol=2e-3)


@pytest.mark.parametrize("M", [1, 2, 3, 8])
@pytest.mark.parametrize("N,K,swiglu", [(512, 4096, False), (768, 11008, False), (2 * 384, 4096, True), (10, 256, False)])
def test_gemv_fused_rmsnorm(ops, M, N, K, swiglu):
    """prologue 1: the TRT-native rms_norm in front of a projection, fused into the GEMV's activation staging."""
    rng = np.random.default_rng(31)
    h = rng.standard_normal((M, K)).astype(np.float16)
    g = (1 + 0.1 * rng.standard_normal(K)).astype(np.float16)
    w = (rng.standard_normal((N, K)) * 0.03).astype(np.float16)
    y = host(ops.gemv(ops.KIND_F16, dev(h), dev(w), swiglu=swiglu, prologue=ops.PRO_RMS, gamma=dev(g), eps=1e-6))
    ref = R.gemm_f16(R.rmsnorm(h, g, 1e-6), w)
    if swiglu:
        ref = R.swiglu(ref[:, :N // 2], ref[:, N // 2:])
    np.testing.assert_allclose(y.astype(np.float32), ref.astype(np.float32), rtol=4e-3, atol=4e-3)


@pytest.mark.parametrize("M", [1, 4, 8])
@pytest.mark.parametrize("prologue", [2, 3])
def test_gemv_fused_quant_prologues(ops, M, prologue):
    """prologue 2 = RmsnormQuantization, 3 = QuantizePerToken, fused in front of the W8A8 GEMV."""
    rng = np.random.default_rng(32)
    N, K = 640, 4096
    h = rng.standard_normal((M, K)).astype(np.float16)
    g = (1 + 0.1 * rng.standard_normal(K)).astype(np.float16)
    b = rng.integers(-128, 128, (N, K), dtype=np.int8)
    sb = (rng.integers(1, 10, (1, N)) * 1e-3).astype(np.float32)
    r = rng.standard_normal((M, N)).astype(np.float16)
    y = host(ops.gemv(ops.KIND_A8W8, dev(h), dev(b), sc=dev(sb), residual=dev(r), prologue=prologue, gamma=dev(g), eps=1e-6))
    q, st = R.rmsnorm_quant(h, g, 1e-6, dynamic=True) if prologue == 2 else R.quantize_per_token(h)
    ref = R.residual_add(R.sq_gemm(q, b, st, sb, np.float16), r)
    # int8 codes may flip by one where rsqrt rounding differs (see test_rmsnorm_quant): compare at output scale
    tol = 4e-3 * float(np.abs(ref.astype(np.float32)).max()) if prologue == 2 else 0.0
    if prologue == 3:
        assert np.array_equal(y, ref)      # same quantiser arithmetic, integer GEMM: bit-exact
    else:
        np.testing.assert_allclose(y.astype(np.float32), ref.astype(np.float32), atol=tol)


# ------------------------------------------------------------------------------------------------
TC_SHAPES = [(1, 256, 512), (8, 384, 4096), (16, 128, 256), (40, 200, 1376), (128, 512, 1024), (300, 256, 2048)]


@pytest.mark.parametrize("M,N,K", TC_SHAPES)
@pytest.mark.parametrize("splits", [0, 1, 3])
def test_gemm_tc_f16(ops, M, N, K, splits):
    rng = np.random.default_rng(9)
    x = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
    w = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
    y = host(ops.gemm_tc(ops.KIND_F16, dev(x), dev(w), force_splits=splits))
    ref = R.gemm_f16(x, w)
    np.testing.assert_allclose(y.astype(np.float32), ref.astype(np.float32), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("M,N,K", TC_SHAPES)
@pytest.mark.parametrize("per_token,per_channel", [(True, True), (False, False)])
@pytest.mark.parametrize("out", ["float16", "float32", "int32"])
def test_gemm_tc_sq(ops, M, N, K, per_token, per_channel, out):
    if K % 16:
        pytest.skip("int8 rows must be 16-byte aligned for TMA")
    rng = np.random.default_rng(10)
    a, b, sa, sb = _sq_inputs(rng, M, N, K, per_token, per_channel)
    tdt = {"float16": torch.float16, "float32": torch.float32, "int32": torch.int32}[out]
    ndt = {"float16": np.float16, "float32": np.float32, "int32": np.int32}[out]
    y = host(ops.gemm_tc(ops.KIND_A8W8, dev(a), dev(b), sc=dev(sb), sr=dev(sa), out_dtype=tdt))
    assert np.array_equal(y, R.sq_gemm(a, b, sa, sb, ndt))           # bit-exact
    y3 = host(ops.gemm_tc(ops.KIND_A8W8, dev(a), dev(b), sc=dev(sb), sr=dev(sa), out_dtype=tdt, force_splits=3))
    assert np.array_equal(y3, R.sq_gemm(a, b, sa, sb, ndt))


@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("M,N,K", [(1, 256, 512), (8, 384, 4096), (128, 256, 1024), (130, 200, 2752)])
def test_gemm_tc_weight_only(ops, bits, M, N, K):
    rng = np.random.default_rng(11)
    x, wp, scales, ref = _wo_inputs(rng, M, N, K, bits)
    y = host(ops.weight_only_quant_matmul(dev(x), dev(wp), dev(scales), 1 if bits == 8 else 2, use_gemv=False))
    np.testing.assert_allclose(y.astype(np.float32), ref.astype(np.float32), atol=_wo_tol(ref, bits))


@pytest.mark.parametrize("bits", [8, 4])
def test_gemm_tc_weight_only_prefill_dequant_route(ops, bits):
    """Prefill-size weight-only GEMMs dequantise the matrix to fp16 once and run the fp16 kernel: identical bits to the
    fused converter kernel (forced with force_nt), and within tolerance of the oracle."""
    rng = np.random.default_rng(17)
    M, N, K = 2100, 384, 512
    x, wp, scales, ref = _wo_inputs(rng, M, N, K, bits)
    kind = ops.KIND_W8 if bits == 8 else ops.KIND_W4
    r = (rng.standard_normal((M, N))).astype(np.float16)
    auto = host(ops.gemm_tc(kind, dev(x), dev(wp), w_scale=dev(scales)))
    fused = host(ops.gemm_tc(kind, dev(x), dev(wp), w_scale=dev(scales), force_nt=256))
    assert np.array_equal(auto, fused)
    np.testing.assert_allclose(auto.astype(np.float32), ref.astype(np.float32), atol=_wo_tol(ref, bits))
    auto_r = host(ops.gemm_tc(kind, dev(x), dev(wp), w_scale=dev(scales), residual=dev(r)))
    fused_r = host(ops.gemm_tc(kind, dev(x), dev(wp), w_scale=dev(scales), residual=dev(r), force_nt=256))
    assert np.array_equal(auto_r, fused_r)


def test_gemm_tc_residual(ops):
    rng = np.random.default_rng(12)
    M, N, K = 8, 256, 512
    x = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
    w = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
    r = rng.standard_normal((M, N)).astype(np.float16)
    y = host(ops.gemm_tc(ops.KIND_F16, dev(x), dev(w), residual=dev(r)))
    ref = R.residual_add(R.gemm_f16(x, w), r)
    np.testing.assert_allclose(y.astype(np.float32), ref.astype(np.float32), rtol=2e-3, atol=2e-3)


# CTA-pair (cta_group::2) kernel, forced with force_nt=512: ragged token / channel tails, odd tile counts
PAIR_SHAPES = [(300, 384, 256), (256, 256, 128), (1000, 640, 512), (513, 1408, 1024), (16, 128, 256)]


@pytest.mark.parametrize("M,N,K", PAIR_SHAPES)
@pytest.mark.parametrize("per_token,per_channel", [(True, True), (False, False)])
@pytest.mark.parametrize("out", ["float16", "float32", "int32"])
def test_gemm_tc_pair_sq(ops, M, N, K, per_token, per_channel, out):
    rng = np.random.default_rng(13)
    a, b, sa, sb = _sq_inputs(rng, M, N, K, per_token, per_channel)
    tdt = {"float16": torch.float16, "float32": torch.float32, "int32": torch.int32}[out]
    ndt = {"float16": np.float16, "float32": np.float32, "int32": np.int32}[out]
    y = host(ops.gemm_tc(ops.KIND_A8W8, dev(a), dev(b), sc=dev(sb), sr=dev(sa), out_dtype=tdt, force_nt=512))
    assert np.array_equal(y, R.sq_gemm(a, b, sa, sb, ndt))           # bit-exact


@pytest.mark.parametrize("M,N,K", PAIR_SHAPES)
@pytest.mark.parametrize("residual", [False, True])
def test_gemm_tc_pair_f16(ops, M, N, K, residual):
    rng = np.random.default_rng(14)
    x = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
    w = (rng.standard_normal((N, K)) * 0.05).astype(np.float16)
    r = rng.standard_normal((M, N)).astype(np.float16) if residual else None
    y = host(ops.gemm_tc(ops.KIND_F16, dev(x), dev(w), residual=dev(r) if residual else None, force_nt=512))
    ref = R.gemm_f16(x, w)
    if residual:
        ref = R.residual_add(ref, r)
    np.testing.assert_allclose(y.astype(np.float32), ref.astype(np.float32), rtol=2e-3, atol=2e-3)
    # same MMA order and epilogue as the one-CTA kernel: identical bits
    y1 = host(ops.gemm_tc(ops.KIND_F16, dev(x), dev(w), residual=dev(r) if residual else None, force_nt=256))
    assert np.array_equal(y, y1)


def test_gemm_tc_sq_prefill_size_auto_pair(ops):
    """A prefill-size SmoothQuant GEMM (auto-dispatched to the CTA-pair kernel) equals the forced one-CTA result."""
    M, N, K = 8192, 4096 + 128, 512
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(-127, 128, (M, K), device="cuda", dtype=torch.int8, generator=g)
    b = torch.randint(-127, 128, (N, K), device="cuda", dtype=torch.int8, generator=g)
    st = torch.rand(M, 1, device="cuda", generator=g) * 0.01 + 1e-3
    sc = torch.rand(1, N, device="cuda", generator=g) * 0.01 + 1e-3
    auto = ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st)
    one = ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st, force_nt=256)
    assert torch.equal(auto, one)
    i32 = ops.gemm_tc(ops.KIND_A8W8, a[:512], b, sc=torch.ones(1, 1, device="cuda"), sr=torch.ones(1, 1, device="cuda"),
                      out_dtype=torch.int32, force_nt=512)
    assert torch.equal(i32, (a[:512].double() @ b.double().t()).to(torch.int32))


# ------------------------------------------------------------------------------------------------
def _mmha_case(rng, B, H, S_max, past, max_in, in_lens, int8_kv, scale=1.0):
    Dh = 128
    qkv = (rng.standard_normal((B, 3 * H * Dh)) * scale).astype(np.float16)
    if int8_kv:
        cache = rng.integers(-127, 128, (B, 2, H, S_max, Dh), dtype=np.int8)
        s_q = np.float32(127.0 / (4.0 * scale))
        s_dq = np.float32(1.0) / s_q
    else:
        cache = (rng.standard_normal((B, 2, H, S_max, Dh)) * scale).astype(np.float16)
        s_q = s_dq = None
    return qkv, cache, s_q, s_dq


@pytest.mark.parametrize("int8_kv", [True, False])
@pytest.mark.parametrize("past,nsplit", [(37, 1), (200, 1), (200, 3), (1100, 0), (1100, 1), (1, 1), (0, 1)])
def test_mmha_decode(ops, int8_kv, past, nsplit):
    rng = np.random.default_rng(13)
    B, H, Dh, S_max = 2, 4, 128, 1280
    max_in = min(past, 24) if past > 0 else 0
    in_lens = np.array([max_in, max(max_in - 7, 1) if max_in > 0 else 0], dtype=np.int32)
    qkv, cache, s_q, s_dq = _mmha_case(rng, B, H, S_max, past, max_in, in_lens, int8_kv)
    cache_ref = cache.copy()
    ref = R.mmha_decode(qkv, cache_ref, past, in_lens, max_in, num_heads=H, head_size=Dh,
                        kv_scale_orig_quant=s_q, kv_scale_quant_orig=s_dq)
    d_cache = dev(cache)
    kw = {}
    if int8_kv:
        kw = dict(kv_scale_orig_quant=dev(np.array([s_q], np.float32)), kv_scale_quant_orig=dev(np.array([s_dq], np.float32)))
    out = ops.mmha_decode(dev(qkv), d_cache, past, num_heads=H, head_size=Dh, max_input_len=max_in,
                          input_lengths=dev(in_lens), nsplit=nsplit, **kw)
    out = host(out)
    # T/tests/attention/test_gpt_attention.py:828-831 uses atol 2e-3 on outputs of magnitude ~1e-3..1;
    # outputs here are O(1) (int8) so compare relative to the output scale
    tol = 2e-3 * max(1.0, float(np.abs(ref.astype(np.float32)).max()))
    np.testing.assert_allclose(out.astype(np.float32), ref.astype(np.float32), atol=tol)
    new_cache = host(d_cache)
    if int8_kv:   # appended K/V rows: int8 within 1 step (RoPE cos/sin ulp differences), rest untouched
        assert np.abs(new_cache.astype(np.int32) - cache_ref.astype(np.int32)).max() <= 1
        assert (new_cache != cache_ref).mean() < 1e-5
    else:
        np.testing.assert_allclose(new_cache.astype(np.float32), cache_ref.astype(np.float32), atol=2e-3)


@pytest.mark.parametrize("use_tc", [True, False])
@pytest.mark.parametrize("int8_kv", [True, False])
@pytest.mark.parametrize("S,lens", [(64, (64, 64)), (96, (96, 50)), (200, (130, 200)), (128, (1, 128)), (300, (300, 257))])
def test_context_attention(ops, int8_kv, S, lens, use_tc):
    rng = np.random.default_rng(14)
    B, H, Dh, S_max = 2, 2, 128, 320
    qkv = (rng.standard_normal((B, S, 3 * H * Dh)) * 0.5).astype(np.float16)
    in_lens = np.array(lens, dtype=np.int32)
    cache_ref = np.zeros((B, 2, H, S_max, Dh), dtype=np.int8 if int8_kv else np.float16)
    s_q = np.float32(127.0 / 2.5) if int8_kv else None
    ref = R.context_attention(qkv, cache_ref, in_lens, num_heads=H, head_size=Dh, kv_scale_orig_quant=s_q)
    d_cache = dev(np.zeros_like(cache_ref))
    kw = dict(kv_scale_orig_quant=dev(np.array([s_q], np.float32))) if int8_kv else {}
    out = host(ops.context_attention(dev(qkv), d_cache, dev(in_lens), num_heads=H, head_size=Dh, use_tc=use_tc, **kw))
    for b in range(B):   # valid rows only (padded rows are unspecified in the reference)
        L = lens[b]
        np.testing.assert_allclose(out[b, :L].astype(np.float32), ref[b, :L].astype(np.float32), atol=5e-3)
    nc = host(d_cache)
    if int8_kv:
        assert np.abs(nc[:, :, :, :S].astype(np.int32) - cache_ref[:, :, :, :S].astype(np.int32)).max() <= 1
    else:
        np.testing.assert_allclose(nc.astype(np.float32), cache_ref.astype(np.float32), atol=2e-3)


@pytest.mark.parametrize("rows,inter", [(7, 384), (33, 11008), (5, 16384), (3, 8)])
def test_swiglu_quant_fused(ops, rows, inter):
    """One-pass SwiGLU + per-token quantisation == the two separate kernels, bit for bit (and the oracle's quantiser on
    the device's own SwiGLU output; the oracle's SwiGLU itself differs from the device by expf rounding only)."""
    rng = np.random.default_rng(16)
    gu = (rng.standard_normal((rows, 2 * inter)) * 2).astype(np.float16)
    q, sc = ops.swiglu_quant(dev(gu))
    act = ops.swiglu(dev(gu))
    q2, sc2 = ops.quantize_per_token(act)
    assert np.array_equal(host(q), host(q2)) and np.array_equal(host(sc), host(sc2))
    rq, rs = R.quantize_per_token(host(act))
    assert np.array_equal(host(q), rq) and np.array_equal(host(sc).reshape(-1), rs.reshape(-1))


# ------------------------------------------------------------------------------------------------
def test_glue(ops):
    rng = np.random.default_rng(15)
    table = rng.standard_normal((100, 256)).astype(np.float16)
    ids = rng.integers(0, 100, (3, 5)).astype(np.int32)
    assert np.array_equal(host(ops.embedding(dev(ids), dev(table))), table[ids])
    gu = rng.standard_normal((7, 2 * 384)).astype(np.float16)
    np.testing.assert_allclose(host(ops.swiglu(dev(gu))).astype(np.float32),
                               R.swiglu(gu[:, :384], gu[:, 384:]).astype(np.float32), rtol=2e-3, atol=1e-3)
    a = rng.standard_normal((4, 512)).astype(np.float16)
    b = rng.standard_normal((4, 512)).astype(np.float16)
    assert np.array_equal(host(ops.add(dev(a), dev(b))), R.residual_add(a, b))
    logits = rng.standard_normal((5, 32000)).astype(np.float32)
    logits[2, 100] = logits[2, 7] = 99.0   # tie -> lowest index
    assert np.array_equal(host(ops.argmax(dev(logits))), R.greedy_argmax(logits))


@pytest.mark.parametrize("kind", ["sq", "fp16"])
@pytest.mark.parametrize("M,inter,K", [(100, 384, 256), (2048 + 77, 1280, 512), (4096, 11008, 4096)])
def test_gate_up_gemm_with_swiglu_epilogue_is_bit_identical_to_the_two_kernels(ops, kind, M, inter, K):
    """tb_gemm_tc_swiglu (CTA-pair tcgen05 kernel, gate and up accumulated side by side in TMEM, silu(gate) * up in the
    drain) against tb_gemm_tc followed by tb_swiglu on the same operands: identical bits.  inter = 384 / 1280 leave the
    second CTA of the last pair without channels; M = 100 / 2125 leave ragged token tiles."""
    g = torch.Generator(device="cuda").manual_seed(M + inter)
    if kind == "sq":
        x = torch.randint(-127, 128, (M, K), device="cuda", dtype=torch.int8, generator=g)
        w = torch.randint(-127, 128, (2 * inter, K), device="cuda", dtype=torch.int8, generator=g)
        sr = torch.rand((M, 1), device="cuda", generator=g) * 2e-3 + 1e-4
        sc = torch.rand((1, 2 * inter), device="cuda", generator=g) * 2e-3 + 1e-4
        two = ops.swiglu(ops.gemm_tc(ops.KIND_A8W8, x, w, sc=sc, sr=sr))
        one = ops.gemm_tc_swiglu(ops.KIND_A8W8, x, w, sc=sc, sr=sr)
    else:
        x = (torch.randn((M, K), device="cuda", generator=g) * 0.5).half()
        w = (torch.randn((2 * inter, K), device="cuda", generator=g) * 0.05).half()
        two = ops.swiglu(ops.gemm_tc(ops.KIND_F16, x, w))
        one = ops.gemm_tc_swiglu(ops.KIND_F16, x, w)
    assert one.shape == (M, inter)
    assert torch.equal(one.view(torch.int16), two.view(torch.int16))
    assert float(one.float().abs().max()) > 0
