"""Parity at the shapes BASELINE.json's configs actually run (LLaMA-7B: 32 heads x 128, hidden 4096, inter 11008; batch 8;
2048-token contexts; M = 16384 prefill rows) — the multi-wave grids, band rasterisation, persistent tile loops, lazy
softmax rescale and split clamps that the small-shape tests in test_kernels_gpu.py never reach.

Checkers: the numpy oracle on sampled (batch, head) pairs / rows (a full CPU evaluation at these sizes would take hours),
the reference's OWN decode-attention kernel compiled for sm_100a (oracle/_ref/libref_cuda.so), and exact integer
accumulators computed in float64 (|acc| <= 127^2 * 11008 < 2^53) pushed through the oracle's epilogue.
Cases follow T/tests/attention/test_gpt_attention.py:30-73 (shapes), :685-695, :828-831 (tolerances)."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_model as RM  # noqa: E402
from oracle import ref_ops as R  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CUDA = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
H, DH, HID, INTER = 32, 128, 4096, 11008


@pytest.fixture(scope="module")
def ops():
    import trtllm_llama_b200  # noqa: F401
    from trtllm_llama_b200 import ops as o
    assert trtllm_llama_b200.lib.tb_check_device() == 0
    return o


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def P(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


# ------------------------------------------------------------------------------------------------
# (a) prefill attention, cfg4 shape: B = 8, H = 32, S = 2048, ragged input lengths, both cache types
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("int8_kv", [True, False])
def test_flash_context_attention_b8_h32_s2048(ops, int8_kv):
    B, S, S_max = 8, 2048, 2176
    g = torch.Generator(device="cuda").manual_seed(41)
    qkv = (torch.randn(B, S, 3 * HID, device="cuda", generator=g) * 0.5).half()
    lens = np.array([2048, 1, 777, 2047, 1025, 64, 1920, 129], dtype=np.int32)
    qkv0 = qkv.clone()                                    # the kernel rotates q, k in place
    cache = torch.zeros((B, 2, H, S_max, DH), dtype=torch.int8 if int8_kv else torch.float16, device="cuda")
    s_q = np.float32(127.0 / 2.5)
    kw = dict(kv_scale_orig_quant=torch.tensor([s_q], device="cuda")) if int8_kv else {}
    out = ops.context_attention(qkv, cache, torch.from_numpy(lens).cuda(), num_heads=H, head_size=DH, use_tc=True, **kw)
    torch.cuda.synchronize()
    # sampled (b, h): every length class incl. the 1-token and the full-length sequence, first / last / middle heads
    for b, h in ((0, 0), (0, 31), (1, 5), (2, 17), (3, 30), (4, 1), (5, 9), (6, 31), (7, 13)):
        sl = [slice(k * HID + h * DH, k * HID + (h + 1) * DH) for k in range(3)]
        one = torch.cat([qkv0[b:b + 1, :, s] for s in sl], dim=-1).cpu().numpy()          # [1, S, 3*Dh], one head
        c_ref = np.zeros((1, 2, 1, S_max, DH), dtype=np.int8 if int8_kv else np.float16)
        ref = R.context_attention(one, c_ref, lens[b:b + 1], num_heads=1, head_size=DH,
                                  kv_scale_orig_quant=s_q if int8_kv else None)
        L = int(lens[b])
        got = out[b, :L, h * DH:(h + 1) * DH].cpu().numpy().astype(np.float32)
        # T/tests/attention/test_gpt_attention.py:685-695: context outputs atol 5e-3
        np.testing.assert_allclose(got, ref[0, :L].astype(np.float32), atol=5e-3, err_msg=f"b={b} h={h}")
        kv = cache[b, :, h, :S].cpu().numpy()
        if int8_kv:
            d = np.abs(kv.astype(np.int32) - c_ref[0, :, 0, :S].astype(np.int32))
            assert d.max() <= 1 and (d != 0).mean() < 1e-3, f"b={b} h={h}"          # cvt.rni of a cos/sin-ulp-different k
        else:
            np.testing.assert_allclose(kv.astype(np.float32), c_ref[0, :, 0, :S].astype(np.float32), atol=2e-3)


# ------------------------------------------------------------------------------------------------
# (b) decode attention, cfg3 shape: B = 8, H = 32, 2047 cached positions, automatic split count — against the
#     reference's own masked_multihead_attention kernel (fast: no CPU oracle needed) and, on samples, the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", ["fp16", "int8_fma", "int8_mma"])
def test_mmha_decode_b8_h32_l2047_vs_reference_kernel(ops, variant):
    if not os.path.exists(REF_CUDA):
        pytest.skip("oracle/_ref/libref_cuda.so not built")
    ref = C.CDLL(REF_CUDA)
    int8_kv = variant != "fp16"
    B, S_max, past, max_in = 8, 2176, 2047, 1920
    g = torch.Generator(device="cuda").manual_seed(43)
    qkv = torch.randn(B, 3 * HID, device="cuda", generator=g).half()
    if int8_kv:
        cache = torch.randint(-127, 128, (B, 2, H, S_max, DH), device="cuda", dtype=torch.int8, generator=g)
        s_q, s_dq = np.float32(127.0 / 4.0), np.float32(4.0 / 127.0)
    else:
        cache = torch.randn(B, 2, H, S_max, DH, device="cuda", generator=g).half()
        s_q = s_dq = np.float32(1.0)
    in_lens = np.array([1920, 1, 1000, 1919, 512, 77, 1920, 1500], dtype=np.int32)
    masked = np.zeros((B, S_max), dtype=np.int32)
    for b in range(B):
        masked[b, in_lens[b]:max_in] = 1
    d_in, d_mask = torch.from_numpy(in_lens).cuda(), torch.from_numpy(masked).cuda()
    d_seq = torch.full((B,), past, dtype=torch.int32, device="cuda")
    d_sq, d_sdq = torch.tensor([s_q], device="cuda"), torch.tensor([s_dq], device="cuda")

    c_ref = cache.clone()
    out_ref = torch.zeros((B, HID), dtype=torch.float16, device="cuda")
    pad_ws = torch.zeros((B,), dtype=torch.int32, device="cuda")
    rc = ref.ref_mmha_decode_half(P(out_ref), P(qkv), P(c_ref), B, H, DH, S_max, past, max_in, P(d_seq), P(d_in), P(d_mask),
                                  P(pad_ws), P(d_sq), P(d_sdq), int(int8_kv), DH, C.c_float(1.0),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    out_ref_h = host(out_ref).astype(np.float32)

    prev = ops.lib.tb_mmha_set_mode({"fp16": -1, "int8_fma": 0, "int8_mma": 1}[variant])
    try:
        kw = dict(kv_scale_orig_quant=d_sq, kv_scale_quant_orig=d_sdq) if int8_kv else {}
        c_my = cache.clone()
        # (1) masked_tokens given, as the reference plugin receives them; (2) derived on the device from input_lengths,
        # sequence_length and max_input_length read from device memory — what the engine's replayed step graph does
        out1 = ops.mmha_decode(qkv, c_my, past, num_heads=H, head_size=DH, max_input_len=max_in, seq_lens=d_seq,
                               input_lengths=d_in, masked_tokens=d_mask, nsplit=0, **kw)
        c_my2 = cache.clone()
        out2 = ops.mmha_decode(qkv, c_my2, 0, num_heads=H, head_size=DH, max_input_len=0, seq_lens=d_seq,
                               input_lengths=d_in, nsplit=0, len_cap=S_max - 1,
                               max_input_len_dev=torch.tensor([max_in], dtype=torch.int32, device="cuda"), **kw)
    finally:
        ops.lib.tb_mmha_set_mode(prev)
    # T/tests/attention/test_gpt_attention.py:828-831: atol 2e-3 on outputs <= 1; relative to the output scale here
    tol = 2e-3 * max(1.0, float(np.abs(out_ref_h).max()))
    np.testing.assert_allclose(host(out1).astype(np.float32), out_ref_h, atol=tol)
    np.testing.assert_allclose(host(out2).astype(np.float32), out_ref_h, atol=tol)
    for c in (c_my, c_my2):
        if int8_kv:
            d = (c.to(torch.int32) - c_ref.to(torch.int32)).abs()
            assert int(d.max()) <= 1 and float((d != 0).float().mean()) < 1e-6
        else:
            assert float((c.float() - c_ref.float()).abs().max()) <= 2e-3
    # the oracle on two sampled sequences pins the reference kernel at this shape as well
    for b in (1, 3):
        c_or = host(cache[b:b + 1]).copy()
        o = R.mmha_decode(host(qkv[b:b + 1]), c_or, past, in_lens[b:b + 1], max_in, num_heads=H, head_size=DH,
                          kv_scale_orig_quant=s_q if int8_kv else None, kv_scale_quant_orig=s_dq if int8_kv else None)
        np.testing.assert_allclose(o.astype(np.float32), out_ref_h[b:b + 1], atol=tol)


# ------------------------------------------------------------------------------------------------
# (c) SmoothQuant tcgen05 GEMMs at M = 16384 (cfg4), one-CTA and CTA-pair kernels, all output types, bit-exact
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,K", [(3 * HID, HID), (HID, INTER)])
@pytest.mark.parametrize("force_nt", [256, 512], ids=["one_cta", "cta_pair"])
def test_sq_gemm_m16384_bit_exact_on_sampled_rows(ops, N, K, force_nt):
    M = 16384
    g = torch.Generator(device="cuda").manual_seed(47)
    a = torch.randint(-128, 128, (M, K), device="cuda", dtype=torch.int8, generator=g)
    b = torch.randint(-128, 128, (N, K), device="cuda", dtype=torch.int8, generator=g)
    # scale distributions of T/tests/quantization/test_smooth_quant_gemm.py:20-127
    st = (torch.randint(1, 10, (M, 1), device="cuda", generator=g).float() * 1e-2).contiguous()
    sc = (torch.randint(1, 10, (1, N), device="cuda", generator=g).float() * 1e-2).contiguous()
    # rows spread over every 128-row tile band incl. the first / last tile and tile edges
    rows = torch.unique(torch.cat([torch.tensor([0, 1, 127, 128, 255, 256, M - 129, M - 128, M - 1]),
                                   torch.randint(0, M, (247,), generator=torch.Generator().manual_seed(3))])).cuda()
    acc = (a[rows].double() @ b.double().t()).to(torch.int32).cpu().numpy()     # exact: |acc| < 2^53
    st_h, sc_h = st[rows].cpu().numpy(), sc.cpu().numpy()
    for tdt, ndt in ((torch.float16, np.float16), (torch.float32, np.float32), (torch.int32, np.int32)):
        y = ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st, out_dtype=tdt, force_nt=force_nt)
        got = host(y[rows])
        assert np.array_equal(got, R.sq_gemm_epilogue(acc, st_h, sc_h, ndt)), f"{tdt} differs from the oracle"
        del y
    # per-tensor scales, automatic kernel choice
    one = torch.tensor([[0.03]], device="cuda")
    y = ops.gemm_tc(ops.KIND_A8W8, a, b, sc=one, sr=one, out_dtype=torch.float16)
    assert np.array_equal(host(y[rows]), R.sq_gemm_epilogue(acc, np.float32([0.03]), np.float32([0.03]), np.float16))


def test_fp16_and_weight_only_gemm_m16384_sampled_rows(ops):
    """fp16 and weight-only int8 / int4 prefill GEMMs at M = 16384 (persistent tile loop, band rasterisation) against the
    oracle on sampled rows."""
    from trtllm_llama_b200.quantization import pack_processed_int4
    M, N, K = 16384, HID, HID
    g = torch.Generator(device="cuda").manual_seed(53)
    x = (torch.rand(M, K, device="cuda", generator=g) * 0.2 - 0.1).half()
    w = (torch.rand(K, N, device="cuda", generator=g) * 2 - 1).half()          # test_weight_only_quant_matmul.py:87
    rows = torch.tensor([0, 127, 128, 4095, 8191, 8192, 12345, M - 1]).cuda()
    xs = host(x[rows])
    y = ops.gemm_tc(ops.KIND_F16, x, w.t().contiguous())
    ref = R.gemm_f16(xs, host(w.t().contiguous()))
    np.testing.assert_allclose(host(y[rows]).astype(np.float32), ref.astype(np.float32), rtol=2e-3, atol=2e-3)
    wn = host(w)
    for bits in (8, 4):
        q, scales = R.symmetric_quantize(wn, bits)
        qt = np.ascontiguousarray(q.T)
        wp = qt if bits == 8 else pack_processed_int4(torch.from_numpy(qt)).numpy()
        ref = R.weight_only_matmul(xs, q, scales)
        y = ops.weight_only_quant_matmul(x, torch.from_numpy(wp).cuda(), torch.from_numpy(scales).cuda(),
                                         1 if bits == 8 else 2, use_gemv=False)
        tol = 2e-3 * float(np.abs(ref.astype(np.float32)).max()) + 1e-3
        np.testing.assert_allclose(host(y[rows]).astype(np.float32), ref.astype(np.float32), atol=tol)


# ------------------------------------------------------------------------------------------------
# (d) one decoder layer at the 7B dimensions through the engine (plugins -> kernels -> CUDA-graph decode) vs the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,int8_kv", [("fp16", True), ("w8", False), ("w4", True), ("sq", True)])
def test_engine_one_layer_at_7b_dimensions(mode, int8_kv):
    from trtllm_llama_b200 import runtime as rt
    from trtllm_llama_b200.quantization import QuantMode
    cfg = RM.LlamaCfg(hidden=HID, heads=H, inter=INTER, layers=1, vocab=2048)   # small vocabulary: CPU oracle time
    w = RM.random_weights(cfg, seed=7, std=0.02)
    B, S, new = 3, 12, 3
    rng = np.random.default_rng(8)
    ids = rng.integers(3, cfg.vocab, (B, S)).astype(np.int32)
    lens = np.array([S, 7, 1], np.int32)
    for b in range(B):
        ids[b, lens[b]:] = 2
    oracle = RM.OracleLlama(cfg, RM.quantize_model(w, mode), mode, int8_kv, kv_scale=4.0 / 127.0, max_seq_len=S + new)
    ref_ids, ref_logits = oracle.generate(ids, lens, new, return_logits=True)

    qm = {"fp16": QuantMode(0), "w8": QuantMode.use_weight_only(False), "w4": QuantMode.use_weight_only(True),
          "sq": QuantMode.use_smooth_quant(True, True)}[mode]
    if int8_kv:
        qm |= QuantMode.INT8_KV_CACHE
    mc = rt.ModelConfig(vocab_size=cfg.vocab, num_layers=1, num_heads=H, hidden_size=HID, inter_size=INTER, rms_eps=cfg.eps,
                        quant_mode=qm, max_batch_size=B, max_input_len=S, max_output_len=new)
    f = lambda a_: torch.from_numpy(a_).cuda()  # noqa: E731
    tw = {k: f(w[k]) for k in ("vocab_embedding", "ln_f", "lm_head")}
    tw["layers"] = [{k: f(v) for k, v in lw.items()} for lw in w["layers"]]
    sess = rt.GenerationSession(mc, rt.build_engine_tensors(tw, mc, kv_scale=4.0 / 127.0))
    sess.setup(B, S, new)
    logits = [sess.context(torch.from_numpy(ids), torch.from_numpy(lens)).cpu().numpy()]
    for _ in range(new - 1):
        logits.append(sess.step().cpu().numpy())
    got = np.stack(logits, 1)
    got_ids = sess.output_ids(new).cpu().numpy()
    # same bound as tests/test_engine_gpu.py (the reference's own model test uses atol 1e-1)
    tol = (3e-2 if mode == "sq" else 1e-2) * max(1.0, float(np.abs(ref_logits).max()))
    for s in range(new):
        np.testing.assert_allclose(got[:, s], ref_logits[:, s], atol=tol, err_msg=f"step {s}")
        top2 = np.sort(ref_logits[:, s], -1)[:, -2:]
        decided = (top2[:, 1] - top2[:, 0]) > 2 * tol
        assert np.array_equal(got_ids[decided, s], ref_ids[decided, s]), f"greedy ids differ at step {s}"
        if not np.array_equal(got_ids[:, s], ref_ids[:, s]):
            assert s >= 1, "diverged already at the context step"
            return
