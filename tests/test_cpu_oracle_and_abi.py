"""CPU suite (`-m "not gpu"`): pins the oracle on the golden vectors generated from the reference's own code
(tests/golden/make_golden.py), checks the product's build-time quantiser against the oracle, the host logic
(QuantMode, plugin creation / serialisation / shape inference — none of which launch kernels), and that the
C-ABI library loads and exports every symbol include/*.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import ref_model as RM
from oracle import ref_ops as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
REF_HOST = os.path.join(ROOT, "oracle", "_ref", "libref_host.so")


# ------------------------------------------------------------------------------------------------
# oracle vs golden vectors produced by the reference's own oracles / C++ quantiser
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("per_token", [0, 1])
@pytest.mark.parametrize("per_channel", [0, 1])
@pytest.mark.parametrize("dt", ["float16", "float32", "int32"])
def test_oracle_sq_gemm_matches_reference_oracle(per_token, per_channel, dt):
    """gt_matmul_smooth_quant (T/tests/quantization/_utils.py:92-121); the reference asserts with rtol 1e-7."""
    key = f"sq_{per_token}{per_channel}_{dt}"
    got = R.sq_gemm(GOLD["sq_a"], GOLD["sq_b"], GOLD[key + "_sa"], GOLD[key + "_sb"], {"float16": np.float16,
                    "float32": np.float32, "int32": np.int32}[dt])
    ref = GOLD[key]
    if dt == "float32":
        # the reference oracle multiplies by fp32(sa*sb) computed as a matmul of the two scale vectors: identical grouping
        np.testing.assert_allclose(got, ref, rtol=1e-7)
    else:
        assert np.array_equal(got, ref)


def test_oracle_per_token_quant_matches_reference_oracle():
    """gt_quantize_per_token (_utils.py:124-129) evaluates (x*127)/amax, the kernel x*(127/amax)
    (K/quantization.cu:93-117, which the oracle follows and tests/test_ref_cuda_parity.py pins bit-exactly against
    the real kernel): the two roundings may differ by one code on isolated elements."""
    q, s = R.quantize_per_token(GOLD["qpt_x"])
    d = np.abs(q.astype(np.int32) - GOLD["qpt_q"].astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 1e-3
    np.testing.assert_allclose(s.reshape(-1, 1), GOLD["qpt_s"], rtol=1e-6)


@pytest.mark.parametrize("bits", [8, 4])
def test_oracle_symmetric_quantize_matches_reference_cpp(bits):
    """the reference's C++ symmetric_quantize (cutlass_preprocessors.cpp:615-721), run through oracle/_ref/libref_host.so
    when the fixture was generated: bit-exact ints and fp16 scales."""
    q, scales = R.symmetric_quantize(GOLD[f"woq{bits}_w_kn"], bits)
    unp = GOLD[f"woq{bits}_unprocessed"]
    expect = unp if bits == 8 else R.unpack_int4(unp)
    assert np.array_equal(q, expect)
    assert np.array_equal(scales.view(np.uint16), GOLD[f"woq{bits}_scales"].view(np.uint16))


@pytest.mark.parametrize("bits", [8, 4])
def test_oracle_weight_only_matmul_within_reference_tolerance(bits):
    """woq_gt_matmul + woq_assert_colwise_near_eq (_utils.py:36-89): per column atol = 1.5 * max / 2^(bits-1)."""
    q, scales = R.symmetric_quantize(GOLD[f"woq{bits}_w_kn"], bits)
    got = R.weight_only_matmul(GOLD[f"woq{bits}_act"], q, scales).astype(np.float32)
    ref = GOLD[f"woq{bits}_ref"].astype(np.float32)
    for col in range(ref.shape[1]):
        atol = 1.5 * ref[:, col].max() / (1 << (bits - 1))
        np.testing.assert_allclose(got[:, col], ref[:, col], atol=max(atol, 1e-2))
    # and far tighter than the reference asks: the oracle is the same arithmetic up to the fp16 scale rounding
    np.testing.assert_allclose(got, ref, rtol=3e-3, atol=3e-3)


@pytest.mark.skipif(not os.path.exists(REF_HOST), reason="oracle/_ref/libref_host.so not built")
@pytest.mark.parametrize("bits", [8, 4])
def test_oracle_symmetric_quantize_live_against_libref_host(bits):
    lib = C.CDLL(REF_HOST)
    rng = np.random.default_rng(3)
    K, N = 128, 192
    w = (rng.standard_normal((K, N)) * 0.05).astype(np.float16)
    w[:, 5] = 0                                           # all-zero column edge case
    processed = np.zeros(K * N * bits // 8, np.int8)
    unprocessed = np.zeros(K * N * bits // 8, np.int8)
    scales = np.zeros(N, np.float16)
    assert lib.ref_symmetric_quantize(w.view(np.uint16).ctypes.data_as(C.c_void_p), C.c_int64(K), C.c_int64(N), bits,
                                      processed.ctypes.data_as(C.c_void_p), unprocessed.ctypes.data_as(C.c_void_p),
                                      scales.view(np.uint16).ctypes.data_as(C.c_void_p)) == 0
    q, s = R.symmetric_quantize(w, bits)
    exp = unprocessed.reshape(K, N) if bits == 8 else R.unpack_int4(unprocessed.reshape(K, N // 2))
    assert np.array_equal(q, exp)
    assert np.array_equal(s.view(np.uint16), scales.view(np.uint16))


def test_oracle_model_matches_hf_llama():
    """HF LlamaForCausalLM fp32 (the oracle of T/tests/model/test_llama.py:153-354, atol 1e-1 there): context logits and
    one generation step of the fp16 oracle model agree within 2e-2."""
    cfg = RM.LlamaCfg.tiny(layers=2, hidden=256, inter=384, vocab=512)
    w = RM.random_weights(cfg, seed=11, std=0.05)
    ids = GOLD["hf_ids"]
    o = RM.OracleLlama(cfg, RM.quantize_model(w, "fp16"), "fp16", False, max_seq_len=16)
    l0 = o.context(ids, np.full((ids.shape[0],), ids.shape[1], np.int32))
    np.testing.assert_allclose(l0, GOLD["hf_logits0"], atol=2e-2)
    assert np.array_equal(R.greedy_argmax(l0), GOLD["hf_tok0"][:, 0])
    l1 = o.step(GOLD["hf_tok0"][:, 0])
    np.testing.assert_allclose(l1, GOLD["hf_logits1"], atol=2e-2)


def test_oracle_layernorm_quant_vs_torch_layernorm():
    """T/tests/quantization/test_smooth_quant_layer_norm.py:22-112: x = randn(8,128,1024) seed 1997, torch LayerNorm oracle,
    int8 atol 1, dynamic scales atol 1e-2."""
    torch = pytest.importorskip("torch")
    torch.manual_seed(1997)
    x = torch.randn(8, 128, 1024, dtype=torch.float32)
    ln = torch.nn.LayerNorm(1024, eps=1e-5)
    torch.nn.init.normal_(ln.weight, 1.0, 0.1)
    torch.nn.init.normal_(ln.bias, 0.0, 0.1)
    with torch.no_grad():
        ref = ln(x)
    amax = ref.abs().amax(dim=-1, keepdim=True)
    ref_q = (ref * (127.0 / amax)).round().clip(-128, 127).numpy().astype(np.int32)
    q, s = R.layernorm_quant(x.numpy().astype(np.float16), ln.weight.detach().numpy().astype(np.float16),
                             ln.bias.detach().numpy().astype(np.float16), 1e-5)
    assert np.abs(q.astype(np.int32) - ref_q).max() <= 1
    np.testing.assert_allclose(s[..., 0], (amax / 127.0).numpy()[..., 0], atol=1e-2)


def test_oracle_rope_matches_hf_rotary():
    """neox pairing (j, j + Dh/2) with base 10000 == HF LlamaRotaryEmbedding + rotate_half."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 7, 128)).astype(np.float16)
    pos = np.arange(7)[None, :].repeat(3, 0)
    got = R.rope_neox(x, pos).astype(np.float32)
    inv = 1.0 / (10000.0 ** (np.arange(0, 128, 2, dtype=np.float64) / 128))
    ang = pos[..., None] * inv
    cos, sin = np.cos(np.concatenate([ang, ang], -1)), np.sin(np.concatenate([ang, ang], -1))
    xf = x.astype(np.float64)
    rot = np.concatenate([-xf[..., 64:], xf[..., :64]], -1)
    np.testing.assert_allclose(got, xf * cos + rot * sin, atol=2e-3)


def test_oracle_mmha_single_step_equals_context_last_row():
    """decode step t over a cache written by the context oracle == row t of a (t+1)-token context pass (fp16 cache)."""
    rng = np.random.default_rng(9)
    B, H, Dh, S = 2, 2, 128, 10
    qkv = (rng.standard_normal((B, S, 3 * H * Dh)) * 0.5).astype(np.float16)
    lens = np.array([S, S], np.int32)
    cache_full = np.zeros((B, 2, H, 16, Dh), np.float16)
    full = R.context_attention(qkv.copy(), cache_full, lens, num_heads=H, head_size=Dh)
    cache = np.zeros((B, 2, H, 16, Dh), np.float16)
    R.context_attention(qkv[:, :S - 1].copy(), cache, lens - 1, num_heads=H, head_size=Dh)
    step = R.mmha_decode(qkv[:, S - 1], cache, S - 1, lens - 1, S - 1, num_heads=H, head_size=Dh)
    np.testing.assert_allclose(step.astype(np.float32), full[:, S - 1].astype(np.float32), atol=3e-3)
    np.testing.assert_allclose(cache[:, :, :, :S].astype(np.float32), cache_full[:, :, :, :S].astype(np.float32), atol=1e-3)


# ------------------------------------------------------------------------------------------------
# product host logic
# ------------------------------------------------------------------------------------------------
def test_product_quantiser_matches_oracle_bit_exact():
    torch = pytest.importorskip("torch")
    from trtllm_llama_b200 import quantization as Q
    rng = np.random.default_rng(21)
    w_nk = (rng.standard_normal((192, 256)) * 0.04).astype(np.float16)
    for bits, qt in ((8, torch.int8), (4, torch.quint4x2)):
        q, s = R.symmetric_quantize(np.ascontiguousarray(w_nk.T), bits)
        unp, proc, sc = Q._symmetric_quantize_last_axis_of_batched_matrix(torch.from_numpy(w_nk).t(), qt)
        q_nk = np.ascontiguousarray(q.T)
        if bits == 8:
            assert np.array_equal(proc.numpy(), q_nk)
        else:
            # processed int4: nibble positions of every 8-element group hold elements (0,2,4,6,1,3,5,7)
            assert np.array_equal(Q.unpack_processed_int4(proc).numpy(), q_nk)
            grp = q_nk.reshape(q_nk.shape[0], -1, 8)[:, :, [0, 2, 4, 6, 1, 3, 5, 7]].reshape(q_nk.shape)
            assert np.array_equal(proc.numpy(), R.pack_int4(grp))
        assert np.array_equal(unp.numpy(), q if bits == 8 else R.pack_int4(q))
        assert np.array_equal(sc.numpy().view(np.uint16), s.view(np.uint16))
        # round trip helpers (thop/weightOnlyQuantOp.cpp:347-356)
        assert np.array_equal(Q.preprocess_weights_for_mixed_gemm(unp, qt).numpy(), proc.numpy())
    p4 = Q.pack_int8_tensor_to_packed_int4(torch.from_numpy(np.clip(q, -8, 7)))
    assert np.array_equal(Q.unpack_int4_packed_tensor_to_int8(p4).numpy(), np.clip(q, -8, 7))
    qs, ss = Q.quantize_per_channel_int8(torch.from_numpy(w_nk))
    ref = RM.quantize_linear(w_nk, "sq")
    assert np.array_equal(qs.numpy(), ref["q"]) and np.array_equal(ss.numpy(), ref["scale_ch"])
    with pytest.raises(ValueError):
        Q.symmetric_quantize_last_axis_of_batched_matrix(torch.from_numpy(w_nk), torch.float16)


def test_quant_mode_flags():
    """T/tensorrt_llm/quantization/mode.py:4-137 behaviour used by build.py (LQ/build.py:276-324)."""
    from trtllm_llama_b200.quantization import QuantMode
    assert QuantMode.use_weight_only().is_int8_weight_only() and not QuantMode.use_weight_only().is_int4_weight_only()
    assert QuantMode.use_weight_only(True).is_int4_weight_only()
    sq = QuantMode.use_smooth_quant(per_token=True, per_channel=True)
    assert sq.has_act_and_weight_quant() and sq.has_per_token_dynamic_scaling() and sq.has_per_channel_scaling()
    assert not sq.is_weight_only() and not sq.has_act_static_scaling()
    assert (QuantMode(0) | QuantMode.INT8_KV_CACHE).has_int8_kv_cache()
    assert not QuantMode(0).has_any_quant()
    with pytest.raises(ValueError):
        QuantMode.from_description(quantize_weights=False, quantize_activations=True)


def test_shard_weights_follow_megatron_rules():
    torch = pytest.importorskip("torch")
    from trtllm_llama_b200.runtime import shard_weights
    cfg = RM.LlamaCfg.tiny(layers=1, hidden=256, inter=384, vocab=512)
    w = RM.random_weights(cfg, seed=1)
    tw = {k: torch.from_numpy(w[k]) for k in ("vocab_embedding", "ln_f", "lm_head")}
    tw["layers"] = [{k: torch.from_numpy(v) for k, v in w["layers"][0].items()}]
    parts = [shard_weights(tw, 2, r, cfg.heads) for r in range(2)]
    q, k, v = np.split(w["layers"][0]["qkv"], 3, axis=0)
    for r in range(2):
        lw = parts[r]["layers"][0]
        exp = np.concatenate([q[r * 128:(r + 1) * 128], k[r * 128:(r + 1) * 128], v[r * 128:(r + 1) * 128]], 0)
        assert np.array_equal(lw["qkv"].numpy(), exp)                       # heads split per Q/K/V (weight.py:95-100)
        assert np.array_equal(lw["dense"].numpy(), w["layers"][0]["dense"][:, r * 128:(r + 1) * 128])   # input dim
        assert np.array_equal(lw["gate"].numpy(), w["layers"][0]["gate"][r * 192:(r + 1) * 192])
        assert np.array_equal(lw["down"].numpy(), w["layers"][0]["down"][:, r * 192:(r + 1) * 192])
        assert np.array_equal(parts[r]["lm_head"].numpy(), w["lm_head"][r * 256:(r + 1) * 256])
    # row-parallel partial sums add up to the unsharded projection
    x = np.random.default_rng(0).standard_normal((3, 256)).astype(np.float32)
    full = x @ w["layers"][0]["dense"].astype(np.float32).T
    s = sum(x[:, r * 128:(r + 1) * 128] @ parts[r]["layers"][0]["dense"].numpy().astype(np.float32).T for r in range(2))
    np.testing.assert_allclose(s, full, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------
# C ABI: the library loads and exports every declared symbol; plugin host logic works without a GPU
# ------------------------------------------------------------------------------------------------
def _declared_symbols():
    names = set()
    for h in ("trtllm_b200.h", "trtllm_b200_plugin.h", "trtllm_b200_runtime.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(tb[pr]?t?_[a-z0-9_]+)\s*\(", src))
    return {n for n in names if not n.endswith("_t")}


def test_library_exports_every_declared_symbol():
    from trtllm_llama_b200 import _lib
    h = _lib.load_library()
    declared = _declared_symbols()
    assert len(declared) > 60
    missing = [n for n in sorted(declared) if not hasattr(h, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    for n in ("initLibNvInferPlugins", "getPluginRegistry", "getInferLibVersion"):     # P/exports.map:19-31
        assert hasattr(h, n)
    not_bound = declared - set(_lib.SIGNATURES)
    assert not not_bound, f"declared but missing a ctypes signature: {sorted(not_bound)}"
    assert b"sm_100a" in h.tb_version()


def _field(name, value, kind):
    from trtllm_llama_b200._lib import TbpField
    arr = {"i32": (C.c_int32 * 1), "f32": (C.c_float * 1), "i8": (C.c_int8 * 1)}[kind](value)
    return TbpField(name.encode(), C.cast(arr, C.c_void_p), {"i32": 5, "f32": 1, "i8": 3}[kind], 1), arr


def _create(lib, name, fields):
    from trtllm_llama_b200._lib import TbpField
    keep = [_field(*f) for f in fields]
    arr = (TbpField * len(keep))(*[k[0] for k in keep])
    return lib.tbp_create(name.encode(), b"1", b"tensorrt_llm", arr, len(keep))


def test_plugin_registry_and_host_side_contracts():
    """creators are registered under ("tensorrt_llm", name, "1") (T/tensorrt_llm/functional.py:2828-2830); creation,
    serialisation round trip, output shapes / dtypes and format checks need no device."""
    from trtllm_llama_b200 import _lib
    from trtllm_llama_b200._lib import TbpDims, TbpTensorDesc
    lib = _lib.load_library()
    assert lib.tbp_init(b"tensorrt_llm") == 0
    names = {lib.tbp_creator_name(i).decode() for i in range(lib.tbp_num_creators())}
    assert {"GPTAttention", "SmoothQuantGemm", "WeightOnlyQuantMatmul", "RmsnormQuantization", "LayernormQuantization",
            "QuantizePerToken", "QuantizeTensor", "AllReduce", "AllGather", "Gemm"} <= names
    fields = (C.c_char_p * 32)()
    n = lib.tbp_creator_fields(b"GPTAttention", fields, 32)
    got = {fields[i].decode() for i in range(n)}
    assert {"num_heads", "head_size", "unidirectional", "q_scaling", "rotary_embedding_dim", "neox_rotary_style",
            "context_fmha_type", "multi_block_mode", "multi_query_mode", "int8_kv_cache", "fp8_kv_cache",
            "remove_input_padding", "mask_type", "paged_kv_cache", "type_id", "in_flight_batching"} <= got

    attn = _create(lib, "GPTAttention", [("num_heads", 32, "i32"), ("head_size", 128, "i32"), ("unidirectional", 1, "i32"),
                                         ("q_scaling", 1.0, "f32"), ("rotary_embedding_dim", 128, "i32"),
                                         ("neox_rotary_style", 1, "i8"), ("int8_kv_cache", 1, "i32"), ("type_id", 1, "i32")])
    assert attn
    assert lib.tbp_type(attn) == b"GPTAttention" and lib.tbp_version(attn) == b"1" and lib.tbp_namespace(attn) == b"tensorrt_llm"
    assert lib.tbp_nb_outputs(attn) == 2
    dims = (TbpDims * 10)()
    shapes = [(8, 1, 12288), (8, 2, 32, 2560, 128), (8,), (2,), (8, 2560), (8,), (2048,), (8, 1, 2560), (1,), (1,)]
    for d, shp in zip(dims, shapes):
        d.nb_dims = len(shp)
        for j, v in enumerate(shp):
            d.d[j] = v
    out = TbpDims()
    assert lib.tbp_output_dims(attn, 0, dims, 10, C.byref(out)) == 0
    assert [out.d[i] for i in range(out.nb_dims)] == [8, 1, 4096]                 # last dim = num_heads * head_size
    assert lib.tbp_output_dims(attn, 1, dims, 10, C.byref(out)) == 0
    assert [out.d[i] for i in range(out.nb_dims)] == [8, 2, 32, 2560, 128]        # present KV == past KV dims
    types = (C.c_int32 * 10)(1, 2, 3, 3, 3, 3, 3, 3, 0, 0)
    assert lib.tbp_output_dtype(attn, 0, types, 10) == 1 and lib.tbp_output_dtype(attn, 1, types, 10) == 2
    io = (TbpTensorDesc * 12)()
    for i, t in enumerate([1, 2, 3, 3, 3, 3, 3, 3, 0, 0, 1, 2]):
        io[i].type, io[i].format = t, 0
    assert all(lib.tbp_supports_format(attn, pos, io, 10, 2) for pos in range(12))
    io[1].type = 1
    assert not lib.tbp_supports_format(attn, 1, io, 10, 2)                        # int8 KV requires an int8 cache tensor
    # serialise -> deserialise -> identical blob
    size = lib.tbp_serialization_size(attn)
    buf = (C.c_char * size)()
    lib.tbp_serialize(attn, buf)
    again = lib.tbp_deserialize(b"GPTAttention", b"1", b"tensorrt_llm", buf, size)
    assert again
    buf2 = (C.c_char * size)()
    lib.tbp_serialize(again, buf2)
    assert bytes(buf) == bytes(buf2)
    clone = lib.tbp_clone(again)
    assert lib.tbp_serialization_size(clone) == size
    for p in (attn, again, clone):
        lib.tbp_destroy(p)

    # unsupported configurations are rejected at creation with NULL (creators never throw)
    assert not _create(lib, "GPTAttention", [("num_heads", 32, "i32"), ("head_size", 64, "i32"), ("type_id", 1, "i32")])
    assert not _create(lib, "WeightOnlyQuantMatmul", [("type_id", 0, "i32"), ("weight_type_id", 1, "i32")])
    assert not lib.tbp_create(b"NoSuchPlugin", b"1", b"tensorrt_llm", None, 0)

    wo = _create(lib, "WeightOnlyQuantMatmul", [("type_id", 1, "i32"), ("weight_type_id", 2, "i32")])
    d2 = (TbpDims * 3)()
    for d, shp in zip(d2, [(8, 1, 4096), (4096, 1536), (12288,)]):       # int4: declared weight [K, N/8]
        d.nb_dims = len(shp)
        for j, v in enumerate(shp):
            d.d[j] = v
    assert lib.tbp_output_dims(wo, 0, d2, 3, C.byref(out)) == 0
    assert [out.d[i] for i in range(out.nb_dims)] == [8, 1, 12288]
    lib.tbp_destroy(wo)

    sqp = _create(lib, "SmoothQuantGemm", [("has_per_channel_scaling", 1, "i32"), ("has_per_token_scaling", 1, "i32"),
                                           ("type_id", 3, "i32")])
    assert sqp and lib.tbp_output_dtype(sqp, 0, (C.c_int32 * 4)(2, 0, 0, 0), 4) == 3      # int32 output type
    lib.tbp_destroy(sqp)

    rq = _create(lib, "RmsnormQuantization", [("eps", 1e-6, "f32"), ("dyn_act_scaling", 1, "i32"), ("type_id", 1, "i32")])
    assert rq and lib.tbp_nb_outputs(rq) == 2
    d4 = (TbpDims * 4)()
    for d, shp in zip(d4, [(8, 128, 4096), (4096,), (4096,), (1,)]):
        d.nb_dims = len(shp)
        for j, v in enumerate(shp):
            d.d[j] = v
    assert lib.tbp_output_dims(rq, 1, d4, 4, C.byref(out)) == 0
    assert [out.d[i] for i in range(out.nb_dims)] == [8, 128, 1]                  # per-token scale [..., 1]
    lib.tbp_destroy(rq)


def test_product_path_fails_loudly_without_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from trtllm_llama_b200 import _lib, runtime as rt
    assert _lib.load_library().tb_check_device() != 0
    with pytest.raises(RuntimeError):
        rt.GenerationSession(rt.ModelConfig(), {})


def test_pad_finished_after_end_id():
    """GenerationSession.decode(..., sampling_config): positions after a sequence's first end_id hold end_id."""
    import torch
    import trtllm_llama_b200  # noqa: F401
    from trtllm_llama_b200.runtime import pad_finished
    ids = torch.tensor([[5, 2, 7, 8, 2, 9], [4, 6, 8, 1, 3, 5], [2, 9, 9, 9, 9, 9], [7, 7, 7, 7, 7, 2]], dtype=torch.int32)
    got = pad_finished(ids.clone(), 2)
    assert got.tolist() == [[5, 2, 2, 2, 2, 2], [4, 6, 8, 1, 3, 5], [2, 2, 2, 2, 2, 2], [7, 7, 7, 7, 7, 2]]


def test_host_side_sizing_functions():
    """Host-only entry points of the kernel ABI (no device work): workspace sizing and launch heuristics."""
    from trtllm_llama_b200 import _lib
    h = _lib.load_library()
    # decode rows per kind: tensor-core GEMV takes up to 8 rows when K is a whole number of k-steps, else the FMA kernel's 4
    for kind in (0, 1, 2, 3):
        assert h.tb_gemv_max_rows(kind, 4096) == 8 and h.tb_gemv_max_rows(kind, 11008) == 8
    assert h.tb_gemv_max_rows(0, 136) == 4
    # GEMM scratch: split-K partials for decode sizes; from 2048 rows on also one weight-only matrix expanded to fp16
    small = h.tb_gemm_tc_workspace_bytes(8, 4096, 4096)
    assert 0 < small < 64 << 20
    for (N, K) in ((12288, 4096), (22016, 4096), (4096, 11008)):
        assert h.tb_gemm_tc_workspace_bytes(2048, N, K) >= N * K * 2
        assert h.tb_gemm_tc_workspace_bytes(16384, N, K) >= N * K * 2
        assert h.tb_gemm_tc_workspace_bytes(128, N, K) < N * K * 2
    assert h.tb_gemm_tc_counter_bytes() >= 4096 * 4
    # decode attention: the splits of one (batch, head) form a thread-block cluster (<= 8), >= 64 cached keys each
    for (b, heads, length) in ((1, 32, 130), (1, 32, 4096), (8, 32, 2048), (64, 32, 2048), (1, 4, 100000)):
        n = h.tb_mmha_num_splits(b, heads, length, 32)
        assert 1 <= n <= 8 and (n == 1 or length // n >= 64)
    assert h.tb_mmha_num_splits(1, 32, 10, 32) == 1
    # prefill attention consumes V in place: nominal scratch only
    assert h.tb_context_attention_workspace_bytes(8, 2048, 32) == 256
