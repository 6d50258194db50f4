"""The plugin boundary itself, driven the way TensorRT drives it (through the C view of the virtual calls,
include/trtllm_b200_plugin.h): creator lookup by (name, "1", "tensorrt_llm") -> createPlugin(fields) ->
getOutputDimensions / getWorkspaceSize -> enqueue(inputDesc, outputDesc, inputs, outputs, workspace, stream), with the
reference's field names and input order and WITHOUT any of this library's extension fields (reference behaviour).
Results are compared with the oracle; every plugin is also serialised, deserialised and enqueued again."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_ops as R  # noqa: E402

DT = {"float32": 0, "float16": 1, "int8": 2, "int32": 3}
TORCH_DT = {torch.float32: 0, torch.float16: 1, torch.int8: 2, torch.int32: 3}


@pytest.fixture(scope="module")
def L():
    from trtllm_llama_b200 import _lib
    lib = _lib.load_library()
    assert lib.tbp_init(b"tensorrt_llm") == 0
    return lib


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


class Plugin:
    """minimal 'TensorRT' for one plugin instance"""

    def __init__(self, lib, name, fields):
        from trtllm_llama_b200._lib import TbpField
        self.lib, self.name = lib, name
        self._keep = []
        arr = (TbpField * max(1, len(fields)))()
        for i, (fname, value, kind) in enumerate(fields):
            buf = {"i32": C.c_int32, "f32": C.c_float, "i8": C.c_int8}[kind](value)
            self._keep.append(buf)
            arr[i] = TbpField(fname.encode(), C.cast(C.pointer(buf), C.c_void_p), {"i32": 5, "f32": 1, "i8": 3}[kind], 1)
        self.h = lib.tbp_create(name.encode(), b"1", b"tensorrt_llm", arr, len(fields))
        assert self.h, f"createPlugin({name}) returned NULL"
        assert lib.tbp_initialize(self.h) == 0

    def roundtrip(self):
        """serialize -> deserializePlugin -> clone: the object TensorRT would hold at run time"""
        n = self.lib.tbp_serialization_size(self.h)
        buf = (C.c_char * max(n, 1))()
        self.lib.tbp_serialize(self.h, buf)
        h2 = self.lib.tbp_deserialize(self.name.encode(), b"1", b"tensorrt_llm", buf, n)
        assert h2
        h3 = self.lib.tbp_clone(h2)
        self.lib.tbp_destroy(h2)
        self.lib.tbp_destroy(self.h)
        self.h = h3
        assert self.lib.tbp_initialize(self.h) == 0
        return self

    def _descs(self, tensors, shapes=None):
        from trtllm_llama_b200._lib import TbpTensorDesc
        d = (TbpTensorDesc * len(tensors))()
        for i, t in enumerate(tensors):
            shp = shapes[i] if shapes and shapes[i] is not None else tuple(t.shape)
            d[i].dims.nb_dims = len(shp)
            for j, v in enumerate(shp):
                d[i].dims.d[j] = v
            d[i].type = TORCH_DT[t.dtype]
            d[i].format, d[i].scale = 0, 1.0
        return d

    def run(self, inputs, out_dtypes, in_shapes=None, in_types=None, host_inputs=()):
        """inputs: list of CUDA tensors (or host tensors for indices in host_inputs, or None); returns output tensors
        shaped by getOutputDimensions and typed by getOutputDataType."""
        from trtllm_llama_b200._lib import TbpDims
        lib = self.lib
        n_in = len(inputs)
        placeholder = torch.zeros(1, dtype=torch.int32)
        tens = [t if t is not None else placeholder for t in inputs]
        idesc = self._descs(tens, in_shapes)
        if in_types:
            for i, ty in in_types.items():
                idesc[i].type = ty
        dims = (TbpDims * n_in)(*[idesc[i].dims for i in range(n_in)])
        types = (C.c_int32 * n_in)(*[idesc[i].type for i in range(n_in)])
        n_out = lib.tbp_nb_outputs(self.h)
        outs = []
        for o in range(n_out):
            od = TbpDims()
            assert lib.tbp_output_dims(self.h, o, dims, n_in, C.byref(od)) == 0
            ty = lib.tbp_output_dtype(self.h, o, types, n_in)
            tdt = {0: torch.float32, 1: torch.float16, 2: torch.int8, 3: torch.int32}[ty]
            assert tdt == out_dtypes[o], f"output {o}: plugin says {tdt}, expected {out_dtypes[o]}"
            outs.append(torch.zeros([od.d[j] for j in range(od.nb_dims)], dtype=tdt, device="cuda"))
        odesc = self._descs(outs)
        ws_bytes = lib.tbp_workspace_size(self.h, idesc, n_in, odesc, n_out)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device="cuda")
        ws.fill_(0xAB)          # TensorRT workspaces are not zeroed
        in_ptrs = (C.c_void_p * n_in)(*[None if inputs[i] is None else inputs[i].data_ptr() for i in range(n_in)])
        out_ptrs = (C.c_void_p * n_out)(*[t.data_ptr() for t in outs])
        rc = lib.tbp_enqueue(self.h, idesc, odesc, in_ptrs, out_ptrs, ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, f"{self.name}::enqueue returned {rc}"
        torch.cuda.synchronize()
        return outs

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.tbp_destroy(self.h)
            self.h = None


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("out", ["float16", "float32", "int32"])
@pytest.mark.parametrize("M", [3, 32])
def test_smooth_quant_gemm_plugin(L, out, M):
    """T/tests/quantization/test_smooth_quant_gemm.py: bit-exact against int32 matmul x (scale_a (x) scale_b)."""
    rng = np.random.default_rng(41)
    N, K = 384, 768
    a = rng.integers(-128, 128, (2, M, K), dtype=np.int8)             # leading dims are flattened into M
    b = rng.integers(-128, 128, (N, K), dtype=np.int8)
    sa = (rng.integers(1, 10, (2 * M, 1)) * 1e-2).astype(np.float32)
    sb = (rng.integers(1, 10, (1, N)) * 1e-2).astype(np.float32)
    p = Plugin(L, "SmoothQuantGemm", [("has_per_channel_scaling", 1, "i32"), ("has_per_token_scaling", 1, "i32"),
                                      ("type_id", DT[out], "i32")]).roundtrip()
    tdt = {"float16": torch.float16, "float32": torch.float32, "int32": torch.int32}[out]
    # weight is declared fp32 [N, K/4] by the reference ("workaround for trt not supporting int8 inputs in plugins")
    (c,) = p.run([dev(a), dev(b), dev(sa), dev(sb)], [tdt], in_shapes=[None, (N, K // 4), None, None], in_types={1: 0})
    ref = R.sq_gemm(a.reshape(-1, K), b, sa, sb, {"float16": np.float16, "float32": np.float32, "int32": np.int32}[out])
    assert c.shape == (2, M, N)
    assert np.array_equal(c.cpu().numpy().reshape(-1, N), ref)


@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("M", [1, 8, 40])
def test_weight_only_quant_matmul_plugin(L, bits, M):
    """T/tests/quantization/test_weight_only_quant_matmul.py:85-122 (column tolerance 1.5 * max / 2^(bits-1)), with
    weights produced by the product's symmetric_quantize_last_axis_of_batched_matrix."""
    from trtllm_llama_b200.quantization import symmetric_quantize_last_axis_of_batched_matrix as sq
    torch.manual_seed(0)
    N, K = 256, 4096
    w_kn = torch.rand((K, N), dtype=torch.float16) * 2 - 1.0
    processed, scales = sq(w_kn, torch.int8 if bits == 8 else torch.quint4x2)
    x = (torch.rand((M, K), dtype=torch.float16) * 2 - 1.0) * 0.1
    p = Plugin(L, "WeightOnlyQuantMatmul", [("type_id", 1, "i32"), ("weight_type_id", 1 if bits == 8 else 2, "i32")]).roundtrip()
    pack = 4 if bits == 8 else 8
    (y,) = p.run([x.cuda(), processed.cuda(), scales.cuda()], [torch.float16], in_shapes=[None, (K, N // pack), None],
                 in_types={1: 0})
    q, s = R.symmetric_quantize(w_kn.numpy(), bits)
    ref = R.weight_only_matmul(x.numpy(), q, s).astype(np.float32)
    got = y.cpu().numpy().astype(np.float32)
    for col in range(0, N, 17):
        atol = 1.5 * np.abs(ref[:, col]).max() / (1 << (bits - 1))
        np.testing.assert_allclose(got[:, col], ref[:, col], atol=max(atol, 1e-2))
    np.testing.assert_allclose(got, ref, atol=3e-3 * np.abs(ref).max() + 1e-3)


@pytest.mark.parametrize("rms,dynamic", [(True, True), (True, False), (False, True), (False, False)])
def test_norm_quantization_plugins(L, rms, dynamic):
    """RmsnormQuantization (new) and LayernormQuantization (T/tests/quantization/test_smooth_quant_layer_norm.py:
    int8 atol 1, dynamic scales atol 1e-2): same fields, same four inputs."""
    rng = np.random.default_rng(1997)
    x = rng.standard_normal((4, 16, 1024)).astype(np.float16)
    g = (1 + 0.1 * rng.standard_normal(1024)).astype(np.float16)
    b = (0.1 * rng.standard_normal(1024)).astype(np.float16) if not rms else np.zeros(1024, np.float16)
    scale = np.array([20.0], np.float32)
    p = Plugin(L, "RmsnormQuantization" if rms else "LayernormQuantization",
               [("eps", 1e-5, "f32"), ("use_diff_of_squares", 0, "i32"), ("dyn_act_scaling", int(dynamic), "i32"),
                ("type_id", 1, "i32")]).roundtrip()
    outs = p.run([dev(x), dev(g), dev(b), dev(scale)], [torch.int8, torch.float32] if dynamic else [torch.int8])
    if rms:
        q_ref, s_ref = R.rmsnorm_quant(x.reshape(-1, 1024), g, 1e-5, scale[0], dynamic)
    else:
        q_ref, s_ref = R.layernorm_quant(x.reshape(-1, 1024), g, b, 1e-5, scale[0], dynamic)
    q = outs[0].cpu().numpy().reshape(-1, 1024)
    assert outs[0].shape == (4, 16, 1024)
    assert np.abs(q.astype(np.int32) - q_ref.astype(np.int32)).max() <= 1
    if dynamic:
        assert outs[1].shape == (4, 16, 1)
        np.testing.assert_allclose(outs[1].cpu().numpy().reshape(-1, 1), s_ref, atol=1e-2, rtol=1e-3)


def test_quantize_plugins(L):
    rng = np.random.default_rng(42)
    x = rng.standard_normal((3, 5, 512)).astype(np.float16)
    q, s = Plugin(L, "QuantizePerToken", []).roundtrip().run([dev(x)], [torch.int8, torch.float32])
    q_ref, s_ref = R.quantize_per_token(x)
    assert q.shape == (3, 5, 512) and s.shape == (3, 5, 1)
    assert np.array_equal(q.cpu().numpy(), q_ref) and np.array_equal(s.cpu().numpy(), s_ref)     # bit-exact
    sc = np.array([11.5], np.float32)
    (q2,) = Plugin(L, "QuantizeTensor", []).roundtrip().run([dev(x), dev(sc)], [torch.int8])
    assert np.array_equal(q2.cpu().numpy(), R.quantize_tensor(x, sc[0]))


def test_gemm_plugin(L):
    rng = np.random.default_rng(43)
    x = (rng.standard_normal((2, 20, 512)) * 0.5).astype(np.float16)
    w = (rng.standard_normal((320, 512)) * 0.05).astype(np.float16)
    p = Plugin(L, "Gemm", [("transa", 0, "i32"), ("transb", 1, "i32"), ("type_id", 1, "i32")]).roundtrip()
    (y,) = p.run([dev(x), dev(w)], [torch.float16])
    ref = R.gemm_f16(x.reshape(-1, 512), w).reshape(2, 20, 320)
    np.testing.assert_allclose(y.cpu().numpy().astype(np.float32), ref.astype(np.float32), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("int8_kv", [True, False])
def test_gpt_attention_plugin_context_then_generation(L, int8_kv):
    """The reference's own protocol (T/tests/attention/test_gpt_attention.py:580-836): one context call with
    past_key_value_length = [0, 1] (HOST tensor), then generation calls with [max_input_len + step, 0]; KV cache updated
    in place through the aliased input 1 / output 1; half of each sequence is padding."""
    rng = np.random.default_rng(44)
    B, H, Dh, S, S_max, steps = 2, 4, 128, 24, 64, 5
    hidden = H * Dh
    in_lens = np.array([S, S // 2], np.int32)
    fields = [("num_heads", H, "i32"), ("head_size", Dh, "i32"), ("unidirectional", 1, "i32"), ("q_scaling", 1.0, "f32"),
              ("rotary_embedding_dim", Dh, "i32"), ("neox_rotary_style", 1, "i8"), ("context_fmha_type", 0, "i8"),
              ("multi_block_mode", 0, "i8"), ("multi_query_mode", 0, "i8"), ("int8_kv_cache", int(int8_kv), "i32"),
              ("fp8_kv_cache", 0, "i32"), ("remove_input_padding", 0, "i8"), ("mask_type", 1, "i32"),
              ("paged_kv_cache", 0, "i32"), ("type_id", 1, "i32"), ("in_flight_batching", 0, "i32")]
    p = Plugin(L, "GPTAttention", fields).roundtrip()
    kv_dt = torch.int8 if int8_kv else torch.float16
    cache = torch.zeros((B, 2, H, S_max, Dh), dtype=kv_dt, device="cuda")
    cache_ref = np.zeros((B, 2, H, S_max, Dh), np.int8 if int8_kv else np.float16)
    s_q, s_dq = (np.float32(127.0 / 3.0), np.float32(3.0 / 127.0)) if int8_kv else (None, None)
    masked = np.zeros((B, S_max), np.int32)
    for b in range(B):
        masked[b, in_lens[b]:S] = 1
    common = dict(lens=dev(in_lens), masked=dev(masked), max_in=torch.zeros(S, dtype=torch.int32, device="cuda"),
                  indir=torch.zeros((B, 1, S_max), dtype=torch.int32, device="cuda"))
    scales = [dev(np.array([s_q], np.float32)), dev(np.array([s_dq], np.float32))] if int8_kv else []

    def call(qkv, seq_len_value, host_len):
        seq = torch.full((B,), seq_len_value, dtype=torch.int32, device="cuda")
        host = torch.tensor(host_len, dtype=torch.int32)           # HOST tensor (gptAttentionPlugin.cpp:261-278)
        ins = [dev(qkv), cache, seq, host, common["masked"], common["lens"], common["max_in"], common["indir"]] + scales
        from trtllm_llama_b200._lib import TbpDims, TbpTensorDesc
        lib = L
        n_in = len(ins)
        idesc = p._descs(ins)
        out = torch.zeros(qkv.shape[:-1] + (hidden,), dtype=torch.float16, device="cuda")
        odesc = p._descs([out, cache])
        ws = torch.empty(max(lib.tbp_workspace_size(p.h, idesc, n_in, odesc, 2), 16), dtype=torch.uint8, device="cuda")
        in_ptrs = (C.c_void_p * n_in)(*[t.data_ptr() for t in ins])
        out_ptrs = (C.c_void_p * 2)(out.data_ptr(), cache.data_ptr())     # present_key_value aliases past_key_value
        assert lib.tbp_enqueue(p.h, idesc, odesc, in_ptrs, out_ptrs, ws.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
        torch.cuda.synchronize()
        return out.cpu().numpy()

    qkv0 = (rng.standard_normal((B, S, 3 * hidden)) * 0.5).astype(np.float16)
    ctx = call(qkv0, S, [0, 1])
    ref = R.context_attention(qkv0, cache_ref, in_lens, num_heads=H, head_size=Dh, kv_scale_orig_quant=s_q)
    for b in range(B):
        np.testing.assert_allclose(ctx[b, :in_lens[b]].astype(np.float32), ref[b, :in_lens[b]].astype(np.float32), atol=5e-3)
    for step in range(steps):
        qkv = (rng.standard_normal((B, 1, 3 * hidden)) * 0.5).astype(np.float16)
        past = S + step
        got = call(qkv, past, [past, 0])
        want = R.mmha_decode(qkv[:, 0], cache_ref, past, in_lens, S, num_heads=H, head_size=Dh, kv_scale_orig_quant=s_q,
                             kv_scale_quant_orig=s_dq)
        tol = 2e-3 * max(1.0, float(np.abs(want.astype(np.float32)).max()))
        np.testing.assert_allclose(got[:, 0].astype(np.float32), want.astype(np.float32), atol=tol, err_msg=f"step {step}")
    kv = cache.cpu().numpy()
    if int8_kv:
        d = np.abs(kv[:, :, :, :S + steps].astype(np.int32) - cache_ref[:, :, :, :S + steps].astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() < 1e-2
    else:
        np.testing.assert_allclose(kv.astype(np.float32), cache_ref.astype(np.float32), atol=4e-3)


def _attention_fields(H, Dh, int8_kv, remove_padding, ifb, paged=0):
    return [("num_heads", H, "i32"), ("head_size", Dh, "i32"), ("unidirectional", 1, "i32"), ("q_scaling", 1.0, "f32"),
            ("rotary_embedding_dim", Dh, "i32"), ("neox_rotary_style", 1, "i8"), ("context_fmha_type", 0, "i8"),
            ("multi_block_mode", 0, "i8"), ("multi_query_mode", 0, "i8"), ("int8_kv_cache", int(int8_kv), "i32"),
            ("fp8_kv_cache", 0, "i32"), ("remove_input_padding", int(remove_padding), "i8"), ("mask_type", 1, "i32"),
            ("paged_kv_cache", paged, "i32"), ("type_id", 1, "i32"), ("in_flight_batching", int(ifb), "i32")]


def _enqueue(L, p, ins, outs):
    """enqueue with caller-owned outputs (the KV cache is aliased); host tensors are passed by their host address"""
    n_in = len(ins)
    idesc, odesc = p._descs(ins), p._descs(outs)
    ws = torch.empty(max(L.tbp_workspace_size(p.h, idesc, n_in, odesc, len(outs)), 16), dtype=torch.uint8, device="cuda")
    ws.fill_(0xAB)
    in_ptrs = (C.c_void_p * n_in)(*[t.data_ptr() for t in ins])
    out_ptrs = (C.c_void_p * len(outs))(*[t.data_ptr() for t in outs])
    assert L.tbp_enqueue(p.h, idesc, odesc, in_ptrs, out_ptrs, ws.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("int8_kv", [True, False])
def test_gpt_attention_plugin_packed_input(L, int8_kv):
    """remove_input_padding (LQ/build.py --remove_input_padding; gptAttentionCommon.cpp:467-478): the context input is
    [1, num_tokens, 3*hidden] with the sequences back to back; same results as the padded protocol, same cache."""
    rng = np.random.default_rng(45)
    B, H, Dh, S, S_max, steps = 3, 4, 128, 20, 48, 3
    hidden = H * Dh
    in_lens = np.array([S, 7, 13], np.int32)
    T = int(in_lens.sum())
    p = Plugin(L, "GPTAttention", _attention_fields(H, Dh, int8_kv, True, False)).roundtrip()
    kv_np = np.int8 if int8_kv else np.float16
    cache = torch.zeros((B, 2, H, S_max, Dh), dtype=torch.int8 if int8_kv else torch.float16, device="cuda")
    cache_ref = np.zeros((B, 2, H, S_max, Dh), kv_np)
    s_q, s_dq = (np.float32(127.0 / 3.0), np.float32(3.0 / 127.0)) if int8_kv else (None, None)
    scales = [dev(np.array([s_q], np.float32)), dev(np.array([s_dq], np.float32))] if int8_kv else []
    masked = np.zeros((B, S_max), np.int32)
    for b in range(B):
        masked[b, in_lens[b]:S] = 1
    fixed = [dev(masked), dev(in_lens), torch.zeros(S, dtype=torch.int32, device="cuda"),
             torch.zeros((B, 1, S_max), dtype=torch.int32, device="cuda")]

    padded = (rng.standard_normal((B, S, 3 * hidden)) * 0.5).astype(np.float16)
    packed = np.concatenate([padded[b, :in_lens[b]] for b in range(B)])[None]
    out = torch.zeros((1, T, hidden), dtype=torch.float16, device="cuda")
    _enqueue(L, p, [dev(packed), cache, torch.full((B,), S, dtype=torch.int32, device="cuda"),
                    torch.tensor([0, 1], dtype=torch.int32)] + fixed + scales, [out, cache])
    ref = R.context_attention(padded, cache_ref, in_lens, num_heads=H, head_size=Dh, kv_scale_orig_quant=s_q)
    ref_packed = np.concatenate([ref[b, :in_lens[b]] for b in range(B)])
    np.testing.assert_allclose(out[0].cpu().numpy().astype(np.float32), ref_packed.astype(np.float32), atol=5e-3)
    for step in range(steps):                        # generation: [1, B, 3*hidden], one token per sequence
        qkv = (rng.standard_normal((1, B, 3 * hidden)) * 0.5).astype(np.float16)
        past = S + step
        o = torch.zeros((1, B, hidden), dtype=torch.float16, device="cuda")
        _enqueue(L, p, [dev(qkv), cache, torch.full((B,), past, dtype=torch.int32, device="cuda"),
                        torch.tensor([past, 0], dtype=torch.int32)] + fixed + scales, [o, cache])
        want = R.mmha_decode(qkv[0], cache_ref, past, in_lens, S, num_heads=H, head_size=Dh, kv_scale_orig_quant=s_q,
                             kv_scale_quant_orig=s_dq)
        tol = 2e-3 * max(1.0, float(np.abs(want.astype(np.float32)).max()))
        np.testing.assert_allclose(o[0].cpu().numpy().astype(np.float32), want.astype(np.float32), atol=tol, err_msg=f"step {step}")


@pytest.mark.parametrize("int8_kv", [True, False])
def test_gpt_attention_plugin_in_flight_batching(L, int8_kv):
    """in_flight_batching (gptAttentionPlugin.cpp:150-200): one enqueue carries requests in different phases — two that are
    generating, an idle slot, and two arriving with their prompts — told apart by the HOST tensors host_input_lengths /
    host_request_types; each group runs on its slice of the packed tokens and of the cache.  Every request must get exactly
    what it gets when it is served alone through the padded protocol."""
    rng = np.random.default_rng(46)
    H, Dh, S_max, max_in = 4, 128, 64, 24
    hidden = H * Dh
    p = Plugin(L, "GPTAttention", _attention_fields(H, Dh, int8_kv, True, True)).roundtrip()
    kv_t = torch.int8 if int8_kv else torch.float16
    s_q, s_dq = (np.float32(127.0 / 3.0), np.float32(3.0 / 127.0)) if int8_kv else (None, None)
    scales = [dev(np.array([s_q], np.float32)), dev(np.array([s_dq], np.float32))] if int8_kv else []
    nseq = 5
    cache = torch.zeros((nseq, 2, H, S_max, Dh), dtype=kv_t, device="cuda")
    ref_cache = np.zeros((nseq, 2, H, S_max, Dh), np.int8 if int8_kv else np.float16)

    def ctx_ref(slot, qkv):
        c = ref_cache[slot:slot + 1]
        o = R.context_attention(qkv[None], c, np.array([qkv.shape[0]], np.int32), num_heads=H, head_size=Dh, kv_scale_orig_quant=s_q)
        return o[0]

    def gen_ref(slot, qkv, past):
        c = ref_cache[slot:slot + 1]
        return R.mmha_decode(qkv[None], c, past, np.array([past], np.int32), past, num_heads=H, head_size=Dh,
                             kv_scale_orig_quant=s_q, kv_scale_quant_orig=s_dq)[0]

    def enqueue(types, host_lens, seq_lens, tokens):
        ins = [dev(tokens[None]), cache, dev(np.asarray(seq_lens, np.int32)), torch.tensor([0, 0], dtype=torch.int32),
               torch.zeros((nseq, S_max), dtype=torch.int32, device="cuda"), dev(np.asarray(host_lens, np.int32)),
               torch.zeros(max_in, dtype=torch.int32, device="cuda"), torch.zeros((nseq, 1, S_max), dtype=torch.int32, device="cuda")]
        ins += scales + [torch.tensor(host_lens, dtype=torch.int32), torch.tensor(types, dtype=torch.int32)]
        out = torch.zeros((1, tokens.shape[0], hidden), dtype=torch.float16, device="cuda")
        _enqueue(L, p, ins, [out, cache])
        return out[0].cpu().numpy()

    def close(got, want, what):
        tol = 5e-3 * max(1.0, float(np.abs(want.astype(np.float32)).max()))
        np.testing.assert_allclose(got.astype(np.float32), want.astype(np.float32), atol=tol, err_msg=what)

    q = lambda n: (rng.standard_normal((n, 3 * hidden)) * 0.5).astype(np.float16)  # noqa: E731
    # round 1: slots 0 and 1 arrive (context), the others are idle
    l0, l1 = 17, 9
    t0, t1 = q(l0), q(l1)
    got = enqueue([0, 0, 2, 2, 2], [l0, l1, 0, 0, 0], [l0, l1, 0, 0, 0], np.concatenate([t0, t1]))
    close(got[:l0], ctx_ref(0, t0), "slot 0 context")
    close(got[l0:], ctx_ref(1, t1), "slot 1 context")
    # round 2: slots 0, 1 generate; slot 2 idle; slots 3, 4 arrive — one enqueue, three groups
    l3, l4 = 24, 5
    g0, g1, t3, t4 = q(1), q(1), q(l3), q(l4)
    got = enqueue([1, 1, 2, 0, 0], [1, 1, 0, l3, l4], [l0, l1, 0, l3, l4], np.concatenate([g0, g1, t3, t4]))
    close(got[0], gen_ref(0, g0[0], l0), "slot 0 generation")
    close(got[1], gen_ref(1, g1[0], l1), "slot 1 generation")
    close(got[2:2 + l3], ctx_ref(3, t3), "slot 3 context")
    close(got[2 + l3:], ctx_ref(4, t4), "slot 4 context")
    # round 3: everybody generates at their own position
    lens = [l0 + 1, l1 + 1, 0, l3, l4]
    toks = [q(1) for _ in range(4)]
    got = enqueue([1, 1, 2, 1, 1], [1, 1, 0, 1, 1], lens, np.concatenate(toks))
    for k, slot in enumerate([0, 1, 3, 4]):
        close(got[k], gen_ref(slot, toks[k][0], lens[slot]), f"slot {slot} generation, round 3")
    kv = cache.cpu().numpy()
    if int8_kv:
        d = np.abs(kv.astype(np.int32) - ref_cache.astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() < 1e-2
    else:
        np.testing.assert_allclose(kv.astype(np.float32), ref_cache.astype(np.float32), atol=4e-3)
    assert np.all(kv[2] == 0)                          # the idle slot's cache is untouched
