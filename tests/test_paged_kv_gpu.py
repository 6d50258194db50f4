"""Paged KV cache: GPTAttention with paged_kv_cache = 1 — the block pool + block_pointers input of
P/gptAttentionPlugin/gptAttentionPlugin.cpp:204-235, KVBlockArray addressing of K/kvCacheUtils.h:34-112, the block
bookkeeping of T/tensorrt_llm/runtime/kv_cache_manager.py — must reproduce the contiguous cache exactly: same kernels, same
arithmetic, only the address of a cached row changes.  The manager deals blocks in shuffled order, so every sequence's blocks
are scattered over the pool and interleaved with the other sequences'."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_model as RM  # noqa: E402
from oracle import ref_ops as R  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    import trtllm_llama_b200  # noqa: F401
    from trtllm_llama_b200 import ops as o
    return o


def _blocks(rng, B, max_blocks, pool_blocks):
    ids = rng.permutation(pool_blocks)[:B * max_blocks].reshape(B, max_blocks).astype(np.int64)
    return ids


def _pointer_table(pool, ids, block_bytes, pool_blocks):
    """[B, 2, max_blocks] int64 device addresses: K blocks in the first half of the pool, V blocks in the second"""
    base = pool.data_ptr()
    k = base + ids * block_bytes
    v = base + (pool_blocks + ids) * block_bytes
    return torch.from_numpy(np.stack([k, v], axis=1).astype(np.int64)).cuda()


@pytest.mark.parametrize("int8_kv", [True, False])
@pytest.mark.parametrize("past,tpb", [(37, 16), (200, 64), (1100, 128), (1100, 16)])
def test_mmha_decode_paged_equals_contiguous(ops, int8_kv, past, tpb):
    import ctypes as C
    lib = ops.lib
    rng = np.random.default_rng(51)
    B, H, Dh = 3, 4, 128
    max_blocks = (past + 1 + tpb - 1) // tpb + 1
    S_max = max_blocks * tpb
    pool_blocks = B * max_blocks + 5
    elt = 1 if int8_kv else 2
    block_elems = H * tpb * Dh
    if int8_kv:
        cache = rng.integers(-127, 128, (B, 2, H, S_max, Dh), dtype=np.int8)
    else:
        cache = rng.standard_normal((B, 2, H, S_max, Dh)).astype(np.float16)
    qkv = rng.standard_normal((B, 3 * H * Dh)).astype(np.float16)
    in_lens = np.array([24, 17, 1], np.int32)
    max_in = 24
    ids = _blocks(rng, B, max_blocks, pool_blocks)
    pool = np.zeros((2, pool_blocks, H, tpb, Dh), dtype=cache.dtype)
    for b in range(B):
        for j in range(max_blocks):
            pool[:, ids[b, j]] = cache[b, :, :, j * tpb:(j + 1) * tpb]           # [2, H, tpb, Dh]
    d_pool = torch.from_numpy(pool).cuda()
    table = _pointer_table(d_pool, ids, block_elems * elt, pool_blocks)
    s_q, s_dq = np.float32(127.0 / 4.0), np.float32(4.0 / 127.0)
    d_sq, d_sdq = torch.tensor([s_q], device="cuda"), torch.tensor([s_dq], device="cuda")
    d_qkv, d_in = torch.from_numpy(qkv).cuda(), torch.from_numpy(in_lens).cuda()
    d_seq = torch.full((B,), past, dtype=torch.int32, device="cuda")
    kw = dict(kv_scale_orig_quant=d_sq, kv_scale_quant_orig=d_sdq) if int8_kv else {}
    for nsplit in (1, 0):
        d_cache = torch.from_numpy(cache).cuda()
        ref = ops.mmha_decode(d_qkv, d_cache, past, num_heads=H, head_size=Dh, max_input_len=max_in, seq_lens=d_seq,
                              input_lengths=d_in, nsplit=nsplit, **kw)
        d_pool2 = d_pool.clone()
        table2 = _pointer_table(d_pool2, ids, block_elems * elt, pool_blocks)
        out = torch.empty((B, H * Dh), dtype=torch.float16, device="cuda")
        ns = nsplit if nsplit > 0 else lib.tb_mmha_num_splits(B, H, past, 32)
        P = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)   # noqa: E731
        rc = lib.tb_mmha_decode_paged(P(out), P(d_qkv), P(table2), tpb, max_blocks, P(d_seq), P(d_in), None, None,
                                      P(d_sq) if int8_kv else None, P(d_sdq) if int8_kv else None, B, H, Dh, past, max_in, past,
                                      Dh, C.c_float(1.0), int(int8_kv), ns, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        torch.cuda.synchronize()
        assert torch.equal(out, ref), f"paged attention output differs from the contiguous cache (nsplit {nsplit})"
        # the appended K / V row landed in the right block, nothing else moved
        after_c, after_p = d_cache.cpu().numpy(), d_pool2.cpu().numpy()
        for b in range(B):
            for j in range(max_blocks):
                assert np.array_equal(after_p[:, ids[b, j]], after_c[b, :, :, j * tpb:(j + 1) * tpb]), (b, j)
    del table


@pytest.mark.parametrize("mode,int8_kv", [("fp16", True), ("w4", False), ("sq", True)])
def test_engine_paged_kv_cache_generates_the_same_ids(mode, int8_kv):
    """The same request through a contiguous-cache engine and a paged-cache engine (tokens_per_block 16: every sequence
    crosses block boundaries during the context phase and again while generating): identical logits and ids."""
    from trtllm_llama_b200 import runtime as rt
    from trtllm_llama_b200.quantization import QuantMode
    cfg = RM.LlamaCfg.tiny(layers=2, hidden=256, inter=384, vocab=512)
    w = RM.random_weights(cfg, seed=41, std=0.05)
    B, S, new = 3, 27, 24
    rng = np.random.default_rng(42)
    ids = rng.integers(3, cfg.vocab, (B, S)).astype(np.int32)
    lens = np.array([S, 13, 2], np.int32)
    for b in range(B):
        ids[b, lens[b]:] = 2
    qm = {"fp16": QuantMode(0), "w4": QuantMode.use_weight_only(True), "sq": QuantMode.use_smooth_quant(True, True)}[mode]
    if int8_kv:
        qm |= QuantMode.INT8_KV_CACHE
    f = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    tw = {k: f(w[k]) for k in ("vocab_embedding", "ln_f", "lm_head")}
    tw["layers"] = [{k: f(v) for k, v in lw.items()} for lw in w["layers"]]
    outs = {}
    for paged in (False, True):
        mc = rt.ModelConfig(vocab_size=cfg.vocab, num_layers=cfg.layers, num_heads=cfg.heads, hidden_size=cfg.hidden,
                            inter_size=cfg.inter, rms_eps=cfg.eps, quant_mode=qm, max_batch_size=4, max_input_len=S,
                            max_output_len=new, paged_kv_cache=paged, tokens_per_block=16)
        sess = rt.GenerationSession(mc, rt.build_engine_tensors(tw, mc, kv_scale=4.0 / 127.0))
        sess.setup(B, S, new)
        logits = [sess.context(torch.from_numpy(ids), torch.from_numpy(lens)).cpu().numpy()]
        for _ in range(new - 1):
            logits.append(sess.step().cpu().numpy())
        outs[paged] = (np.stack(logits, 1), sess.output_ids(new).cpu().numpy())
        if paged:
            m = sess.kv_cache_manager
            table = m.get_block_table(B).numpy()
            used = table[table >= 0]
            assert len(set(used.tolist())) == len(used), "a pool block was handed to two sequences"
            assert (np.diff(table[0][table[0] >= 0]) != 1).any(), "blocks came out contiguous: the indirection is not exercised"
            assert (table[0] >= 0).sum() == -(-(S + new) // 16)                 # grew block by block while generating
            host = sess.decode(torch.from_numpy(ids).pin_memory(), torch.from_numpy(lens).pin_memory()).numpy()
            assert np.array_equal(host, outs[paged][1])
    assert np.array_equal(outs[True][1], outs[False][1]), "paged and contiguous engines generated different ids"
    assert np.array_equal(outs[True][0], outs[False][0]), "paged and contiguous engines produced different logits"
