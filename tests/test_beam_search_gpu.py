"""Beam search: tb_beam_search_step / tb_gather_tree / the cache-indirection read of the decode attention
against the numpy restatement of the reference's beam layer (oracle/ref_ops.py beam_search_step, gather_tree,
mmha_decode_beams), and through the engine against the oracle model's beam loop.

Token ids, parents, finished flags, lengths and the cache indirection must be identical; cum_log_probs within 2e-5
(__expf / block-wide sums vs numpy).  A step whose selection margin (gap between the normalised scores of neighbouring
candidates) is below the score tolerance may legitimately order two candidates differently: such steps end the comparison
of that case, and the test requires that most cases run to the end."""
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


@pytest.mark.parametrize("W", [2, 4, 7, 16])
@pytest.mark.parametrize("V", [512, 32000])
@pytest.mark.parametrize("lp", [0.0, 1.0, 0.6])
def test_beam_step_and_gather_tree_match_the_restated_reference(ops, W, V, lp):
    rng = np.random.default_rng(W * 1000 + V)
    B, max_in, n, end_id = 3, 5, 7, 2
    S_max = max_in + n + 1
    st = ops.BeamSearchState(B, W, n, S_max, max_in)
    cum = np.tile(np.array([0.0] + [-1e20] * (W - 1), np.float32), B)
    fin, lens = np.zeros(B * W, bool), np.full(B * W, max_in, np.int64)
    indir = np.zeros((B, W, S_max), np.int32)
    ids_t, par_t, steps_checked = [], [], 0
    for s in range(n):
        rows = B if s == 0 else B * W
        logits = rng.standard_normal((rows, V)).astype(np.float32) * 3.0
        logits[:, end_id] += 4.0                      # end_id is often among the candidates: beams do finish
        full = np.repeat(logits, W, axis=0) if s == 0 else logits
        tok, par, cum, fin, lens, indir, margin = R.beam_search_step(full, cum, fin, lens, indir, max_in + s, beam_width=W,
                                                                     end_id=end_id, length_penalty=lp)
        ids_t.append(tok)
        par_t.append(par)
        st.advance(torch.from_numpy(logits).cuda(), end_id, lp, broadcast=(s == 0))
        if margin < 3e-5:
            break
        steps_checked += 1
        np.testing.assert_array_equal(st.ids_t[s].cpu().numpy(), tok, err_msg=f"tokens, step {s}")
        np.testing.assert_array_equal(st.parent_t[s].cpu().numpy(), par, err_msg=f"parents, step {s}")
        np.testing.assert_array_equal(st.next_ids.cpu().numpy(), tok)
        np.testing.assert_array_equal(st.finished.cpu().numpy().astype(bool), fin)
        np.testing.assert_array_equal(st.lens.cpu().numpy(), lens)
        np.testing.assert_allclose(st.cum.cpu().numpy(), cum, rtol=0, atol=2e-5 * (s + 1) + 1e-6 * np.abs(cum).max())
        live = ~fin.reshape(B, W)
        got = st.indir[0].cpu().numpy()
        np.testing.assert_array_equal(got[live][:, :max_in + s + 1], indir[live][:, :max_in + s + 1], err_msg=f"indirection, step {s}")
    assert steps_checked >= 2
    k = steps_checked
    ref = R.gather_tree(np.stack(ids_t[:k]), np.stack(par_t[:k]), W, end_id)
    np.testing.assert_array_equal(st.gather_tree(k, end_id).cpu().numpy(), ref)


@pytest.mark.parametrize("int8_kv", [False, True])
@pytest.mark.parametrize("W,past", [(2, 9), (4, 70), (3, 300)])
def test_decode_attention_reads_the_cache_through_the_indirection(ops, int8_kv, W, past):
    rng = np.random.default_rng(past + W)
    B, H, Dh = 2, 4, 128
    rows, S_max, max_in = B * W, past + 4, 6
    lens = np.repeat(np.array([6, 4], np.int32), W)
    qkv = (rng.standard_normal((rows, 3 * H * Dh)) * 0.5).astype(np.float16)
    if int8_kv:
        cache = rng.integers(-127, 128, (rows, 2, H, S_max, Dh)).astype(np.int8)
        s_oq, s_qo = np.float32(127.0 / 4.0), np.float32(4.0 / 127.0)
    else:
        cache = (rng.standard_normal((rows, 2, H, S_max, Dh)) * 0.5).astype(np.float16)
        s_oq = s_qo = None
    indir = rng.integers(0, W, (B, W, S_max)).astype(np.int32)
    ref_cache = cache.copy()
    ref = R.mmha_decode_beams(qkv, ref_cache, indir, W, past, lens, max_in, num_heads=H, head_size=Dh,
                              kv_scale_orig_quant=s_oq, kv_scale_quant_orig=s_qo)
    dcache = torch.from_numpy(cache).cuda()
    f = lambda v: None if v is None else torch.tensor([v], dtype=torch.float32, device="cuda")  # noqa: E731
    out = ops.mmha_decode_beams(torch.from_numpy(qkv).cuda(), dcache, torch.from_numpy(indir).cuda(), past, num_heads=H,
                                head_size=Dh, max_input_len=max_in, input_lengths=torch.from_numpy(lens).cuda(),
                                kv_scale_orig_quant=f(s_oq), kv_scale_quant_orig=f(s_qo))
    scale = float(np.abs(ref.astype(np.float32)).max())
    np.testing.assert_allclose(out.cpu().numpy().astype(np.float32), ref.astype(np.float32), rtol=0, atol=2e-3 * max(scale, 1.0))
    got = dcache.cpu().numpy()
    if int8_kv:
        assert np.abs(got.astype(np.int32) - ref_cache.astype(np.int32)).max() <= 1      # appended row: int8 codes +-1
    else:                                                                               # appended K row: RoPE within 1 fp16 ulp
        np.testing.assert_allclose(got.astype(np.float32), ref_cache.astype(np.float32), rtol=2e-3, atol=1e-4)
        mask = np.ones(S_max, bool)
        mask[past] = False
        np.testing.assert_array_equal(got[:, :, :, mask], cache[:, :, :, mask])         # nothing else is touched
    # the same call with the identity indirection equals the plain kernel on every row
    ident = np.broadcast_to(np.arange(W, dtype=np.int32)[None, :, None], (B, W, S_max)).copy()
    c1, c2 = torch.from_numpy(cache).cuda(), torch.from_numpy(cache).cuda()
    a = ops.mmha_decode_beams(torch.from_numpy(qkv).cuda(), c1, torch.from_numpy(ident).cuda(), past, num_heads=H, head_size=Dh,
                              max_input_len=max_in, input_lengths=torch.from_numpy(lens).cuda(), kv_scale_orig_quant=f(s_oq),
                              kv_scale_quant_orig=f(s_qo), nsplit=1)
    b = ops.mmha_decode(torch.from_numpy(qkv).cuda(), c2, past, num_heads=H, head_size=Dh, max_input_len=max_in,
                        input_lengths=torch.from_numpy(lens).cuda(), kv_scale_orig_quant=f(s_oq), kv_scale_quant_orig=f(s_qo),
                        nsplit=1)
    if not (int8_kv and past >= 512):
        assert torch.equal(a, b)


@pytest.mark.parametrize("mode,int8_kv", [("fp16", False), ("fp16", True), ("w8", True), ("sq", True)])
@pytest.mark.parametrize("W", [2, 4])
def test_engine_beam_search_matches_the_oracle_model(mode, int8_kv, W):
    from test_engine_gpu import _prompts, _session
    cfg = RM.LlamaCfg.tiny(layers=2, hidden=256, inter=384, vocab=512)
    w = RM.random_weights(cfg, seed=11, std=0.08)
    B, S, new, end_id = 2, 10, 6, 2
    rng = np.random.default_rng(W)
    ids, lens = _prompts(rng, cfg, B, S, [S, S - 3])
    oracle = RM.OracleLlama(cfg, RM.quantize_model(w, mode), mode, int8_kv, kv_scale=4.0 / 127.0, max_seq_len=S + new)
    ref_ids, ref_cum, margin = oracle.generate_beams(ids, lens, new, W, end_id, length_penalty=1.0)

    from trtllm_llama_b200 import runtime as rt
    sess, mc = _session(cfg, w, mode, int8_kv, max_batch=B * W, max_in=S, max_out=new)
    sess.setup(B, S, new, beam_width=W)
    sc = rt.SamplingConfig(end_id=end_id, pad_id=end_id, num_beams=W, length_penalty=1.0)
    out = sess.decode(torch.from_numpy(ids), torch.from_numpy(lens), sc, max_new_tokens=new)
    assert tuple(out.shape) == (B, W, new)
    # the candidate scores are sums of log-probabilities of logits that agree within the engine tolerance; compare ids only
    # when every selection of the oracle run was clear of that tolerance
    tol = (3e-2 if mode == "sq" else 1e-2) * new / S
    if margin > tol:
        np.testing.assert_array_equal(out.numpy(), ref_ids)
        np.testing.assert_allclose(sess.cum_log_probs.numpy()[:, 0], ref_cum[:, 0], rtol=0, atol=0.15 if mode == "sq" else 0.05)
    else:   # a near-tie may pick another path: the best beam's score still has to be as good as the oracle's, within tol
        assert np.all(sess.cum_log_probs.numpy()[:, 0] / (S + new) >= ref_cum[:, 0] / (S + new) - tol)
    # beam search with the same request twice gives the same answer (the step graph is replayed)
    out2 = sess.decode(torch.from_numpy(ids), torch.from_numpy(lens), sc, max_new_tokens=new)
    assert torch.equal(out, out2)
