"""Host-side logic that needs no GPU: the paged-KV block bookkeeping (mirror of
T/tensorrt_llm/runtime/kv_cache_manager.py) and the numpy restatement of the sampling kernels, pinned by Random123's
known-answer vectors for Philox4x32-10 and by the sampling distributions it must produce."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ref_ops as R  # noqa: E402


def test_kv_cache_manager_bookkeeping():
    from trtllm_llama_b200.runtime import GenerationSequence, KVCacheManager
    m = KVCacheManager(blocks=12, tokens_per_block=16, max_blocks_per_seq=4)
    a, b = GenerationSequence(0, 0), GenerationSequence(1, 1)
    m.add_sequence(a, 15)       # 15 + 1 positions -> one block
    m.add_sequence(b, 16)       # 16 + 1 -> two blocks
    t = m.get_block_table(2).numpy()
    assert (t[0] >= 0).sum() == 1 and (t[1] >= 0).sum() == 2 and len(m.free_blocks) == 9
    assert m.step([False, False])                       # a: len 15 -> crosses into block 1; b: 16 -> 17, no new block
    t = m.get_block_table(2).numpy()
    assert (t[0] >= 0).sum() == 2 and (t[1] >= 0).sum() == 2
    for _ in range(14):
        m.step([False, False])
    assert m.step([False, True])                        # b finishes: its blocks return to the free list, a is batch idx 0
    assert len(m.sequences) == 1 and m.sequences[0].get_batch_idx() == 0 and len(m.free_blocks) == 12 - len(m.allocated[a])
    with pytest.raises(ValueError):
        KVCacheManager(blocks=4, tokens_per_block=24, max_blocks_per_seq=2)


def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32 with 10 rounds (the generator behind curand's Philox and tb_sample)."""
    assert R.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert R.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert R.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = np.array([R.sampling_uniform(7, s, b) for s in range(40) for b in range(50)], np.float64)
    assert (u > 0).all() and (u <= 1).all() and abs(u.mean() - 0.5) < 0.02 and len(np.unique(u)) == len(u)


def test_sampling_restatement_distributions():
    """top-k / top-p / top-k + top-p of the restated reference on a 6-token distribution (chi-square against the
    renormalised masses the reference's walk implies: the last kept token only gets the mass below the threshold)."""
    base = np.full(64, -20.0, np.float32)
    probs = [0.4, 0.25, 0.15, 0.1, 0.06, 0.04]
    toks = [3, 9, 17, 21, 40, 63]
    base[toks] = np.log(np.array(probs, np.float32))
    n = 2000
    u = np.array([R.sampling_uniform(99, 0, b) for b in range(n)], np.float32)
    for top_k, top_p, kept in ((4, 1.0, 4), (0, 0.7, 3), (6, 0.6, 2), (1, 1.0, 1)):
        ids, _ = R.sample_top_k_top_p(np.tile(base, (n, 1)), top_k, top_p, 1.0, u)
        counts = np.array([(ids == t).sum() for t in toks[:kept]], np.float64)
        assert counts.sum() == n
        mass = np.array(probs[:kept], np.float64)
        limit = top_p * (sum(probs[:top_k]) if top_k > 0 else 1.0)
        mass[-1] = limit - mass[:-1].sum()
        expected = n * mass / limit
        assert float(((counts - expected) ** 2 / np.maximum(expected, 1e-9)).sum()) < 25.0, (top_k, top_p, counts, expected)
    # temperature -> 0 sharpens to greedy
    ids, _ = R.sample_top_k_top_p(np.tile(base, (50, 1)), 6, 1.0, 1e-3, u[:50])
    assert (ids == 3).all()


def test_bench_roofline_reports_shards_outside_the_decode_gemv():
    """bench.py times the decode GEMV class on rank 0 only; at tp = 4 / 8 the int4 down projection's K (2752 / 1376) is not a
    whole number of the GEMV's k-steps, 8 rows are then outside the decode-shape GEMV (the engine uses the tcgen05 GEMM) and the
    timing function must say so instead of raising on one rank (which once left rank 0 alone in a barrier: a hung N = 4 run)."""
    torch = pytest.importorskip("torch")
    import bench
    from trtllm_llama_b200 import _lib
    _lib.load_library()
    cfg = dict(bench.LLAMA7B)
    cfg["layers"] = 1
    tp = 4
    hid, inter = cfg["hidden"], cfg["inter"]
    shapes = {"attention.qkv": (3 * hid // tp, hid), "attention.dense": (hid, hid // tp),
              "mlp.fc_gate": (2 * inter // tp, hid), "mlp.proj": (hid, inter // tp)}
    tensors = {}
    for name, (N, K) in shapes.items():
        tensors[f"layers.0.{name}.weight"] = torch.zeros((N, K // 2), dtype=torch.int8)          # int4: two weights per byte
        tensors[f"layers.0.{name}.per_channel_scale"] = torch.ones(N, dtype=torch.float16)
    r = bench.gemv_roofline(torch, tensors, cfg, "w4", 6500.0, "measured", rows=8, tp=tp)
    assert r["achieved"] is None and "2752" in r["kernel"] and "gemm_tc_kernel" in r["kernel"]


def _run_bench_main_mocked(monkeypatch, capsys, rank, world, fail_parity=False, fail_side=False):
    """bench.main() with the GPU work mocked out: returns (printed line or None, the sequence of collectives this rank entered)."""
    import json
    import sys
    import types
    torch = pytest.importorskip("torch")
    import bench
    ops_log = []

    class FakeDist:
        def barrier(self): ops_log.append("barrier")
        def destroy_process_group(self): ops_log.append("destroy")

    class FakeCtx:
        def __init__(self, args):
            self.args, self.rank, self.world, self.local = args, rank, world, rank
            self.torch, self.lib, self.hbm, self.which = torch, None, 6500.0, "measured"
            self.dist = FakeDist() if world > 1 else None
        def barrier(self):
            if self.dist is not None:
                self.dist.barrier()

    def fake_workload(cx, name, steps, warmup, with_e2e=True, with_roofline=True, return_ids=False):
        ops_log.append("workload:" + name)
        if fail_side and name == "cfg5_b8" and cx.rank == 0:
            raise RuntimeError("rank-0-only failure inside a side workload")
        if cx.rank != 0:
            return None, None
        head = {"value": 1.0, "ms_per_request": 1.0, "e2e": {}, "gpu_launches": 1, "context_ms": 1.0, "roofline": {}, "clocks": {},
                "decode_step": {"algorithmic_bytes": 1e9}}
        return head, [[1, 2, 3]]

    def fake_parity(cx, name, ids):
        if fail_parity:
            raise RuntimeError("tp = 1 engine could not be built")
        return {"ok": True}

    def fake_forced(cx, name, n_steps=32):
        ops_log.append("forced")
        return {"ok": True}

    monkeypatch.setattr(bench, "Ctx", FakeCtx)
    monkeypatch.setattr(bench, "run_decode_workload", fake_workload)
    monkeypatch.setattr(bench, "run_prefill_workload", lambda cx, steps, warmup: {"prefill_ms": 1.0})
    monkeypatch.setattr(bench, "tp_parity", fake_parity)
    monkeypatch.setattr(bench, "tp_parity_forced", fake_forced)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    fake_ref = types.ModuleType("tools.ref_kernel_bench")
    fake_ref.reference_kernels = lambda: {"rows": []}
    monkeypatch.setitem(sys.modules, "tools.ref_kernel_bench", fake_ref)
    fake_hf = types.ModuleType("oracle.hf_baseline")
    fake_hf.time_hf_cpu = lambda **kw: {"value": 1.0, "cores": 1, "sample": "mock", "t_prefill_s": 1.0, "t_step_s": 1.0}
    monkeypatch.setitem(sys.modules, "oracle.hf_baseline", fake_hf)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", str(world), "--steps", "3", "--warmup", "3"])
    bench.main()
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    return (json.loads(out[-1]) if out else None), ops_log


def test_bench_main_single_gpu_line_has_the_contract_keys(monkeypatch, capsys):
    line, _ = _run_bench_main_mocked(monkeypatch, capsys, rank=0, world=1)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "workloads", "reference_kernels", "cpu_baseline"):
        assert k in line, k
    assert "assembly_error" not in line and "cfg4_prefill" in line["workloads"]


@pytest.mark.parametrize("fail_parity", [False, True])
def test_bench_main_every_rank_enters_the_same_collectives(monkeypatch, capsys, fail_parity):
    """torchrun ranks must agree on the sequence of collectives even when a rank-0-only part fails (the tp = 1 comparison
    engine): otherwise rank 0 pairs a barrier with a later one of its peers and the job hangs at exit."""
    line0, ops0 = _run_bench_main_mocked(monkeypatch, capsys, rank=0, world=4, fail_parity=fail_parity)
    line1, ops1 = _run_bench_main_mocked(monkeypatch, capsys, rank=1, world=4, fail_parity=fail_parity)
    assert ops0 == ops1 and ops0[-2:] == ["barrier", "destroy"]
    assert line1 is None and line0["n_gpus"] == 4 and line0["tp_parity"] is (not fail_parity)
    assert set(line0["workloads"]) == {"cfg5", "cfg5_b8"}


@pytest.mark.parametrize("fail", [None, "build", "step"])
def test_tp_parity_forced_keeps_ranks_in_step_when_the_tp1_engine_fails(monkeypatch, fail):
    """bench.tp_parity_forced is a collective: rank 0's private tp = 1 engine may fail to build or to step without changing the
    sequence of broadcasts / engine steps rank 0 runs (else the peers wait in a broadcast forever)."""
    torch = pytest.importorskip("torch")
    import types
    import bench

    def run(rank):
        log = []

        class FakeSess:
            def __init__(self, tp):
                self.tp, self._e, self.n = tp, None, 0
            def context(self, ids, lens):
                log.append(f"context tp{self.tp}")
                return torch.zeros(1, 8)
            def step(self):
                self.n += 1
                if self.tp == 1 and fail == "step" and self.n == 3:
                    raise RuntimeError("boom")
                log.append(f"step tp{self.tp}")
                return torch.zeros(1, 8)
            def force_ids(self, tok):
                pass

        def fake_build(cx, mode, int8_kv, B, in_len, out_len, tp, rank_, graph=True, peer_ar=True):
            if tp == 1 and fail == "build":
                raise RuntimeError("no memory for the tp = 1 engine")
            return FakeSess(tp), {}

        class FakeDist:
            def broadcast(self, t, src):
                log.append("broadcast")

        cx = types.SimpleNamespace(torch=torch, dist=FakeDist(), world=4, rank=rank, args=types.SimpleNamespace(no_graph=False, nccl_only=False),
                                   lib=types.SimpleNamespace(tbrt_last_launches=lambda e: 1))
        monkeypatch.setattr(bench, "build_session", fake_build)
        monkeypatch.setattr(torch, "zeros", lambda *a, **k: torch.ones(*a, **{kk: v for kk, v in k.items() if kk != "device"}) * 0)
        monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
        r = bench.tp_parity_forced(cx, "cfg2", n_steps=5)
        return r, [x for x in log if "tp1" not in x]

    r0, ops0 = run(0)
    r1, ops1 = run(1)
    assert ops0 == ops1 and ops0.count("broadcast") == 6 and ops0.count("step tp4") == 5
    assert r0["ok"] is (fail is None)
    if fail:
        assert "tp = 1 engine" in r0["error"]
