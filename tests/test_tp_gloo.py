"""Tensor-parallel host logic on CPU with the gloo backend, world_size 2: the Megatron split of
runtime.shard_weights + one all-reduce after each row-parallel projection + one all-gather of the vocab-parallel
logits (T/examples/llama/weight.py:71-178, T/tensorrt_llm/layers/linear.py:78-139) reproduces the
unsharded oracle decoder."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from oracle import ref_model as RM  # noqa: E402
from oracle import ref_ops as R  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tp_layer(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from trtllm_llama_b200.runtime import shard_weights
    cfg = RM.LlamaCfg.tiny(layers=2, hidden=256, inter=384, vocab=512)
    w = RM.random_weights(cfg, seed=2, std=0.05)
    tw = {k: torch.from_numpy(w[k]) for k in ("vocab_embedding", "ln_f", "lm_head")}
    tw["layers"] = [{k: torch.from_numpy(v) for k, v in lw.items()} for lw in w["layers"]]
    sw = shard_weights(tw, world, rank, cfg.heads)
    rng = np.random.default_rng(4)
    B, S = 2, 6
    ids = rng.integers(3, cfg.vocab, (B, S)).astype(np.int32)
    lens = np.array([S, S - 2], np.int32)
    Hl = cfg.heads // world
    cache = [np.zeros((B, 2, Hl, 16, 128), np.float16) for _ in range(cfg.layers)]
    h = w["vocab_embedding"][ids.reshape(-1)]

    def allreduce(x):   # fp16 payload, summed in fp32 then rounded: what ncclAllReduce(sum, fp16) delivers up to order
        t = torch.from_numpy(x.astype(np.float32))
        dist.all_reduce(t)
        return t.numpy().astype(np.float16)

    for li in range(cfg.layers):
        lw = {k: v.numpy() for k, v in sw["layers"][li].items()}
        x = R.rmsnorm(h, lw["input_layernorm"], cfg.eps)
        qkv = R.gemm_f16(x, lw["qkv"])
        a = R.context_attention(qkv.reshape(B, S, -1), cache[li], lens, num_heads=Hl, head_size=128).reshape(B * S, -1)
        h = R.residual_add(allreduce(R.gemm_f16(a, lw["dense"])), h)
        x = R.rmsnorm(h, lw["post_layernorm"], cfg.eps)
        act = R.swiglu(R.gemm_f16(x, lw["gate"]), R.gemm_f16(x, lw["up"]))
        h = R.residual_add(allreduce(R.gemm_f16(act, lw["down"])), h)
    hl = h.reshape(B, S, -1)[np.arange(B), lens - 1]
    local = R.gemm_f16(R.rmsnorm(hl, w["ln_f"], cfg.eps), sw["lm_head"].numpy())          # [B, V/tp]
    parts = [torch.zeros(local.shape, dtype=torch.float16) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(local))
    logits = np.concatenate([p.numpy() for p in parts], axis=1).astype(np.float32)
    if rank == 0:
        ret["logits"] = logits
    dist.destroy_process_group()


def test_tp2_schedule_matches_unsharded_oracle():
    cfg = RM.LlamaCfg.tiny(layers=2, hidden=256, inter=384, vocab=512)
    w = RM.random_weights(cfg, seed=2, std=0.05)
    rng = np.random.default_rng(4)
    ids = rng.integers(3, cfg.vocab, (2, 6)).astype(np.int32)
    lens = np.array([6, 4], np.int32)
    ref = RM.OracleLlama(cfg, RM.quantize_model(w, "fp16"), "fp16", False, max_seq_len=16).context(ids, lens)
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_tp_layer, args=(2, port, ret), nprocs=2, join=True)
        got = ret["logits"]
    np.testing.assert_allclose(got, ref, atol=1e-2 * max(1.0, float(np.abs(ref).max())))
    assert np.array_equal(got.argmax(-1), ref.argmax(-1))
