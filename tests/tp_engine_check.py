"""Run under torchrun with N ranks (one per GPU): the tensor-parallel engine must reproduce the CPU oracle's logits and
greedy ids on BOTH decode paths — the fused step kernel (partial sums / flags / arg-max candidates exchanged through peer
memory inside the one persistent kernel) and the per-operator plugin schedule (one-shot NVLink all-reduce kernel or the
NCCL AllReduce / AllGather plugins).  Driven by tests/test_tp_gpu.py when >= 2 GPUs are visible."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from oracle import ref_model as RM
    from trtllm_llama_b200 import runtime as rt
    from trtllm_llama_b200._lib import lib
    from trtllm_llama_b200.quantization import QuantMode
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        assert lib.tb_comm_unique_id(idbuf.data_ptr()) == 0
    idd = idbuf.cuda()
    dist.broadcast(idd, 0)
    idbuf = idd.cpu()
    assert lib.tb_comm_init(idbuf.data_ptr(), (C.c_int32 * world)(*range(world)), world, rank) == 0

    mode = sys.argv[1] if len(sys.argv) > 1 else "w4"
    int8_kv = True
    cfg = RM.LlamaCfg.tiny(layers=2, hidden=128 * 2 * world, inter=256 * world, vocab=256 * world)
    w = RM.random_weights(cfg, seed=8, std=0.05)
    B, S, new = 2, 9, 5
    rng = np.random.default_rng(9)
    ids = rng.integers(3, cfg.vocab, (B, S)).astype(np.int32)
    lens = np.array([S, S - 3], np.int32)
    ids[1, S - 3:] = 2
    qm = {"fp16": QuantMode(0), "w8": QuantMode.use_weight_only(False), "w4": QuantMode.use_weight_only(True)}[mode]
    qm |= QuantMode.INT8_KV_CACHE
    mc = rt.ModelConfig(vocab_size=cfg.vocab, num_layers=cfg.layers, num_heads=cfg.heads, hidden_size=cfg.hidden,
                        inter_size=cfg.inter, quant_mode=qm, max_batch_size=B, max_input_len=S, max_output_len=new,
                        tp_size=world, tp_rank=rank)
    tw = {k: torch.from_numpy(w[k]).cuda() for k in ("vocab_embedding", "ln_f", "lm_head")}
    tw["layers"] = [{k: torch.from_numpy(v).cuda() for k, v in lw.items()} for lw in w["layers"]]
    sess = rt.GenerationSession(mc, rt.build_engine_tensors(rt.shard_weights(tw, world, rank, cfg.heads), mc))
    if os.environ.get("TB_TP_NCCL_ONLY") != "1":
        sess.enable_peer_allreduce()      # fused NVLink all-reduce + residual on the decode path
    sess.setup(B, S, new)
    runs = {}
    for path, fused in (("fused", True), ("plugins", False)):
        sess.set_decode_mode(fused)
        logits = [sess.context(torch.from_numpy(ids), torch.from_numpy(lens)).cpu().numpy()]
        launches = []
        for _ in range(new - 1):
            logits.append(sess.step().cpu().numpy())
            launches.append(int(sess.last_launches))
        runs[path] = (np.stack(logits, 1), sess.output_ids(new).cpu().numpy(), launches)
        dist.barrier()
    if os.environ.get("TB_TP_NCCL_ONLY") != "1" and B <= sess.fused_step_max_batch:
        assert set(runs["fused"][2]) == {1}, f"the fused step did not run under tensor parallelism: {runs['fused'][2]}"
    assert min(runs["plugins"][2]) > 10
    # every rank must hold the same ids (the arg-max candidates of all vocabulary shards reach every rank)
    for path in runs:
        t = torch.from_numpy(runs[path][1]).cuda()
        ref_t = t.clone()
        dist.broadcast(ref_t, 0)
        assert torch.equal(t, ref_t), f"{path}: rank {rank} generated different ids than rank 0"
    if rank == 0:
        # the reference quantises each rank's shard separately (LQ/weight_quant.py:264-271); the oracle must do the same:
        # per-output-channel scales of column-parallel weights are shard-independent, row-parallel ones are not
        class ShardedOracle(RM.OracleLlama):
            pass
        ow = RM.quantize_model(w, "fp16")
        for li, lw in enumerate(w["layers"]):
            for name, dim in (("qkv", 0), ("gate", 0), ("up", 0), ("dense", 1), ("down", 1)):
                if mode == "fp16":
                    continue
                if dim == 0:
                    ql = RM.quantize_linear(lw[name], mode)
                    deq = (ql["q"].astype(np.float16) * ql["scales"][:, None]).astype(np.float16)
                else:
                    parts = np.split(lw[name], world, axis=1)
                    deq = np.concatenate([(RM.quantize_linear(p, mode)["q"].astype(np.float16) *
                                           RM.quantize_linear(p, mode)["scales"][:, None]).astype(np.float16) for p in parts], 1)
                ow["layers"][li][name] = {"w": deq}
        ref_ids, ref = RM.OracleLlama(cfg, ow, "fp16", int8_kv, kv_scale=4.0 / 127.0, max_seq_len=S + new).generate(
            ids, lens, new, return_logits=True)
        tol = 2e-2 * max(1.0, float(np.abs(ref).max()))
        for path, (got, got_ids, _) in runs.items():
            for s in range(new):
                np.testing.assert_allclose(got[:, s], ref[:, s], atol=tol, err_msg=f"{path} step {s}")
                top2 = np.sort(ref[:, s], -1)[:, -2:]
                dec = (top2[:, 1] - top2[:, 0]) > 2 * tol
                assert np.array_equal(got_ids[dec, s], ref_ids[dec, s]), f"{path} step {s}"
                if not np.array_equal(got_ids[:, s], ref_ids[:, s]):
                    break
        # both paths round at the same points (fp16 partial -> fp32 rank-ordered sum -> fp16 -> + residual)
        same = np.array_equal(runs["fused"][1], runs["plugins"][1])
        print(f"TP{world} {mode} OK maxdiff fused {np.abs(runs['fused'][0] - ref).max():.4f} plugins "
              f"{np.abs(runs['plugins'][0] - ref).max():.4f} tol {tol:.4f} ids_equal_between_paths {same}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
