#!/usr/bin/env python
"""examples/llama_quant/summarize.py of the reference (LQ/summarize.py:65-362), degraded as SURVEY.md §7 prescribes:
cnn_dailymail, the ``rouge`` metric, LLaMA weights and ``tokenizer.model`` are not available offline, so the
"articles" are seeded synthetic token sequences and the accuracy check is token-level agreement between the engine
and the reference model (the CPU oracle with the same weights, standing in for HF) instead of ROUGE-1."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def main(args):
    import torch
    from trtllm_llama_b200 import builder as B
    from trtllm_llama_b200 import runtime as rt
    from trtllm_llama_b200._lib import lib
    # flags of the reference CLI this path does not honour are REJECTED, never silently ignored
    if args.num_beams != 1 and args.top_k != 1:
        raise SystemExit("--num_beams > 1 does not combine with --top_k sampling (as in the reference's decoder)")
    if args.test_hf:
        raise SystemExit("--test_hf: no HF checkpoint / tokenizer offline; run_hf.py times the HF path on synthetic weights")
    if args.top_k < 0:
        raise SystemExit("--top_k must be >= 0")
    # one process per rank, as the reference under mpirun (LQ/summarize.py:65-110); here torchrun provides RANK / WORLD_SIZE
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    mc = B.model_config_from_json(os.path.join(args.engine_dir, "config.json"), rank)
    assert world == mc.tp_size, f'Engine world size ({mc.tp_size}) != Runtime world size ({world})'
    torch.cuda.set_device(rank % torch.cuda.device_count())
    if world > 1:
        from run import setup_tp
        setup_tp(lib, world, rank)
    tensors = B.deserialize_engine(os.path.join(args.engine_dir, B.get_engine_name("llama", "float16", world, rank)))
    session = rt.GenerationSession(mc, tensors)
    if world > 1:
        session.enable_peer_allreduce()
    # LQ/summarize.py:125-140: top_k = 1 is greedy; larger values sample (seeded) among the k best
    sampling = rt.SamplingConfig(end_id=None, top_k=args.top_k, random_seed=args.random_seed) if args.top_k != 1 else None
    if args.num_beams != 1:       # LQ/summarize.py:125-140 passes num_beams to the decoder; the best beam is summarised
        sampling = rt.SamplingConfig(end_id=2, pad_id=2, num_beams=args.num_beams)
    rng = np.random.default_rng(0)
    max_in = min(args.max_input_len, mc.max_input_len)
    out_len = min(args.output_len, mc.max_output_len)
    total, agree, n_tok = 0.0, 0, 0
    for ite in range(args.max_ite):
        lens = rng.integers(max_in // 2, max_in + 1, args.batch_size).astype(np.int32)
        ids = np.full((args.batch_size, int(lens.max())), 2, np.int32)
        for b, L in enumerate(lens):
            ids[b, :L] = rng.integers(3, mc.vocab_size, L)
        session.setup(args.batch_size, ids.shape[1], out_len, beam_width=args.num_beams)
        t0 = time.time()
        out = session.decode(torch.from_numpy(ids).pin_memory(), torch.from_numpy(lens).pin_memory(), sampling).numpy()
        total += time.time() - t0
        if args.num_beams != 1:
            out = out[:, 0]
        if args.check_accuracy and args.oracle_weights and rank == 0 and sampling is None:
            from oracle import ref_model as RM            # checker only (tests/bench infrastructure)
            w = np.load(args.oracle_weights, allow_pickle=True).item()
            mode = {0: "fp16", 1: "w8", 2: "w4", 3: "sq"}[mc.mode]
            cfg = RM.LlamaCfg(hidden=mc.hidden_size, heads=mc.num_heads, inter=mc.inter_size, layers=mc.num_layers,
                              vocab=mc.vocab_size)
            ref = RM.OracleLlama(cfg, RM.quantize_model(w, mode), mode, mc.quant_mode.has_int8_kv_cache(),
                                 max_seq_len=ids.shape[1] + out_len).generate(ids, lens, out_len)
            agree += int((ref == out).sum())
            n_tok += out.size
    if rank != 0:
        return
    print(f'TensorRT-LLM (total latency: {total:.3f} sec)')
    print(f'TensorRT-LLM tokens/s: {args.max_ite * args.batch_size * out_len / total:.1f}')
    if n_tok:
        rate = 100.0 * agree / n_tok
        print(f'token agreement with the reference model: {rate:.2f} %')
        if args.check_accuracy:
            assert rate >= args.agreement_threshold, f"agreement {rate:.2f} % below {args.agreement_threshold} %"


if __name__ == '__main__':
    p = argparse.ArgumentParser()
    p.add_argument('--hf_model_location', type=str, default=None)
    p.add_argument('--test_hf', action='store_true')
    p.add_argument('--test_trt_llm', action='store_true')
    p.add_argument('--data_type', type=str, choices=['fp32', 'fp16'], default='fp16')
    p.add_argument('--dataset_path', type=str, default='')
    p.add_argument('--log_level', type=str, default='info')
    p.add_argument('--engine_dir', type=str, default='llama_outputs')
    p.add_argument('--batch_size', type=int, default=1)
    p.add_argument('--max_ite', type=int, default=20)
    p.add_argument('--check_accuracy', action='store_true')
    p.add_argument('--tensorrt_llm_rouge1_threshold', type=float, default=15.0)
    p.add_argument('--agreement_threshold', type=float, default=90.0)
    p.add_argument('--oracle_weights', type=str, default=None, help=".npy dict of fp16 weights for the agreement check")
    p.add_argument('--num_beams', type=int, default=1)
    p.add_argument('--top_k', type=int, default=1)
    p.add_argument('--random_seed', type=int, default=0)
    p.add_argument('--max_input_len', type=int, default=923)     # LQ/summarize.py:91-92
    p.add_argument('--output_len', type=int, default=100)
    main(p.parse_args())
