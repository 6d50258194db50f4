#!/usr/bin/env python
"""examples/llama_quant/build.py of the reference, on the B200 plugin engine.

Keeps the reference's flags (LQ/build.py:40-273) and outputs (``llama_{dtype}_tp{N}_rank{r}.engine``, ``config.json``).
Flags that only steer TensorRT (timing cache, builder_opt, *_plugin dtype selectors) are accepted and recorded."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

MODEL_NAME = "llama"


def parse_arguments():
    p = argparse.ArgumentParser()
    p.add_argument('--world_size', type=int, default=1, help='world size, only support tensor parallelism now')
    p.add_argument('--model_dir', type=str, default=None)
    p.add_argument('--dtype', type=str, default='float16', choices=['float16', 'float32', 'bfloat16'])
    p.add_argument('--timing_cache', type=str, default='model.cache')
    p.add_argument('--log_level', type=str, default='info')
    p.add_argument('--vocab_size', type=int, default=32000)
    p.add_argument('--n_layer', type=int, default=32)
    p.add_argument('--n_positions', type=int, default=2048)
    p.add_argument('--n_embd', type=int, default=4096)
    p.add_argument('--n_head', type=int, default=32)
    p.add_argument('--n_kv_head', type=int, default=None)
    p.add_argument('--hidden_act', type=str, default='silu')
    p.add_argument('--inter_size', type=int, default=11008)
    p.add_argument('--no_bias', action="store_false", default=True)
    p.add_argument('--max_batch_size', type=int, default=8)
    p.add_argument('--max_input_len', type=int, default=2048)
    p.add_argument('--max_output_len', type=int, default=512)
    p.add_argument('--max_beam_width', type=int, default=1)
    p.add_argument('--use_gpt_attention_plugin', nargs='?', const='float16', default=False,
                   choices=['float16', 'float32', 'bfloat16'])
    p.add_argument('--use_gemm_plugin', nargs='?', const='float16', default=False, choices=['float16', 'float32', 'bfloat16'])
    p.add_argument('--parallel_build', default=False, action='store_true')
    p.add_argument('--gpus_per_node', type=int, default=8)
    p.add_argument('--builder_opt', type=int, default=None)
    p.add_argument('--output_dir', type=str, default='llama_outputs')
    p.add_argument('--remove_input_padding', default=False, action='store_true')
    p.add_argument('--use_smooth_quant', default=False, action="store_true")
    p.add_argument('--use_weight_only', default=False, action="store_true")
    p.add_argument('--weight_only_precision', const='int8', type=str, nargs='?', default='int8', choices=['int8', 'int4'])
    p.add_argument('--per_channel', default=False, action="store_true")
    p.add_argument('--per_token', default=False, action="store_true")
    p.add_argument('--int8_kv_cache', default=False, action="store_true")
    p.add_argument('--random_seed', type=int, default=None)
    p.add_argument('--paged_kv_cache', action="store_true", default=False)
    p.add_argument('--tokens_per_block', type=int, default=64, help='paged KV cache block size (a power of two >= 16)')
    args = p.parse_args()
    if args.dtype != 'float16':
        p.error("only --dtype float16 is built on this path")
    if args.n_kv_head not in (None, args.n_head):
        p.error("multi-query attention is out of scope (DESIGN.md)")
    if not 1 <= args.max_beam_width <= 16:
        p.error("--max_beam_width must be in [1, 16]")
    if args.max_beam_width > 1 and args.paged_kv_cache:
        p.error("beam search reads the contiguous KV cache: --paged_kv_cache needs --max_beam_width 1")
    if args.use_smooth_quant and not (args.per_token and args.per_channel):
        p.error("SmoothQuant is built for --per_token --per_channel")
    return args


def main():
    args = parse_arguments()
    import torch
    from trtllm_llama_b200 import builder as B
    from trtllm_llama_b200.runtime import ModelConfig
    tik = time.time()
    os.makedirs(args.output_dir, exist_ok=True)
    qm = B.quant_mode_from_args(args)
    mc = ModelConfig(vocab_size=args.vocab_size, num_layers=args.n_layer, num_heads=args.n_head, hidden_size=args.n_embd,
                     inter_size=args.inter_size, quant_mode=qm,
                     max_batch_size=args.max_batch_size * args.max_beam_width,     # rows = batch entries x beams
                     max_input_len=args.max_input_len, max_output_len=args.max_output_len, tp_size=args.world_size,
                     paged_kv_cache=args.paged_kv_cache, tokens_per_block=args.tokens_per_block,
                     remove_input_padding=args.remove_input_padding)
    dev = "cuda" if torch.cuda.is_available() else "cpu"       # quantisation is build-time work; a GPU only makes it fast
    weights = (B.load_from_ft_llama(args.model_dir, mc, dev) if args.model_dir
               else B.random_llama_weights(mc, seed=args.random_seed or 0, device=dev))
    for rank in range(args.world_size):
        tensors = B.build_rank_engine(weights, mc, rank, require_kv_scale=bool(args.model_dir))
        name = B.get_engine_name(MODEL_NAME, args.dtype, args.world_size, rank)
        B.serialize_engine(tensors, os.path.join(args.output_dir, name))
        print(f"[build] serialized {name}: {sum(t.numel() * t.element_size() for t in tensors.values()) / 2**30:.2f} GiB")
    B.save_config(os.path.join(args.output_dir, "config.json"), precision=args.dtype, world_size=args.world_size, mc=mc,
                  plugin_config={"gpt_attention_plugin": args.use_gpt_attention_plugin or "float16",
                                 "gemm_plugin": args.use_gemm_plugin or "float16",
                                 "smooth_quant_gemm_plugin": "float16" if args.use_smooth_quant else False,
                                 "weight_only_quant_matmul_plugin": "float16" if args.use_weight_only else False,
                                 "rmsnorm_quantization_plugin": "float16" if args.use_smooth_quant else False,
                                 "nccl_plugin": "float16" if args.world_size > 1 else False,
                                 "remove_input_padding": bool(args.remove_input_padding), "paged_kv_cache": bool(args.paged_kv_cache),
                                 "tokens_per_block": args.tokens_per_block})
    print(f"Total time of building all {args.world_size} engines: {time.strftime('%H:%M:%S', time.gmtime(time.time() - tik))}")


if __name__ == '__main__':
    main()
