#!/usr/bin/env python
"""examples/llama_quant/run_hf.py of the reference (LQ/run_hf.py:20-108): HF ``generate`` latency loop, ``top_k=1``.
BASELINE config 1 re-targets it to fp32 on the host CPU; without a checkpoint directory a seeded random-init
LLaMA of the requested size is used and the prompt is synthetic token ids."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def parse_arguments():
    p = argparse.ArgumentParser()
    p.add_argument('--max_output_len', type=int, required=True)
    p.add_argument('--log_level', type=str, default='error')
    p.add_argument('--hf_model_location', type=str, default=None)
    p.add_argument('--tokenizer_dir', type=str, default=".")
    p.add_argument('--input_text', type=str, default='Born in north-east France, Soyer trained as a')
    p.add_argument('--num_beams', type=int, default=1)
    p.add_argument('--input_len', type=int, default=128, help="synthetic prompt length when no tokenizer is available")
    p.add_argument('--iterations', type=int, default=2)
    p.add_argument('--n_layer', type=int, default=32)
    return p.parse_args()


def main():
    args = parse_arguments()
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    if args.hf_model_location and os.path.isdir(args.hf_model_location):
        from transformers import AutoModelForCausalLM
        model = AutoModelForCausalLM.from_pretrained(args.hf_model_location).float().eval()
    else:
        from oracle.hf_baseline import build_hf_llama       # measurement infrastructure, random-init weights
        model = build_hf_llama(layers=args.n_layer)
    tok_path = os.path.join(args.tokenizer_dir, "tokenizer.model")
    if os.path.exists(tok_path):
        from transformers import LlamaTokenizer
        ids = torch.tensor([LlamaTokenizer.from_pretrained(args.tokenizer_dir, legacy=False).encode(args.input_text)])
    else:
        ids = torch.randint(3, model.config.vocab_size, (1, args.input_len), generator=torch.Generator().manual_seed(1234))
    lat = []
    for _ in range(args.iterations):
        t0 = time.time()
        with torch.no_grad():
            out = model.generate(ids, max_new_tokens=args.max_output_len, do_sample=False, num_beams=args.num_beams,
                                 top_k=None, temperature=None, top_p=None, pad_token_id=2, eos_token_id=None)
        lat.append(time.time() - t0)
    print(f'Output ids: {out[0, ids.shape[1]:].tolist()}')
    print(f'HF mean latency: {np.mean(lat[1:] or lat):.5f} sec on {torch.get_num_threads()} CPU threads '
          f'({args.max_output_len / np.mean(lat[1:] or lat):.2f} tokens/s)')


if __name__ == '__main__':
    main()
