"""CPU baseline: the reference's ``run_hf.py`` path (HF ``AutoModelForCausalLM.generate``, ``top_k=1``,
``num_beams=1``; LQ/run_hf.py:50-84) in fp32 on the host cores, with seeded random-init LLaMA-7B weights
(no checkpoint exists offline).  TEST / MEASUREMENT INFRASTRUCTURE ONLY: used by ``bench.py``'s
``cpu_baseline`` leg and ``--impl reference`` arm, never by the product path.

The full 128-in/128-out request would take minutes on host cores, so a BOUNDED sample is timed — the 128-token
prefill plus ``new_tokens`` greedy steps — and the rate for the full request is extrapolated as
``out_len / (t_prefill + out_len * t_step)`` (stated in the JSON ``sample`` field)."""
from __future__ import annotations

import os
import time


def build_hf_llama(hidden=4096, inter=11008, layers=32, heads=32, vocab=32000, seed=0):
    import torch
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers, num_attention_heads=heads,
                      num_key_value_heads=heads, vocab_size=vocab, rms_norm_eps=1e-6, max_position_embeddings=2048,
                      tie_word_embeddings=False)
    # meta-device construction + vectorised fill: HF's own init loop is several minutes for 6.7 B parameters
    with torch.device("meta"):
        model = LlamaForCausalLM(cfg)
    model = model.to_empty(device="cpu")
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1:
                p.fill_(1.0)
            else:
                p.uniform_(-0.035, 0.035, generator=g)     # std ~= 0.02 (HF initializer_range)
        for name, b in model.named_buffers():
            if "inv_freq" in name:
                dim = hidden // heads
                b.copy_(1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim)))
    model.eval()
    return model


def time_hf_cpu(batch=1, in_len=128, out_len=128, new_tokens=8, threads=None, seed=1234, **model_kw):
    """returns dict(value tokens/s extrapolated, cores, sample, t_prefill, t_step)."""
    import torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = build_hf_llama(**model_kw)
    g = torch.Generator().manual_seed(seed)
    vocab = model.config.vocab_size
    ids = torch.randint(3, vocab, (batch, in_len), generator=g)
    with torch.no_grad():
        model(input_ids=ids[:, :4], use_cache=True)        # untimed warm-up (thread pool, allocator, rotary cache)
        t0 = time.perf_counter()
        out = model(input_ids=ids, use_cache=True)
        t_prefill = time.perf_counter() - t0
        past = out.past_key_values
        tok = out.logits[:, -1].argmax(-1, keepdim=True)
        t1 = time.perf_counter()
        for _ in range(new_tokens):
            out = model(input_ids=tok, past_key_values=past, use_cache=True)
            past = out.past_key_values
            tok = out.logits[:, -1].argmax(-1, keepdim=True)
        t_step = (time.perf_counter() - t1) / new_tokens
    value = batch * out_len / (t_prefill + out_len * t_step)
    return {"value": value, "cores": threads, "t_prefill_s": t_prefill, "t_step_s": t_step,
            "sample": f"HF LlamaForCausalLM fp32 (random-init 7B), batch {batch}: {in_len}-token prefill + {new_tokens} greedy "
                      f"steps timed; rate extrapolated to {out_len} new tokens as out/(t_prefill + out*t_step)"}
