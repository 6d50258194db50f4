"""Development aid: the prefill attention (ctx_prep + flash_ctx_tc) alone at the cfg4 shape (B=8, S=2048, H=32, int8 KV):
CUDA events around 10 back-to-back calls, prints us per call; with `ncu` as the parent only the kernels matter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import trtllm_llama_b200  # noqa
from trtllm_llama_b200 import ops
B, S, H, D = 8, 2048, 32, 128
qkv = (torch.randn(B, S, 3 * H * D, device="cuda") * 0.5).half()
lens = torch.full((B,), S, dtype=torch.int32, device="cuda")
cache = torch.zeros((B, 2, H, S + 128, D), dtype=torch.int8, device="cuda")
sc = torch.tensor([127.0 / 2.5], device="cuda")
def call():
    return ops.context_attention(qkv, cache, lens, num_heads=H, head_size=D, use_tc=True, kv_scale_orig_quant=sc)
for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    call()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
flops = 4 * D * H * B * (S * (S + 1) / 2)
print({"us_per_call_prep_plus_flash": round(us, 1), "causal_TFLOPs_over_whole_call": round(flops / us / 1e6, 1)})
