"""Development aid: decode attention at cfg3 sizes (B = 8, 32 heads, 2040 cached positions) for ncu.  argv: int8|fp16"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import trtllm_llama_b200  # noqa
from trtllm_llama_b200 import ops

int8 = (sys.argv[1] if len(sys.argv) > 1 else "int8") == "int8"
B, H, Dh, S_max, past = 8, 32, 128, 2048, 2040
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B, 3 * H * Dh, device="cuda", generator=g)).half()
if int8:
    cache = torch.randint(-127, 128, (B, 2, H, S_max, Dh), device="cuda", dtype=torch.int8, generator=g)
    s_q = torch.tensor([127.0 / 4.0], device="cuda")
    s_dq = torch.tensor([4.0 / 127.0], device="cuda")
else:
    cache = torch.randn(B, 2, H, S_max, Dh, device="cuda", generator=g).half()
    s_q = s_dq = None
for _ in range(4):
    ops.mmha_decode(qkv, cache, past, num_heads=H, head_size=Dh, max_input_len=1920, kv_scale_orig_quant=s_q,
                    kv_scale_quant_orig=s_dq)
torch.cuda.synchronize()
