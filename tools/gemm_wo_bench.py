"""Development aid: prefill-size projections (M = 15360 = cfg3's 8 x 1920 prompt rows) through the fp16, weight-only
int8 and weight-only int4 tcgen05 GEMMs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import trtllm_llama_b200  # noqa
from trtllm_llama_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 15360
for name, N, K in [("qkv", 12288, 4096), ("dense", 4096, 4096), ("gate_up", 22016, 4096), ("down", 4096, 11008)]:
    x = (torch.randn(M, K, device="cuda") * 0.5).half()
    out = {}
    for mode in ("f16", "w8", "w4"):
        if mode == "f16":
            w = (torch.randn(N, K, device="cuda") * 0.05).half()
            f = lambda: ops.gemm_tc(ops.KIND_F16, x, w)
        else:
            w = torch.randint(-127, 127, (N, K if mode == "w8" else K // 2), device="cuda", dtype=torch.int8)
            sc = torch.rand(N, device="cuda").half() * 0.01
            kind = ops.KIND_W8 if mode == "w8" else ops.KIND_W4
            f = lambda: ops.gemm_tc(kind, x, w, w_scale=sc)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out[mode] = {"ms": round(ms, 3), "TFLOPS": round(2.0 * M * N * K / (ms * 1e-3) / 1e12, 1)}
        del w
    print(json.dumps({name: out}), flush=True)
