"""Development aid: the tensor-core decode GEMV (gemv_mma.cu) on the LLaMA-7B projection shapes, timed by replaying a
CUDA graph of 20 back-to-back launches over 8 rotating weight copies (no L2 reuse).  argv: kind (w8|w4|f16|sq) M"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import trtllm_llama_b200  # noqa
from trtllm_llama_b200 import ops

mode = sys.argv[1] if len(sys.argv) > 1 else "w8"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kind = {"f16": ops.KIND_F16, "w8": ops.KIND_W8, "w4": ops.KIND_W4, "sq": ops.KIND_A8W8}[mode]
bpw = {"f16": 2.0, "w8": 1.0, "w4": 0.5, "sq": 1.0}[mode]
os.environ.setdefault("TB_GEMV_MMA_MIN_M", "1")
only = sys.argv[3] if len(sys.argv) > 3 else None
for name, N, K, swiglu, pro in [s for s in [("qkv", 12288, 4096, False, 1), ("dense", 4096, 4096, False, 0),
                                 ("gate_up", 22016, 4096, True, 1), ("down", 4096, 11008, False, 0)] if only in (None, s[0])]:
    copies = 6
    if mode == "f16":
        ws = [(torch.randn(N, K, device="cuda") * 0.05).half() for _ in range(copies)]
    else:
        ws = [torch.randint(-127, 127, (N, int(K * bpw)), device="cuda", dtype=torch.int8) for _ in range(copies)]
    scale = torch.rand(N, device="cuda").half() * 0.01
    sc = torch.rand(1, N, device="cuda") * 0.01
    x = (torch.randn(M, K, device="cuda") * 0.5).half()
    gamma = torch.ones(K, device="cuda").half()
    if mode == "sq":
        pro = 2 if pro == 1 else 3
    def call(w):
        if mode == "sq":
            return ops.gemv(kind, x, w, sc=sc, swiglu=swiglu, prologue=pro, gamma=gamma)
        return ops.gemv(kind, x, w, w_scale=None if mode == "f16" else scale, swiglu=swiglu, prologue=pro, gamma=gamma)
    for w in ws:
        call(w)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for i in range(24):
                call(ws[i % copies])
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * 24)
    nbytes = N * K * bpw
    print(json.dumps({"mode": mode, "M": M, "shape": name, "us": round(us, 2), "GBps": round(nbytes / us / 1e3, 0)}), flush=True)
    del ws
