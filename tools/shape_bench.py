"""Development aid: time the decode-shape projections of LLaMA-7B through tb_gemv and tb_gemm_tc (CUDA events,
rotating over distinct weight copies so nothing is L2-resident)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import trtllm_llama_b200  # noqa
from trtllm_llama_b200 import ops

def bench(fn, n=40):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

shapes = [("qkv", 12288, 4096), ("dense", 4096, 4096), ("gate_up", 22016, 4096), ("down", 4096, 11008), ("lm_head", 32000, 4096)]
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp16", "w8", "w4", "sq"]
Ms = [int(a) for a in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "8"])]
for mode in modes:
    for name, N, K in shapes:
        copies = max(2, int(400e6 // (N * K * {"fp16": 2, "w8": 1, "w4": 0.5, "sq": 1}[mode])) + 1)
        if mode == "fp16":
            ws = [(torch.randn(N, K, device="cuda") * 0.02).half() for _ in range(copies)]
        elif mode == "w4":
            ws = [torch.randint(-128, 127, (N, K // 2), device="cuda", dtype=torch.int8) for _ in range(copies)]
        else:
            ws = [torch.randint(-128, 127, (N, K), device="cuda", dtype=torch.int8) for _ in range(copies)]
        sc16 = torch.ones(N, device="cuda", dtype=torch.float16) * 0.01
        sc32 = torch.ones(1, N, device="cuda", dtype=torch.float32) * 0.01
        nbytes = ws[0].numel() * ws[0].element_size()
        for M in Ms:
            x16 = (torch.randn(M, K, device="cuda") * 0.1).half()
            x8 = torch.randint(-127, 127, (M, K), device="cuda", dtype=torch.int8)
            st = torch.ones(M, 1, device="cuda", dtype=torch.float32)
            kind = {"fp16": ops.KIND_F16, "w8": ops.KIND_W8, "w4": ops.KIND_W4, "sq": ops.KIND_A8W8}[mode]
            def tc(i):
                w = ws[i % copies]
                if mode == "fp16": ops.gemm_tc(kind, x16, w)
                elif mode == "sq": ops.gemm_tc(kind, x8, w, sc=sc32, sr=st)
                else: ops.gemm_tc(kind, x16, w, w_scale=sc16)
            def gv(i):
                w = ws[i % copies]
                if mode == "fp16": ops.gemv(kind, x16, w)
                elif mode == "sq": ops.gemv(kind, x8, w, sc=sc32, sr=st)
                else: ops.gemv(kind, x16, w, w_scale=sc16)
            t_tc = bench(tc)
            line = f"{mode:5s} {name:8s} M={M}: gemm_tc {t_tc:7.2f} us {nbytes / t_tc / 1e3:7.0f} GB/s"
            if M <= 4:
                t_gv = bench(gv)
                line += f" | gemv {t_gv:7.2f} us {nbytes / t_gv / 1e3:7.0f} GB/s"
            print(line, flush=True)
        del ws
        torch.cuda.empty_cache()
