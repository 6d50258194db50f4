"""Development aid: CTA-pair tcgen05 GEMM (gemm_tc2.cu, force_nt=512) against the one-CTA kernel (force_nt=256) and a
torch reference, then timings of the LLaMA-7B prefill projections (M = 16384) through both."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import trtllm_llama_b200  # noqa
from trtllm_llama_b200 import ops

torch.manual_seed(0)
ok = True
for (M, N, K) in [(300, 384, 256), (256, 256, 128), (1000, 640, 512), (2048, 4096, 4096), (513, 11008, 1024)]:
    a = torch.randint(-127, 128, (M, K), device="cuda", dtype=torch.int8)
    b = torch.randint(-127, 128, (N, K), device="cuda", dtype=torch.int8)
    st = torch.rand(M, 1, device="cuda") * 0.01 + 1e-3
    sc = torch.rand(1, N, device="cuda") * 0.01 + 1e-3
    one = ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st, force_nt=256)
    two = ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st, force_nt=512)
    i32 = ops.gemm_tc(ops.KIND_A8W8, a, b, sc=torch.ones(1, 1, device="cuda"), sr=torch.ones(1, 1, device="cuda"),
                      out_dtype=torch.int32, force_nt=512)
    ref = (a.double() @ b.double().t()).to(torch.int32) if M * N * K < 2e10 else None
    e_i8 = bool(torch.equal(one, two))
    e_ref = bool(torch.equal(i32, ref)) if ref is not None else None
    x = (torch.randn(M, K, device="cuda") * 0.5).half()
    w = (torch.randn(N, K, device="cuda") * 0.05).half()
    r = (torch.randn(M, N, device="cuda")).half()
    f1 = ops.gemm_tc(ops.KIND_F16, x, w, residual=r, force_nt=256)
    f2 = ops.gemm_tc(ops.KIND_F16, x, w, residual=r, force_nt=512)
    e_f16 = bool(torch.equal(f1, f2))
    err = (f2.float() - (x.float() @ w.float().t() + r.float())).abs().max().item()
    print(json.dumps({"shape": [M, N, K], "i8_pair_eq_single": e_i8, "i32_eq_exact": e_ref, "f16_pair_eq_single": e_f16,
                      "f16_max_err_vs_fp32": round(err, 5)}), flush=True)
    ok = ok and e_i8 and (e_ref is not False) and e_f16
print("PARITY", "OK" if ok else "FAILED", flush=True)

M = 16384
for name, N, K in [("qkv", 12288, 4096), ("dense", 4096, 4096), ("gate_up", 22016, 4096), ("down", 4096, 11008)]:
    a = torch.randint(-127, 127, (M, K), device="cuda", dtype=torch.int8)
    b = torch.randint(-127, 127, (N, K), device="cuda", dtype=torch.int8)
    st = torch.rand(M, 1, device="cuda") * 0.01
    sc = torch.rand(1, N, device="cuda") * 0.01
    out = {}
    for label, nt in [("single", 256), ("pair", 512)]:
        for _ in range(3):
            ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st, force_nt=nt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st, force_nt=nt)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out[label] = {"ms": round(ms, 3), "int8_TOPS": round(2.0 * M * N * K / (ms * 1e-3) / 1e12, 1)}
    print(json.dumps({name: out}), flush=True)
    del a, b
x = (torch.randn(M, 4096, device="cuda") * 0.5).half()
w = (torch.randn(12288, 4096, device="cuda") * 0.05).half()
for label, nt in [("single", 256), ("pair", 512)]:
    for _ in range(3):
        ops.gemm_tc(ops.KIND_F16, x, w, force_nt=nt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm_tc(ops.KIND_F16, x, w, force_nt=nt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"fp16_qkv_" + label: {"ms": round(ms, 3), "TFLOPS": round(2.0 * M * 12288 * 4096 / (ms * 1e-3) / 1e12, 1)}}), flush=True)
