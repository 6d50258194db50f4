"""A/B for the SwiGLU-in-epilogue gate / up GEMM at the BASELINE configs[3] shape (M = 16384, inter 11008, K 4096, SmoothQuant):
  two kernels:  tb_gemm_tc (gate|up, fp16 out) + tb_swiglu_quant
  fused:        tb_gemm_tc_swiglu + tb_quantize_per_token
CUDA events, 10 timed iterations each after 3 warm-ups."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import trtllm_llama_b200  # noqa
from trtllm_llama_b200 import ops

M, inter, K = 16384, 11008, 4096
a = torch.randint(-127, 127, (M, K), device="cuda", dtype=torch.int8)
b = torch.randint(-127, 127, (2 * inter, K), device="cuda", dtype=torch.int8)
st = torch.rand(M, 1, device="cuda") * 0.001
sc = torch.rand(1, 2 * inter, device="cuda") * 0.001


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10


res = {"gemm_gate_up_ms": timed(lambda: ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st)),
       "gemm_swiglu_fused_ms": timed(lambda: ops.gemm_tc_swiglu(ops.KIND_A8W8, a, b, sc=sc, sr=st))}
gu = ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st)
act = ops.gemm_tc_swiglu(ops.KIND_A8W8, a, b, sc=sc, sr=st)
res["swiglu_quant_ms"] = timed(lambda: ops.swiglu_quant(gu))
res["per_token_quant_ms"] = timed(lambda: ops.quantize_per_token(act))
res["two_kernels_total_ms"] = res["gemm_gate_up_ms"] + res["swiglu_quant_ms"]
res["fused_total_ms"] = res["gemm_swiglu_fused_ms"] + res["per_token_quant_ms"]
ops_ = 2.0 * M * 2 * inter * K
res["gemm_TOPS"] = ops_ / res["gemm_gate_up_ms"] / 1e9
res["fused_gemm_TOPS"] = ops_ / res["gemm_swiglu_fused_ms"] / 1e9
q1, s1 = ops.swiglu_quant(gu)
q2, s2 = ops.quantize_per_token(act)
res["identical"] = bool(torch.equal(q1, q2) and torch.equal(s1.flatten(), s2.flatten()))
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))
