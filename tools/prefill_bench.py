"""BASELINE config 4: LLaMA-7B SmoothQuant per-token/per-channel int8, batch 8, prefill 2048 (M = 16384 token rows).
Times (CUDA events) each projection GEMM of a layer through the SmoothQuantGemm path (tcgen05 kind::i8) and the whole
context phase through the engine; reports int8 TOPS against 2 x the measured bf16 peak (the int8 peak itself is not in
MEASURED_PEAKS.json) and the datasheet 4.5 POPS."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import trtllm_llama_b200  # noqa
from trtllm_llama_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
only_gemm = len(sys.argv) > 2 and sys.argv[2] == "gemm"
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
res = {"M": M, "gemms": {}}
tot_ops, tot_ms = 0.0, 0.0
for name, N, K in [("qkv", 12288, 4096), ("dense", 4096, 4096), ("gate_up", 22016, 4096), ("down", 4096, 11008)]:
    a = torch.randint(-127, 127, (M, K), device="cuda", dtype=torch.int8)
    b = torch.randint(-127, 127, (N, K), device="cuda", dtype=torch.int8)
    st = torch.rand(M, 1, device="cuda") * 0.01
    sc = torch.rand(1, N, device="cuda") * 0.01
    for _ in range(3):
        ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tops = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    res["gemms"][name] = {"ms": round(ms, 3), "int8_TOPS": round(tops, 1)}
    tot_ops += 2.0 * M * N * K; tot_ms += ms
    del a, b
res["layer_gemm_ms"] = round(tot_ms, 3)
res["layer_int8_TOPS"] = round(tot_ops / (tot_ms * 1e-3) / 1e12, 1)
res["frac_of_2x_bf16_measured"] = round(res["layer_int8_TOPS"] / (2 * peaks["bf16_tflops"]), 4)
res["frac_of_4500_datasheet"] = round(res["layer_int8_TOPS"] / 4500.0, 4)
print(json.dumps(res), flush=True)
if only_gemm:
    sys.exit(0)

# whole context phase through the engine (SQ + int8 KV), batch 8 x 2048
from bench import LLAMA7B, make_weights
from trtllm_llama_b200 import runtime as rt
from trtllm_llama_b200._lib import lib
from trtllm_llama_b200.quantization import QuantMode
B, S = 8, M // 8
qm = QuantMode.use_smooth_quant(True, True) | QuantMode.INT8_KV_CACHE
mc = rt.ModelConfig(vocab_size=32000, num_layers=32, num_heads=32, hidden_size=4096, inter_size=11008, quant_mode=qm,
                    max_batch_size=B, max_input_len=S, max_output_len=8)
w = make_weights(torch, LLAMA7B, 0, 1)
tensors = rt.build_engine_tensors(w, mc)
del w
torch.cuda.empty_cache()
sess = rt.GenerationSession(mc, tensors)
ids = torch.randint(3, 32000, (B, S), dtype=torch.int32, device="cuda")
lens = torch.full((B,), S, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    assert lib.tbrt_context(sess._e, ids.data_ptr(), lens.data_ptr(), B, S, st) == 0, lib.tbrt_last_error()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    lib.tbrt_context(sess._e, ids.data_ptr(), lens.data_ptr(), B, S, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
gemm_ops = 2.0 * M * 6476005376
print(json.dumps({"context_ms": round(ms, 2), "prefill_tokens_per_s": round(M / (ms * 1e-3), 0),
                  "int8_TOPS_incl_attention_and_glue": round(gemm_ops / (ms * 1e-3) / 1e12, 1),
                  "roofline_ms_at_2x_bf16": round(gemm_ops / (2 * peaks["bf16_tflops"] * 1e12) * 1e3, 1)}), flush=True)
