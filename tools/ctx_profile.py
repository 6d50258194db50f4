import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import LLAMA7B, make_weights
from trtllm_llama_b200 import runtime as rt
from trtllm_llama_b200._lib import lib
from trtllm_llama_b200.quantization import QuantMode
B, S = 8, 2048
cfg = dict(LLAMA7B); cfg["layers"] = 2
qm = QuantMode.use_smooth_quant(True, True) | QuantMode.INT8_KV_CACHE
mc = rt.ModelConfig(vocab_size=32000, num_layers=2, num_heads=32, hidden_size=4096, inter_size=11008, quant_mode=qm, max_batch_size=B, max_input_len=S, max_output_len=8)
w = make_weights(torch, cfg, 0, 1)
sess = rt.GenerationSession(mc, rt.build_engine_tensors(w, mc))
ids = torch.randint(3, 32000, (B, S), dtype=torch.int32, device="cuda"); lens = torch.full((B,), S, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    assert lib.tbrt_context(sess._e, ids.data_ptr(), lens.data_ptr(), B, S, st) == 0
torch.cuda.synchronize()
