import sys, torch
sys.path.insert(0, "/root/repo")
import trtllm_llama_b200  # noqa
from trtllm_llama_b200 import ops
nt = int(sys.argv[1])
M, N, K = 15360, 12288, 4096
x = (torch.randn(M, K, device="cuda") * 0.5).half()
w = (torch.randn(N, K, device="cuda") * 0.05).half()
for _ in range(3):
    ops.gemm_tc(ops.KIND_F16, x, w, force_nt=nt)
torch.cuda.synchronize()
