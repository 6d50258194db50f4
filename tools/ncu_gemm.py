import sys, torch
sys.path.insert(0, "/root/repo")
import trtllm_llama_b200  # noqa
from trtllm_llama_b200 import ops
nt = int(sys.argv[1])
M, N, K = 16384, 12288, 4096
a = torch.randint(-127, 127, (M, K), device="cuda", dtype=torch.int8)
b = torch.randint(-127, 127, (N, K), device="cuda", dtype=torch.int8)
st = torch.rand(M, 1, device="cuda") * 0.01
sc = torch.rand(1, N, device="cuda") * 0.01
for _ in range(3):
    ops.gemm_tc(ops.KIND_A8W8, a, b, sc=sc, sr=st, force_nt=nt)
torch.cuda.synchronize()
