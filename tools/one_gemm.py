"""Launches the step's dominant GEMM shape a few times (for an `ncu --set full` capture of gemm_tc_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sheeprl_b200.lib import CudaOps
cu = CudaOps("cuda")
M, N, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16384, 512, 1536)))
A, B, C = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.empty(M, N, device="cuda")
for _ in range(6):
    cu.gemm(A, B, C, False, True)
torch.cuda.synchronize()
print("ok")
