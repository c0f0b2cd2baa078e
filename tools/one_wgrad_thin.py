"""One launch each of the thin-channel conv kernels at the Dreamer-V3 S shapes (for an ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sheeprl_b200.lib import CudaOps
cu = CudaOps("cuda")
NB, h, w, Cs, Cb = 1024, 32, 32, 32, 3
small = torch.randn(NB, h, w, Cs, device="cuda")
big = torch.randn(NB, 2 * h, 2 * w, Cb, device="cuda")
W = torch.randn(Cs, Cb, 4, 4, device="cuda")
dW = torch.empty_like(W)
out = torch.empty_like(big)
osm = torch.empty_like(small)
for _ in range(3):
    cu.conv_wgrad(small, big, dW)
    cu.conv_up(small, W, out, None)
    cu.conv_down(big, W, osm)
torch.cuda.synchronize()
print("ok")
