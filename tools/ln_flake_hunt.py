"""Repeats the LayerNorm(+tanh) forward/backward parity check many times and reports the worst error of every output
(hunting a once-seen failure of tests/test_gpu_ops.py::test_ln_act_tanh_relu[2-1000-32])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.ops_emul import EmulOps
from sheeprl_b200.lib import CudaOps


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


cu, em = CudaOps("cuda"), EmulOps()
M, C, act = 1000, 32, 2
X, gam, bet, dY = rnd(M, C, seed=1, scale=2.0), rnd(C, seed=101) + 1.0, rnd(C, seed=201), rnd(M, C, seed=301)
Yc = torch.empty(M, C); em.ln_act_fwd(X, gam, bet, 1e-5, act, Yc)
dXc, dgc, dbc = torch.empty(M, C), torch.empty(C), torch.empty(C)
em.ln_act_bwd(X, gam, bet, 1e-5, act, dY, dXc, dgc, dbc)
worst = {"Y": 0.0, "dX": 0.0, "dg": 0.0, "db": 0.0}
junk = []
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 400):
    if it % 7 == 0:   # churn the caching allocator / leave other kernels in flight
        junk = [torch.randn(1 << (10 + (it % 9)), device="cuda") for _ in range(3)]
        A, B = torch.randn(2048, 512, device="cuda"), torch.randn(512, 512, device="cuda")
        Cm = torch.empty(2048, 512, device="cuda"); cu.gemm(A, B, Cm, False, True)
    Xg, gg, bg, dYg = X.cuda(), gam.cuda(), bet.cuda(), dY.cuda()
    Yg = torch.empty(M, C, device="cuda"); cu.ln_act_fwd(Xg, gg, bg, 1e-5, act, Yg)
    dXg, dgg, dbg = torch.empty(M, C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    cu.ln_act_bwd(Xg, gg, bg, 1e-5, act, dYg, dXg, dgg, dbg)
    for k, (g_, c_) in {"Y": (Yg, Yc), "dX": (dXg, dXc), "dg": (dgg, dgc), "db": (dbg, dbc)}.items():
        e = float((g_.cpu() - c_).abs().max() / c_.abs().max())
        if e > worst[k]:
            worst[k] = e
            print(f"iter {it}: new worst {k} rel-to-max error {e:.3e}", flush=True)
print("worst", worst)
