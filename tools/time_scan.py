"""Times the RSSM scan forward (fused persistent kernel vs per-step kernels) on the BASELINE config."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sheeprl_b200.configs import make_dv3_cfg
from sheeprl_b200.engine import DV3Engine
from bench import synthetic_batch

cfg = make_dv3_cfg("S"); adim = (2,)
eng = DV3Engine(cfg, adim, device="cuda")
for g in (eng.wm, eng.actor, eng.critic):
    g.flat.normal_(0, 0.02)
data = synthetic_batch(cfg, adim, 1, device="cuda")
data = {k: (v if v.dtype == torch.uint8 else v.float()) for k, v in data.items()}
eng.train_step(data, None)
first = data["is_first"].reshape(-1)
for fused in (False, True):
    eng.fused_scan = fused
    for _ in range(2):
        eng._scan_forward(first)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng._scan_forward(first)
    e1.record(); torch.cuda.synchronize()
    print("fused" if fused else "per-step", "scan fwd ms:", e0.elapsed_time(e1) / 5, "err flag", eng.ops.rssm_scan_error(eng._scan_ws) if eng._scan_ws is not None else None)

# backward scan timing (needs a fused forward right before each fused backward)
for fused in (False, True):
    eng.fused_scan = fused
    tot = 0.0
    for it in range(4):
        eng._scan_forward(first)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng._scan_backward(first)
        e1.record(); torch.cuda.synchronize()
        if it > 0:
            tot += e0.elapsed_time(e1)
    print("fused" if fused else "per-step", "scan bwd (+deferred wgrad GEMMs) ms:", tot / 3, "err flag", eng.ops.rssm_scan_error(eng._scan_ws))
