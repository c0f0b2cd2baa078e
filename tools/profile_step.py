"""One eager Dreamer-V3 update between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sheeprl_b200.configs import make_dv3_cfg
from sheeprl_b200.engine import DV3Engine
from sheeprl_b200.algos.dreamer_v3.agent import initial_state
from bench import synthetic_batch

cfg = make_dv3_cfg("S"); adim = (2,)
eng = DV3Engine(cfg, adim, device="cuda")
g = torch.Generator().manual_seed(0)
for grp in (eng.wm, eng.actor, eng.critic):
    grp.load(initial_state(grp, {}, g))
eng.target.load(eng.critic.state_dict())
data = synthetic_batch(cfg, adim, 1, device="cuda")
data = {k: (v if v.dtype == torch.uint8 else v.float()) for k, v in data.items()}
for _ in range(3):
    eng.train_step(data, None)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.train_step(data, None)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", eng.metrics[:3].tolist())
