"""Dreamer-V3 S with CONTINUOUS actions (6-dim, BASELINE batch): finite-ness + step time (eager and CUDA graph)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import dv3_oracle as O
from sheeprl_b200.configs import make_dv3_cfg
from sheeprl_b200.engine import DV3Engine

cfg = make_dv3_cfg("S"); adim = (6,)
eng = DV3Engine(cfg, adim, device="cuda", is_continuous=True)
wm, actor, critic, target = O.init_params(cfg, adim, seed=0, is_continuous=True)
eng.wm.load(wm), eng.actor.load(actor), eng.critic.load(critic), eng.target.load(target)
data = {k: v.cuda() for k, v in O.make_batch(cfg, adim, seed=1, as_uint8=True, is_continuous=True).items()}
for _ in range(3):
    eng.train_step(data, None)
torch.cuda.synchronize()
assert torch.isfinite(eng.metrics).all() and torch.isfinite(eng.actor.flat).all(), eng.metrics
def timed(f, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
eager = timed(lambda: eng.train_step(data, None))
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s): eng.train_step(data, None)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): eng.train_step(data, None)
g.replay()
graph = timed(g.replay)
print(json.dumps({"metric": "Dreamer-V3 S continuous(6) train step", "ms_eager": eager, "ms_graph": graph,
                  "steps_per_s": 1e3 / graph, "metrics": [round(float(x), 4) for x in eng.metrics[:10]]}))
