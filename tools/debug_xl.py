import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import dv3_oracle as O
from oracle.make_golden import perturbed_oracle_init
from sheeprl_b200.configs import make_dv3_cfg
from tests.helpers import oracle_run
from tests.test_gpu_engine import make_engine, to_cuda
size = sys.argv[1] if len(sys.argv) > 1 else "XL"
kw = {}
for a in sys.argv[2:]:
    k, v = a.split("="); kw[k] = int(v)
cfg = make_dv3_cfg(size, per_rank_batch_size=2, per_rank_sequence_length=4, horizon=3, **kw)
adim = (3,); a, w = cfg.algo, cfg.algo.world_model
init = perturbed_oracle_init(cfg, adim, 21, 0.02)
data = [O.make_batch(cfg, adim, seed=22)]
noise = [O.draw_noise(4, 2, 3, w.stochastic_size, w.discrete_size, adim, seed=23)]
st, o, ms, _ = oracle_run(cfg, adim, init, data, noise, 1, condition_margin=1e-3, keep=True)
eng = make_engine(cfg, adim, init)
eng.train_step({k: v.clone().cuda() for k, v in data[0].items()}, to_cuda(noise[0]))
torch.cuda.synchronize()
rows = []
for grp, nm in (("wm", "world_model"), ("actor", "actor"), ("critic", "critic")):
    og = o[0][f"grads/{grp}"]
    mx = {"wm": cfg.algo.world_model.clip_gradients, "actor": cfg.algo.actor.clip_gradients, "critic": cfg.algo.critic.clip_gradients}[grp]
    coef = min(1.0, mx / (float(o[0]["Grads/" + nm]) + 1e-6))
    for k, v in og.items():
        g = getattr(eng, grp).gviews[k].cpu() * coef
        rel = float((g - v).norm() / (v.norm() + 1e-30))
        rows.append((rel, grp, k, float(v.norm()), float(g.norm())))
rows.sort(reverse=True)
for r in rows[:14]:
    print("%.2e %s %s ref=%.4g got=%.4g" % r)
print({k: (float(v), float(o[0][k])) for k, v in eng.metrics_dict().items() if k.startswith("Grads")})
for grp in ("wm", "actor", "critic"):
    G = getattr(eng, grp)
    mine = float(G.grad.double().norm())
    views = float(torch.sqrt(sum((v.double() ** 2).sum() for v in G.gviews.values())))
    orc = float(torch.sqrt(sum((v.double() ** 2).sum() for v in o[0][f"grads/{grp}"].values())))
    print(grp, "flat-buffer norm", mine, "views norm", views, "oracle post-clip norm", orc)
