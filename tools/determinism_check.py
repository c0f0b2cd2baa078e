"""Run-to-run reproducibility of the Dreamer-V3 update at the BASELINE config: two engines, same seeds, two steps each
(on-device Philox noise); prints the largest relative metric difference, the fraction of differing imagined samples and
the largest parameter difference.  Bit-equal replays print zeros."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bench import synthetic_batch
    from oracle import dv3_oracle as O          # initial parameters only (tool, not product)
    from sheeprl_b200.configs import make_dv3_cfg
    from sheeprl_b200.engine import DV3Engine

    cfg = make_dv3_cfg("S")
    adim = (2,)
    init = O.init_params(cfg, adim, seed=0)
    data = synthetic_batch(cfg, adim, seed=3, device="cuda:0")
    runs = []
    for rep in range(2):
        eng = DV3Engine(cfg, adim, in_channels=3, device="cuda:0")
        for g, p in zip((eng.wm, eng.actor, eng.critic, eng.target), init):
            g.load(p)
        eng.rng_seed = 99
        for _ in range(2):
            eng.train_step({k: v.clone() for k, v in data.items()}, None)
        torch.cuda.synchronize()
        runs.append({"metrics": eng.metrics.clone(), "traj": eng.traj[:, :, : eng.Z].clone(), "wm": eng.wm.flat.clone(),
                     "actor": eng.actor.flat.clone(), "critic": eng.critic.flat.clone(), "wm_grad": eng.wm.grad.clone()})
    a, b = runs
    out = {
        "metrics_max_rel_diff": float(((a["metrics"] - b["metrics"]).abs() / a["metrics"].abs().clamp_min(1e-12)).max()),
        "imagined_samples_differing_frac": float((a["traj"] != b["traj"]).float().mean()),
        "wm_param_max_abs_diff": float((a["wm"] - b["wm"]).abs().max()),
        "actor_param_max_abs_diff": float((a["actor"] - b["actor"]).abs().max()),
        "critic_param_max_abs_diff": float((a["critic"] - b["critic"]).abs().max()),
        "wm_grad_max_rel_diff": float((a["wm_grad"] - b["wm_grad"]).abs().max() / a["wm_grad"].abs().max()),
        "bit_equal": bool(all(torch.equal(a[k], b[k]) for k in a)),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
