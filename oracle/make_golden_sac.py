"""TEST INFRASTRUCTURE — writes tests/golden/sac_*.pt by EXECUTING THE REAL REFERENCE `sac.train` (container only):

    python -m oracle.make_golden_sac

Each fixture: dims, initial parameters (reference `build_agent` under torch.manual_seed), the batches, the injected
N(0,1) noise of the two `Normal.rsample` calls per update, and what the unmodified reference produced after every
update: all parameters (actor, critics, targets, log_alpha) and the three logged losses.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402
from oracle import sac_oracle as SO  # noqa: E402
from sheeprl_b200.utils.utils import dotdict  # noqa: E402

FIXTURES = {
    "sac_tiny": dict(obs_dim=5, act_dim=3, hidden=16, n_critics=2, B=8, low=-2.0, high=1.0, updates=3, seed=3),
    "sac_c4": dict(obs_dim=17, act_dim=6, hidden=256, n_critics=2, B=256, low=-1.0, high=1.0, updates=3, seed=4),
}


def sac_cfg(hidden: int, n_critics: int):
    opt = {"lr": 3e-4, "eps": 1e-4, "weight_decay": 0, "betas": [0.9, 0.999]}
    return dotdict({"algo": {"mlp_keys": {"encoder": ["state"]}, "cnn_keys": {"encoder": []}, "gamma": 0.99, "tau": 0.005,
                             "actor": {"hidden_size": hidden, "optimizer": dict(opt)},
                             "critic": {"hidden_size": hidden, "n": n_critics, "target_network_frequency": 1,
                                        "optimizer": dict(opt)},
                             "alpha": {"alpha": 1.0, "optimizer": dict(opt)}},
                    "distribution": {}})


def export(agent):
    out = {"actor": {k: v.detach().clone() for k, v in agent.actor.module.state_dict().items()
                     if k not in ("action_scale", "action_bias")}, "qf": {}, "qf_target": {},
           "log_alpha": {"log_alpha": agent.log_alpha.detach().clone()}}
    for i, (q, t) in enumerate(zip(agent.qfs, agent.qfs_target)):
        for k, v in q.module.state_dict().items():
            out["qf"][f"{i}.{k}"] = v.detach().clone()
        for k, v in t.module.state_dict().items():
            out["qf_target"][f"{i}.{k}"] = v.detach().clone()
    return out


def run(spec):
    import sheeprl.algos.sac.agent as A
    import sheeprl.algos.sac.sac as S
    import torch.distributions.normal as TN

    A.get_single_device_fabric = lambda f: f
    cfg = sac_cfg(spec["hidden"], spec["n_critics"])

    class Box:
        shape = (spec["act_dim"],)
        low = np.full(spec["act_dim"], spec["low"], np.float32)
        high = np.full(spec["act_dim"], spec["high"], np.float32)

    fab = H.FakeFabric()
    fab.all_reduce = lambda x, group=None: x
    torch.manual_seed(spec["seed"])
    agent, _ = A.build_agent(fab, cfg, {"state": H.Shape((spec["obs_dim"],))}, Box)
    init = export(agent)
    qf_opt = torch.optim.Adam(agent.qfs.parameters(), lr=3e-4, eps=1e-4)
    actor_opt = torch.optim.Adam(agent.actor.parameters(), lr=3e-4, eps=1e-4)
    alpha_opt = torch.optim.Adam([agent.log_alpha], lr=3e-4, eps=1e-4)
    g = torch.Generator().manual_seed(spec["seed"] + 100)
    steps = []
    orig = TN._standard_normal
    for u in range(1, spec["updates"] + 1):
        data = SO.make_batch(spec["B"], spec["obs_dim"], spec["act_dim"], seed=spec["seed"] * 10 + u)
        eps = [torch.randn(spec["B"], spec["act_dim"], generator=g) for _ in range(2)]
        queue = list(eps)
        TN._standard_normal = lambda shape, dtype, device: queue.pop(0).reshape(shape)
        agg = H.RecordingAggregator()
        try:
            S.train(fab, agent, actor_opt, qf_opt, alpha_opt, {k: v.clone() for k, v in data.items()}, agg, u, cfg, 1)
        finally:
            TN._standard_normal = orig
        assert not queue
        keep = spec["B"] <= 16 or u == spec["updates"]          # big fixture: parameters after the last update only
        steps.append({"data": data, "eps_next": eps[0], "eps_cur": eps[1], "update": u,
                      "after": export(agent) if keep else None, "losses": dict(agg.values)})
    return {"spec": spec, "init": init, "steps": steps}


def main():
    H.install()
    for name, spec in FIXTURES.items():
        fx = run(spec)
        path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
        torch.save(fx, path)
        print(name, os.path.getsize(path), fx["steps"][-1]["losses"])


if __name__ == "__main__":
    main()
