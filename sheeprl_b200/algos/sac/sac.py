"""`train()` of SAC on the B200 engine — the reference's signature and side effects
(sheeprl/algos/sac/sac.py:32-78): parameters / optimiser states / target networks updated in place, three
`aggregator.update` calls.  The body is one call into `SACEngine.train_step` (csrc/mlp.cu + csrc/sac.cu)."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from sheeprl_b200.algos.dreamer_v3.dreamer_v3 import B200Adam
from sheeprl_b200.utils.registry import register_algorithm

METRIC_ORDER = ("Loss/value_loss", "Loss/policy_loss", "Loss/alpha_loss")


def make_optimizers(agent, cfg):
    """(actor_optimizer, qf_optimizer, alpha_optimizer) handles in the order `sac.main` builds them (sac.py:149-157)"""
    e = agent._b200_engine
    mk = lambda g, o: B200Adam(g, list(g.shapes), o["lr"], o["eps"], o["betas"])  # noqa: E731
    return mk(e.actor, e.opt["actor"]), mk(e.qf, e.opt["qf"]), mk(e.alpha, e.opt["alpha"])


def train(fabric, agent, actor_optimizer, qf_optimizer, alpha_optimizer, data: Dict[str, torch.Tensor], aggregator,
          update: int, cfg: Dict[str, Any], policy_steps_per_iter: int, group=None,
          noise: Optional[Dict[str, torch.Tensor]] = None) -> None:
    """One SAC update on `data` = {observations, next_observations, actions, rewards, terminated}, each `[B, ...]`
    float32 on `fabric.device`.  `noise` (extra, optional): injected N(0,1) draws for parity tests."""
    eng = getattr(agent, "_b200_engine", None)
    if eng is None:
        raise TypeError("train() needs the agent returned by sheeprl_b200.algos.sac.agent.build_agent")
    B = data["observations"].shape[0]
    if B != eng.B:
        eng.B = B
        eng._alloc()
    do_ema = update % (cfg.algo.critic.target_network_frequency // policy_steps_per_iter + 1) == 0   # sac.py:56
    eng.train_step(data, do_ema, noise)
    if aggregator and not aggregator.disabled:
        md = eng.metrics_dict()
        for k in METRIC_ORDER:
            aggregator.update(k, md[k])


def _optimizer_factory(agents):
    from sheeprl_b200.utils.delegate import group_of

    def make(config, params):
        if not agents:
            return None
        e = agents[-1]._b200_engine
        groups = {"actor": e.actor, "qf": e.qf, "alpha": e.alpha}
        name = group_of(params, groups)
        if name is None:
            return None
        target = str(config.get("_target_", "torch.optim.Adam"))
        if not target.endswith("Adam"):
            raise NotImplementedError(f"optimizer {target}: the fused update kernel implements torch.optim.Adam")
        g = groups[name]
        return B200Adam(g, list(g.shapes), float(config["lr"]), float(config.get("eps", 1e-8)),
                        tuple(config.get("betas", (0.9, 0.999))), float(config.get("weight_decay", 0.0) or 0.0))

    return make


def reference_substitutions(cfg, agents):
    """names of `sheeprl/algos/sac/sac.py` replaced while the reference's `main` runs: build_agent (:146-150), train
    (:343-356)"""
    from sheeprl_b200.algos.sac import agent as A

    def build_agent(*a, **k):
        out = A.build_agent(*a, **k)
        agents.append(out[0])
        return out

    return {"build_agent": build_agent, "train": train}


@register_algorithm()
def main(fabric, cfg: Dict[str, Any]):
    """Entry point registered for `algo.name=sac` (sheeprl/cli.py:82-98, 199): the reference's own interaction loop
    (sac.py:81-330) with this package's `build_agent` / `train` / Adam handles."""
    from sheeprl_b200.utils.delegate import run_reference_main

    agents = []
    return run_reference_main("sheeprl.algos.sac.sac", fabric, cfg, reference_substitutions(cfg, agents),
                              _optimizer_factory(agents))
