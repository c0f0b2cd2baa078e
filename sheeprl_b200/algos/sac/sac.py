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


@register_algorithm()
def main(fabric, cfg: Dict[str, Any]):
    raise NotImplementedError(
        "the environment-interaction loop (sheeprl/algos/sac/sac.py:81-330) is outside this round's hot path "
        "(SURVEY.md §8); call build_agent()/train() from the reference's main().")
