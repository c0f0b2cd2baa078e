"""Dreamer-V3 `train()` on the B200 engine — drop-in for `sheeprl/algos/dreamer_v3/dreamer_v3.py:48-357`.

Same positional signature, same side effects (parameters / optimiser state / Moments buffers updated in
place, 13 `aggregator.update` calls, `data["is_first"][0]` forced to 1), but the body is one call into
`DV3Engine.train_step`: ~2.5k hand-written CUDA kernel launches (optionally one CUDA-graph replay), no
autograd, no torch arithmetic.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence

import torch

from sheeprl_b200.engine import DV3Engine
from sheeprl_b200.utils.registry import register_algorithm

METRIC_ORDER = (
    "Loss/world_model_loss", "Loss/observation_loss", "Loss/reward_loss", "Loss/state_loss", "Loss/continue_loss",
    "State/kl", "State/post_entropy", "State/prior_entropy", "Loss/policy_loss", "Loss/value_loss",
    "Grads/world_model", "Grads/actor", "Grads/critic",
)


class B200Adam:
    """Handle standing where the reference passes a `torch.optim.Adam` (dreamer_v3.py:448-457).  The update
    itself is the fused clip+Adam kernel inside the engine; this object only exposes the state in torch's
    `Optimizer.state_dict()` layout so that reference checkpoints round-trip."""

    def __init__(self, group, names: Sequence[str], lr: float, eps: float, betas=(0.9, 0.999), weight_decay=0.0):
        if weight_decay:
            raise NotImplementedError("weight_decay != 0 is not supported by the fused Adam kernel")
        self.group, self.names = group, list(names)
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False)

    def zero_grad(self, set_to_none: bool = True):  # gradients live in the engine's flat buffer
        return None

    def step(self):
        raise RuntimeError("B200Adam.step() is fused into DV3Engine.train_step()")

    def state_dict(self) -> Dict[str, Any]:
        m, v = self.group.optimizer_views()
        state = {}
        if self.group.step > 0:
            for i, n in enumerate(self.names):
                state[i] = {"step": torch.tensor(float(self.group.step)), "exp_avg": m[n].detach().clone(),
                            "exp_avg_sq": v[n].detach().clone()}
        return {"state": state, "param_groups": [dict(self.defaults, params=list(range(len(self.names))))]}

    def load_state_dict(self, sd: Dict[str, Any]):
        m, v = self.group.optimizer_views()
        steps = set()
        with torch.no_grad():
            for i, n in enumerate(self.names):
                st = sd["state"].get(i)
                if st is None:
                    continue
                m[n].copy_(st["exp_avg"])
                v[n].copy_(st["exp_avg_sq"])
                steps.add(int(st["step"]))
        if len(steps) > 1:
            raise ValueError("per-parameter Adam steps differ; the fused kernel keeps one step per group")
        self.group.step = steps.pop() if steps else 0
        self.group.step_t.fill_(self.group.step)


def make_optimizers(engine: DV3Engine, cfg):
    a = cfg.algo
    mk = lambda g, o: B200Adam(g, list(g.shapes), float(o.lr), float(o.eps), tuple(o.betas), float(o.weight_decay))  # noqa
    return (mk(engine.wm, a.world_model.optimizer), mk(engine.actor, a.actor.optimizer),
            mk(engine.critic, a.critic.optimizer))


def _engine_of(module) -> DV3Engine:
    eng = getattr(module, "_b200_engine", None)
    if eng is None:
        raise TypeError("train() needs the modules returned by sheeprl_b200.algos.dreamer_v3.agent.build_agent")
    return eng


def train(
    fabric,
    world_model,
    actor,
    critic,
    target_critic,
    world_optimizer,
    actor_optimizer,
    critic_optimizer,
    data: Dict[str, torch.Tensor],
    aggregator,
    cfg: Dict[str, Any],
    is_continuous: bool,
    actions_dim: Sequence[int],
    moments,
    noise: Optional[Dict[str, torch.Tensor]] = None,
) -> None:
    """One Dreamer-V3 update.  `data`: dict of `[T, B, ...]` tensors on `fabric.device` (float32 as the
    reference passes them; the image key may also be uint8).  `noise` (extra, optional): injected Exp(1)
    sampling noise for parity tests; None -> on-device Philox."""
    eng = _engine_of(world_model)
    if bool(is_continuous) != eng.is_continuous:
        raise ValueError("is_continuous differs from the value build_agent() was called with")
    if moments is not None and getattr(moments, "low", None) is not None and moments.low.data_ptr() != eng.moments_state.data_ptr():
        moments.bind(eng.moments_state)
    eng.train_step(data, noise)
    if aggregator and not aggregator.disabled:
        md = eng.metrics_dict()
        for k in METRIC_ORDER:
            aggregator.update(k, md[k])


@register_algorithm()
def main(fabric, cfg: Dict[str, Any]):
    raise NotImplementedError(
        "the environment-interaction loop (sheeprl/algos/dreamer_v3/dreamer_v3.py:361-780) is outside this "
        "round's hot path (SURVEY.md §8); call build_agent()/train() from the reference's main().")
