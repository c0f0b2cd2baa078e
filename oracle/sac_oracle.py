"""TEST INFRASTRUCTURE (oracle) — functional fp32 restatement of one SAC update on the CPU.

Follows: `train` sheeprl/algos/sac/sac.py:32-78; `SACAgent.get_next_target_q_values / get_q_values /
qfs_target_ema` sac/agent.py:245-267; `SACActor.forward / _get_actions_and_log_probs` sac/agent.py:94-142;
`SACCritic` sac/agent.py:19-52 (MLP: Linear-ReLU-Linear-ReLU-Linear, models/models.py:16-119); losses
sac/loss.py:9-29; torch.optim.Adam (configs/algo/sac.yaml: lr 3e-4, eps 1e-4, no clipping).

Parity PINNED: tests/golden/sac_*.pt are produced by the executed reference `train()` with the Normal.rsample
noise injected (oracle/make_golden_sac.py); tests/test_sac_cpu.py checks this file against them.
Only tests/, __graft_entry__.smoke() and the CPU-baseline legs of the bench scripts may import this module.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
from torch import Tensor

from oracle.dv3_oracle import AdamState

LOG_STD_MAX, LOG_STD_MIN = 2.0, -5.0          # sac/agent.py:15-16


def init_params(obs_dim: int, act_dim: int, hidden: int, n_critics: int, seed: int, alpha: float = 1.0):
    """torch.nn.Linear default init (kaiming-uniform weights, uniform bias), reference key names."""
    g = torch.Generator().manual_seed(seed)

    def linear(prefix, fan_in, fan_out, out):
        bound = 1.0 / math.sqrt(fan_in)
        out[f"{prefix}.weight"] = (torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound
        out[f"{prefix}.bias"] = (torch.rand(fan_out, generator=g) * 2 - 1) * bound

    actor: Dict[str, Tensor] = {}
    linear("model._model.0", obs_dim, hidden, actor)
    linear("model._model.2", hidden, hidden, actor)
    linear("fc_mean", hidden, act_dim, actor)
    linear("fc_logstd", hidden, act_dim, actor)
    qf: Dict[str, Tensor] = {}
    for i in range(n_critics):
        linear(f"{i}.model._model.0", obs_dim + act_dim, hidden, qf)
        linear(f"{i}.model._model.2", hidden, hidden, qf)
        linear(f"{i}.model._model.4", hidden, 1, qf)
    target = {k: v.clone() for k, v in qf.items()}
    return {"actor": actor, "qf": qf, "qf_target": target, "log_alpha": {"log_alpha": torch.log(torch.tensor([alpha]))}}


def _lin(p, prefix, x):
    return x @ p[f"{prefix}.weight"].t() + p[f"{prefix}.bias"]


def critic_forward(p: Dict[str, Tensor], i: int, obs: Tensor, act: Tensor) -> Tensor:
    x = torch.cat([obs, act], -1)
    x = torch.relu(_lin(p, f"{i}.model._model.0", x))
    x = torch.relu(_lin(p, f"{i}.model._model.2", x))
    return _lin(p, f"{i}.model._model.4", x)


def actor_forward(p: Dict[str, Tensor], obs: Tensor, eps: Tensor, scale: Tensor, bias: Tensor):
    """tanh-squashed reparameterised Normal sample and its log-prob (sac/agent.py:94-142, Eq. 26 of 1812.05905)"""
    x = torch.relu(_lin(p, "model._model.0", obs))
    x = torch.relu(_lin(p, "model._model.2", x))
    mean, log_std = _lin(p, "fc_mean", x), _lin(p, "fc_logstd", x)
    std = log_std.clamp(LOG_STD_MIN, LOG_STD_MAX).exp()
    x_t = mean + std * eps
    y_t = torch.tanh(x_t)
    action = y_t * scale + bias
    logp = -((x_t - mean) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))
    logp = logp - torch.log(scale * (1 - y_t.pow(2)) + 1e-6)
    return action, logp.sum(-1, keepdim=True)


def sac_train_step(P, opt_qf: AdamState, opt_actor: AdamState, opt_alpha: AdamState, data: Dict[str, Tensor],
                   eps_next: Tensor, eps_cur: Tensor, gamma: float, tau: float, do_ema: bool, n_critics: int,
                   scale: Tensor, bias: Tensor, target_entropy: float):
    """One `train()` call.  P is mutated in place; returns the three logged losses (sac.py:74-78)."""
    actor, qf, tgt, la = P["actor"], P["qf"], P["qf_target"], P["log_alpha"]
    alpha = float(la["log_alpha"].exp())
    # ---- critics (sac.py:45-53)
    with torch.no_grad():
        a2, logp2 = actor_forward(actor, data["next_observations"], eps_next, scale, bias)
        q2 = torch.cat([critic_forward(tgt, i, data["next_observations"], a2) for i in range(n_critics)], -1)
        y = data["rewards"] + (1 - data["terminated"]) * gamma * (q2.min(-1, keepdim=True)[0] - alpha * logp2)
    qp = {k: v.detach().requires_grad_(True) for k, v in qf.items()}
    q = torch.cat([critic_forward(qp, i, data["observations"], data["actions"]) for i in range(n_critics)], -1)
    qf_loss = sum(((q[..., i:i + 1] - y) ** 2).mean() for i in range(n_critics))
    grads = torch.autograd.grad(qf_loss, list(qp.values()))
    opt_qf.step(qf, dict(zip(qp.keys(), grads)))
    # ---- target EMA (sac.py:55-57, agent.py:264-267)
    if do_ema:
        with torch.no_grad():
            for k in tgt:
                tgt[k].copy_(tau * qf[k] + (1 - tau) * tgt[k])
    # ---- actor (sac.py:59-66)
    ap = {k: v.detach().requires_grad_(True) for k, v in actor.items()}
    act, logp = actor_forward(ap, data["observations"], eps_cur, scale, bias)
    qn = torch.cat([critic_forward(qf, i, data["observations"], act) for i in range(n_critics)], -1)
    actor_loss = (alpha * logp - qn.min(-1, keepdim=True)[0]).mean()
    grads = torch.autograd.grad(actor_loss, list(ap.values()))
    opt_actor.step(actor, dict(zip(ap.keys(), grads)))
    # ---- temperature (sac.py:68-73)
    lp = la["log_alpha"].detach().requires_grad_(True)
    alpha_loss = (-lp * (logp.detach() + target_entropy)).mean()
    (g,) = torch.autograd.grad(alpha_loss, [lp])
    opt_alpha.step(la, {"log_alpha": g})
    return {"Loss/value_loss": float(qf_loss), "Loss/policy_loss": float(actor_loss), "Loss/alpha_loss": float(alpha_loss)}


def make_batch(B: int, obs_dim: int, act_dim: int, seed: int) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    return {
        "observations": torch.randn(B, obs_dim, generator=g),
        "next_observations": torch.randn(B, obs_dim, generator=g),
        "actions": torch.rand(B, act_dim, generator=g) * 2 - 1,
        "rewards": torch.randn(B, 1, generator=g),
        "terminated": (torch.rand(B, 1, generator=g) < 0.05).float(),
    }
