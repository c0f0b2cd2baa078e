"""TEST INFRASTRUCTURE (oracle) — CPU restatements of the PPO pieces on the hot path.

`gae_oracle` restates `sheeprl/utils/utils.py:63-100` (`gae`).  Pinned against the imported reference in
tests/test_oracle_pin.py (container only) and against tests/golden/gae_small.pt everywhere.
"""
from __future__ import annotations

import torch


def gae_oracle(rewards, values, dones, next_value, num_steps, gamma, gae_lambda):
    """Generalised advantage estimation, reverse scan over `num_steps` (reference: utils/utils.py:63-100).
    `dones[t]` masks the bootstrap from step t+1 into step t; the last step bootstraps `next_value`."""
    alive = (dones == 0).to(rewards.dtype)
    adv = torch.zeros_like(rewards)
    running = torch.zeros_like(rewards[0])
    for t in range(num_steps - 1, -1, -1):
        if t == num_steps - 1:
            nv, mask = next_value, alive[-1]
        else:
            nv, mask = values[t + 1], alive[t]
        delta = rewards[t] + nv * mask * gamma - values[t]
        running = delta + mask * running * gamma * gae_lambda
        adv[t] = running
    return adv + values, adv


# ---------------------------------------------------------------------------------------------------------
# PPO update (reference: sheeprl/algos/ppo/ppo.py:30-102, ppo/agent.py:84-239, ppo/loss.py:6-75,
# models/models.py:288-328 NatureCNN, utils/utils.py:121-130 normalize_tensor)
# Parity PINNED: tests/golden/ppo_*.pt come from the executed reference train() (oracle/make_golden_ppo.py).
# ---------------------------------------------------------------------------------------------------------
import math  # noqa: E402
from typing import Dict, List, Sequence  # noqa: E402

import torch.nn.functional as F  # noqa: E402

from oracle.dv3_oracle import AdamState  # noqa: E402,F401

CONVS = ((8, 4), (4, 2), (3, 1))          # NatureCNN (kernel, stride), channels 32/64/64 (models.py:301-309)


def net_cfg(spec, which: str):
    """(dense_units, mlp_layers, layer_norm) of `which` in {"encoder", "actor", "critic"}: the per-net overrides
    spec["nets"][which] = (dense, layers) and spec["layer_norm"] (bool or per-net dict) over spec["dense"/"layers"]."""
    dense, layers = (spec.get("nets") or {}).get(which, (spec["dense"], spec["layers"]))
    ln = spec.get("layer_norm", False)
    ln = bool(ln.get(which, False)) if isinstance(ln, dict) else bool(ln)
    return int(dense), int(layers), ln


def _stack_shapes(out, prefix, d, dense, layers, ln, last):
    """keys of an `MLP` (models/models.py:17-126 over utils/model.py:34-88 miniblocks): per hidden layer
    Linear [, LayerNorm], activation -> Sequential index stride 3 with norm, 2 without; then the output Linear."""
    st = 3 if ln else 2
    for i in range(layers):
        out[f"{prefix}._model.{st * i}.weight"] = (dense, d)
        out[f"{prefix}._model.{st * i}.bias"] = (dense,)
        if ln:
            out[f"{prefix}._model.{st * i + 1}.weight"] = (dense,)
            out[f"{prefix}._model.{st * i + 1}.bias"] = (dense,)
        d = dense
    if last is not None:
        out[f"{prefix}._model.{st * layers}.weight"] = (last, d)
        out[f"{prefix}._model.{st * layers}.bias"] = (last,)
        d = last
    return d


def ppo_param_shapes(spec) -> "Dict[str, tuple]":
    """Reference state-dict keys/shapes of PPOAgent for `spec` = dict(cnn_channels (0 = no image; the sum over the
    image keys, which the encoder concatenates on the channel axis, ppo/agent.py:34-36), screen, mlp_dim (0 = no vector
    obs; sum over the vector keys), dense, layers, cnn_features, mlp_features, actions_dim, is_continuous
    [, layer_norm, nets, dist])."""
    out = {}
    feat = 0
    if spec["cnn_channels"]:
        c, s = spec["cnn_channels"], spec["screen"]
        for i, ((k, st), co) in enumerate(zip(CONVS, (32, 64, 64))):
            out[f"feature_extractor.cnn_encoder.model._model.{2 * i}.weight"] = (co, c, k, k)
            out[f"feature_extractor.cnn_encoder.model._model.{2 * i}.bias"] = (co,)
            c, s = co, (s - k) // st + 1
        out["feature_extractor.cnn_encoder.model.fc.weight"] = (spec["cnn_features"], c * s * s)
        out["feature_extractor.cnn_encoder.model.fc.bias"] = (spec["cnn_features"],)
        feat += spec["cnn_features"]
    if spec["mlp_dim"]:
        dense, layers, ln = net_cfg(spec, "encoder")
        feat += _stack_shapes(out, "feature_extractor.mlp_encoder.model", spec["mlp_dim"], dense, layers, ln, spec["mlp_features"])
    dense, layers, ln = net_cfg(spec, "critic")
    _stack_shapes(out, "critic", feat, dense, layers, ln, 1)
    dense, layers, ln = net_cfg(spec, "actor")
    _stack_shapes(out, "actor.actor_backbone", feat, dense, layers, ln, None)
    heads = [2 * sum(spec["actions_dim"])] if spec["is_continuous"] else list(spec["actions_dim"])
    for i, a in enumerate(heads):
        out[f"actor.actor_heads.{i}.weight"] = (a, dense)       # in_features = actor dense_units (ppo/agent.py:180-183)
        out[f"actor.actor_heads.{i}.bias"] = (a,)
    return out


def _mlp(p, prefix, x, cfg, act, last=True):
    dense, layers, ln = cfg
    st = 3 if ln else 2
    for i in range(layers):
        x = F.linear(x, p[f"{prefix}._model.{st * i}.weight"], p[f"{prefix}._model.{st * i}.bias"])
        if ln:
            x = F.layer_norm(x, (dense,), p[f"{prefix}._model.{st * i + 1}.weight"], p[f"{prefix}._model.{st * i + 1}.bias"], 1e-5)
        x = act(x)
    if last:
        x = F.linear(x, p[f"{prefix}._model.{st * layers}.weight"], p[f"{prefix}._model.{st * layers}.bias"])
    return x


SAFE_LIM = 1.0 - 1e-6       # safetanh / safeatanh clamp, eps = finfo(float32).resolution (utils/utils.py:304-313)


def tanh_correction(tanh_actions):
    """the term the reference SUBTRACTS from Normal.log_prob for `tanh_normal` (ppo/agent.py:201-205, 264-268)"""
    return 2.0 * (math.log(2.0) - tanh_actions - F.softplus(-2.0 * tanh_actions)).sum(-1, keepdim=True)


def ppo_forward(p, spec, obs: Dict[str, torch.Tensor], actions: torch.Tensor):
    """PPOAgent.forward with given actions (ppo/agent.py:208-239): (logprobs [B,1], entropy [B,1], values [B,1]).
    obs["rgb"] already normalised (x/255 - 0.5, ppo/utils.py:69-72)."""
    act = torch.tanh if spec.get("act", "tanh") == "tanh" else torch.relu
    feats = []
    if spec["cnn_channels"]:
        x = obs["rgb"]
        for i, (k, st) in enumerate(CONVS):
            pre = f"feature_extractor.cnn_encoder.model._model.{2 * i}"
            x = torch.relu(F.conv2d(x, p[f"{pre}.weight"], p[f"{pre}.bias"], stride=st))
        x = x.flatten(1)
        feats.append(torch.relu(F.linear(x, p["feature_extractor.cnn_encoder.model.fc.weight"],
                                         p["feature_extractor.cnn_encoder.model.fc.bias"])))
    if spec["mlp_dim"]:
        feats.append(_mlp(p, "feature_extractor.mlp_encoder.model", obs["state"], net_cfg(spec, "encoder"), act))
    feat = torch.cat(feats, -1)
    values = _mlp(p, "critic", feat, net_cfg(spec, "critic"), act)
    h = _mlp(p, "actor.actor_backbone", feat, net_cfg(spec, "actor"), act, last=False)
    if spec["is_continuous"]:
        out = F.linear(h, p["actor.actor_heads.0.weight"], p["actor.actor_heads.0.bias"])
        mean, log_std = out.chunk(2, -1)
        std = log_std.exp()
        corr = 0.0
        if spec.get("dist", "normal") == "tanh_normal":                 # stored actions are tanh-squashed (agent.py:194-206)
            corr = tanh_correction(actions)
            actions = torch.atanh(actions.clamp(-SAFE_LIM, SAFE_LIM))
        lp = (-((actions - mean) ** 2) / (2 * std ** 2) - log_std - math.log(math.sqrt(2 * math.pi))).sum(-1, keepdim=True)
        ent = (0.5 + 0.5 * math.log(2 * math.pi) + log_std).sum(-1, keepdim=True)
        return lp - corr, ent, values
    lps, ents, off = [], [], 0
    for i, a in enumerate(spec["actions_dim"]):
        logits = F.linear(h, p[f"actor.actor_heads.{i}.weight"], p[f"actor.actor_heads.{i}.bias"])
        logp = logits - logits.logsumexp(-1, keepdim=True)            # OneHotCategorical(logits=...) normalisation
        probs = logp.exp()
        lps.append((logp * actions[:, off:off + a]).sum(-1))
        ents.append(-(torch.clamp(logp, min=torch.finfo(logp.dtype).min) * probs).sum(-1))
        off += a
    return torch.stack(lps, -1).sum(-1, keepdim=True), torch.stack(ents, -1).sum(-1, keepdim=True), values


def ppo_losses(lp, ent, values, batch, hp):
    """policy_loss / value_loss / entropy_loss (ppo/loss.py:6-75), reduction = mean."""
    adv = batch["advantages"]
    if hp["normalize_advantages"]:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    ratio = (lp - batch["logprobs"]).exp()
    pg = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - hp["clip_coef"], 1 + hp["clip_coef"])).mean()
    if hp["clip_vloss"]:
        vc = batch["values"] + torch.clamp(values - batch["values"], -hp["clip_coef"], hp["clip_coef"])
        v = 0.5 * torch.max((values - batch["returns"]) ** 2, (vc - batch["returns"]) ** 2).mean()
    else:
        v = ((values - batch["returns"]) ** 2).mean()
    e = (-ent).mean()
    return pg, v, e


def ppo_train(p: Dict[str, torch.Tensor], opt: "AdamState", spec, data: Dict[str, torch.Tensor],
              index_batches: Sequence[Sequence[int]], hp) -> List[Dict[str, float]]:
    """`train()` over the given minibatch index lists (the reference draws them with RandomSampler + BatchSampler,
    ppo.py:49-56).  data: flat [N, ...] float tensors (rgb raw 0..255 as float).  Mutates p / opt."""
    logs = []
    for idx in index_batches:
        idx = torch.as_tensor(idx)
        batch = {k: v[idx] for k, v in data.items()}
        obs = {}
        if spec["cnn_channels"]:
            obs["rgb"] = batch["rgb"] / 255 - 0.5
        if spec["mlp_dim"]:
            obs["state"] = batch["state"]
        q = {k: v.detach().requires_grad_(True) for k, v in p.items()}
        lp, ent, values = ppo_forward(q, spec, obs, batch["actions"])
        pg, v, e = ppo_losses(lp, ent, values, batch, hp)
        loss = pg + hp["vf_coef"] * v + hp["ent_coef"] * e
        grads = dict(zip(q.keys(), torch.autograd.grad(loss, list(q.values()))))
        if hp["max_grad_norm"] > 0:
            total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
            coef = torch.clamp(hp["max_grad_norm"] / (total + 1e-6), max=1.0)
            grads = {k: g * coef for k, g in grads.items()}
        opt.step(p, grads)
        logs.append({"Loss/policy_loss": float(pg), "Loss/value_loss": float(v), "Loss/entropy_loss": float(e)})
    return logs


def make_rollout(spec, N: int, seed: int) -> Dict[str, torch.Tensor]:
    """Synthetic flat rollout [N, ...] with the keys ppo.train reads (ppo.py:399-407)."""
    g = torch.Generator().manual_seed(seed)
    d = {}
    if spec["cnn_channels"]:
        d["rgb"] = torch.randint(0, 256, (N, spec["cnn_channels"], spec["screen"], spec["screen"]), generator=g).float()
    if spec["mlp_dim"]:
        d["state"] = torch.randn(N, spec["mlp_dim"], generator=g)
    if spec["is_continuous"]:
        d["actions"] = torch.randn(N, sum(spec["actions_dim"]), generator=g)
        if spec.get("dist", "normal") == "tanh_normal":
            d["actions"] = torch.tanh(d["actions"]).clamp(-SAFE_LIM, SAFE_LIM)
    else:
        d["actions"] = torch.cat([F.one_hot(torch.randint(0, a, (N,), generator=g), a).float()
                                  for a in spec["actions_dim"]], -1)
    d["logprobs"] = -torch.rand(N, 1, generator=g) * 2
    d["values"] = torch.randn(N, 1, generator=g)
    d["returns"] = d["values"] + torch.randn(N, 1, generator=g) * 0.5
    d["advantages"] = torch.randn(N, 1, generator=g)
    return d
