"""TEST INFRASTRUCTURE (oracle) — CPU restatement of one Plan2Explore / Dreamer-V3 exploration update
(`sheeprl/algos/p2e_dv3/p2e_dv3_exploration.py:41-520`, discrete actions), on top of the Dreamer-V3 oracle's pieces.

Phases, in the reference's order:
  1. dynamic learning              == dreamer_v3 (`world_model_phase`, :113-205) except that the reward / continue
                                              heads read the DETACHED latent (:157,160)
  2. ensemble learning             :212-240   N MLPs predict the next posterior from [z_t, h_t, a_t]
  3. behaviour learning exploration :242-392  rollout with the exploration actor; one critic per entry of
                                              `critics_exploration` (intrinsic reward = ensemble disagreement, or the
                                              task reward), advantages mixed by weight, each critic regressed
  4. behaviour learning task        :397-474  the plain Dreamer-V3 behaviour step with the task actor / critic

Parity PINNED: tests/golden/p2e_tiny.pt is written by oracle/make_golden_p2e.py from the EXECUTED reference train().
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
from torch import Tensor

from oracle.dv3_oracle import (AdamState, actor_logits, categorical_normalise, clip_grad_norm, dense_stack, recurrent_step,
                               st_sample, transition_logits, twohot_log_prob, twohot_mean, world_model_phase)


def draw_noise(T: int, B: int, H: int, S: int, D: int, actions_dim: Sequence[int], seed: int) -> Dict[str, Tensor]:
    """Exp(1) draws of one update: the scan's posterior samples and, per behaviour phase (`_expl`, `_task`), the imagined
    states and one tensor per action head."""
    g = torch.Generator().manual_seed(seed)
    N = T * B

    def exp1(*shape):
        return torch.empty(*shape).exponential_(1.0, generator=g)

    out = {"prior": exp1(T, B, S, D), "post": exp1(T, B, S, D)}
    for ph in ("expl", "task"):
        out[f"img_state_{ph}"] = exp1(H, N, S, D)
        out[f"img_action_{ph}"] = [exp1(H + 1, N, ad) for ad in actions_dim]
    return out


def reference_noise_order(noise: Dict[str, Tensor], T: int, H: int, n_heads: int) -> List[Tensor]:
    """the order in which the reference's train() calls torch.multinomial (scan: prior then posterior per step;
    per behaviour phase: actor at step 0, then (transition, actor) per imagined step, then the actor's re-evaluation
    on the whole trajectory, whose samples are discarded)"""
    out = []
    for t in range(T):
        out += [noise["prior"][t], noise["post"][t]]
    for ph in ("expl", "task"):
        acts = noise[f"img_action_{ph}"]
        out += [acts[k][0] for k in range(n_heads)]
        for i in range(1, H + 1):
            out.append(noise[f"img_state_{ph}"][i - 1])
            out += [acts[k][i] for k in range(n_heads)]
        out += [torch.ones(acts[k].shape).reshape(-1, acts[k].shape[-1]) for k in range(n_heads)]
    return out


def _moments(state: Dict[str, Tensor], lam: Tensor, mo):
    """Moments.forward (dreamer_v3/utils.py:56-63) -> (offset, invscale)"""
    lo = torch.quantile(lam.detach().flatten(), mo.percentile.low)
    hi = torch.quantile(lam.detach().flatten(), mo.percentile.high)
    state["low"] = mo.decay * state["low"] + (1 - mo.decay) * lo
    state["high"] = mo.decay * state["high"] + (1 - mo.decay) * hi
    return state["low"], torch.maximum(torch.tensor(1.0 / mo.max), state["high"] - state["low"])


def _lambda_values(rew, values, cont, gamma, lmbda):
    """compute_lambda_values (dreamer_v3/utils.py:66-77) on [H+1,N,1] inputs -> [H,N,1]"""
    c = cont[1:] * gamma
    interm = rew[1:] + c * values[1:] * (1 - lmbda)
    nxt = values[-1]
    lam = []
    for t in reversed(range(rew.shape[0] - 1)):
        nxt = interm[t] + c[t] * lmbda * nxt
        lam.append(nxt)
    return torch.stack(list(reversed(lam)))


def ensemble_forward(ens: Dict[str, Tensor], i: int, x: Tensor, n_hid: int, eps: float) -> Tensor:
    return dense_stack(ens, f"{i}._model.", x, n_hid, eps, True)


def _rollout(cfg, wm_c, actor, zs, hs, img_state, img_action, actions_dim, condition_margin):
    """H imagined steps from every posterior state with `actor` (p2e_dv3_exploration.py:242-269 / :397-424)"""
    a, w = cfg.algo, cfg.algo.world_model
    S, D = w.stochastic_size, w.discrete_size
    eps, um, n_hid, H = a.mlp_layer_norm.kw.eps, a.unimix, a.mlp_layers, a.horizon
    N = zs.shape[0] * zs.shape[1]
    zi, hi = zs.detach().reshape(N, -1), hs.detach().reshape(N, -1)

    def act(state, i):
        ls = actor_logits(actor, state, n_hid, actions_dim, um, eps)
        return torch.cat([st_sample(l, 1, ad, img_action[k][i], condition_margin)
                          for k, (l, ad) in enumerate(zip(ls, actions_dim))], -1)

    traj, acts = [torch.cat((zi, hi), -1)], []
    acts.append(act(traj[0], 0))
    for i in range(1, H + 1):
        hi = recurrent_step(wm_c, zi, acts[-1], hi, eps)
        zi = st_sample(transition_logits(wm_c, hi, S, D, um, eps), S, D, img_state[i - 1], condition_margin)
        traj.append(torch.cat((zi, hi), -1))
        acts.append(act(traj[-1], i))
    return torch.stack(traj), torch.stack(acts)


def _policy_loss(cfg, actor, traj, acts, advantage, discount, actions_dim):
    a = cfg.algo
    ls = actor_logits(actor, traj, a.mlp_layers, actions_dim, a.unimix, a.mlp_layer_norm.kw.eps)
    logp, ent = 0.0, 0.0
    for l, av in zip(ls, torch.split(acts, list(actions_dim), -1)):
        lg, pr = categorical_normalise(l)
        logp = logp + lg.gather(-1, av.argmax(-1, keepdim=True))
        ent = ent + (-(torch.clamp(lg, min=torch.finfo(lg.dtype).min) * pr).sum(-1))
    objective = logp[:-1] * advantage.detach()
    entropy = a.actor.ent_coef * ent
    return -torch.mean(discount[:-1].detach() * (objective + entropy.unsqueeze(-1)[:-1]))


def _critic_update(cfg, critic, target, opt, traj, lam, discount):
    a = cfg.algo
    eps, n_hid = a.mlp_layer_norm.kw.eps, a.mlp_layers
    qv = dense_stack(critic, "_model.", traj[:-1], n_hid, eps, True)
    with torch.no_grad():
        tgt = twohot_mean(dense_stack(target, "_model.", traj[:-1], n_hid, eps, True))
    loss = -twohot_log_prob(qv, lam.detach()) - twohot_log_prob(qv, tgt)
    loss = torch.mean(loss * discount[:-1].squeeze(-1))
    loss.backward()
    with torch.no_grad():
        norm = clip_grad_norm([v.grad for v in critic.values()], a.critic.clip_gradients)
        opt.step(critic, {k: v.grad for k, v in critic.items()})
    return loss.detach(), norm


def p2e_train_step(cfg, wm, ensembles, actor_task, critic_task, target_task, actor_expl, critics_expl,
                   opts: Dict[str, AdamState], data, noise, moments_task, actions_dim, condition_margin: float = 0.0):
    """One exploration update.  Parameter dicts are mutated in place.
    ensembles: one dict with the reference's ModuleList keys "{i}._model.{j}.weight|bias";
    critics_expl: {name: {"weight", "reward_type", "module": params, "target_module": params, "moments": {"low","high"}}}
    (insertion order = the reference's); opts: {"wm","ens","actor_task","critic_task","actor_expl","critic_expl_<name>"}."""
    a, w = cfg.algo, cfg.algo.world_model
    T, B, H = a.per_rank_sequence_length, a.per_rank_batch_size, a.horizon
    S, D = w.stochastic_size, w.discrete_size
    Z = S * D
    N = T * B
    eps, n_hid = a.mlp_layer_norm.kw.eps, a.mlp_layers
    out: Dict[str, Tensor] = {}
    trainable = [wm, ensembles, actor_task, critic_task, actor_expl] + [c["module"] for c in critics_expl.values()]
    for d in trainable:
        for v in d.values():
            v.requires_grad_(True)
            v.grad = None

    # ---- 1. dynamic learning
    zs, hs, cont_target = world_model_phase(cfg, wm, opts["wm"], data, noise, condition_margin, False, out,
                                            detach_heads=True)
    zs, hs = zs.detach(), hs.detach()

    # ---- 2. ensemble learning (:212-240).  NB the clip covers the LAST member only (`module=ens` after the loop)
    n_ens = a.ensembles.n
    ens_in = torch.cat((zs, hs, data["actions"].float()), -1)
    loss = 0.0
    for i in range(n_ens):
        pred = ensemble_forward(ensembles, i, ens_in, a.ensembles.mlp_layers, eps)[:-1]
        loss = loss + ((pred - zs[1:]) ** 2).sum(-1).mean()
    loss.backward()
    with torch.no_grad():
        last = [v.grad for k, v in ensembles.items() if k.startswith(f"{n_ens - 1}.")]
        out["Grads/ensemble"] = clip_grad_norm(last, a.ensembles.clip_gradients)
        opts["ens"].step(ensembles, {k: v.grad for k, v in ensembles.items()})
    out["Loss/ensemble_loss"] = loss.detach()

    wm_c = {k: v.detach() for k, v in wm.items()}
    ens_c = {k: v.detach() for k, v in ensembles.items()}
    true_cont = cont_target.reshape(1, N, 1)

    def continues(traj):
        c = (torch.sigmoid(dense_stack(wm_c, "continue_model._model.", traj, n_hid, eps, True)) > 0.5).float()
        return torch.cat((true_cont, c[1:]), 0)

    # ---- 3. behaviour learning: exploration (:242-392)
    with torch.no_grad():
        traj, acts = _rollout(cfg, wm_c, actor_expl, zs, hs, noise["img_state_expl"], noise["img_action_expl"], actions_dim,
                              condition_margin)
        cont = continues(traj)
        discount = torch.cumprod(cont * a.gamma, 0) / a.gamma
        weights_sum = sum(c["weight"] for c in critics_expl.values())
        advantage = 0.0
        for name, c in critics_expl.items():
            values = twohot_mean(dense_stack(c["module"], "_model.", traj, n_hid, eps, True))
            if c["reward_type"] == "intrinsic":
                emb = torch.stack([ensemble_forward(ens_c, i, torch.cat((traj, acts), -1), a.ensembles.mlp_layers, eps)
                                   for i in range(n_ens)])
                rew = emb.var(0).mean(-1, keepdim=True) * a.intrinsic_reward_multiplier
                out[f"Rewards/intrinsic_{name}"] = rew.mean()
            else:
                rew = twohot_mean(dense_stack(wm_c, "reward_model._model.", traj, n_hid, eps, True))
            lam = _lambda_values(rew, values, cont, a.gamma, a.lmbda)
            c["lambda_values"] = lam
            offset, invscale = _moments(c["moments"], lam, a.actor.moments)
            advantage = advantage + ((lam - offset) / invscale - (values[:-1] - offset) / invscale) * c["weight"] / weights_sum
            out[f"Values_exploration/predicted_values_{name}"] = values.mean()
            out[f"Values_exploration/lambda_values_{name}"] = lam.mean()
    policy_loss = _policy_loss(cfg, actor_expl, traj, acts, advantage, discount, actions_dim)
    policy_loss.backward()
    with torch.no_grad():
        out["Grads/actor_exploration"] = clip_grad_norm([v.grad for v in actor_expl.values()], a.actor.clip_gradients)
        opts["actor_expl"].step(actor_expl, {k: v.grad for k, v in actor_expl.items()})
    out["Loss/policy_loss_exploration"] = policy_loss.detach()
    for name, c in critics_expl.items():
        vl, norm = _critic_update(cfg, c["module"], c["target_module"], opts[f"critic_expl_{name}"], traj, c["lambda_values"],
                                  discount)
        out[f"Loss/value_loss_exploration_{name}"], out[f"Grads/critic_exploration_{name}"] = vl, norm

    # ---- 4. behaviour learning: task (:397-474)
    with torch.no_grad():
        traj, acts = _rollout(cfg, wm_c, actor_task, zs, hs, noise["img_state_task"], noise["img_action_task"], actions_dim,
                              condition_margin)
        values = twohot_mean(dense_stack(critic_task, "_model.", traj, n_hid, eps, True))
        rew = twohot_mean(dense_stack(wm_c, "reward_model._model.", traj, n_hid, eps, True))
        cont = continues(traj)
        lam = _lambda_values(rew, values, cont, a.gamma, a.lmbda)
        discount = torch.cumprod(cont * a.gamma, 0) / a.gamma
        offset, invscale = _moments(moments_task, lam, a.actor.moments)
        advantage = (lam - offset) / invscale - (values[:-1] - offset) / invscale
    policy_loss = _policy_loss(cfg, actor_task, traj, acts, advantage, discount, actions_dim)
    policy_loss.backward()
    with torch.no_grad():
        out["Grads/actor_task"] = clip_grad_norm([v.grad for v in actor_task.values()], a.actor.clip_gradients)
        opts["actor_task"].step(actor_task, {k: v.grad for k, v in actor_task.items()})
    out["Loss/policy_loss_task"] = policy_loss.detach()
    out["Loss/value_loss_task"], out["Grads/critic_task"] = _critic_update(cfg, critic_task, target_task, opts["critic_task"],
                                                                           traj, lam, discount)
    for d in trainable:
        for v in d.values():
            v.grad = None
            v.requires_grad_(False)
    for c in critics_expl.values():
        c.pop("lambda_values", None)
    return out
