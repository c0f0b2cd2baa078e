"""TEST INFRASTRUCTURE — writes tests/golden/p2e_tiny.pt by EXECUTING THE REAL REFERENCE
`sheeprl.algos.p2e_dv3.p2e_dv3_exploration.train` (container only):

    python -m oracle.make_golden_p2e

Fixture: config kwargs, initial parameters of every module (reference `build_agent`, perturbed so that LayerNorm
affine / biases / the zero-initialised heads are exercised), two replay batches, the Exp(1) noise of both updates
(conditioned by the oracle so no draw sits on a near-tie), the metrics the reference logged and every module's
parameters + the five Moments after two updates.
"""
from __future__ import annotations

import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import dv3_oracle as O  # noqa: E402
from oracle import p2e_oracle as P  # noqa: E402
from oracle import ref_harness, ref_run  # noqa: E402
from sheeprl_b200.configs import make_p2e_dv3_cfg  # noqa: E402

CFG = dict(size="S", per_rank_batch_size=3, per_rank_sequence_length=5, horizon=4, dense_units=32, mlp_layers=2,
           cnn_channels_multiplier=4, recurrent_state_size=24, hidden_size=32, stochastic_size=6, discrete_size=5, bins=31,
           n_ensembles=3, intrinsic_weight=0.4, extrinsic_weight=1.0, intrinsic_reward_multiplier=2.0,
           algo__world_model__kl_free_nats=0.05)      # KL term active (the default 1.0 clamps it away at this tiny state size)
ACTIONS_DIM = (3, 2)
STEPS = 2


def _sd(m):
    return {k.replace("_forward_module.", ""): v.detach().clone() for k, v in m.state_dict().items()}


def build_reference(cfg, seed=0):
    ref_harness.install()
    import contextlib

    import sheeprl.algos.p2e_dv3.agent as PA

    # Fabric plumbing the harness's FakeFabric does not have: single-device re-wrap, RNG isolation while the ensemble
    # members are seeded one by one (agent.py:117, 177-180)
    PA.get_single_device_fabric = lambda f: f
    PA.isolate_rng = contextlib.nullcontext
    build_agent = PA.build_agent
    rcfg = ref_run.to_ref_cfg(cfg)
    fab = ref_harness.FakeFabric()
    fab.seed_everything = lambda s: torch.manual_seed(s)
    sz = cfg.env.screen_size
    space = {k: ref_harness.Shape((3, sz, sz)) for k in cfg.algo.cnn_keys.encoder}
    torch.manual_seed(seed)
    wm, ens, actor_task, critic_task, target_task, actor_expl, critics_expl, _ = build_agent(
        fab, ACTIONS_DIM, False, rcfg, space)
    return fab, rcfg, wm, ens, actor_task, critic_task, target_task, actor_expl, critics_expl


def export(wm, ens, actor_task, critic_task, target_task, actor_expl, critics_expl):
    out = {"wm": _sd(wm), "ens": _sd(ens), "actor_task": _sd(actor_task), "critic_task": _sd(critic_task),
           "target_task": _sd(target_task), "actor_expl": _sd(actor_expl)}
    for k, c in critics_expl.items():
        out[f"critic_expl_{k}"], out[f"target_expl_{k}"] = _sd(c["module"]), _sd(c["target_module"])
    return out


def oracle_state(cfg, sd):
    """parameter dicts / Adam states / Moments for oracle.p2e_oracle.p2e_train_step from exported state dicts"""
    a = cfg.algo
    p = {k: {n: v.clone() for n, v in d.items()} for k, d in sd.items()}
    critics = {}
    for k, v in a.critics_exploration.items():
        if v.weight > 0:
            critics[k] = {"weight": v.weight, "reward_type": v.reward_type, "module": p[f"critic_expl_{k}"],
                          "target_module": p[f"target_expl_{k}"], "moments": {"low": torch.zeros(()), "high": torch.zeros(())}}
    w = a.world_model
    opts = {"wm": O.AdamState(p["wm"], w.optimizer.lr, w.optimizer.eps),
            "ens": O.AdamState(p["ens"], a.ensembles.optimizer.lr, a.ensembles.optimizer.eps),
            "actor_task": O.AdamState(p["actor_task"], a.actor.optimizer.lr, a.actor.optimizer.eps),
            "critic_task": O.AdamState(p["critic_task"], a.critic.optimizer.lr, a.critic.optimizer.eps),
            "actor_expl": O.AdamState(p["actor_expl"], a.actor.optimizer.lr, a.actor.optimizer.eps)}
    for k in critics:
        opts[f"critic_expl_{k}"] = O.AdamState(p[f"critic_expl_{k}"], a.critic.optimizer.lr, a.critic.optimizer.eps)
    return p, critics, opts, {"low": torch.zeros(()), "high": torch.zeros(())}


def run_oracle(cfg, sd, data, noise, margin=0.0):
    p, critics, opts, mt = oracle_state(cfg, sd)
    metrics = []
    for s in range(len(data)):
        metrics.append(P.p2e_train_step(cfg, p["wm"], p["ens"], p["actor_task"], p["critic_task"], p["target_task"],
                                        p["actor_expl"], critics, opts, data[s], noise[s], mt, ACTIONS_DIM, margin))
    moments = {"task": mt, **{k: c["moments"] for k, c in critics.items()}}
    return p, metrics, moments


def run_reference(cfg, sd, data, noise):
    ref_harness.install()
    from sheeprl.algos.dreamer_v3.utils import Moments
    from sheeprl.algos.p2e_dv3 import p2e_dv3_exploration as X

    fab, rcfg, wm, ens, actor_task, critic_task, target_task, actor_expl, critics_expl = build_reference(cfg)
    mods = {"wm": wm, "ens": ens, "actor_task": actor_task, "critic_task": critic_task, "target_task": target_task,
            "actor_expl": actor_expl}
    for k, c in critics_expl.items():
        mods[f"critic_expl_{k}"], mods[f"target_expl_{k}"] = c["module"], c["target_module"]
    for k, m in mods.items():
        ref_run._load(m, sd[k])
    a = cfg.algo

    def adam(params, o):
        return torch.optim.Adam(params, lr=o.lr, eps=o.eps, weight_decay=o.weight_decay, betas=tuple(o.betas))

    wo, eo = adam(wm.parameters(), a.world_model.optimizer), adam(ens.parameters(), a.ensembles.optimizer)
    ato, cto = adam(actor_task.parameters(), a.actor.optimizer), adam(critic_task.parameters(), a.critic.optimizer)
    aeo = adam(actor_expl.parameters(), a.actor.optimizer)
    for c in critics_expl.values():
        c["optimizer"] = adam(c["module"].parameters(), a.critic.optimizer)
    mo = a.actor.moments
    new_m = lambda: Moments(mo.decay, mo.max, mo.percentile.low, mo.percentile.high)  # noqa: E731
    m_task, m_expl = new_m(), {k: new_m() for k in critics_expl}
    T, H = a.per_rank_sequence_length, a.horizon
    metrics = []
    for s in range(len(data)):
        agg = ref_harness.RecordingAggregator()
        batch = {k: v.clone().float() for k, v in data[s].items()}
        with ref_harness.NoiseQueue(P.reference_noise_order(noise[s], T, H, len(ACTIONS_DIM))):
            X.train(fab, wm, actor_task, critic_task, target_task, wo, ato, cto, batch, agg, rcfg, ens, eo, actor_expl,
                    critics_expl, aeo, m_expl, m_task, False, ACTIONS_DIM)
        metrics.append(agg.values)
    moments = {"task": m_task, **m_expl}
    moments = {k: {"low": v.low.detach().clone(), "high": v.high.detach().clone()} for k, v in moments.items()}
    return export(wm, ens, actor_task, critic_task, target_task, actor_expl, critics_expl), metrics, moments


def main():
    cfg = make_p2e_dv3_cfg(**CFG)
    built = build_reference(cfg)
    sd = export(*built[2:])
    g = torch.Generator().manual_seed(5)
    for name, d in sd.items():
        if name.startswith("target_"):
            continue
        for v in d.values():
            v.add_(torch.randn(v.shape, generator=g) * 0.05)
    sd["target_task"] = {k: v + 0.01 for k, v in sd["critic_task"].items()}
    for k in list(sd):
        if k.startswith("critic_expl_"):
            sd["target_expl_" + k[len("critic_expl_"):]] = {n: v - 0.01 for n, v in sd[k].items()}
    a, w = cfg.algo, cfg.algo.world_model
    T, B, H = a.per_rank_sequence_length, a.per_rank_batch_size, a.horizon
    data = [O.make_batch(cfg, ACTIONS_DIM, seed=1 + s) for s in range(STEPS)]
    noise = [P.draw_noise(T, B, H, w.stochastic_size, w.discrete_size, ACTIONS_DIM, seed=10 + s) for s in range(STEPS)]
    run_oracle(cfg, copy.deepcopy(sd), data, noise, margin=1e-3)          # conditions `noise` in place
    after, metrics, moments = run_reference(cfg, sd, data, noise)
    # the oracle must reproduce the executed reference before the fixture is written
    p, om, omom = run_oracle(cfg, copy.deepcopy(sd), data, noise)
    worst = 0.0
    for s in range(STEPS):
        for k, v in metrics[s].items():
            if k in om[s]:
                err = abs(float(om[s][k]) - float(v)) / max(1.0, abs(float(v)))
                worst = max(worst, err)
                assert err < 2e-4, (s, k, float(om[s][k]), float(v))
    print("oracle vs reference: worst relative metric error", worst)
    out = {"cfg": CFG, "actions_dim": ACTIONS_DIM, "init": sd, "data": data, "noise": noise, "after": after,
           "metrics": [{k: float(v) for k, v in m.items()} for m in metrics], "moments": moments}
    path = os.path.join(ROOT, "tests", "golden", "p2e_tiny.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path), sorted(metrics[-1]))


if __name__ == "__main__":
    main()
