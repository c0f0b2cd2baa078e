"""TEST INFRASTRUCTURE — run the REAL reference train() (container only) with injected noise.

Used by tests/test_oracle_pin.py and oracle/make_golden.py.
"""
from __future__ import annotations

import copy
from typing import Dict, Sequence

import torch

from oracle import ref_harness
from oracle.dv3_oracle import reference_noise_order, reference_normal_order, vec_dims


def to_ref_cfg(cfg):
    ref_harness.install()
    from sheeprl.utils.utils import dotdict as ref_dotdict

    return ref_dotdict(copy.deepcopy(cfg.as_dict()))


def build_reference_agent(cfg, actions_dim: Sequence[int], in_channels: int = 3, seed: int = 0,
                          is_continuous: bool = False):
    ref_harness.install()
    from sheeprl.algos.dreamer_v3.agent import build_agent

    rcfg = to_ref_cfg(cfg)
    fab = ref_harness.FakeFabric()
    sz = cfg.env.screen_size
    cch = dict(cfg.env.get("cnn_channels", {}) or {})
    multi = len(cfg.algo.cnn_keys.encoder) > 1
    obs_space = {k: ref_harness.Shape((cch[k] if multi else in_channels, sz, sz)) for k in cfg.algo.cnn_keys.encoder}
    obs_space.update({k: ref_harness.Shape((d,)) for k, d in vec_dims(cfg).items()})
    torch.manual_seed(seed)
    wm, actor, critic, target, player = build_agent(fab, tuple(actions_dim), is_continuous, rcfg, obs_space)
    return fab, rcfg, wm, actor, critic, target, player


def reference_state_dicts(wm, actor, critic, target) -> Dict[str, Dict[str, torch.Tensor]]:
    def sd(m):
        return {k.replace("_forward_module.", ""): v.detach().clone() for k, v in m.state_dict().items()}

    return {"wm": sd(wm), "actor": sd(actor), "critic": sd(critic), "target": sd(target)}


def run_reference_train(cfg, actions_dim, data, noise, n_steps: int = 1, in_channels: int = 3, seed: int = 0,
                        state=None, moments_state=None, is_continuous: bool = False):
    """Returns (state_dicts_after, metrics list, moments(low,high)).  `state`: optional dict of state
    dicts to load before stepping."""
    ref_harness.install()
    from sheeprl.algos.dreamer_v3 import dreamer_v3 as D
    from sheeprl.algos.dreamer_v3.utils import Moments

    fab, rcfg, wm, actor, critic, target, _ = build_reference_agent(cfg, actions_dim, in_channels, seed, is_continuous)
    if state is not None:
        for mod, name in ((wm, "wm"), (actor, "actor"), (critic, "critic"), (target, "target")):
            _load(mod, state[name])
    a = cfg.algo

    def adam(params, o):
        return torch.optim.Adam(params, lr=o.lr, eps=o.eps, weight_decay=o.weight_decay, betas=tuple(o.betas))

    wo = adam(wm.parameters(), a.world_model.optimizer)
    ao = adam(actor.parameters(), a.actor.optimizer)
    co = adam(critic.parameters(), a.critic.optimizer)
    mo = a.actor.moments
    moments = Moments(mo.decay, mo.max, mo.percentile.low, mo.percentile.high)
    if moments_state is not None:
        moments.low = moments_state["low"].clone()
        moments.high = moments_state["high"].clone()
    metrics = []
    T, H = a.per_rank_sequence_length, a.horizon
    for s in range(n_steps):
        agg = ref_harness.RecordingAggregator()
        batch = {k: v.clone().float() for k, v in data[s].items()}
        if is_continuous:
            # categorical draws: prior/post per step + the imagined states; Normal.rsample: the actions
            import torch.distributions.normal as TN

            cat = []
            for t in range(T):
                cat += [noise[s]["prior"][t], noise[s]["post"][t]]
            cat += [noise[s]["img_state"][i] for i in range(H)]
            normal = reference_normal_order(noise[s], H)
            orig = TN._standard_normal
            TN._standard_normal = lambda shape, dtype, device: normal.pop(0).reshape(shape)
            try:
                with ref_harness.NoiseQueue(cat):
                    D.train(fab, wm, actor, critic, target, wo, ao, co, batch, agg, rcfg, True, tuple(actions_dim), moments)
            finally:
                TN._standard_normal = orig
            assert not normal, "the reference drew fewer Normal samples than expected"
        else:
            with ref_harness.NoiseQueue(reference_noise_order(noise[s], T, H, len(actions_dim))):
                D.train(fab, wm, actor, critic, target, wo, ao, co, batch, agg, rcfg, False, tuple(actions_dim), moments)
        metrics.append(agg.values)
    return (reference_state_dicts(wm, actor, critic, target), metrics,
            {"low": moments.low.detach().clone(), "high": moments.high.detach().clone()})


def _load(module, sd):
    own = module.state_dict()
    with torch.no_grad():
        for k, v in own.items():
            v.copy_(sd[k.replace("_forward_module.", "")])
