"""Shared helpers for the parity tests (test infrastructure)."""
import os

import torch

from oracle import dv3_oracle as O
from sheeprl_b200.configs import make_dv3_cfg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    fx = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    cfg = make_dv3_cfg(**fx["cfg_kwargs"])
    return fx, cfg


def clone_state(sd):
    return {n: {k: v.clone() for k, v in d.items()} for n, d in sd.items()}


def oracle_run(cfg, adim, init, data, noise, steps, condition_margin=0.0, keep=False, is_continuous=False):
    st = clone_state(init)
    a, w = cfg.algo, cfg.algo.world_model
    opts = [O.AdamState(st["wm"], w.optimizer.lr, w.optimizer.eps),
            O.AdamState(st["actor"], a.actor.optimizer.lr, a.actor.optimizer.eps),
            O.AdamState(st["critic"], a.critic.optimizer.lr, a.critic.optimizer.eps)]
    ms = {"low": torch.zeros(()), "high": torch.zeros(())}
    outs = []
    for s in range(steps):
        outs.append(O.dv3_train_step(cfg, st["wm"], st["actor"], st["critic"], st["target"], *opts, data[s], noise[s],
                                     ms, adim, condition_margin=condition_margin, keep=keep,
                                     is_continuous=is_continuous))
    return st, outs, ms, opts


def max_diff(a, b):
    return max(float((a[k].float() - b[k].float()).abs().max()) for k in a)


def assert_params_close(got, want, lr, steps, tol=1e-6, frac=2e-3, label=""):
    """Post-Adam parameters.  Adam's first steps move every element by ~lr * sign(g): an element whose
    gradient is numerically zero (|g| ~ 1e-9, e.g. directions a following LayerNorm cancels) can take
    either sign, so a handful of elements may legitimately differ by 2*lr per step.  Everything else must
    agree to `tol`; no element may differ by more than the sign-flip bound."""
    worst_frac = 0.0
    for k in want:
        d = (got[k].float() - want[k].float()).abs()
        bad = d > tol
        assert float(d.max()) <= 2.5 * lr * steps + tol, (label, k, float(d.max()))
        f = float(bad.float().mean())
        worst_frac = max(worst_frac, f)
        assert f <= frac or int(bad.sum()) <= 3, (label, k, f, int(bad.sum()))
    return worst_frac


def image_channels(cfg) -> int:
    """total channels of the image keys of a fixture config (3 for the single-key fixtures)"""
    cch = dict(cfg.env.get("cnn_channels", {}) or {})
    return sum(cch.values()) if len(cch) > 1 else 3
