"""SAC on a GPU-less host: the oracle against the executed reference (tests/golden/sac_*.pt), and the engine's kernel
schedule (hand-derived backward) against the same fixtures with the torch test double in place of the CUDA ops."""
import os

import pytest
import torch

from oracle import sac_oracle as SO
from oracle.dv3_oracle import AdamState

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)


def clone_params(P):
    return {g: {k: v.clone() for k, v in d.items()} for g, d in P.items()}


def assert_group_close(got, want, what, rtol=1e-4, atol=1e-5, lr=3e-4):
    """Adam's first steps move every weight by ~lr*sign(g): a gradient that is 0 up to rounding may flip sign, so a
    handful of entries may differ by up to 2*lr (same rule as tests/helpers.assert_params_close)."""
    for k, w in want.items():
        g = got[k].detach().cpu()
        bad = (g - w).abs() > atol + rtol * w.abs()
        if bad.any():
            assert bad.float().mean() < 2e-3 and (g - w).abs().max() <= 2.2 * lr * 3, (what, k, int(bad.sum()), float((g - w).abs().max()))


def run_oracle(fx):
    sp = fx["spec"]
    P = clone_params(fx["init"])
    opts = [AdamState(P[g], 3e-4, 1e-4) for g in ("qf", "actor", "log_alpha")]
    scale = torch.full((sp["act_dim"],), (sp["high"] - sp["low"]) / 2.0)
    bias = torch.full((sp["act_dim"],), (sp["high"] + sp["low"]) / 2.0)
    for st in fx["steps"]:
        losses = SO.sac_train_step(P, *opts, st["data"], st["eps_next"], st["eps_cur"], 0.99, 0.005,
                                   st["update"] % 2 == 0, sp["n_critics"], scale, bias, -float(sp["act_dim"]))
        yield st, P, losses


@pytest.mark.parametrize("name", ["sac_tiny", "sac_c4"])
def test_oracle_matches_reference(name):
    fx = load(name)
    for st, P, losses in run_oracle(fx):
        for k, v in st["losses"].items():
            assert abs(losses[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, losses[k], v)
        if st["after"] is not None:
            for g in ("actor", "qf", "qf_target", "log_alpha"):
                assert_group_close(P[g], st["after"][g], f"{name}/u{st['update']}/{g}")


# ---------------------------------------------------------------------------------------------------------
# engine schedule (hand-derived backward) with the torch test double
# ---------------------------------------------------------------------------------------------------------
def make_engine(fx, device="cpu", ops=None):
    from oracle.ops_emul import EmulOps
    from sheeprl_b200.algos.sac.engine import SACEngine

    sp = fx["spec"]
    opt = {"lr": 3e-4, "eps": 1e-4, "betas": (0.9, 0.999)}
    eng = SACEngine(sp["obs_dim"], sp["act_dim"], sp["hidden"], sp["hidden"], sp["n_critics"], sp["B"], 0.99, 0.005, 1.0,
                    sp["low"], sp["high"], opt, opt, opt, device, ops or EmulOps())
    init = fx["init"]
    eng.load_reference_state(init["actor"], init["qf"], init["qf_target"], init["log_alpha"]["log_alpha"])
    return eng


def check_engine(fx, eng, name):
    dev = eng.device
    for st in fx["steps"]:
        data = {k: v.to(dev) for k, v in st["data"].items()}
        noise = {"eps_next": st["eps_next"].to(dev), "eps_cur": st["eps_cur"].to(dev)}
        eng.train_step(data, st["update"] % 2 == 0, noise)
        md = {k: float(v) for k, v in eng.metrics_dict().items()}
        for k, v in st["losses"].items():
            assert abs(md[k] - v) <= 1e-4 * max(1.0, abs(v)), (name, st["update"], k, md[k], v)
        if st["after"] is not None:
            got = eng.export_reference_state()
            for g in ("actor", "qf", "qf_target", "log_alpha"):
                assert_group_close(got[g], st["after"][g], f"{name}/u{st['update']}/{g}")


@pytest.mark.parametrize("name", ["sac_tiny", "sac_c4"])
def test_engine_schedule_matches_reference(name):
    fx = load(name)
    check_engine(fx, make_engine(fx), name)


def test_public_api_and_state_dict_roundtrip():
    from oracle.ops_emul import EmulOps
    from oracle.make_golden_sac import sac_cfg
    from sheeprl_b200.algos.sac.agent import build_agent
    from sheeprl_b200.algos.sac.sac import make_optimizers, train
    import numpy as np

    class Fab:
        device, world_size, global_rank = torch.device("cpu"), 1, 0

    class Space:
        def __init__(self, shape):
            self.shape = shape

    class Box:
        shape, low, high = (3,), np.full(3, -2.0, np.float32), np.full(3, 1.0, np.float32)

    cfg = sac_cfg(16, 2)
    cfg.algo.per_rank_batch_size = 8
    agent, player = build_agent(Fab, cfg, {"state": Space((5,))}, Box, ops=EmulOps())
    fx = load("sac_tiny")
    # a reference-format state dict (keys of SACAgent.state_dict() in the reference) loads and round-trips
    sd = agent.state_dict()
    assert set(sd) >= {"_actor.fc_mean.weight", "_actor.fc_logstd.bias", "_qfs.1.model._model.4.weight",
                       "_qfs_target.0.model._model.0.bias", "_log_alpha"}
    ref_sd = {}
    for k, v in fx["init"]["actor"].items():
        ref_sd[f"_actor._forward_module.{k}"] = v
    for k, v in fx["init"]["qf"].items():
        ref_sd[f"_qfs.{k}"] = v
    for k, v in fx["init"]["qf_target"].items():
        ref_sd[f"_qfs_target.{k}"] = v
    ref_sd["_log_alpha"] = fx["init"]["log_alpha"]["log_alpha"]
    agent.load_state_dict(ref_sd)
    opts = make_optimizers(agent, cfg)

    class Agg:
        disabled = False

        def __init__(self):
            self.v = {}

        def update(self, k, v):
            self.v[k] = float(v)

    agg = Agg()
    st = fx["steps"][0]
    train(Fab, agent, *opts, dict(st["data"]), agg, st["update"], cfg, 1,
          noise={"eps_next": st["eps_next"], "eps_cur": st["eps_cur"]})
    for k, v in st["losses"].items():
        assert abs(agg.v[k] - v) <= 1e-4 * max(1.0, abs(v))
    assert opts[0].state_dict()["state"][0]["step"] == 1


def test_player_greedy_and_sampled_actions():
    """SACPlayer (sac/agent.py:270-314): greedy = tanh(mean)*scale+bias; sampled actions stay inside the bounds"""
    from sheeprl_b200.algos.sac.agent import SACPlayer

    fx = load("sac_tiny")
    eng = make_engine(fx)
    player = SACPlayer(eng)
    obs = fx["steps"][0]["data"]["observations"]
    P = fx["init"]["actor"]
    x = torch.relu(obs @ P["model._model.0.weight"].t() + P["model._model.0.bias"])
    x = torch.relu(x @ P["model._model.2.weight"].t() + P["model._model.2.bias"])
    mean = x @ P["fc_mean.weight"].t() + P["fc_mean.bias"]
    sp = fx["spec"]
    scale, bias = (sp["high"] - sp["low"]) / 2.0, (sp["high"] + sp["low"]) / 2.0
    got = player.get_actions(obs, greedy=True)
    assert got.shape == (sp["B"], sp["act_dim"])
    assert float((got - (torch.tanh(mean) * scale + bias)).abs().max()) < 1e-5
    a1, a2 = player.get_actions(obs), player.get_actions(obs)
    assert not torch.equal(a1, a2) and float(a1.min()) >= sp["low"] - 1e-6 and float(a1.max()) <= sp["high"] + 1e-6
