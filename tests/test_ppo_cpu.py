"""PPO on a GPU-less host: the oracle against the executed reference (tests/golden/ppo_*.pt), and the engine's kernel
schedule (hand-derived backward) against the same fixtures with the torch test double in place of the CUDA ops."""
import os

import pytest
import torch

from oracle import ppo_oracle as PO
from oracle.dv3_oracle import AdamState

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ["ppo_vector", "ppo_branches", "ppo_continuous", "ppo_pixel", "ppo_tanh_ln", "ppo_multikey"]


def load(name):
    fx = torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)
    if "rgb" in fx["data"]:
        fx["data"]["rgb"] = fx["data"]["rgb"].float()
    return fx


def assert_params_close(got, want, what, rtol=1e-4, atol=2e-5, lr=1e-3, steps=1):
    """Adam's first steps move a weight by ~lr*sign(g): entries whose gradient is 0 up to rounding may differ by up to
    2*lr per step; everything else must agree to 1e-4 relative."""
    for k, w in want.items():
        g = got[k].detach().cpu()
        err = (g - w).abs()
        bad = err > atol + rtol * w.abs()
        if bad.any():
            assert bad.float().mean() < 5e-3 and float(err.max()) <= 2.2 * lr * steps, (what, k, int(bad.sum()), float(err.max()))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name):
    fx = load(name)
    p = {k: v.clone() for k, v in fx["init"].items()}
    assert {k: tuple(v.shape) for k, v in p.items()} == PO.ppo_param_shapes(fx["spec"])
    opt = AdamState(p, 1e-3, 1e-4)
    logs = PO.ppo_train(p, opt, fx["spec"], fx["data"], fx["index_batches"], fx["hp"])
    for got, want in zip(logs, fx["losses"]):
        for k, v in want.items():
            assert abs(got[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, got[k], v)
    assert_params_close(p, fx["after"], name, steps=len(logs))


# ---------------------------------------------------------------------------------------------------------
# engine schedule with the torch test double
# ---------------------------------------------------------------------------------------------------------
def make_engine(fx, device="cpu", ops=None):
    from oracle.ops_emul import EmulOps
    from sheeprl_b200.algos.ppo.engine import PPOEngine

    eng = PPOEngine(fx["spec"], fx["hp"], {"lr": 1e-3, "eps": 1e-4, "betas": (0.9, 0.999)}, device, ops or EmulOps())
    assert dict(eng.reference_shapes()) == PO.ppo_param_shapes(fx["spec"])
    eng.load_reference_state(fx["init"])
    return eng


def check_engine(fx, eng, name, uint8_image=False):
    dev = eng.device
    data = {k: v.to(dev) for k, v in fx["data"].items()}
    if uint8_image and "rgb" in data:
        data["rgb"] = data["rgb"].to(torch.uint8)
    logs = []
    eng.train(data, fx["index_batches"], lambda l: logs.append(l.cpu()))
    assert len(logs) == len(fx["losses"])
    for got, want in zip(logs, fx["losses"]):
        for i, k in enumerate(("Loss/policy_loss", "Loss/value_loss", "Loss/entropy_loss")):
            assert abs(float(got[i]) - want[k]) <= 1e-4 * max(1.0, abs(want[k])), (name, k, float(got[i]), want[k])
    assert_params_close(eng.export_reference_state(), fx["after"], name, steps=len(logs))


@pytest.mark.parametrize("name", NAMES)
def test_engine_schedule_matches_reference(name):
    fx = load(name)
    check_engine(fx, make_engine(fx), name)


@pytest.mark.parametrize("name", ["ppo_branches", "ppo_tanh_ln", "ppo_multikey"])
def test_public_api_draws_the_reference_minibatches(name):
    """train() without explicit indices uses torch's RandomSampler/BatchSampler exactly as the reference: under the
    fixture's sampler seed it visits the recorded minibatches and lands on the reference's parameters.  The config
    tree is the one the reference was run with (layer_norm / distribution / per-network sizes / several obs keys) and
    the rollout holds one tensor per observation key."""
    from oracle.make_golden_ppo import obs_space, ppo_cfg, split_obs
    from oracle.ops_emul import EmulOps
    from sheeprl_b200.algos.ppo.agent import build_agent
    from sheeprl_b200.algos.ppo.ppo import make_optimizer, train

    fx = load(name)

    class Fab:
        device, world_size, global_rank = torch.device("cpu"), 1, 0

    cfg = ppo_cfg(fx["spec"], fx["hp"], fx["batch"], fx["epochs"])
    agent, _ = build_agent(Fab, fx["spec"]["actions_dim"], fx["spec"]["is_continuous"], cfg, obs_space(fx["spec"]),
                           agent_state=fx["init"], ops=EmulOps())
    opt = make_optimizer(agent, cfg)

    class _Agg:
        disabled = False
        rows = []

        def update(self, k, v):
            self.rows.append((k, float(v)))

    Agg = _Agg()
    torch.manual_seed(fx["sampler_seed"])
    train(Fab, agent, opt, split_obs(fx["spec"], fx["data"]), Agg, cfg)
    assert len(Agg.rows) == 3 * len(fx["losses"])
    assert abs(Agg.rows[-1][1] - fx["losses"][-1]["Loss/entropy_loss"]) < 1e-4
    assert_params_close(agent.state_dict(), fx["after"], "public", steps=len(fx["losses"]))
    assert opt.state_dict()["state"][0]["step"] == len(fx["losses"])


# ---------------------------------------------------------------------------------------------------------
# PPOPlayer against the executed reference player (tests/golden/ppo_player.pt, oracle/make_golden_ppo_player.py)
# ---------------------------------------------------------------------------------------------------------
def check_player(name, device="cpu", ops=None, uint8_image=False):
    from sheeprl_b200.algos.ppo.agent import PPOPlayer

    pf = torch.load(os.path.join(GOLDEN, "ppo_player.pt"), weights_only=False)[name]
    fx = load(name)
    eng = make_engine(fx, device=device, ops=ops)
    player = PPOPlayer(eng)
    from oracle.make_golden_ppo import split_obs

    obs = {k: v.to(device) for k, v in pf["obs"].items()}
    if uint8_image and "rgb" in obs:
        obs["rgb"] = torch.round((obs["rgb"] + 0.5) * 255).to(torch.uint8)
    obs = split_obs(fx["spec"], obs)                       # one tensor per observation key, as the rollout loop passes
    assert player.actor.distribution == (fx["spec"].get("dist", "normal") if fx["spec"]["is_continuous"] else "discrete")
    actions, logp, values = player(obs, noise=pf["noise"].to(device))
    cont = fx["spec"]["is_continuous"]
    for got, want in zip(actions, pf["actions"]):
        if cont:
            assert float((got.cpu() - want).abs().max()) <= 1e-5
        else:
            assert torch.equal(got.cpu(), want)
    assert float((logp.cpu() - pf["logp"]).abs().max()) <= 1e-4 * max(1.0, float(pf["logp"].abs().max()))
    assert float((values.cpu() - pf["values"]).abs().max()) <= 1e-4 * max(1.0, float(pf["values"].abs().max()))
    assert float((player.get_values(obs).cpu() - pf["values2"]).abs().max()) <= 1e-4 * max(1.0, float(pf["values2"].abs().max()))
    for got, want in zip(player.get_actions(obs, greedy=True), pf["greedy"]):
        assert float((got.cpu() - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
    if pf.get("sampled") is not None:                      # get_actions(greedy=False) on the same injected noise
        for got, want in zip(player.get_actions(obs, greedy=False, noise=pf["noise"].to(device)), pf["sampled"]):
            assert float((got.cpu() - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max()))


PLAYER_NAMES = ["ppo_branches", "ppo_continuous", "ppo_pixel", "ppo_tanh_ln", "ppo_multikey"]


@pytest.mark.parametrize("name", PLAYER_NAMES)
def test_player_matches_reference(name):
    check_player(name)
