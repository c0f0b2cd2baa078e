"""PlayerDV3 acting path against the executed reference PlayerDV3 (tests/golden/dv3_player_*.pt, written by
oracle/make_golden_player.py): states and actions after every call of the script (init, steps, partial reset, steps).
Here on a GPU-less host with the torch test double; tests/test_gpu_player.py runs the same through the C-ABI."""
import os

import pytest
import torch

from tests.helpers import GOLDEN, load_fixture


def run_player(name, device="cpu", ops=None, uint8_obs=False):
    from oracle.ops_emul import EmulOps
    from sheeprl_b200.algos.dreamer_v3.player import PlayerDV3
    from sheeprl_b200.engine import DV3Engine

    fx = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    tf, cfg = load_fixture(fx["train_fixture"])
    cont = tf.get("is_continuous", False)
    eng = DV3Engine(cfg, tf["actions_dim"], in_channels=3, device=device, ops=ops or EmulOps(), is_continuous=cont)
    eng.wm.load(tf["init"]["wm"]), eng.actor.load(tf["init"]["actor"])
    player = PlayerDV3(eng, fx["num_envs"])
    key = cfg.algo.cnn_keys.encoder[0] if cfg.algo.cnn_keys.encoder else None
    player.init_states()
    got = [{"h": player.recurrent_state.clone().cpu(), "z": player.stochastic_state.clone().cpu(), "a": player.actions.clone().cpu()}]
    for s, obs in enumerate(fx["obs"]):
        if s == fx["reset_at"]:
            player.init_states(fx["reset_envs"])
        o = {k: v.to(device) for k, v in obs.items()} if isinstance(obs, dict) else {key: obs.to(device)}
        if uint8_obs and key is not None:                       # raw pixels: the kernel normalises
            o[key] = torch.round((o[key] + 0.5) * 255).to(torch.uint8)
        acts = player.get_actions(o, noise={"z": fx["noise_z"][s].to(device), "a": fx["noise_a"][s].to(device)})
        assert torch.equal(torch.cat(acts, -1), player.actions)
        got.append({"h": player.recurrent_state.clone().cpu(), "z": player.stochastic_state.clone().cpu(),
                    "a": player.actions.clone().cpu()})
    return fx, got, cont


def check(fx, got, cont):
    for i, (g, w) in enumerate(zip(got, fx["log"])):
        assert torch.equal(g["z"], w["z"].reshape(g["z"].shape)), f"stochastic state differs after call {i}"
        assert float((g["h"] - w["h"]).abs().max()) <= 1e-5, f"recurrent state differs after call {i}"
        if cont:
            assert float((g["a"] - w["a"]).abs().max()) <= 1e-5, f"actions differ after call {i}"
        else:
            assert torch.equal(g["a"], w["a"]), f"actions differ after call {i}"


PLAYER_FIXTURES = ["dv3_player_discrete", "dv3_player_continuous", "dv3_player_vector", "dv3_player_vector_only"]


@pytest.mark.parametrize("name", PLAYER_FIXTURES)
def test_player_matches_reference(name):
    fx, got, cont = run_player(name)
    check(fx, got, cont)


def test_player_shares_the_trainer_parameters():
    """the acting engine adopts the trainer's flat groups: an in-place parameter change is seen by the next action"""
    from oracle.ops_emul import EmulOps
    from sheeprl_b200.algos.dreamer_v3.player import PlayerDV3
    from sheeprl_b200.engine import DV3Engine

    tf, cfg = load_fixture("dv3_tiny_a")
    eng = DV3Engine(cfg, tf["actions_dim"], in_channels=3, device="cpu", ops=EmulOps())
    player = PlayerDV3(eng, 2)
    assert player.eng.wm is eng.wm and player.eng.actor is eng.actor
    assert player.eng.wm.flat.data_ptr() == eng.wm.flat.data_ptr()
