"""TEST INFRASTRUCTURE — writes tests/golden/dv3_player_*.pt by EXECUTING THE REAL REFERENCE PlayerDV3 (container only):

    python -m oracle.make_golden_player

Weights: the initial state of the matching train fixture (dv3_tiny_a discrete / dv3_tiny_c continuous / dv3_tiny_v image +
two vector keys / dv3_tiny_vo vector only; `obs` is a dict per observation key for those two).  Script: init_states(),
three env steps, a partial reset of env 1, two more steps; every categorical / Normal draw consumes injected noise.
Stored: the normalised observations, the noise, and after every call the player's actions / recurrent_state /
stochastic_state.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402
from oracle import ref_run  # noqa: E402
from oracle.dv3_oracle import vec_dims  # noqa: E402
from tests.helpers import load_fixture  # noqa: E402

NUM_ENVS, STEPS, RESET_AT, RESET_ENVS = 3, 5, 3, [1]


def run(train_fixture: str):
    import torch.distributions.normal as TN

    fx, cfg = load_fixture(train_fixture)
    cont = fx.get("is_continuous", False)
    adim = fx["actions_dim"]
    cfg.env.num_envs = NUM_ENVS
    fab, rcfg, wm, actor, critic, target, player = ref_run.build_reference_agent(cfg, adim, is_continuous=cont)
    for mod, name in ((wm, "wm"), (actor, "actor")):
        ref_run._load(mod, fx["init"][name])
    w = cfg.algo.world_model
    S, D, A = w.stochastic_size, w.discrete_size, int(sum(adim))
    g = torch.Generator().manual_seed(77)
    player.num_envs = NUM_ENVS
    player.init_states()
    log = [{"h": player.recurrent_state.clone(), "z": player.stochastic_state.clone(), "a": player.actions.clone()}]
    obs_l, nz_l, na_l = [], [], []
    for s in range(STEPS):
        if s == RESET_AT:
            player.init_states(RESET_ENVS)
        obs = {}
        if cfg.algo.cnn_keys.encoder:
            obs[cfg.algo.cnn_keys.encoder[0]] = torch.randint(
                0, 256, (1, NUM_ENVS, 3, cfg.env.screen_size, cfg.env.screen_size), generator=g).float() / 255 - 0.5
        nz = torch.empty(NUM_ENVS, S, D).exponential_(1.0, generator=g)
        na = torch.randn(NUM_ENVS, A, generator=g) if cont else torch.empty(NUM_ENVS, A).exponential_(1.0, generator=g)
        cat = [nz.reshape(-1, D)]
        if not cont:
            off = 0
            for ad in adim:
                cat.append(na[:, off:off + ad])
                off += ad
        for k, d in vec_dims(cfg).items():                   # vector observations (drawn after the noise: older fixtures keep their streams)
            obs[k] = torch.randn(1, NUM_ENVS, d, generator=g) * 3.0
        normal = [na] if cont else []
        orig = TN._standard_normal
        TN._standard_normal = lambda shape, dtype, device: normal.pop(0).reshape(shape)
        try:
            with H.NoiseQueue(cat):
                player.get_actions(obs)
        finally:
            TN._standard_normal = orig
        obs_l.append(obs), nz_l.append(nz), na_l.append(na)
        log.append({"h": player.recurrent_state.clone(), "z": player.stochastic_state.reshape(1, NUM_ENVS, -1).clone(),
                    "a": player.actions.clone()})
    return {"train_fixture": train_fixture, "num_envs": NUM_ENVS, "reset_at": RESET_AT, "reset_envs": RESET_ENVS,
            "obs": obs_l, "noise_z": nz_l, "noise_a": na_l, "log": log}


def main():
    H.install()
    only = sys.argv[1:]
    for name, src in (("dv3_player_discrete", "dv3_tiny_a"), ("dv3_player_continuous", "dv3_tiny_c"),
                      ("dv3_player_vector", "dv3_tiny_v"), ("dv3_player_vector_only", "dv3_tiny_vo")):
        if only and name not in only:
            continue
        out = run(src)
        path = os.path.join(ROOT, "tests", "golden", name + ".pt")
        torch.save(out, path)
        print(name, os.path.getsize(path), out["log"][-1]["a"].flatten().tolist())


if __name__ == "__main__":
    main()
