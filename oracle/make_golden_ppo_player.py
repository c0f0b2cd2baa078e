"""TEST INFRASTRUCTURE — writes tests/golden/ppo_player.pt by EXECUTING THE REAL REFERENCE PPOPlayer (container only):

    python -m oracle.make_golden_ppo_player

For the agents of the train fixtures (ppo_branches: two discrete heads; ppo_continuous: Normal; ppo_pixel: NatureCNN +
vector key; ppo_tanh_ln: tanh_normal + LayerNorm MLPs; ppo_multikey: two image + two vector keys) with their initial weights: forward(obs) with injected sampling noise -> (actions, logprobs, values),
get_values(obs), get_actions(obs, greedy=True).
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402
from oracle.make_golden_ppo import obs_space, ppo_cfg, split_obs  # noqa: E402

E = 5


def run(name):
    import sheeprl.algos.ppo.agent as A

    A.get_single_device_fabric = lambda f: f
    fx = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), weights_only=False)
    spec = fx["spec"]
    cfg = ppo_cfg(spec, fx["hp"], fx["batch"], fx["epochs"])
    space = obs_space(spec)
    agent, player = A.build_agent(H.FakeFabric(), spec["actions_dim"], spec["is_continuous"], cfg, space, fx["init"])
    g = torch.Generator().manual_seed(123)
    obs = {}
    if spec["cnn_channels"]:
        obs["rgb"] = torch.randint(0, 256, (E, spec["cnn_channels"], spec["screen"], spec["screen"]), generator=g).float() / 255 - 0.5
    if spec["mlp_dim"]:
        obs["state"] = torch.randn(E, spec["mlp_dim"], generator=g)
    A_tot = sum(spec["actions_dim"])
    cat_obs, obs = obs, split_obs(spec, obs)                 # the reference reads one tensor per key
    sampled = None
    if spec["is_continuous"]:
        noise = torch.randn(E, A_tot, generator=g)
        orig = torch.normal
        torch.normal = lambda loc, scale, **kw: loc + scale * noise.reshape(loc.shape)
        try:
            with torch.no_grad():
                actions, logp, values = player(obs)
                sampled = [a.clone() for a in player.get_actions(obs, greedy=False)]
        finally:
            torch.normal = orig
    else:
        noise = torch.empty(E, A_tot).exponential_(1.0, generator=g)
        q, off = [], 0
        for ad in spec["actions_dim"]:
            q.append(noise[:, off:off + ad])
            off += ad
        with H.NoiseQueue(q), torch.no_grad():
            actions, logp, values = player(obs)
    with torch.no_grad():
        vals2 = player.get_values(obs)
        greedy = player.get_actions(obs, greedy=True)
    return {"train_fixture": name, "obs": cat_obs, "sampled": sampled, "noise": noise, "actions": [a.clone() for a in actions], "logp": logp.clone(),
            "values": values.clone(), "values2": vals2.clone(), "greedy": [a.clone() for a in greedy]}


def main():
    H.install()
    out = {n: run(n) for n in ("ppo_branches", "ppo_continuous", "ppo_pixel", "ppo_tanh_ln", "ppo_multikey")}
    path = os.path.join(ROOT, "tests", "golden", "ppo_player.pt")
    torch.save(out, path)
    print(os.path.getsize(path), {n: out[n]["logp"].flatten().tolist()[:3] for n in out})


if __name__ == "__main__":
    main()
