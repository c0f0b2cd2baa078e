"""Checkpoint / resume parity (reference: dreamer_v3.py:736-763 state layout): the objects returned by
build_agent() / make_optimizers() expose the reference's state-dict layouts; saving them after step 1 and loading them
into a fresh agent reproduces step 2 of an uninterrupted run bit for bit, and the optimiser state is interchangeable
with torch.optim.Adam's."""
import copy

import pytest
import torch

from oracle.ops_emul import EmulOps
from sheeprl_b200.algos.dreamer_v3.dreamer_v3 import B200Adam, make_optimizers
from sheeprl_b200.engine import DV3Engine
from tests.helpers import load_fixture


def fresh(cfg, adim, cont):
    eng = DV3Engine(cfg, adim, in_channels=3, device="cpu", ops=EmulOps(), is_continuous=cont)
    return eng, make_optimizers(eng, cfg)


@pytest.mark.parametrize("name", ["dv3_tiny_a", "dv3_tiny_c"])
def test_resume_reproduces_uninterrupted_run(name):
    fx, cfg = load_fixture(name)
    adim, cont = fx["actions_dim"], fx.get("is_continuous", False)
    batches = [{k: v.clone().float() for k, v in d.items()} for d in fx["data"]]
    # uninterrupted: two steps
    e1, _ = fresh(cfg, adim, cont)
    for g, n in ((e1.wm, "wm"), (e1.actor, "actor"), (e1.critic, "critic"), (e1.target, "target")):
        g.load(fx["init"][n])
    e1.train_step(copy.deepcopy(batches[0]), fx["noise"][0])
    # checkpoint after step 1 (reference layout: model state dicts + optimiser state dicts + moments)
    _, opts1 = None, make_optimizers(e1, cfg)
    ckpt = {"world_model": e1.wm.state_dict(), "actor": e1.actor.state_dict(), "critic": e1.critic.state_dict(),
            "target_critic": e1.target.state_dict(), "moments": e1.moments_state.clone(),
            "world_optimizer": opts1[0].state_dict(), "actor_optimizer": opts1[1].state_dict(),
            "critic_optimizer": opts1[2].state_dict()}
    e1.train_step(copy.deepcopy(batches[1]), fx["noise"][1])
    # resumed: fresh agent, load, one step
    e2, opts2 = fresh(cfg, adim, cont)
    e2.wm.load(ckpt["world_model"]), e2.actor.load(ckpt["actor"]), e2.critic.load(ckpt["critic"])
    e2.target.load(ckpt["target_critic"])
    e2.moments_state.copy_(ckpt["moments"])
    for o, k in zip(opts2, ("world_optimizer", "actor_optimizer", "critic_optimizer")):
        o.load_state_dict(ckpt[k])
    e2.train_step(copy.deepcopy(batches[1]), fx["noise"][1])
    for ga, gb in ((e1.wm, e2.wm), (e1.actor, e2.actor), (e1.critic, e2.critic)):
        assert torch.equal(ga.flat, gb.flat) and torch.equal(ga.exp_avg, gb.exp_avg) and ga.step == gb.step == 2
    assert torch.equal(e1.metrics, e2.metrics)


def test_optimizer_state_interchanges_with_torch_adam():
    fx, cfg = load_fixture("dv3_tiny_a")
    eng, opts = fresh(cfg, fx["actions_dim"], False)
    for g, n in ((eng.wm, "wm"), (eng.actor, "actor"), (eng.critic, "critic"), (eng.target, "target")):
        g.load(fx["init"][n])
    eng.train_step({k: v.clone().float() for k, v in fx["data"][0].items()}, fx["noise"][0])
    sd = opts[1].state_dict()                                           # actor optimiser, torch layout
    params = [torch.nn.Parameter(v.clone()) for v in eng.actor.views.values()]
    ref = torch.optim.Adam(params, lr=cfg.algo.actor.optimizer.lr, eps=cfg.algo.actor.optimizer.eps)
    ref.load_state_dict(sd)                                             # torch accepts it as its own
    back = ref.state_dict()
    assert set(back["state"]) == set(sd["state"]) and float(back["state"][0]["step"]) == 1.0
    for i in sd["state"]:
        assert torch.equal(back["state"][i]["exp_avg"], sd["state"][i]["exp_avg"])
    o2 = B200Adam(eng.actor, list(eng.actor.shapes), 1e-4, 1e-8)
    o2.load_state_dict(back)                                            # and torch's own dict loads back
    assert eng.actor.step == 1
