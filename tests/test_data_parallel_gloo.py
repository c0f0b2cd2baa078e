"""N>1 path on CPU: world_size-2 gloo processes drive the engine (with the op specification as test double)
through `sheeprl_b200.parallel.attach_data_parallel`.  Checks the reference's DDP semantics: gradients averaged
over ranks before clip+Adam (dreamer_v3.py:191,298,318), Moments computed on the all-gathered lambda values
(dreamer_v3/utils.py:57), replicas bit-identical afterwards."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    from oracle.ops_emul import EmulOps
    from sheeprl_b200.engine import DV3Engine
    from sheeprl_b200.parallel import attach_data_parallel, init_process_group_from_env
    from tests.helpers import load_fixture

    init_process_group_from_env("gloo")
    fx, cfg = load_fixture("dv3_tiny_a")
    eng = DV3Engine(cfg, fx["actions_dim"], in_channels=3, device="cpu", ops=EmulOps())
    eng.wm.load(fx["init"]["wm"]), eng.actor.load(fx["init"]["actor"]), eng.critic.load(fx["init"]["critic"])
    eng.target.load(fx["init"]["target"])
    attach_data_parallel(eng)
    data = {k: v.clone().float() for k, v in fx["data"][rank].items()}       # a different batch per rank
    eng.train_step(data, fx["noise"][rank])
    out[rank] = {"wm": eng.wm.flat.clone(), "actor": eng.actor.flat.clone(), "critic": eng.critic.flat.clone(),
                 "moments": eng.moments_state.clone(), "wm_grad": eng.wm.grad.clone()}
    dist.destroy_process_group()


def test_two_rank_gloo_step_keeps_replicas_identical_and_averages_gradients():
    mp.set_start_method("spawn", force=True)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    a, b = out[0], out[1]
    for k in ("wm", "actor", "critic", "moments", "wm_grad"):
        assert torch.equal(a[k], b[k]), k            # replicas stay bit-identical
    # the averaged world-model gradient equals the mean of the two single-rank gradients
    sys.path.insert(0, ROOT)
    from oracle.ops_emul import EmulOps
    from sheeprl_b200.engine import DV3Engine
    from tests.helpers import load_fixture

    fx, cfg = load_fixture("dv3_tiny_a")
    grads = []
    for r in range(2):
        eng = DV3Engine(cfg, fx["actions_dim"], in_channels=3, device="cpu", ops=EmulOps())
        eng.wm.load(fx["init"]["wm"]), eng.actor.load(fx["init"]["actor"]), eng.critic.load(fx["init"]["critic"])
        eng.target.load(fx["init"]["target"])
        eng.train_step({k: v.clone().float() for k, v in fx["data"][r].items()}, fx["noise"][r])
        grads.append(eng.wm.grad.clone())
    want = 0.5 * (grads[0] + grads[1])
    assert float((a["wm_grad"] - want).abs().max()) <= 1e-5 * float(want.abs().max())


def _worker_sac_ppo(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    from sheeprl_b200.parallel import attach_data_parallel, init_process_group_from_env
    from tests.test_ppo_cpu import load as load_ppo, make_engine as make_ppo
    from tests.test_sac_cpu import load as load_sac, make_engine as make_sac

    init_process_group_from_env("gloo")
    # SAC: each rank gets a different batch of the fixture (sac.py trains DDP-wrapped actor / critics, :68-73 reduces alpha)
    fx = load_sac("sac_tiny")
    eng = make_sac(fx)
    attach_data_parallel(eng)
    st = fx["steps"][rank]
    eng.train_step(dict(st["data"]), True, {"eps_next": st["eps_next"], "eps_cur": st["eps_cur"]})
    res = {"sac_actor": eng.actor.flat.clone(), "sac_qf": eng.qf.flat.clone(), "sac_alpha": eng.alpha.flat.clone(),
           "sac_qf_grad": eng.qf.grad.clone()}
    # PPO: ranks train on different minibatches of the rollout
    fp = load_ppo("ppo_branches")
    pe = make_ppo(fp)
    attach_data_parallel(pe)
    pe.train(fp["data"], [fp["index_batches"][rank]])
    res["ppo"] = pe.group.flat.clone()
    res["ppo_grad"] = pe.group.grad.clone()
    out[rank] = res
    dist.destroy_process_group()


def test_two_rank_gloo_sac_and_ppo_average_gradients():
    """SAC / PPO engines through the same data-parallel hook: replicas identical, gradient = mean over ranks"""
    mp.set_start_method("spawn", force=True)
    out = mp.Manager().dict()
    port = 30100 + (os.getpid() % 500)
    mp.spawn(_worker_sac_ppo, args=(2, port, out), nprocs=2, join=True)
    a, b = out[0], out[1]
    for k in a.keys():
        assert torch.equal(a[k], b[k]), k
    sys.path.insert(0, ROOT)
    from tests.test_ppo_cpu import load as load_ppo, make_engine as make_ppo
    from tests.test_sac_cpu import load as load_sac, make_engine as make_sac

    fx, grads = load_sac("sac_tiny"), []
    for r in range(2):
        eng = make_sac(fx)
        st = fx["steps"][r]
        eng.train_step(dict(st["data"]), True, {"eps_next": st["eps_next"], "eps_cur": st["eps_cur"]})
        grads.append(eng.qf.grad.clone())
    # the critic gradient is computed before any parameter changes: the DP gradient is the mean of the two
    want = 0.5 * (grads[0] + grads[1])
    assert float((a["sac_qf_grad"] - want).abs().max()) <= 1e-5 * float(want.abs().max())
    fp, pg = load_ppo("ppo_branches"), []
    for r in range(2):
        pe = make_ppo(fp)
        pe.train(fp["data"], [fp["index_batches"][r]])
        pg.append(pe.group.grad.clone())
    want = 0.5 * (pg[0] + pg[1])
    assert float((a["ppo_grad"] - want).abs().max()) <= 1e-5 * float(want.abs().max())


def _worker_p2e(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    from sheeprl_b200.parallel import attach_data_parallel, init_process_group_from_env
    from tests.test_p2e_cpu import load, make_engine

    init_process_group_from_env("gloo")
    fx, cfg = load()
    eng = make_engine(fx, cfg)
    attach_data_parallel(eng)
    eng.train_step({k: v.clone().float() for k, v in fx["data"][rank].items()}, fx["noise"][rank])   # a different batch per rank
    res = {n: g.flat.clone() for n, g in eng.groups().items()}
    res["ens_last"], res["ens_rest"] = eng.ens_last.flat.clone(), eng.ens_rest.flat.clone()
    res["moments_intrinsic"] = eng.critics_expl["intrinsic"]["moments_state"].clone()
    out[rank] = res
    dist.destroy_process_group()


def test_two_rank_gloo_plan2explore_replicas_stay_identical():
    """every optimiser group of the Plan2Explore engine (ensembles, exploration actor / critics, task actor / critic,
    world model) goes through the same all-reduce hook; the exploration critics' Moments see the gathered lambda-values"""
    mp.set_start_method("spawn", force=True)
    out = mp.Manager().dict()
    mp.spawn(_worker_p2e, args=(2, 30700 + (os.getpid() % 500), out), nprocs=2, join=True)
    a, b = out[0], out[1]
    for k in a.keys():
        assert torch.equal(a[k], b[k]), k
    sys.path.insert(0, ROOT)
    from tests.test_p2e_cpu import load, make_engine

    fx, cfg = load()
    solo = make_engine(fx, cfg)
    solo.train_step({k: v.clone().float() for k, v in fx["data"][0].items()}, fx["noise"][0])
    assert not torch.equal(solo.actor_expl.flat, a["actor_expl"])        # the other rank's batch did contribute


def _worker_public(rank, world, port, out):
    """only the public surface: build_agent(fabric with world_size 2) + train(); no explicit attach call"""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    from oracle.ops_emul import EmulOps
    from sheeprl_b200.algos.dreamer_v3.agent import build_agent
    from sheeprl_b200.algos.dreamer_v3.dreamer_v3 import make_optimizers, train
    from sheeprl_b200.algos.dreamer_v3.utils import Moments
    from sheeprl_b200.parallel import init_process_group_from_env
    from tests.helpers import load_fixture

    init_process_group_from_env("gloo")            # what fabric.launch does before calling the entry point

    class Fab:
        device, world_size, global_rank = torch.device("cpu"), world, rank

    class Space:
        shape = (3, 64, 64)

    fx, cfg = load_fixture("dv3_tiny_a")
    wm, actor, critic, target, _ = build_agent(Fab, fx["actions_dim"], False, cfg, {"rgb": Space}, fx["init"]["wm"],
                                               fx["init"]["actor"], fx["init"]["critic"], fx["init"]["target"], ops=EmulOps())
    eng = wm._b200_engine
    assert eng.world_size == world and eng.allreduce is not None and eng.allgather is not None
    opts = make_optimizers(eng, cfg)
    mo = cfg.algo.actor.moments
    moments = Moments(mo.decay, mo.max, mo.percentile.low, mo.percentile.high)
    data = {k: v.clone().float() for k, v in fx["data"][rank].items()}
    train(Fab, wm, actor, critic, target, *opts, data, None, cfg, False, fx["actions_dim"], moments, noise=fx["noise"][rank])
    out[rank] = {"wm": eng.wm.flat.clone(), "actor": eng.actor.flat.clone(), "critic": eng.critic.flat.clone(),
                 "moments": torch.stack([moments.low.clone(), moments.high.clone()]), "seed": eng.rng_seed}
    dist.destroy_process_group()


def test_two_rank_gloo_through_build_agent_only():
    """the reference gets data parallelism from `fabric.setup_module` inside build_agent (agent.py:1205-1214); here
    `build_agent` installs the all-reduce / all-gather hooks itself when fabric.world_size > 1"""
    mp.set_start_method("spawn", force=True)
    out = mp.Manager().dict()
    mp.spawn(_worker_public, args=(2, 31300 + (os.getpid() % 500), out), nprocs=2, join=True)
    a, b = out[0], out[1]
    for k in ("wm", "actor", "critic", "moments"):
        assert torch.equal(a[k], b[k]), k
    assert a["seed"] != b["seed"]                   # ranks draw different sampling-noise streams
