"""The registered entry point `sheeprl_b200.algos.dreamer_v3.dreamer_v3.main` (what `sheeprl exp=dreamer_v3` launches,
sheeprl/cli.py:82-98,199) drives the UNMODIFIED reference interaction loop (dreamer_v3.py:361-780) with the B200
`build_agent` / `train` / optimizer handles / Moments / replay rings substituted.  Container-only: the reference is imported
from /root/reference through the stub harness, the environment is a dummy, the kernels are the torch test double."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_harness
from oracle.ops_emul import EmulOps
from sheeprl_b200.configs import make_dv3_cfg
from tests import fake_gym

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present")


class Fabric(ref_harness.FakeFabric):
    def __init__(self, tmp):
        super().__init__("cpu")
        self.tmp, self.logged, self.checkpoints = tmp, {}, []
        self.loggers = []

    def load(self, path):
        return torch.load(path, weights_only=False)

    def log_dict(self, d, step):
        self.logged.update(d)

    def log(self, k, v, step):
        self.logged[k] = v

    def call(self, hook, **kw):
        assert hook == "on_checkpoint_coupled"
        self.checkpoints.append(kw)
        torch.save(kw["state"], kw["ckpt_path"]) if os.path.isdir(os.path.dirname(kw["ckpt_path"])) else None


def _loop_cfg(tmp, **over):
    cfg = make_dv3_cfg("S", per_rank_batch_size=2, per_rank_sequence_length=8, horizon=4, dense_units=32, mlp_layers=1,
                       cnn_channels_multiplier=4, recurrent_state_size=32, hidden_size=32, stochastic_size=4,
                       discrete_size=4, bins=15)
    d = cfg.as_dict()
    d["seed"] = 3
    d["dry_run"] = False
    d["root_dir"], d["run_name"] = str(tmp), "run"
    d["checkpoint"] = {"resume_from": None, "every": 0, "save_last": True, "keep_last": 1}
    d["metric"] = {"log_level": 0, "log_every": 1000, "sync_on_compute": False, "aggregator": {}}
    d["model_manager"] = {"disabled": True}
    d["buffer"] = {"size": 64, "memmap": False, "checkpoint": False, "validate_args": False, "from_numpy": False}
    d["env"].update({"num_envs": 1, "sync_env": True, "clip_rewards": False, "action_repeat": 1, "frame_stack": -1,
                     "wrapper": {"_target_": "tests.fake_gym.DummyImageEnv"}})
    d["algo"].update({"total_steps": 40, "learning_starts": 24, "replay_ratio": 0.25, "per_rank_pretrain_steps": 0,
                      "run_test": False, "name": "dreamer_v3"})
    d["algo"]["critic"]["per_rank_target_network_update_freq"] = 1
    d["algo"].update(over)
    ref_harness.install()
    from sheeprl.utils.utils import dotdict

    return dotdict(d)


def _harness(tmp):
    """what is absent on this host and NOT part of the drop-in: the environment factory, loggers, metric plumbing"""
    ref_harness.install()
    import sheeprl.algos.dreamer_v3.dreamer_v3 as R
    from sheeprl.utils.metric import MetricAggregator
    from sheeprl.utils.timer import timer

    R.gym = fake_gym.module()
    R.make_env = lambda cfg, seed, rank_off, log_dir, prefix, vector_env_idx=0: (lambda: fake_gym.DummyImageEnv(seed=seed))
    R.RestartOnException = lambda fn: fn()
    R.get_logger = lambda fabric, cfg: None
    R.get_log_dir = lambda fabric, root, run: os.path.join(root, run)
    R.save_configs = lambda cfg, log_dir: None
    MetricAggregator.disabled = True
    timer.disabled = True
    return R


def test_registered_main_runs_the_reference_loop_on_the_b200_engine(tmp_path):
    R = _harness(tmp_path)
    import sheeprl_b200.algos.dreamer_v3.agent as A
    import sheeprl_b200.algos.dreamer_v3.dreamer_v3 as B
    from sheeprl_b200.data import buffers as Bf
    from sheeprl_b200.utils.registry import find_algorithm

    found = find_algorithm("dreamer_v3")
    assert found is not None and found[0] == "sheeprl_b200.algos.dreamer_v3" and found[1]["entrypoint"] == "main"
    os.makedirs(tmp_path / "run" / "checkpoint", exist_ok=True)
    cfg, fab = _loop_cfg(tmp_path), Fabric(tmp_path)
    ref_train, ref_build = R.train, R.build_agent
    seen = {"train": 0, "engines": []}
    orig_train = B.train

    def counting_train(*a, **k):
        seen["train"] += 1
        seen["engines"].append(a[1]._b200_engine)
        return orig_train(*a, **k)

    A.DEFAULT_OPS, Bf.DEFAULTS["ops"] = EmulOps(), EmulOps()
    B.train = counting_train
    try:
        B.main(fab, cfg)
    finally:
        B.train = orig_train
        A.DEFAULT_OPS, Bf.DEFAULTS["ops"], Bf.DEFAULTS["device"] = None, None, "cuda"
    # the reference module is left exactly as it was
    assert R.train is ref_train and R.build_agent is ref_build
    # 40 policy steps, learning starts at 24, replay ratio 0.25 -> a handful of updates through OUR train()
    assert seen["train"] >= 3
    eng = seen["engines"][0]
    assert eng.wm.step == seen["train"] and eng.actor.step == seen["train"]
    # final checkpoint: reference key layout, optimizer handles in torch's state-dict layout
    (ck,) = fab.checkpoints
    st = ck["state"]
    assert set(st) >= {"world_model", "actor", "critic", "target_critic", "world_optimizer", "actor_optimizer",
                       "critic_optimizer", "moments", "ratio", "iter_num", "batch_size"}
    assert "rssm.recurrent_model.rnn.linear.weight" in st["world_model"]
    wo = st["world_optimizer"]
    assert len(wo["state"]) == len(st["world_model"]) and int(wo["state"][0]["step"]) == seen["train"]
    assert wo["param_groups"][0]["lr"] == pytest.approx(float(cfg.algo.world_model.optimizer.lr))
    assert all(torch.isfinite(v).all() for v in st["world_model"].values())
    assert float(st["moments"]["high"]) != 0.0 or float(st["moments"]["low"]) != 0.0


def test_main_resumes_from_its_own_checkpoint(tmp_path):
    """resume through the entry point: state dicts, optimizer moments and Moments go back into the engine's flat groups
    (dreamer_v3.py:365-366, 437-462)"""
    _harness(tmp_path)
    import sheeprl_b200.algos.dreamer_v3.agent as A
    import sheeprl_b200.algos.dreamer_v3.dreamer_v3 as B
    from sheeprl_b200.data import buffers as Bf

    os.makedirs(tmp_path / "run" / "checkpoint", exist_ok=True)
    A.DEFAULT_OPS, Bf.DEFAULTS["ops"] = EmulOps(), EmulOps()
    try:
        fab = Fabric(tmp_path)
        B.main(fab, _loop_cfg(tmp_path))
        path = fab.checkpoints[0]["ckpt_path"]
        first = torch.load(path, weights_only=False)
        cfg2 = _loop_cfg(tmp_path, total_steps=100)   # the reference shifts learning_starts by the resumed iteration and keeps the Ratio state
        cfg2.checkpoint.resume_from = path
        fab2 = Fabric(tmp_path)
        B.main(fab2, cfg2)
    finally:
        A.DEFAULT_OPS, Bf.DEFAULTS["ops"], Bf.DEFAULTS["device"] = None, None, "cuda"
    second = fab2.checkpoints[0]["state"]
    n0, n1 = int(first["world_optimizer"]["state"][0]["step"]), int(second["world_optimizer"]["state"][0]["step"])
    assert n1 > n0 >= 3, (n0, n1)                                # the Adam step count continued from the checkpoint
    assert second["iter_num"] > first["iter_num"]
