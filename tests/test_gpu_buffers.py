"""Device replay buffers through the C-ABI kernels (b200rl_replay_gather / b200rl_replay_scatter): bit-exact
against what the executed reference sampled (tests/golden/buffers.npz) and, at the Dreamer-V3 batch size, against
the numpy oracle on the same seeded rows."""
import numpy as np
import pytest
import torch

from tests.buffer_scenarios import SCENARIOS, run_scenario, seed_rngs
from tests.test_buffers_cpu import _check, make_product

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    from sheeprl_b200.lib import CudaOps

    return CudaOps()


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_device_buffers_match_reference(cu, name):
    got = run_scenario(name, lambda cls, kw: make_product(cls, kw, device="cuda", ops=cu))
    _check(name, got)


def test_dreamer_batch_against_oracle(cu):
    """EnvIndependentReplayBuffer(SequentialReplayBuffer), 4 envs x 2048 rows of 64x64x3 uint8, B=16, T=64 (the
    BASELINE batch): same bytes as the numpy oracle, output already in [n_samples, T, B, ...]."""
    from oracle.buffers_oracle import EnvIndependentOracle
    from sheeprl_b200.data.buffers import EnvIndependentReplayBuffer, SequentialReplayBuffer

    size, n_envs, T, B = 2048, 4, 64, 16
    rb = EnvIndependentReplayBuffer(size, n_envs, buffer_cls=SequentialReplayBuffer, device="cuda", ops=cu)
    orc = EnvIndependentOracle(size, n_envs, sequential=True)
    seed_rngs(rb, 99)
    seed_rngs(orc, 99)
    g = np.random.default_rng(3)
    for chunk in (1500, 900):                                     # second add wraps every ring
        data = {"rgb": g.integers(0, 256, size=(chunk, n_envs, 3, 64, 64), dtype=np.uint8),
                "rewards": g.standard_normal((chunk, n_envs, 1)),
                "is_first": (g.random((chunk, n_envs, 1)) < 0.02).astype(np.float64)}
        rb.add(data)
        orc.add(data)
    for n_samples in (1, 2):
        got = rb.sample_tensors(B, n_samples=n_samples, sequence_length=T)
        want = orc.sample(B, n_samples=n_samples, sequence_length=T)
        assert got["rgb"].shape == (n_samples, T, B, 3, 64, 64) and got["rgb"].dtype == torch.uint8
        for k in want:
            assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    assert rb.full == (True,) * n_envs


def test_buffer_feeds_train(cu):
    """sample_tensors -> train(): the uint8 image goes straight into the engine (obs_prep kernel), no host hop."""
    from sheeprl_b200.algos.dreamer_v3.agent import build_agent
    from sheeprl_b200.algos.dreamer_v3.dreamer_v3 import make_optimizers, train
    from sheeprl_b200.algos.dreamer_v3.utils import Moments
    from sheeprl_b200.configs import make_dv3_cfg
    from sheeprl_b200.data.buffers import EnvIndependentReplayBuffer, SequentialReplayBuffer

    cfg = make_dv3_cfg("S", per_rank_batch_size=4, per_rank_sequence_length=8, horizon=3, dense_units=32, mlp_layers=1,
                       cnn_channels_multiplier=2, recurrent_state_size=32, hidden_size=32, stochastic_size=4,
                       discrete_size=4, bins=31)
    dev = torch.device("cuda")

    class Fab:
        device, world_size, global_rank = dev, 1, 0

    class Space:
        shape = (3, 64, 64)

    wm, actor, critic, target, _ = build_agent(Fab, (3,), False, cfg, {"rgb": Space})
    eng = wm._b200_engine
    opts = make_optimizers(eng, cfg)
    mo = cfg.algo.actor.moments
    moments = Moments(mo.decay, mo.max, mo.percentile.low, mo.percentile.high)
    rb = EnvIndependentReplayBuffer(64, 2, buffer_cls=SequentialReplayBuffer, device="cuda", ops=cu)
    g = np.random.default_rng(0)
    rb.add({"rgb": g.integers(0, 256, size=(40, 2, 3, 64, 64), dtype=np.uint8),
            "actions": np.eye(3, dtype=np.int64)[g.integers(0, 3, size=(40, 2))],
            "rewards": g.standard_normal((40, 2, 1)), "terminated": np.zeros((40, 2, 1)),
            "truncated": np.zeros((40, 2, 1)), "is_first": np.zeros((40, 2, 1))})
    local = rb.sample_tensors(4, sequence_length=8, n_samples=2)
    for i in range(2):
        batch = {k: (v[i] if v.dtype == torch.uint8 else v[i].float()) for k, v in local.items()}
        train(Fab, wm, actor, critic, target, *opts, batch, None, cfg, False, (3,), moments)
    assert torch.isfinite(eng.metrics).all()


@pytest.mark.parametrize("kind", ["uniform", "sequential", "env_independent"])
def test_device_ring_pickles_and_returns_to_the_gpu(cu, kind, tmp_path):
    """checkpoint round trip of an HBM-resident ring (host tensors in the file, storage back on cuda:0 afterwards) and
    the checkpoint callback's truncated fix-up on device storage"""
    from sheeprl_b200.utils.callback import CheckpointCallback, load_replay_buffer
    from tests.test_buffer_checkpoint_cpu import _Fabric, _heads, _make, _same_samples

    a, kw = _make(kind, device="cuda", ops=cu)
    CheckpointCallback().on_checkpoint_coupled(_Fabric(), str(tmp_path / "c.ckpt"), {"iter_num": 1}, a)
    b = load_replay_buffer(torch.load(tmp_path / "c.ckpt", weights_only=False)["rb"], ops=cu)
    rings = b.buffer if kind == "env_independent" else [b]
    assert all(t.is_cuda for r in rings for t in r.buffer.values()) and _heads(a) == _heads(b)
    for r in rings:                                       # undo the checkpoint's truncation mark, then compare draws
        r["truncated"][(r._pos - 1) % r.buffer_size] = 0
    _same_samples(a, b, kw)
