"""Replay buffers on a GPU-less host: the numpy oracle and the product's host-side index plans (with the torch test
double standing in for the two byte-moving kernels) must reproduce what the EXECUTED REFERENCE sampled
(tests/golden/buffers.npz, oracle/make_golden_buffers.py) bit for bit."""
import os

import numpy as np
import pytest
import torch

from oracle import buffers_oracle as BO
from oracle.ops_emul import EmulOps
from sheeprl_b200.data import buffers as RB
from tests.buffer_scenarios import SCENARIOS, run_scenario

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "buffers.npz"))


def _check(name, got):
    keys = [k for k in GOLDEN.files if k.startswith(name + "/")]
    assert keys and sorted(keys) == sorted(got), (sorted(keys)[:5], sorted(got)[:5])
    for k in keys:
        want, have = GOLDEN[k], got[k]
        assert want.shape == have.shape and want.dtype == have.dtype, (k, want.shape, have.shape, want.dtype, have.dtype)
        assert np.array_equal(want, have), k


def make_oracle(cls, kw):
    if cls == "EnvIndependentReplayBuffer":
        return _OracleAdapter(BO.EnvIndependentOracle(kw["buffer_size"], kw["n_envs"],
                                                      sequential=kw["buffer_cls"] == "SequentialReplayBuffer"))
    return _OracleAdapter(BO.RingOracle(kw["buffer_size"], kw["n_envs"], sequential=cls == "SequentialReplayBuffer"))


class _OracleAdapter:
    def __init__(self, o):
        self.o = o

    @property
    def _buf(self):
        return self.o._buf

    @property
    def _rng(self):
        return self.o._rng

    @_rng.setter
    def _rng(self, v):
        self.o._rng = v

    def add(self, data, indices=None):
        self.o.add(data, indices) if indices is not None else self.o.add(data)

    def sample(self, **kw):
        r = self.o.sample(**kw)
        return r[0] if isinstance(r, tuple) else r


def make_product(cls, kw, device="cpu", ops=None):
    kw = dict(kw)
    if "buffer_cls" in kw:
        kw["buffer_cls"] = getattr(RB, kw["buffer_cls"])
    return getattr(RB, cls)(device=device, ops=ops or EmulOps(), **kw)


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_oracle_matches_reference(name):
    _check(name, run_scenario(name, make_oracle))


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_host_plans_match_reference(name):
    _check(name, run_scenario(name, make_product))


def test_sequences_never_straddle_the_write_head():
    """the reference's invariant (tests/test_data/test_sequential_buffer.py:83-105) at every head position"""
    size, T = 37, 9
    for pos_adds in range(size + 1, 2 * size + 1):
        rb = RB.SequentialReplayBuffer(size, 1, device="cpu", ops=EmulOps())
        rb._rng = np.random.default_rng(pos_adds)
        rb.add({"t": np.arange(pos_adds, dtype=np.int64).reshape(-1, 1, 1)})
        s = rb.sample(64, sequence_length=T)["t"][0, :, :, 0]          # [T, B] of global step numbers
        assert (np.diff(s, axis=0) == 1).all(), pos_adds               # consecutive steps => no straddle
        assert s.min() >= pos_adds - size and s.max() < pos_adds


def test_errors_follow_the_reference():
    with pytest.raises(ValueError):
        RB.ReplayBuffer(-1, device="cpu")
    with pytest.raises(ValueError):
        RB.SequentialReplayBuffer(1, -1, device="cpu")
    with pytest.raises(ValueError):
        RB.EnvIndependentReplayBuffer(0, device="cpu")
    with pytest.raises(ValueError):
        RB.ReplayBuffer(4, memmap=True, memmap_dir=None, device="cpu")
    with pytest.raises(ValueError):
        RB.ReplayBuffer(4, memmap=True, memmap_dir="/tmp/x", memmap_mode="r", device="cpu")
    rb = RB.SequentialReplayBuffer(10, 1, device="cpu", ops=EmulOps())
    with pytest.raises(ValueError):
        rb.sample(1, sequence_length=2)                                 # nothing added
    rb.add({"a": np.zeros((3, 1, 1))})
    with pytest.raises(ValueError):
        rb.sample(1, sequence_length=4)                                 # longer than the data added so far
    with pytest.raises(ValueError):
        rb.sample(0)
    rb.add({"a": np.zeros((10, 1, 1))})
    with pytest.raises(ValueError):
        rb.sample(1, sequence_length=11)                                # longer than the ring
    with pytest.raises(RuntimeError):
        rb.add({"a": np.zeros((3,))}, validate_args=True)
    with pytest.raises(RuntimeError):
        rb.add({"a": np.zeros((3, 1, 1)), "b": np.zeros((2, 1, 1))}, validate_args=True)
    with pytest.raises(ValueError):
        rb.add([1, 2], validate_args=True)
    u = RB.ReplayBuffer(5, 1, device="cpu", ops=EmulOps())
    u.add({"observations": np.zeros((1, 1, 2))})
    with pytest.raises(RuntimeError):
        u.sample(1, sample_next_obs=True)                               # a single row has no successor
    with pytest.raises(TypeError):
        u[0]
    e = RB.EnvIndependentReplayBuffer(5, 2, device="cpu", ops=EmulOps())
    with pytest.raises(ValueError):
        e.add({"a": np.zeros((1, 2, 1))}, indices=[0])


def test_dtypes_and_item_access():
    rb = RB.ReplayBuffer(6, 2, device="cpu", ops=EmulOps())
    rb.add({"observations": np.ones((2, 2, 3), dtype=np.uint8), "rewards": np.zeros((2, 2, 1))})
    assert rb["observations"].dtype == torch.uint8 and rb["rewards"].dtype == torch.float64   # dtype of the first add
    rb["rewards"] = np.full((6, 2, 1), 7.0)
    t = rb.sample_tensors(4, dtype=torch.float32, n_samples=2)
    assert t["rewards"].dtype == torch.float32 and t["rewards"].shape == (2, 4, 1) and float(t["rewards"].min()) == 7.0
    assert len(rb) == 6 and not rb.full and not rb.empty and rb.n_envs == 2


def test_memmap_spill_and_restore(tmp_path):
    """§8f-2: a device ring written with to_memmap() is (a) readable as the reference's raw [size, n_envs, ...] memmap
    files and (b) restored bit-for-bit, write head included, so sampling continues identically."""
    from tests.buffer_scenarios import synth_rows

    def make():
        rb = RB.EnvIndependentReplayBuffer(16, 3, buffer_cls=RB.SequentialReplayBuffer, device="cpu", ops=EmulOps())
        return rb

    a = make()
    a.add(synth_rows(11, 3, 1))
    a.add(synth_rows(9, 2, 2), indices=[0, 2])                     # wraps two of the rings
    a.to_memmap(tmp_path)
    raw = np.memmap(tmp_path / "env_2" / "observations.memmap", dtype=np.uint8, mode="r", shape=(16, 1, 3, 4, 4))
    assert np.array_equal(raw, a.buffer[2]["observations"].numpy())
    b = make()
    b.add(synth_rows(1, 3, 3))                                      # allocates storage; contents are overwritten
    b.load_memmap(tmp_path)
    assert [r._pos for r in b.buffer] == [r._pos for r in a.buffer] and b.full == a.full
    for ra, rb_ in zip(a.buffer, b.buffer):
        for k in ra.buffer:
            assert torch.equal(ra[k], rb_[k]), k
    from tests.buffer_scenarios import seed_rngs
    seed_rngs(a, 5), seed_rngs(b, 5)
    sa, sb = a.sample(8, sequence_length=4, n_samples=2), b.sample(8, sequence_length=4, n_samples=2)
    for k in sa:
        assert np.array_equal(sa[k], sb[k])
