"""Scripted add/sample scenarios for the replay buffers, shared by
  * oracle/make_golden_buffers.py  (drives the EXECUTED REFERENCE, writes tests/golden/buffers.npz),
  * tests/test_buffers_cpu.py      (numpy oracle and host index plans vs the golden file),
  * tests/test_gpu_buffers.py      (device buffers through the C-ABI kernels vs the golden file).

A scenario = constructor kwargs + a list of ops.  ``("add", length, seed[, indices])`` adds seeded synthetic rows,
``("sample", kwargs)`` samples.  Edge cases follow the reference's own buffer tests (SURVEY §8c:
tests/test_data/test_sequential_buffer.py:21-44,83-127; test_env_independent_rb.py; test_buffers.py:140-235):
wrap-around adds, an add longer than the ring, windows that must not straddle the write head when full (including
pos < sequence_length, where the first valid range is empty), the not-full upper bound, next-observation rows.
"""
from __future__ import annotations

import numpy as np

SEED = 20240917


def synth_rows(length: int, n_envs: int, seed: int):
    """Dreamer-like rows: uint8 image, float64 scalars (np.zeros default dtype, dreamer_v3.py:543-546), int64 one-hot
    actions (prefill, dreamer_v3.py:565-571)."""
    g = np.random.default_rng(seed)
    return {
        "observations": g.integers(0, 256, size=(length, n_envs, 3, 4, 4), dtype=np.uint8),
        "rewards": g.standard_normal((length, n_envs, 1)),
        "actions": np.eye(3, dtype=np.int64)[g.integers(0, 3, size=(length, n_envs))],
        "vec": g.standard_normal((length, n_envs, 5)).astype(np.float32),          # 20-byte rows: unaligned gather
    }


SCENARIOS = {
    # uniform rows (SAC / PPO storage), 3 envs
    "uniform": dict(
        cls="ReplayBuffer", kwargs=dict(buffer_size=20, n_envs=3), ops=[
            ("add", 7, 1), ("sample", dict(batch_size=5)), ("sample", dict(batch_size=4, sample_next_obs=True, n_samples=2)),
            ("add", 7, 2), ("add", 7, 3), ("sample", dict(batch_size=6, n_samples=3)),
            ("sample", dict(batch_size=6, sample_next_obs=True)),
            ("add", 19, 4), ("sample", dict(batch_size=9, sample_next_obs=True, n_samples=2)),
            ("add", 1, 5), ("sample", dict(batch_size=3, sample_next_obs=True)),    # pos == 0 when full
        ]),
    "uniform_long_add": dict(
        cls="ReplayBuffer", kwargs=dict(buffer_size=8, n_envs=2), ops=[
            ("add", 11, 6), ("sample", dict(batch_size=16)), ("add", 3, 7), ("sample", dict(batch_size=16, sample_next_obs=True)),
        ]),
    # Dreamer sequences, one env
    "seq_1env": dict(
        cls="SequentialReplayBuffer", kwargs=dict(buffer_size=32, n_envs=1), ops=[
            ("add", 9, 10), ("sample", dict(batch_size=4, sequence_length=5)),
            ("sample", dict(batch_size=3, sequence_length=9, n_samples=2)),          # not-full upper bound
            ("add", 20, 11), ("add", 6, 12),                                           # wraps: pos = 3 < T
            ("sample", dict(batch_size=8, sequence_length=5, n_samples=2)),           # first range empty
            ("sample", dict(batch_size=8, sequence_length=5, sample_next_obs=True)),
            ("add", 10, 13), ("sample", dict(batch_size=8, sequence_length=5, n_samples=3)),
            ("sample", dict(batch_size=2, sequence_length=32)),                       # whole ring
        ]),
    "seq_4env": dict(
        cls="SequentialReplayBuffer", kwargs=dict(buffer_size=16, n_envs=4), ops=[
            ("add", 10, 20), ("sample", dict(batch_size=6, sequence_length=4, n_samples=2)),
            ("add", 10, 21), ("sample", dict(batch_size=6, sequence_length=4, n_samples=2, sample_next_obs=True)),
            ("add", 16, 22), ("sample", dict(batch_size=5, sequence_length=7)),
        ]),
    # Dreamer-V3's buffer: EnvIndependentReplayBuffer(buffer_cls=SequentialReplayBuffer) (dreamer_v3.py:478-485)
    "envind_seq": dict(
        cls="EnvIndependentReplayBuffer", kwargs=dict(buffer_size=24, n_envs=4, buffer_cls="SequentialReplayBuffer"), ops=[
            ("add", 8, 30), ("sample", dict(batch_size=8, sequence_length=6, n_samples=2)),
            ("add", 5, 31, [1, 3]),                                                    # desynchronise the write heads
            ("sample", dict(batch_size=16, sequence_length=6)),
            ("add", 14, 32), ("add", 7, 33, [0, 2]),
            ("sample", dict(batch_size=16, sequence_length=6, n_samples=2)),
            ("sample", dict(batch_size=5, sequence_length=6, sample_next_obs=True)),
            ("sample", dict(batch_size=1, sequence_length=3)),                        # bincount shorter than n_envs
        ]),
    "envind_uniform": dict(
        cls="EnvIndependentReplayBuffer", kwargs=dict(buffer_size=12, n_envs=3, buffer_cls="ReplayBuffer"), ops=[
            ("add", 5, 40), ("sample", dict(batch_size=7)), ("add", 9, 41),
            ("sample", dict(batch_size=7, n_samples=2, sample_next_obs=True)),
        ]),
}


def seed_rngs(rb, seed: int) -> None:
    """Deterministic Generators on a buffer object of either implementation (the reference leaves them unseeded,
    buffers.py:79,589): the container's ``_rng`` from child 0, ring i's ``_rng`` from child i+1."""
    rings = rb._buf if isinstance(rb._buf, (list, tuple)) else None
    if rings is None:
        rb._rng = np.random.default_rng(seed)
        return
    kids = np.random.SeedSequence(seed).spawn(len(rings) + 1)
    rb._rng = np.random.default_rng(kids[0])
    for r, s in zip(rings, kids[1:]):
        r._rng = np.random.default_rng(s)


def run_scenario(name: str, make_buffer, to_numpy=lambda x: np.asarray(x)):
    """Drive ``make_buffer(cls_name, kwargs)`` through a scenario; returns {f"{name}/op{j}/{key}": array}."""
    sc = SCENARIOS[name]
    rb = make_buffer(sc["cls"], dict(sc["kwargs"]))
    seed_rngs(rb, SEED)
    n_envs = sc["kwargs"]["n_envs"]
    out = {}
    for j, op in enumerate(sc["ops"]):
        if op[0] == "add":
            idx = op[3] if len(op) > 3 else None
            data = synth_rows(op[1], n_envs if idx is None else len(idx), op[2])
            rb.add(data, indices=idx) if idx is not None else rb.add(data)
        else:
            s = rb.sample(**op[1])
            for k, v in s.items():
                out[f"{name}/op{j}/{k}"] = to_numpy(v)
    return out
