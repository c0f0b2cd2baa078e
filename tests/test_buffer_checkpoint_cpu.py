"""Buffer checkpoints: the device rings pickle like the reference's buffer objects
(`state["rb"] = rb`, utils/callback.py:37-41), `CheckpointCallback` mirrors the reference's truncated fix-up, and —
where the reference tree is present — buffers convert to / from the reference's own classes with identical
subsequent samples."""
import io
import os

import numpy as np
import pytest
import torch

from oracle import ref_harness
from oracle.ops_emul import EmulOps
from sheeprl_b200.data import buffers as RB
from sheeprl_b200.utils.callback import CheckpointCallback, load_replay_buffer
from tests.buffer_scenarios import seed_rngs, synth_rows


def _rows(length, n_envs, seed):
    d = synth_rows(length, n_envs, seed)
    d["truncated"] = np.zeros((length, n_envs, 1))
    return d


def _make(kind, device="cpu", ops=None):
    ops = ops or EmulOps()
    if kind == "uniform":
        rb = RB.ReplayBuffer(12, 2, device=device, ops=ops)
        rb.add(_rows(9, 2, 1)), rb.add(_rows(7, 2, 2))                       # wrapped, full
        kw = dict(batch_size=6, sample_next_obs=True, n_samples=2)
    elif kind == "sequential":
        rb = RB.SequentialReplayBuffer(16, 2, device=device, ops=ops)
        rb.add(_rows(11, 2, 3))                                              # not full
        kw = dict(batch_size=5, sequence_length=4, n_samples=2)
    else:
        rb = RB.EnvIndependentReplayBuffer(16, 3, buffer_cls=RB.SequentialReplayBuffer, device=device, ops=ops)
        rb.add(_rows(11, 3, 4)), rb.add(_rows(9, 2, 5), indices=[0, 2])      # two of the three rings wrapped
        kw = dict(batch_size=8, sequence_length=4, n_samples=2)
    seed_rngs(rb, 77)
    return rb, kw


def _heads(rb):
    rings = rb.buffer if isinstance(rb, RB.EnvIndependentReplayBuffer) else [rb]
    return [(r._pos, r._full) for r in rings]


def _same_samples(a, b, kw):
    sa, sb = a.sample(**kw), b.sample(**kw)
    assert sorted(sa) == sorted(sb)
    for k in sa:
        assert sa[k].dtype == sb[k].dtype and np.array_equal(sa[k], sb[k]), k


@pytest.mark.parametrize("kind", ["uniform", "sequential", "env_independent"])
def test_pickle_round_trip_continues_identically(kind):
    a, kw = _make(kind)
    bio = io.BytesIO()
    torch.save({"rb": a}, bio)
    bio.seek(0)
    b = load_replay_buffer(torch.load(bio, weights_only=False)["rb"], device="cpu", ops=EmulOps())
    assert type(b) is type(a) and _heads(a) == _heads(b)
    _same_samples(a, b, kw)
    extra = _rows(5, a.n_envs, 9)                                            # both keep working after the restore
    a.add(extra), b.add(extra)
    assert _heads(a) == _heads(b)
    _same_samples(a, b, kw)


class _Fabric:
    world_size, global_rank, is_global_zero = 1, 0, True

    def save(self, path, state):
        torch.save(state, path)


@pytest.mark.parametrize("kind", ["uniform", "env_independent"])
def test_checkpoint_callback_marks_the_last_step_truncated_only_in_the_file(kind, tmp_path):
    rb, kw = _make(kind)
    rings = rb.buffer if kind == "env_independent" else [rb]
    before = [r["truncated"].clone() for r in rings]
    cb = CheckpointCallback(keep_last=2)
    for i in range(3):
        cb.on_checkpoint_coupled(_Fabric(), str(tmp_path / f"ckpt_{i}.ckpt"), {"iter_num": i}, rb)
        os.utime(tmp_path / f"ckpt_{i}.ckpt", (i + 1, i + 1))
    assert sorted(p.name for p in tmp_path.glob("*.ckpt")) == ["ckpt_1.ckpt", "ckpt_2.ckpt"]          # keep_last
    for r, t in zip(rings, before):
        assert torch.equal(r["truncated"], t)                                # live buffer untouched afterwards
    state = torch.load(tmp_path / "ckpt_2.ckpt", weights_only=False)
    saved = load_replay_buffer(state["rb"], device="cpu", ops=EmulOps())
    for r, s in zip(rings, saved.buffer if kind == "env_independent" else [saved]):
        row = (r._pos - 1) % r.buffer_size
        assert bool((s["truncated"][row] == 1).all())                        # the episode ends in the checkpoint
        mask = torch.ones(r.buffer_size, dtype=torch.bool)
        mask[row] = False
        assert torch.equal(s["truncated"][mask], r["truncated"][mask])


needs_reference = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree absent")


@needs_reference
@pytest.mark.parametrize("kind", ["uniform", "sequential", "env_independent"])
def test_conversion_to_and_from_the_reference_classes(kind, tmp_path):
    ref_harness.install()
    import sheeprl.data.buffers as SB

    mine, kw = _make(kind)
    # ours -> reference object (what a reference run would resume from), also through its pickle and memmap storage
    for memmap in (False, True):
        ref = mine.to_reference(memmap=memmap, memmap_dir=tmp_path / f"mm_{kind}_{memmap}" if memmap else None)
        assert isinstance(ref, getattr(SB, type(mine).__name__))
        bio = io.BytesIO()
        torch.save({"rb": ref}, bio)
        bio.seek(0)
        ref2 = torch.load(bio, weights_only=False)["rb"]
        # reference object (as found in a reference checkpoint) -> ours
        back = load_replay_buffer(ref2, device="cpu", ops=EmulOps())
        assert type(back) is type(mine) and _heads(back) == _heads(mine)
        want = ref.sample(**kw)
        got = back.sample(**kw)
        for k in want:
            assert np.array_equal(np.asarray(want[k]), got[k]), (kind, memmap, k)
