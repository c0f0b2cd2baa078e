"""TEST INFRASTRUCTURE — writes tests/golden/buffers.npz by EXECUTING THE REAL REFERENCE buffers (container only):

    python -m oracle.make_golden_buffers

Runs every scenario of tests/buffer_scenarios.py on sheeprl.data.buffers.{ReplayBuffer, SequentialReplayBuffer,
EnvIndependentReplayBuffer} with seeded Generators and stores every sampled array.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from tests.buffer_scenarios import SCENARIOS, run_scenario  # noqa: E402


def main():
    ref_harness.install()
    from sheeprl.data import buffers as RB  # the unmodified reference

    def make(cls, kw):
        if "buffer_cls" in kw:
            kw["buffer_cls"] = getattr(RB, kw["buffer_cls"])
        return getattr(RB, cls)(**kw)

    out = {}
    for name in SCENARIOS:
        out.update(run_scenario(name, make))
    path = os.path.join(ROOT, "tests", "golden", "buffers.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
