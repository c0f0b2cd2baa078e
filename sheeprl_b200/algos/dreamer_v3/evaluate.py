"""`sheeprl-eval` entry for Dreamer-V3 checkpoints on the B200 player (reference: sheeprl/algos/dreamer_v3/evaluate.py:16-63).
The evaluation loop is the reference's (`evaluate` -> `utils.test`); only `build_agent` is substituted, so the player that
acts is `PlayerDV3` on the CUDA kernels, loaded from the checkpoint's `world_model` / `actor` state dicts."""
from __future__ import annotations

from typing import Any, Dict

from sheeprl_b200.utils.delegate import import_reference, substituted


def evaluate(fabric, cfg: Dict[str, Any], state: Dict[str, Any]):
    from sheeprl_b200.algos.dreamer_v3 import agent as A
    from sheeprl_b200.algos.dreamer_v3 import utils as U

    ref = import_reference("sheeprl.algos.dreamer_v3.evaluate")
    ref_utils = import_reference("sheeprl.algos.dreamer_v3.utils")
    with substituted(ref, {"build_agent": A.build_agent}), substituted(ref_utils, {"prepare_obs": U.prepare_obs}):
        return ref.evaluate(fabric, cfg, state)


try:  # register under the reference's evaluation registry when it is importable (sheeprl/utils/registry.py:111-120)
    from sheeprl.utils.registry import register_evaluation  # type: ignore

    evaluate = register_evaluation(algorithms="dreamer_v3")(evaluate)
except Exception:  # pragma: no cover - real sheeprl not installed
    pass
