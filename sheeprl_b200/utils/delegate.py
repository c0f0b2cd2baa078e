"""Entry-point delegation: run the REFERENCE's own `main(fabric, cfg)` with the B200 pieces substituted.

The reference resolves `cfg.algo.name` to a registered `main` and launches it with `fabric.launch(main, cfg)`
(sheeprl/cli.py:82-98, 199).  Environment interaction, logging, checkpoint cadence, the replay-ratio governor —
everything in that loop that is not the update step — stays the reference's code: the B200 `main` imports the
reference module of the same algorithm and, for the duration of the call, rebinds the names the loop looks up in its
module globals (`build_agent`, `train`, `Moments`, the buffer classes, ...) to this package's implementations and
routes `hydra.utils.instantiate(<optimizer cfg>, params=...)` to the fused-Adam handles.  Nothing of the reference is
copied or re-implemented here.
"""
from __future__ import annotations

import contextlib
import importlib
from typing import Any, Callable, Dict, Optional


class _InstantiateProxy:
    """Stands where the reference module holds `hydra`: `hydra.utils.instantiate(cfg, params=...)` builds the
    optimizer handle of the flat group those parameters live in; every other attribute is the real hydra's."""

    def __init__(self, real, make_optimizer: Callable[[Any, list], Any]):
        self._real, self._make = real, make_optimizer
        self.utils = self

    def instantiate(self, config, *args, **kwargs):
        params = kwargs.get("params", None)
        if params is not None:
            opt = self._make(config, list(params))
            if opt is not None:
                return opt
        return self._real.utils.instantiate(config, *args, **kwargs)

    def __getattr__(self, name):
        real = object.__getattribute__(self, "_real")
        try:
            return getattr(real.utils, name)
        except AttributeError:
            return getattr(real, name)


@contextlib.contextmanager
def substituted(module, names: Dict[str, Any]):
    """Temporarily rebinds module-level names (only those the module actually defines)."""
    saved = {}
    try:
        for k, v in names.items():
            if hasattr(module, k):
                saved[k] = getattr(module, k)
                setattr(module, k, v)
        yield module
    finally:
        for k, v in saved.items():
            setattr(module, k, v)


def import_reference(module_name: str):
    try:
        return importlib.import_module(module_name)
    except ImportError as e:  # the drop-in is meant to run under an installed sheeprl
        raise ImportError(
            f"{module_name} is not importable: the B200 entry point runs the reference's own interaction loop "
            "(sheeprl must be installed; see INTEGRATION.md) and substitutes build_agent()/train() in it") from e


def run_reference_main(module_name: str, fabric, cfg, names: Dict[str, Any],
                       make_optimizer: Optional[Callable[[Any, list], Any]] = None, entry: str = "main"):
    ref = import_reference(module_name)
    subs = dict(names)
    if make_optimizer is not None and hasattr(ref, "hydra"):
        subs["hydra"] = _InstantiateProxy(ref.hydra, make_optimizer)
    with substituted(ref, subs):
        return getattr(ref, entry)(fabric, cfg)


def group_of(params: list, groups: Dict[str, Any]):
    """Name of the flat group the first tensor of `params` is a view of (None: not one of ours)."""
    if not params:
        return None
    p = params[0].data_ptr()
    for name, g in groups.items():
        lo = g.flat.data_ptr()
        if lo <= p < lo + 4 * g.numel:
            return name
    return None
