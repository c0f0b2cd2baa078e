"""Algorithm registry, mirroring the reference plugin interface (sheeprl/utils/registry.py:11-35,97-108).

When the real `sheeprl` package is importable, registrations are forwarded to ITS registries so that
`sheeprl exp=dreamer_v3` resolves to this package (the reference's lookup keeps the LAST registered
module with a matching algorithm name, sheeprl/cli.py:82-88).  Otherwise a local registry with the same
structure is used.
"""
from __future__ import annotations

import sys
from typing import Any, Callable, Dict, List

try:  # pragma: no cover - depends on the environment
    from sheeprl.utils.registry import algorithm_registry, evaluation_registry  # type: ignore
except Exception:  # real sheeprl (lightning/hydra) not installed
    algorithm_registry: Dict[str, List[Dict[str, Any]]] = {}
    evaluation_registry: Dict[str, List[Dict[str, Any]]] = {}


def _register(fn: Callable, decoupled: bool) -> Callable:
    module = fn.__module__                       # e.g. sheeprl_b200.algos.dreamer_v3.dreamer_v3
    package, _, algorithm = module.rpartition(".")
    entry = {"name": algorithm, "entrypoint": fn.__name__, "decoupled": decoupled}
    registered = algorithm_registry.setdefault(package, [])
    if any(e["name"] == algorithm for e in registered):
        raise ValueError(f"The algorithm `{algorithm}` has already been registered in the module `{package}`")
    registered.append(entry)
    mod = sys.modules.get(module)
    if mod is not None:
        names = getattr(mod, "__all__", None)
        if names is None:
            mod.__all__ = [fn.__name__]
        elif fn.__name__ not in names:
            names.append(fn.__name__)
    return fn


def register_algorithm(decoupled: bool = False):
    def deco(fn):
        return _register(fn, decoupled)

    return deco


def find_algorithm(name: str):
    """Last registered match wins (reference semantics, sheeprl/cli.py:82-88)."""
    found = None
    for package, algos in algorithm_registry.items():
        for a in algos:
            if a["name"] == name:
                found = (package, a)
    return found
