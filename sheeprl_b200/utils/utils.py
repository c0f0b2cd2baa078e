"""Small host-side helpers mirroring the reference's `sheeprl/utils/utils.py` surface.

Reference: sheeprl/utils/utils.py:34-60 (dotdict).
"""
from __future__ import annotations

import copy
from typing import Any, Dict, Mapping


class dotdict(dict):
    """Nested dict with attribute access (reference: sheeprl/utils/utils.py:34-60)."""

    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in list(self.items()):
            if isinstance(v, Mapping) and not isinstance(v, dotdict):
                self[k] = dotdict(v)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __deepcopy__(self, memo):
        return dotdict({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def as_dict(self) -> Dict[str, Any]:
        return {k: (v.as_dict() if isinstance(v, dotdict) else v) for k, v in self.items()}
