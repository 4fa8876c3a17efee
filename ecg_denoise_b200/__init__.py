"""ecg_denoise_b200 -- B200-native (sm_100a) forward/backward of RA-LENet (caprilovel/ECG_Denoise).

    from ecg_denoise_b200.model.transformer import ralenet      # same API as the reference's model.transformer
    ecg_denoise_b200.install()                                   # or: make `from model.transformer import ralenet`
                                                                 # in the reference's main.py resolve to this package
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install(force: bool = True):
    """Register the B200 modules under the reference's import names (`model.transformer`,
    `model.raletransformer`, `model.ralenet_12leads`) so main.py / Transfer_learning.py / test_cls.py run unchanged."""
    import types
    from . import model as _m
    pkg = sys.modules.get("model")
    if pkg is None or force:
        pkg = types.ModuleType("model")
        pkg.__path__ = []          # mark as package
        sys.modules["model"] = pkg
    for name in ("transformer", "raletransformer", "ralenet_12leads"):
        mod = getattr(_m, name)
        sys.modules[f"model.{name}"] = mod
        setattr(pkg, name, mod)
    return pkg
