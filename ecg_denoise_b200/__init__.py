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
    `model.raletransformer`, `model.ralenet_12leads`) so main.py / Transfer_learning.py / test_cls.py run unchanged.

    Only those three sub-modules are overridden.  The reference's own `model` package (a namespace package next to
    main.py) is kept when it is importable, so its sibling modules -- `model.UNet`, `model.DAM`, `model.ACDAE`,
    `model.ResNet_cls` (main.py:62-83, test_cls.py:11) -- still import from the reference.  `force=False` leaves an
    already imported `model.<name>` alone."""
    import importlib
    import types
    from . import model as _m
    pkg = sys.modules.get("model")
    if pkg is None:
        try:
            pkg = importlib.import_module("model")      # the reference's package, when its root is on sys.path
        except ImportError:
            pkg = types.ModuleType("model")
            pkg.__path__ = []          # mark as package
            sys.modules["model"] = pkg
    for name in ("transformer", "raletransformer", "ralenet_12leads"):
        if not force and f"model.{name}" in sys.modules:
            continue
        mod = getattr(_m, name)
        sys.modules[f"model.{name}"] = mod
        setattr(pkg, name, mod)
    return pkg
