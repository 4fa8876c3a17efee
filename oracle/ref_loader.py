"""Locate and import the UNMODIFIED reference (caprilovel/ECG_Denoise).  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python.  In the build container it lives under /root/reference; `oracle/build_ref.py`
mirrors the files of the RA-LENet path into the git-ignored `baseline/_ref/` so that they travel to the GPU box with
the repository snapshot (nothing of it enters the git history).  Only tests/, bench.py's baseline arms
(`--impl reference`, `--impl reference-gpu`, the `eager_gpu_baseline` / `cpu_baseline` records) and the golden
generators import this module; the product package never does.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]
# the three model files of the path (+ the metrics the north star names)
PATH_FILES = ["model/transformer.py", "model/raletransformer.py", "model/ralenet_12leads.py",
              "local_utils/evaluate.py"]


def find_reference(explicit: str | None = None) -> str | None:
    """directory holding the reference sources, or None."""
    for root in ([explicit] if explicit else []) + CANDIDATES:
        if root and all(os.path.exists(os.path.join(root, f)) for f in PATH_FILES):
            return root
    return None


class Reference:
    """the reference's modules of the path, imported from `root` without modification."""

    def __init__(self, root: str):
        self.root = root
        saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items()
                                                 if k == "model" or k.startswith("model.")
                                                 or k == "local_utils" or k.startswith("local_utils.")}
        for k in saved_mods:
            del sys.modules[k]
        sys.path.insert(0, root)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                self.transformer = importlib.import_module("model.transformer")
                self.raletransformer = importlib.import_module("model.raletransformer")
                self.evaluate = importlib.import_module("local_utils.evaluate")
            # model/ralenet_12leads.py ends in a body-less `if __name__ == "__main__":`: exec its
            # own text with `pass` appended in memory
            src = open(os.path.join(root, "model", "ralenet_12leads.py")).read() + "\n    pass\n"
            mod = types.ModuleType("ralenet_12leads_patched")
            with contextlib.redirect_stdout(io.StringIO()):
                exec(compile(src, os.path.join(root, "model", "ralenet_12leads.py"), "exec"), mod.__dict__)
            self.ralenet_12leads = mod
        finally:
            # leave the interpreter's import state as it was (ecg_denoise_b200.install() may own `model.*`)
            for k in [k for k in sys.modules if k == "model" or k.startswith("model.")
                      or k == "local_utils" or k.startswith("local_utils.")]:
                del sys.modules[k]
            sys.modules.update(saved_mods)
            sys.path[:] = saved_path

    def quiet(self, fn, *a, **k):
        with contextlib.redirect_stdout(io.StringIO()):       # reference Mlp.__init__ prints a flag
            return fn(*a, **k)


_cached: dict = {}


def load_reference(explicit: str | None = None) -> Reference | None:
    root = find_reference(explicit)
    if root is None:
        return None
    if root not in _cached:
        _cached[root] = Reference(root)
    return _cached[root]
