"""Mirror the reference's Python sources into the git-ignored `baseline/_ref/`.  TEST / BASELINE INFRASTRUCTURE ONLY.

    python -m oracle.build_ref [--ref /root/reference]

The reference (caprilovel/ECG_Denoise) is pure Python with no build system, so "building" it is a file copy: the
.py files of `model/`, `local_utils/` and the driver scripts, byte for byte, into `baseline/_ref/` -- which is listed
in .gitignore (no reference source ever enters this repository's history) but NOT in .gpurunignore, so the directory
travels to the GPU box with the snapshot, exactly like the built `.so`.  There it is what `bench.py --impl reference`
(CPU) and the `eager_gpu_baseline` record (the same modules, eager, on the B200) run, and what
`tests/test_dropin_driver.py` drives.  `__graft_entry__.build()` calls this whenever /root/reference is present.
A MANIFEST with sha256 sums is written so a stale or edited mirror is detectable.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
DIRS = ["model", "local_utils"]
TOP = ["main.py", "denoise_train.py", "Transfer_learning.py", "test_cls.py", "requirements.txt"]


def build_ref(ref: str = "/root/reference", dest: str = DEST) -> bool:
    """returns True if the mirror exists afterwards."""
    if not os.path.isdir(ref):
        return os.path.exists(os.path.join(dest, "MANIFEST.json"))
    files = [f for f in TOP if os.path.exists(os.path.join(ref, f))]
    for d in DIRS:
        for name in sorted(os.listdir(os.path.join(ref, d))):
            if name.endswith(".py"):
                files.append(f"{d}/{name}")
    manifest = {}
    for f in files:
        src, dst = os.path.join(ref, f), os.path.join(dest, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        data = open(src, "rb").read()
        manifest[f] = hashlib.sha256(data).hexdigest()
        if not os.path.exists(dst) or open(dst, "rb").read() != data:
            shutil.copyfile(src, dst)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump({"source": ref, "files": manifest}, fh, indent=1, sort_keys=True)
    return True


def verify(dest: str = DEST) -> bool:
    """every mirrored file still has the recorded sha256 (i.e. the reference is unmodified)."""
    try:
        man = json.load(open(os.path.join(dest, "MANIFEST.json")))["files"]
    except (OSError, ValueError, KeyError):
        return False
    for f, h in man.items():
        try:
            if hashlib.sha256(open(os.path.join(dest, f), "rb").read()).hexdigest() != h:
                return False
        except OSError:
            return False
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    ok = build_ref(a.ref)
    print("baseline/_ref:", "ok" if ok and verify() else "MISSING")
