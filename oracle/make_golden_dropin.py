"""Generate tests/golden/dropin_golden.json: the UNMODIFIED reference's `denoise_train.train` driving the UNMODIFIED
reference `transformer.ralenet(high_level_enhence=True)` on the CPU (float32, as users run it).  TEST INFRASTRUCTURE.

    python -m oracle.make_golden_dropin [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dropin_case, ref_loader, synth_weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    R = ref_loader.load_reference(a.ref)
    torch.set_num_threads(1)             # deterministic reduction order
    m = R.quiet(R.transformer.ralenet, high_level_enhence=True)
    m.load_state_dict(synth_weights.make_state_dict("rw", 1, 2023), strict=True)
    res, line = dropin_case.run_train(R.root, m, use_gpu=False)
    out = {"train_snr": res[0], "test_snr": res[1], "train_rmse": res[2], "test_rmse": res[3], "output_txt": line,
           "epochs": dropin_case.EPOCHS, "batch": dropin_case.BATCH}
    path = os.path.join(dropin_case.ROOT, "tests", "golden", "dropin_golden.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, out)


if __name__ == "__main__":
    main()
