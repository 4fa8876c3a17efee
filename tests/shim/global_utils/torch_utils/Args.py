import argparse


class TorchArgs(argparse.ArgumentParser):
    """main.py:30-40 adds three ints and reads `batch_size` and `epochs` from the parsed namespace; the defaults
    follow the reference's dead local copy (local_utils/Args.py:17-32: batch 32, epochs 100)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.add_argument("--batch_size", type=int, default=32)
        self.add_argument("--epochs", type=int, default=100)
        self.add_argument("--lr", type=float, default=1e-4)
        self.add_argument("--seed", type=int, default=2023)
