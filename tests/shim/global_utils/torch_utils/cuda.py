import os


def find_gpus(nums=1):
    """main.py:1-3 assigns the result to CUDA_VISIBLE_DEVICES: keep whatever the launcher chose, else GPU 0."""
    return os.environ.get("CUDA_VISIBLE_DEVICES", "0")
