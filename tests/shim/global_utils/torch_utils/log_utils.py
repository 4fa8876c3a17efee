import functools
import os


def train_log(*dargs, **dkwargs):
    """denoise_train.py:14 `@train_log()`: a decorator factory around train(); logging is out of scope."""
    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*a, **k):
            return fn(*a, **k)
        return wrapper
    return deco


def mkdir(path):
    """denoise_train.py:91-92"""
    os.makedirs(path, exist_ok=True)


class Logger:
    def __init__(self, *a, **k):
        pass

    def write(self, *a, **k):
        pass

    def flush(self):
        pass


def easymail(*a, **k):
    return None
