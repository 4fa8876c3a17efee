import random

import numpy as np
import torch


def random_seed(seed=2023):
    """main.py:24"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
