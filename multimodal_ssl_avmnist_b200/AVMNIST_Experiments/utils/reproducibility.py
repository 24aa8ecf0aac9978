"""Seeding helper with the reference's name and effect (utils/reproducibility.py:8-22)."""
import os
import random

import numpy as np
import torch

os.environ.setdefault("CUBLAS_WORKSPACE_CONFIG", ":4096:8")


def set_seed(seed=1):
    """Seed Python, numpy and torch (CPU + every CUDA device) and ask for deterministic library kernels.  The B200 kernels
    themselves are deterministic up to the fp64 atomics of the BatchNorm reductions."""
    for fn in (random.seed, np.random.seed, torch.manual_seed):
        fn(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    torch.use_deterministic_algorithms(True, warn_only=True)
