"""Minimal stand-in for lightning.pytorch: just enough surface for the reference modules to import and for
their nn.Module logic to run on CPU.  Test infrastructure only (used by tests/golden/make_golden.py)."""
import torch


class LightningModule(torch.nn.Module):
    def save_hyperparameters(self, *a, **k):
        return None

    def log(self, *a, **k):
        return None

    @property
    def device(self):
        return next(self.parameters()).device


class LightningDataModule:
    def __init__(self):
        pass


class Callback:
    pass


class Trainer:
    def __init__(self, *a, **k):
        raise RuntimeError("refshim Trainer is a placeholder")


def seed_everything(seed, workers=False):
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed
