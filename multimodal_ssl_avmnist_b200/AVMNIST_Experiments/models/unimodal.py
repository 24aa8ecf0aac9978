"""CentralNet-style unimodal CNNs (parameter containers).  Mirrors the reference's models/unimodal.py:105-221 by name and
shape so that state_dicts interchange; the arithmetic of the training step runs in libavmnist_b200.so (the fused
conv / BatchNorm / ReLU / max-pool kernels), driven by multimodal_ssl_avmnist_b200.engine.  Calling one of these modules
directly runs the same kernels in inference form through `b200_module_forward`."""
import torch.nn as nn


def _stage(cin, cout, pad):
    return nn.Conv2d(cin, cout, kernel_size=5, padding=pad), nn.BatchNorm2d(cout)


class _CentralCNN(nn.Module):
    CHANNELS = ()
    PADS = ()
    FLAT = 0

    def __init__(self, dropout_prob=0.5, with_head=False):
        super().__init__()
        self.with_head = with_head
        chans = self.CHANNELS
        for k in range(len(chans) - 1):
            conv, bn = _stage(chans[k], chans[k + 1], self.PADS[k])
            setattr(self, f"conv{k + 1}", conv)
            setattr(self, f"bn{k + 1}", bn)
        self.dropout = nn.Dropout(dropout_prob)
        # classifier head of the supervised CentralNet baseline: kept because it is part of the reference's parameter
        # list (EMA'd, check-pointed, in the optimiser) although the DINO path never calls it
        self.fc1 = nn.Linear(self.FLAT, 1024)
        self.fc2 = nn.Linear(1024, 10)

    def forward(self, x):
        from multimodal_ssl_avmnist_b200.module_forward import central_cnn_forward
        return central_cnn_forward(self, x)


class CentralUnimodalImage(_CentralCNN):
    """1x28x28 -> conv5(32, pad 2) -> 14x14 -> conv5(64, pad 0) -> 5x5 -> 1600 features."""
    CHANNELS = (1, 32, 64)
    PADS = (2, 0)
    FLAT = 64 * 5 * 5


class CentralUnimodalAudio(_CentralCNN):
    """1x112x112 -> four conv5(pad 2)+pool stages (8, 16, 32, 64 channels) -> 7x7 -> 3136 features."""
    CHANNELS = (1, 8, 16, 32, 64)
    PADS = (2, 2, 2, 2)
    FLAT = 64 * 7 * 7
