"""CPU baseline for the augmentation stage: the same transform chains, built from the torchvision / torchaudio
classes the reference itself composes (utils/get_data.py:121-231).  TEST INFRASTRUCTURE / BASELINE ONLY: used by
bench.py's cpu_baseline and `--impl reference` legs to time what the reference executes per sample in its DataLoader
workers.  The three custom modules (GaussianNoise :21-27, TimeWarpWithStretch :29-58, GroupedMasking :60-108) are
restated here; the parity oracle for the arithmetic is oracle/augment_ref.py."""
import random

import torch
import torch.nn.functional as F


def _tv():
    import torchvision.transforms as T
    import torchaudio.transforms as AT
    return T, AT


class AddGaussianNoise(torch.nn.Module):
    def __init__(self, std=0.1):
        super().__init__()
        self.std = std

    def forward(self, x):
        return x + self.std * torch.randn_like(x)


class StretchTime(torch.nn.Module):
    """Phase-vocoder time stretch at a random rate, cropped / zero-padded back to `length` frames, magnitude only."""

    def __init__(self, min_factor=0.8, max_factor=1.2, length=112):
        super().__init__()
        _, AT = _tv()
        self.lo, self.hi, self.length = min_factor, max_factor, length
        self.vocoder = AT.TimeStretch(n_freq=length)

    def forward(self, spec):
        rate = random.uniform(self.lo, self.hi)
        out = self.vocoder(torch.complex(spec, torch.zeros_like(spec)), rate)
        n = out.shape[-1]
        if n > self.length:
            out = out[..., :self.length]
        elif n < self.length:
            out = F.pad(out, (0, self.length - n))
        return out.abs()


class MaskGroups(torch.nn.Module):
    def __init__(self, mask_ratio=0.5, group_size=4):
        super().__init__()
        self.ratio, self.g = mask_ratio, group_size

    def forward(self, spec):
        _, h, w = spec.shape
        gh, gw = h // self.g, w // self.g
        keep = torch.ones(gh * gw)
        keep[torch.randperm(gh * gw)[:int(self.ratio * gh * gw)]] = 0
        keep = keep.view(gh, gw).repeat_interleave(self.g, 0).repeat_interleave(self.g, 1)
        return spec * keep


def build_chains(augment_values=None):
    """(global_image, global_audio, local_image, local_audio) Compose objects."""
    T, AT = _tv()
    gi = T.Compose([T.RandomResizedCrop(28, scale=(0.75, 1.0), antialias=True), T.RandomRotation(5),
                    T.RandomAffine(0, translate=(0.1, 0.1))])
    li = T.Compose([T.RandomResizedCrop(28, scale=(0.3, 0.75), antialias=True), T.RandomRotation(15),
                    T.RandomAffine(0, translate=(0.2, 0.2), scale=(0.8, 1.2)), T.RandomErasing(p=0.3, scale=(0.02, 0.15))])
    makers = {
        "time_warp": lambda a: StretchTime(a.get("min_factor", 0.8), a.get("max_factor", 1.2)),
        "frequency_mask": lambda a: AT.FrequencyMasking(a["freq_mask_param"]),
        "time_mask": lambda a: AT.TimeMasking(a["time_mask_param"]),
        "grouped_masking": lambda a: MaskGroups(a.get("mask_ratio", 0.5)),
        "gaussian_noise": lambda a: AddGaussianNoise(a.get("std", 0.1)),
        "random_affine": lambda a: T.RandomAffine(a.get("degrees", 0), translate=tuple(a["translate"]) if "translate" in a else None,
                                                  scale=tuple(a["scale"]) if "scale" in a else None),
        "random_resized_crop": lambda a: T.RandomResizedCrop(tuple(a.get("size", (112, 112))), scale=tuple(a.get("scale", (0.08, 1.0))),
                                                             antialias=True),
    }
    if augment_values is None:
        ga = T.Compose([T.RandomApply([T.RandomResizedCrop((112, 112), scale=(0.8, 1.0), antialias=True)], 0.5),
                        T.RandomApply([StretchTime(0.9, 1.1)], 0.3), T.RandomApply([AT.FrequencyMasking(15)], 0.3),
                        T.RandomApply([AT.TimeMasking(15)], 0.3),
                        T.RandomApply([T.RandomAffine(0, translate=(0, 0.1), scale=(0.9, 1.1))], 0.5),
                        T.RandomApply([MaskGroups(0.15)], 0.5)])
        la = T.Compose([T.RandomApply([T.RandomResizedCrop((112, 112), scale=(0.5, 0.9), antialias=True)], 0.7),
                        T.RandomApply([StretchTime(0.7, 1.3)], 0.7), T.RandomApply([AT.FrequencyMasking(25)], 0.7),
                        T.RandomApply([AT.TimeMasking(25)], 0.7),
                        T.RandomApply([T.RandomAffine(0, translate=(0, 0.2), scale=(0.7, 1.3))], 0.7),
                        T.RandomApply([AddGaussianNoise(0.1)], 0.7), T.RandomApply([MaskGroups(0.6)], 0.9)])
    else:
        chains = {}
        for view in ("global_views", "local_views"):
            ops = []
            for name, args in augment_values["augmentations"][view].items():
                ops.append(T.RandomApply([makers[name](args)], augment_values["augmentation_probabilities"][view][name]))
            chains[view] = T.Compose(ops)
        ga, la = chains["global_views"], chains["local_views"]
    return gi, ga, li, la


@torch.no_grad()
def multicrop(image, audio, chains, n_global=2, n_local=4):
    """image [1,28,28], audio [1,112,112] -> (gi [Vg,1,28,28], ga, li, la), the reference's per-sample call order."""
    gi, ga, li, la = chains
    g_i, g_a, l_i, l_a = [], [], [], []
    for _ in range(n_global):
        g_i.append(gi(image))
        g_a.append(ga(audio))
    for _ in range(n_local):
        l_i.append(li(image))
        l_a.append(la(audio))
    return torch.stack(g_i), torch.stack(g_a), torch.stack(l_i), torch.stack(l_a)


def _worker(args):
    seed, n, augment_values = args
    torch.manual_seed(seed)
    random.seed(seed)
    torch.set_num_threads(1)
    chains = build_chains(augment_values)
    g = torch.Generator().manual_seed(seed)
    import time
    img = torch.rand(n, 1, 28, 28, generator=g)
    aud = torch.rand(n, 1, 112, 112, generator=g)
    multicrop(img[0], aud[0], chains)          # warm-up (imports, lazy kernels)
    t0 = time.perf_counter()
    for i in range(n):
        multicrop(img[i], aud[i], chains)
    return time.perf_counter() - t0


def time_augmentation(n_per_worker, workers, augment_values=None):
    """Wall time of `workers` processes each augmenting n_per_worker samples (how the reference's DataLoader runs it).
    Returns samples / second over all workers."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        times = pool.map(_worker, [(1000 + w, n_per_worker, augment_values) for w in range(workers)])
    return workers * n_per_worker / max(times)
