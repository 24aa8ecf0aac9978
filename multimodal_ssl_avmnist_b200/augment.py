"""Multi-crop augmentation: chain specification, host-side parameter sampling and parameter packing.

The CUDA augmentation kernels (csrc/augment.cu) are *parameter driven*: every (sample, view) carries a short
list of already-sampled op records (`b200_aug_op`, include/avmnist_b200.h).  Two producers fill those records:

* `HostSampler` (this file) draws them on the host from torch's CPU generator and Python's `random` in
  exactly the order the reference's transform objects would (utils/get_data.py:121-257 + the torchvision /
  torchaudio `get_params` call order), so that, seeded identically, the reference API
  (`MultiModalAugmentation.__call__`) reproduces the reference's crops, rotations, masks and noise.
* the device sampler kernel (`b200_aug_sample`, csrc/augment.cu) draws the same distributions from Philox
  for throughput runs (B*6 views per step never touch the host).

Both feed the same apply kernels, so parity tests of the apply path cover the production path.
"""
import math
import random

import numpy as np
import torch

# ---- applied-op kinds (b200_aug_op.kind) -----------------------------------------------------------------
OP_NOP = 0
OP_CROP_RESIZE = 1
OP_AFFINE = 2
OP_ERASE = 3
OP_FREQ_MASK = 4
OP_TIME_MASK = 5
OP_NOISE = 6
OP_GROUP_MASK = 7
OP_TIME_WARP = 8
OP_BLUR3 = 9          # 3x3 Gaussian blur, payload = the three normalised 1-D taps (SimCLR image chain, utils/get_data.py:337)
OP_ELASTIC = 10       # elastic deformation; the sampling grid (identity + displacement) lives in a side array (get_data.py:330)

# ---- chain-spec kinds (b200_aug_spec_op.kind): what to *sample* ------------------------------------------
SPEC_RRC = 1          # a = scale_lo, scale_hi, log_ratio_lo, log_ratio_hi
SPEC_ROTATE = 2       # a = degrees
SPEC_AFFINE = 3       # a = degrees, translate_x, translate_y, scale_lo, scale_hi, has_scale
SPEC_ERASE = 4        # a = scale_lo, scale_hi, log_ratio_lo, log_ratio_hi      (own probability p)
SPEC_FREQ_MASK = 5    # a = mask_param
SPEC_TIME_MASK = 6    # a = mask_param
SPEC_NOISE = 7        # a = std
SPEC_GROUP_MASK = 8   # a = number of masked groups (int(ratio * 784)), group size
SPEC_TIME_WARP = 9    # a = min_factor, max_factor
SPEC_BLUR = 10        # a = sigma_lo, sigma_hi  (kernel size 3)
SPEC_ELASTIC = 11     # a = alpha, sigma

MAX_OPS = 8
OP_WORDS = 8          # int32 words per op record: kind + 7 payload words
GROUP_WORDS = 28      # 784 group bits -> 25 words, padded to 28 (16-byte multiple)
ALWAYS = -1.0         # "p" of an op that is not wrapped in RandomApply


class OpSpec:
    __slots__ = ("kind", "p", "a")

    def __init__(self, kind, p, *a):
        self.kind = kind
        self.p = float(p)
        self.a = [float(v) for v in a] + [0.0] * (6 - len(a))

    def __repr__(self):
        return f"OpSpec(kind={self.kind}, p={self.p}, a={self.a})"


def _log_ratio(ratio):
    # torchvision computes torch.log(torch.tensor(ratio)) in fp32 and feeds the fp32 values to uniform_
    lr = torch.log(torch.tensor([float(ratio[0]), float(ratio[1])]))
    return float(lr[0]), float(lr[1])


def rrc_spec(scale, p=ALWAYS, ratio=(3.0 / 4.0, 4.0 / 3.0)):
    lo, hi = _log_ratio(ratio)
    return OpSpec(SPEC_RRC, p, scale[0], scale[1], lo, hi)


def rotate_spec(degrees, p=ALWAYS):
    return OpSpec(SPEC_ROTATE, p, degrees)


def affine_spec(degrees=0.0, translate=None, scale=None, p=ALWAYS):
    tx, ty = (translate if translate is not None else (0.0, 0.0))
    has_t = 1.0 if translate is not None else 0.0
    if scale is None:
        return OpSpec(SPEC_AFFINE, p, degrees, tx, ty, 1.0, 1.0, 0.0 + 2.0 * has_t)
    return OpSpec(SPEC_AFFINE, p, degrees, tx, ty, scale[0], scale[1], 1.0 + 2.0 * has_t)


def erase_spec(p, scale=(0.02, 0.33), ratio=(0.3, 3.3)):
    lo, hi = _log_ratio(ratio)
    return OpSpec(SPEC_ERASE, p, scale[0], scale[1], lo, hi)


def image_chains():
    """The fixed image chains of the reference (utils/get_data.py:122-132)."""
    g = [rrc_spec((0.75, 1.0)), rotate_spec(5.0), affine_spec(0.0, (0.1, 0.1), None)]
    l = [rrc_spec((0.3, 0.75)), rotate_spec(15.0), affine_spec(0.0, (0.2, 0.2), (0.8, 1.2)),
         erase_spec(0.3, (0.02, 0.15))]
    return g, l


def simclr_chains():
    """The fixed chains of SimCLRMultiModalAugmentation (utils/get_data.py:311-363): (image chain, spectrogram chain)."""
    img = [rrc_spec((0.5, 1.0), ratio=(0.8, 1.2)), rotate_spec(5.0), affine_spec(0.0, (0.1, 0.1), None),
           OpSpec(SPEC_ELASTIC, 0.3, 20.0, 3.0), OpSpec(SPEC_BLUR, 0.3, 0.1, 0.5)]
    aud = [rrc_spec((0.5, 1.0)), OpSpec(SPEC_TIME_WARP, 0.5, 0.9, 1.1), OpSpec(SPEC_FREQ_MASK, 0.5, 10), OpSpec(SPEC_TIME_MASK, 0.5, 10),
           OpSpec(SPEC_NOISE, 0.3, 0.05)]
    return img, aud


def gaussian_taps(ksize, sigma):
    """torchvision _get_gaussian_kernel1d (fp32): the normalised taps of a Gaussian blur."""
    half = (ksize - 1) * 0.5
    x = torch.linspace(-half, half, steps=ksize, dtype=torch.float32)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    return pdf / pdf.sum()


def elastic_grid(alpha, sigma, H, W):
    """ElasticTransform.get_params + the grid of F.elastic_transform (torchvision), drawing the two uniform fields from torch's CPU
    generator exactly like the transform: displacement = gaussian_blur(2 rand - 1) * alpha / size, grid = identity + displacement.
    The blur is torchvision's (reflect padding + depth-wise conv2d with the outer-product kernel).  Returns fp32 [2, H, W] (x, y)."""
    out = []
    for size, axis in ((W, 1), (H, 0)):
        d = torch.rand([1, 1, H, W]) * 2 - 1
        if sigma > 0.0:
            k = int(8 * sigma + 1)
            k += 1 - k % 2
            t = gaussian_taps(k, sigma)
            kernel = torch.mm(t[:, None], t[None, :])[None, None]
            d = torch.nn.functional.conv2d(torch.nn.functional.pad(d, [k // 2] * 4, mode="reflect"), kernel)
        d = d * alpha / size
        ident = torch.linspace((-size + 1) / size, (size - 1) / size, size)
        out.append((ident[None, :] if axis == 1 else ident[:, None]) + d[0, 0])
    return torch.stack(out).numpy()


def default_audio_chains():
    """Hard-coded audio chains used when augment_values is None (utils/get_data.py:133-193)."""
    g = [rrc_spec((0.8, 1.0), 0.5),
         OpSpec(SPEC_TIME_WARP, 0.3, 0.9, 1.1),
         OpSpec(SPEC_FREQ_MASK, 0.3, 15),
         OpSpec(SPEC_TIME_MASK, 0.3, 15),
         affine_spec(0.0, (0.0, 0.1), (0.9, 1.1), 0.5),
         OpSpec(SPEC_GROUP_MASK, 0.5, int(0.15 * 784), 4)]
    l = [rrc_spec((0.5, 0.9), 0.7),
         OpSpec(SPEC_TIME_WARP, 0.7, 0.7, 1.3),
         OpSpec(SPEC_FREQ_MASK, 0.7, 25),
         OpSpec(SPEC_TIME_MASK, 0.7, 25),
         affine_spec(0.0, (0.0, 0.2), (0.7, 1.3), 0.7),
         OpSpec(SPEC_NOISE, 0.7, 0.1),
         OpSpec(SPEC_GROUP_MASK, 0.9, int(0.6 * 784), 4)]
    return g, l


def audio_chains_from_values(augment_values):
    """Chains from the YAML `best_augments` dict as produced by process_augment_config
    (hyperparameter_tuning/objective_augment.py:70-96); op order = dict key order (get_data.py:205-220)."""
    out = {}
    for view in ("global_views", "local_views"):
        chain = []
        augs = augment_values["augmentations"][view]
        probs = augment_values["augmentation_probabilities"][view]
        for name, args in augs.items():
            p = probs[name]
            if name == "time_warp":
                chain.append(OpSpec(SPEC_TIME_WARP, p, args.get("min_factor", 0.8), args.get("max_factor", 1.2)))
            elif name == "frequency_mask":
                chain.append(OpSpec(SPEC_FREQ_MASK, p, args["freq_mask_param"]))
            elif name == "time_mask":
                chain.append(OpSpec(SPEC_TIME_MASK, p, args["time_mask_param"]))
            elif name == "grouped_masking":
                gs = args.get("group_size", 4)
                chain.append(OpSpec(SPEC_GROUP_MASK, p, int(args.get("mask_ratio", 0.5) * (112 // gs) ** 2), gs))
            elif name == "gaussian_noise":
                chain.append(OpSpec(SPEC_NOISE, p, args.get("std", 0.1)))
            elif name == "random_affine":
                chain.append(affine_spec(args.get("degrees", 0), args.get("translate"), args.get("scale"), p))
            elif name == "random_resized_crop":
                chain.append(rrc_spec(tuple(args.get("scale", (0.08, 1.0))), p,
                                      tuple(args.get("ratio", (3.0 / 4.0, 4.0 / 3.0)))))
            else:
                raise KeyError(f"unknown augmentation '{name}'")
        if len(chain) > MAX_OPS:
            raise ValueError(f"{view}: at most {MAX_OPS} audio ops are supported")
        out[view] = chain
    return out["global_views"], out["local_views"]


def values_from_config(config):
    """YAML `best_augments` -> {'augmentations', 'augmentation_probabilities'} (final-training branch of
    process_augment_config, hyperparameter_tuning/objective_augment.py:70-96)."""
    if "best_augments" not in config:
        raise ValueError("best_augments not found in config for final training")
    augs = {"global_views": {}, "local_views": {}}
    probs = {"global_views": {}, "local_views": {}}
    for view in augs:
        for name, params in config["best_augments"][view].items():
            args = {k: v for k, v in params.items() if k != "p"}
            if args:
                augs[view][name] = args
            if "p" in params:
                probs[view][name] = params["p"]
    return {"augmentations": augs, "augmentation_probabilities": probs}


def inverse_affine_matrix(angle, tx, ty, scale):
    """Inverse affine matrix for center (0,0), no shear (torchvision convention), Python doubles."""
    rot = math.radians(angle)
    a, b, c, d = math.cos(rot), -math.sin(rot), math.sin(rot), math.cos(rot)
    m = [d / scale, -b / scale, 0.0, -c / scale, a / scale, 0.0]
    m[2] += m[0] * (-tx) + m[1] * (-ty)
    m[5] += m[3] * (-tx) + m[4] * (-ty)
    return m


def _u(lo, hi):
    return torch.empty(1).uniform_(lo, hi).item()


class HostSampler:
    """Samples op records for one view on the host, consuming torch's CPU generator (and Python's `random`
    for the time-warp rate) in the same order as the reference transform objects."""

    @staticmethod
    def _rrc(op, H, W):
        area = H * W
        for _ in range(10):
            target = area * _u(op.a[0], op.a[1])
            ar = torch.exp(torch.empty(1).uniform_(op.a[2], op.a[3])).item()
            w = int(round(math.sqrt(target * ar)))
            h = int(round(math.sqrt(target / ar)))
            if 0 < w <= W and 0 < h <= H:
                i = torch.randint(0, H - h + 1, size=(1,)).item()
                j = torch.randint(0, W - w + 1, size=(1,)).item()
                return i, j, h, w
        ratio = (math.exp(op.a[2]), math.exp(op.a[3]))
        in_ratio = float(W) / float(H)
        if in_ratio < min(ratio):
            w = W
            h = int(round(w / min(ratio)))
        elif in_ratio > max(ratio):
            h = H
            w = int(round(h * max(ratio)))
        else:
            w, h = W, H
        return (H - h) // 2, (W - w) // 2, h, w

    @staticmethod
    def _erase(op, H, W):
        area = H * W
        for _ in range(10):
            ea = area * _u(op.a[0], op.a[1])
            ar = torch.exp(torch.empty(1).uniform_(op.a[2], op.a[3])).item()
            h = int(round(math.sqrt(ea * ar)))
            w = int(round(math.sqrt(ea / ar)))
            if not (h < H and w < W):
                continue
            i = torch.randint(0, H - h + 1, size=(1,)).item()
            j = torch.randint(0, W - w + 1, size=(1,)).item()
            return i, j, h, w
        return None

    @staticmethod
    def _mask(param, size):
        param = int(param)
        if param < 1:
            return None
        value = torch.rand(1) * param
        min_value = torch.rand(1) * (size - value)
        start = int(min_value.long())
        end = start + int(value.long())
        return start, end

    def sample_view(self, chain, H, W, batch=None):
        """Returns (ops, group_bits, noise): ops = [(kind, params)], group_bits = np.uint8[(H/4)*(W/4)] or None,
        noise = torch.float32[H,W] or None.  batch = B: the chain is applied to a whole [B,1,H,W] tensor at once like the SimCLR
        transforms (one parameter set for the batch, but GaussianNoise draws randn_like of the batch: noise is [B,H,W]).
        self.last_grid: the elastic sampling grid [2,H,W] of this view, or None."""
        ops, bits, noise = [], None, None
        self.last_grid = None
        for op in chain:
            if op.kind == SPEC_ERASE:
                if not bool(torch.rand(1) < op.p):
                    continue
                box = self._erase(op, H, W)
                if box is not None:
                    ops.append((OP_ERASE, box))
                continue
            if op.p != ALWAYS and bool(op.p < torch.rand(1)):
                continue
            if op.kind == SPEC_RRC:
                ops.append((OP_CROP_RESIZE, self._rrc(op, H, W)))
            elif op.kind == SPEC_ROTATE:
                angle = float(_u(-op.a[0], op.a[0]))
                ops.append((OP_AFFINE, tuple(inverse_affine_matrix(-angle, 0.0, 0.0, 1.0))))
            elif op.kind == SPEC_AFFINE:
                angle = float(_u(-op.a[0], op.a[0]))
                flags = int(op.a[5])
                tx = ty = 0
                if flags & 2:
                    max_dx, max_dy = float(op.a[1] * W), float(op.a[2] * H)
                    tx = int(round(_u(-max_dx, max_dx)))
                    ty = int(round(_u(-max_dy, max_dy)))
                scale = float(_u(op.a[3], op.a[4])) if flags & 1 else 1.0
                ops.append((OP_AFFINE, tuple(inverse_affine_matrix(angle, float(tx), float(ty), scale))))
            elif op.kind == SPEC_FREQ_MASK:
                m = self._mask(op.a[0], H)
                if m is not None:
                    ops.append((OP_FREQ_MASK, m))
            elif op.kind == SPEC_TIME_MASK:
                m = self._mask(op.a[0], W)
                if m is not None:
                    ops.append((OP_TIME_MASK, m))
            elif op.kind == SPEC_NOISE:
                if noise is not None:
                    raise ValueError("at most one gaussian_noise op per chain")
                noise = torch.randn(1, H, W)[0] if batch is None else torch.randn(batch, 1, H, W)[:, 0]
                ops.append((OP_NOISE, (op.a[0],)))
            elif op.kind == SPEC_GROUP_MASK:
                if bits is not None:
                    raise ValueError("at most one grouped_masking op per chain")
                gs = int(op.a[1]) or 4
                n_groups = (H // gs) * (W // gs)
                bits = np.zeros(n_groups, dtype=np.uint8)
                bits[torch.randperm(n_groups)[:int(op.a[0])].numpy()] = 1
                ops.append((OP_GROUP_MASK, ()))
            elif op.kind == SPEC_TIME_WARP:
                ops.append((OP_TIME_WARP, (random.uniform(op.a[0], op.a[1]),)))
            elif op.kind == SPEC_ELASTIC:
                if self.last_grid is not None:
                    raise ValueError("at most one elastic op per chain")
                self.last_grid = elastic_grid(op.a[0], op.a[1], H, W)
                ops.append((OP_ELASTIC, ()))
            elif op.kind == SPEC_BLUR:
                sigma = _u(op.a[0], op.a[1])
                ops.append((OP_BLUR3, tuple(float(v) for v in gaussian_taps(3, sigma))))
            else:
                raise ValueError(f"unknown spec kind {op.kind}")
        return ops, bits, noise


# ---- packing ---------------------------------------------------------------------------------------------

def pack_ops(view_ops, out):
    """view_ops: list of (kind, params) for one view; out: int32[MAX_OPS, OP_WORDS] numpy view to fill."""
    if len(view_ops) > MAX_OPS:
        raise ValueError("too many ops")
    out[:] = 0
    for k, (kind, p) in enumerate(view_ops):
        out[k, 0] = kind
        if kind in (OP_CROP_RESIZE, OP_ERASE, OP_FREQ_MASK, OP_TIME_MASK):
            for t, v in enumerate(p):
                out[k, 1 + t] = int(v)
        elif kind in (OP_AFFINE, OP_NOISE, OP_BLUR3):
            vals = np.asarray(p, dtype=np.float32)
            out[k, 1:1 + len(vals)] = vals.view(np.int32)
        elif kind == OP_TIME_WARP:            # the rate is a Python double (random.uniform): keep all 64 bits
            out[k, 1:3] = np.asarray([p[0]], dtype=np.float64).view(np.int32)


def pack_group_bits(bits):
    words = np.zeros(GROUP_WORDS, dtype=np.uint32)
    if bits is not None:
        idx = np.nonzero(bits)[0]
        np.bitwise_or.at(words, idx // 32, (np.uint32(1) << (idx % 32).astype(np.uint32)))
    return words


def pack_spec(chain):
    """Chain spec -> (int32[MAX_OPS, 8]) table for the device sampler: kind, p(bits), a0..a5(bits)."""
    tab = np.zeros((MAX_OPS, 8), dtype=np.int32)
    for k, op in enumerate(chain):
        tab[k, 0] = op.kind
        tab[k, 1:8] = np.asarray([op.p] + op.a, dtype=np.float32).view(np.int32)
    return tab
