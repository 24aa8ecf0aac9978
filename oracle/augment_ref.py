"""Oracle (CPU, numpy) restatement of the multi-crop augmentation arithmetic.  TEST INFRASTRUCTURE ONLY.

Follows the reference call sites utils/get_data.py:21-27 (GaussianNoise), :29-58 (TimeWarpWithStretch),
:60-108 (GroupedMasking), :121-231 (chains), :233-257 (view loop) and the third-party transforms those call
(torchvision RandomResizedCrop / RandomRotation / RandomAffine / RandomErasing, torchaudio
FrequencyMasking / TimeMasking / TimeStretch — un-vendored; algorithms restated from SURVEY.md Appendix A1-A5).

Every op takes an already *sampled* parameter record (see `OP_*`), so the functions are deterministic; the
record layout is the same one the CUDA kernels consume (include/avmnist_b200.h `b200_aug_op`).
"""
import math

import numpy as np

f32 = np.float32

# op kinds -- must match include/avmnist_b200.h
OP_NOP = 0
OP_CROP_RESIZE = 1    # ints  i, j, h, w
OP_AFFINE = 2         # f32   m0..m5 (inverse affine matrix, torchvision convention)
OP_ERASE = 3          # ints  i, j, h, w
OP_FREQ_MASK = 4      # ints  start, end   (rows)
OP_TIME_MASK = 5      # ints  start, end   (columns)
OP_NOISE = 6          # f32   std
OP_GROUP_MASK = 7     # bits live in a side array (one bit per 4x4 group, row-major 28x28)
OP_TIME_WARP = 8      # f32   rate
OP_BLUR3 = 9          # f32   k0, k1, k2 (normalised 1-D Gaussian taps; SimCLR chain, utils/get_data.py:337)
OP_ELASTIC = 10       # the sampling grid (identity + displacement, normalised units) lives in a side array [2, H, W] (get_data.py:330)


def inverse_affine_matrix(angle, tx, ty, scale):
    """torchvision.transforms.functional._get_inverse_affine_matrix with center=[0,0], shear=[0,0]
    (the only form reached from get_data.py:124-125,129-130,151-155,184-187); Python doubles."""
    rot = math.radians(angle)
    a = math.cos(rot)
    b = -math.sin(rot)
    c = math.sin(rot)
    d = math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m = [x / scale for x in m]
    m[2] += m[0] * (-tx) + m[1] * (-ty)
    m[5] += m[3] * (-tx) + m[4] * (-ty)
    return m


def affine_index_map(m, H, W):
    """Integer source index (or -1 = fill) for every output pixel of a NEAREST rotate/affine
    (torchvision _gen_affine_grid + grid_sample(nearest, zeros, align_corners=False))."""
    th = np.asarray(m, dtype=f32).reshape(2, 3)
    xs = np.arange(W, dtype=f32) + f32(-W * 0.5 + 0.5)
    ys = np.arange(H, dtype=f32) + f32(-H * 0.5 + 0.5)
    X, Y = np.meshgrid(xs, ys)
    r00, r10, r20 = (th[0, :] / f32(0.5 * W)).astype(f32)
    r01, r11, r21 = (th[1, :] / f32(0.5 * H)).astype(f32)
    # torch's CPU bmm accumulates k=0..2 with FMAs: acc = x*r0; acc = fma(y, r1, acc); acc = fma(1, r2, acc)
    # (pinned empirically against torch 2.11 CPU: 0 mismatching grid values in 1.3 M; the un-fused form
    # differs in the last ulp for ~14 % of grid values and flips ~1e-5 of the rounded indices)
    def fma(a, b, c):
        return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(f32)
    gx = fma(Y, r10, (X * r00).astype(f32)) + r20
    gy = fma(Y, r11, (X * r01).astype(f32)) + r21
    ix = (((gx + f32(1)) * f32(W)).astype(f32) - f32(1)) / f32(2)
    iy = (((gy + f32(1)) * f32(H)).astype(f32) - f32(1)) / f32(2)
    jx = np.rint(ix).astype(np.int64)
    jy = np.rint(iy).astype(np.int64)
    ok = (jx >= 0) & (jx < W) & (jy >= 0) & (jy < H)
    return np.where(ok, jy * W + jx, -1)


def op_affine(x, m):
    H, W = x.shape
    idx = affine_index_map(m, H, W)
    flat = np.concatenate([x.reshape(-1), np.zeros(1, dtype=f32)])
    return flat[idx].astype(f32)           # idx == -1 picks the appended zero


def aa_weights(in_size, out_size):
    """ATen _upsample_bilinear2d_aa weight table.  Returns (xmin[out], n[out], w[out,3])."""
    scale = f32(in_size) / f32(out_size)
    support = f32(scale) if scale >= 1.0 else f32(1.0)
    invscale = f32(1.0 / scale) if scale >= 1.0 else f32(1.0)
    xmin = np.zeros(out_size, dtype=np.int64)
    cnt = np.zeros(out_size, dtype=np.int64)
    wts = np.zeros((out_size, 3), dtype=f32)
    for i in range(out_size):
        center = f32(float(scale) * (i + 0.5))
        lo = max(int(float(center) - float(support) + 0.5), 0)
        n = min(int(float(center) + float(support) + 0.5), in_size) - lo
        n = max(0, min(n, 3))
        total = f32(0.0)
        for j in range(n):
            t = f32((j + lo - float(center) + 0.5) * float(invscale))
            w = f32(1.0) - abs(t) if abs(t) < 1.0 else f32(0.0)
            wts[i, j] = w
            total = f32(total + w)
        if total != 0:
            norm = f32(1.0 / float(total))
            wts[i, :n] = (wts[i, :n] * norm).astype(f32)
        xmin[i], cnt[i] = lo, n
    return xmin, cnt, wts


def op_crop_resize(x, i, j, h, w):
    """RandomResizedCrop body: crop [i:i+h, j:j+w] then antialiased bilinear resize to x.shape
    (horizontal pass first, then vertical; fp32 intermediates)."""
    H, W = x.shape
    c = x[i:i + h, j:j + w].astype(f32)
    xm, xn, xw = aa_weights(w, W)
    tmp = np.zeros((h, W), dtype=f32)
    def fma(a, b, acc):      # ATen's AA loops contract to FMAs (bit-exact against torch 2.11 CPU)
        return (a.astype(np.float64) * np.float64(b) + acc.astype(np.float64)).astype(f32)
    for o in range(W):
        acc = np.zeros(h, dtype=f32)
        for t in range(xn[o]):
            acc = (c[:, xm[o] + t] * xw[o, t]).astype(f32) if t == 0 else fma(c[:, xm[o] + t], xw[o, t], acc)
        tmp[:, o] = acc
    ym, yn, yw = aa_weights(h, H)
    out = np.zeros((H, W), dtype=f32)
    for o in range(H):
        acc = np.zeros(W, dtype=f32)
        for t in range(yn[o]):
            acc = (tmp[ym[o] + t, :] * yw[o, t]).astype(f32) if t == 0 else fma(tmp[ym[o] + t, :], yw[o, t], acc)
        out[o, :] = acc
    return out


def op_erase(x, i, j, h, w):
    out = x.copy()
    out[i:i + h, j:j + w] = 0
    return out


def op_freq_mask(x, start, end):
    out = x.copy()
    out[start:end, :] = 0
    return out


def op_time_mask(x, start, end):
    out = x.copy()
    out[:, start:end] = 0
    return out


def op_noise(x, std, noise):
    return (x + (noise * f32(std)).astype(f32)).astype(f32)


def op_group_mask(x, bits, group=4):
    """bits: iterable of 0/1, one per group (row-major over (H/4)x(W/4)); 1 = zero the group
    (get_data.py:86-106)."""
    H, W = x.shape
    m = np.asarray(bits, dtype=bool).reshape(H // group, W // group)
    keep = ~np.kron(m, np.ones((group, group), dtype=bool))
    return (x * keep.astype(f32)).astype(f32)


def arange_f32(n, step):
    """torch.arange(0, stop, step, dtype=float32) as ATen's CPU kernel evaluates it (RangeFactoriesKernel: the
    first floor(n/16)*16 values come from 8-lane vectors `float(double(float(step*i8)) + lane*step)`, the tail from
    the scalar `float(step*i)`; pinned empirically against torch 2.11 in the build container, 0 mismatches)."""
    k = np.arange(n, dtype=np.int64)
    scalar = (k.astype(np.float64) * step).astype(f32)
    i8 = k - (k % 8)
    base = (i8.astype(np.float64) * step).astype(f32)
    vec = (base.astype(np.float64) + (k % 8).astype(np.float64) * step).astype(f32)
    return np.where(k < (n // 16) * 16, vec, scalar)


def op_time_warp(x, rate):
    """TimeWarpWithStretch (get_data.py:42-58) == linear interpolation of |x| along time."""
    H, W = x.shape
    n_frames = int(math.ceil(W / rate))
    t = arange_f32(n_frames, rate)                 # torch.arange(0, W, rate, dtype=float32)
    alpha = np.fmod(t, f32(1.0)).astype(f32)
    i0 = t.astype(np.int64)
    xp = np.concatenate([np.abs(x), np.zeros((H, 2), dtype=f32)], axis=1)
    n0 = xp[:, i0]
    n1 = xp[:, i0 + 1]
    mag = ((alpha * n1).astype(f32) + ((f32(1.0) - alpha).astype(f32) * n0).astype(f32)).astype(f32)
    out = np.zeros((H, W), dtype=f32)
    n = min(n_frames, W)
    out[:, :n] = mag[:, :n]
    return out


def gaussian_kernel1d(ksize, sigma):
    """torchvision _get_gaussian_kernel1d in fp32: linspace(-h, h, k), exp(-0.5 (x / sigma)^2), normalised.  Evaluated with torch's own
    fp32 exp (ATen's vectorised exp is not numpy's expf bit for bit, and the taps feed bit-exact comparisons)."""
    import torch
    half = (ksize - 1) * 0.5
    x = torch.linspace(-half, half, steps=ksize, dtype=torch.float32)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    return (pdf / pdf.sum()).numpy().astype(f32)


def blur2d_reflect(x, k1d):
    """torchvision gaussian_blur: reflect padding, depth-wise conv2d with kernel2d = k (outer) k; fp32 row-major FMA accumulation
    (bit-exact against torch 2.11 CPU for the 25-tap displacement blur; within 1 ulp of oneDNN's 3x3 kernels)."""
    k = np.asarray(k1d, dtype=f32)
    n = k.shape[0]
    k2 = (k[:, None] * k[None, :]).astype(f32)
    H, W = x.shape
    xp = np.pad(x.astype(f32), n // 2, mode="reflect")
    acc = np.zeros((H, W), dtype=f32)
    for i in range(n):
        for j in range(n):
            acc = (xp[i:i + H, j:j + W].astype(np.float64) * np.float64(k2[i, j]) + acc.astype(np.float64)).astype(f32)
    return acc


def op_blur3(x, k0, k1, k2):
    return blur2d_reflect(x, [k0, k1, k2])


def elastic_grid(u_dx, u_dy, alpha, sigma):
    """torchvision ElasticTransform.get_params + F.elastic_transform's grid from the transform's two uniform draws u = torch.rand(H, W):
    displacement = blur(2u - 1) * alpha / size (kernel size int(8 sigma + 1) made odd), grid = identity + displacement with the identity
    from torch.linspace((-s+1)/s, (s-1)/s, s) (ATen's vectorised linspace is used as is: its rounding is not numpy's).
    Returns [2, H, W] fp32 (x, y), the layout the kernels consume."""
    import torch
    H, W = u_dx.shape
    out = []
    for u, size, axis in ((u_dx, W, 1), (u_dy, H, 0)):
        d = (u.astype(f32) * f32(2) - f32(1)).astype(f32)
        if sigma > 0.0:
            ks = int(8 * sigma + 1)
            ks += 1 - ks % 2
            d = blur2d_reflect(d, gaussian_kernel1d(ks, sigma))
        d = (d * f32(alpha) / f32(size)).astype(f32)
        ident = torch.linspace((-size + 1) / size, (size - 1) / size, size).numpy().astype(f32)
        out.append((ident[None, :] + d if axis == 1 else ident[:, None] + d).astype(f32))
    return np.stack(out)


def op_elastic(x, grid):
    """torchvision F.elastic_transform(img, displacement, BILINEAR, fill=0) given grid = identity + displacement:
    grid_sample(bilinear, zeros, align_corners=False) of the image AND of a ones mask; out = img * mask."""
    H, W = x.shape
    gx, gy = grid[0].astype(f32), grid[1].astype(f32)
    # ATen GridSamplerKernel.cpp (vectorised CPU path), align_corners=False: unnormalize(g) = (g + 1) * (size / 2) - 0.5; corner
    # weights w = x - floor(x), e = 1 - w, n = y - floor(y), s = 1 - n; nw = s*e, ne = s*w, sw = n*e, se = n*w; sum in that order
    ix = (((gx + f32(1)).astype(f32) * f32(W / 2)).astype(f32) - f32(0.5)).astype(f32)
    iy = (((gy + f32(1)).astype(f32) * f32(H / 2)).astype(f32) - f32(0.5)).astype(f32)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    w_ = (ix - x0).astype(f32)
    e_ = (f32(1) - w_).astype(f32)
    n_ = (iy - y0).astype(f32)
    s_ = (f32(1) - n_).astype(f32)
    xs = np.concatenate([x.astype(f32).reshape(-1), np.zeros(1, dtype=f32)])
    img = np.zeros((H, W), dtype=f32)
    mask = np.zeros((H, W), dtype=f32)
    for dy, dx, wgt in ((0, 0, (s_ * e_).astype(f32)), (0, 1, (s_ * w_).astype(f32)), (1, 0, (n_ * e_).astype(f32)), (1, 1, (n_ * w_).astype(f32))):
        xi = (x0 + dx).astype(np.int64)
        yi = (y0 + dy).astype(np.int64)
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        idx = np.where(ok, yi * W + xi, H * W)
        img = (xs[idx].astype(np.float64) * wgt.astype(np.float64) + img.astype(np.float64)).astype(f32)          # fmadd, like ATen's Vec path
        mask = (ok.astype(np.float64) * wgt.astype(np.float64) + mask.astype(np.float64)).astype(f32)
    return (img * mask).astype(f32)


def apply_chain(src, ops, group_bits=None, noise=None, grid=None):
    """Run one view's op list over a [H,W] fp32 array.  `ops` = list of (kind, params-tuple)."""
    x = np.asarray(src, dtype=f32)
    for kind, p in ops:
        if kind == OP_NOP:
            continue
        elif kind == OP_CROP_RESIZE:
            x = op_crop_resize(x, *[int(v) for v in p[:4]])
        elif kind == OP_AFFINE:
            x = op_affine(x, p[:6])
        elif kind == OP_ERASE:
            x = op_erase(x, *[int(v) for v in p[:4]])
        elif kind == OP_FREQ_MASK:
            x = op_freq_mask(x, int(p[0]), int(p[1]))
        elif kind == OP_TIME_MASK:
            x = op_time_mask(x, int(p[0]), int(p[1]))
        elif kind == OP_NOISE:
            x = op_noise(x, p[0], noise)
        elif kind == OP_GROUP_MASK:
            x = op_group_mask(x, group_bits)
        elif kind == OP_TIME_WARP:
            x = op_time_warp(x, p[0])
        elif kind == OP_BLUR3:
            x = op_blur3(x, p[0], p[1], p[2])
        elif kind == OP_ELASTIC:
            x = op_elastic(x, grid)
        else:
            raise ValueError(f"unknown op kind {kind}")
    return x
