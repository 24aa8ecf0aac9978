"""Inference-form forward of the reference-shaped parameter containers through the CUDA kernels (no autograd tape):
used when an encoder / head module is called on its own (feature extraction, `DownstreamClassifier`-style probes).
Train-mode modules use batch statistics and update their running statistics, eval-mode modules use the running ones."""
import torch

from . import ops


def _conv_block(x, conv, bn, training):
    N, Cin, H, W = x.shape
    Cout, K, pad = conv.out_channels, conv.kernel_size[0], conv.padding[0]
    Ho = H + 2 * pad - K + 1
    dev = x.device
    z = torch.empty(N, Cout, Ho, Ho, device=dev)
    stats = torch.zeros(1, Cout, 2, dtype=torch.float64, device=dev) if training else None
    ops.conv_fwd(x.contiguous().float(), conv.weight.detach(), conv.bias.detach(), z, stats, N, pad)
    scale, shift, mean, invstd = (torch.empty(1, Cout, device=dev) for _ in range(4))
    ops.bn_finalize(stats, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.num_batches_tracked, scale, shift,
                    mean, invstd, 1, N * Ho * Ho, train=training, momentum=bn.momentum, eps=bn.eps)
    out = torch.empty(N, Cout, Ho // 2, Ho // 2, device=dev)
    ops.bn_relu_pool_fwd(z, scale, shift, out, N)
    return out


def _linear(x, lin, act=0):
    y = torch.empty(x.shape[0], lin.out_features, device=x.device)
    ops.linear_fwd(x.contiguous(), lin.weight.detach(), lin.bias.detach(), y, act=act)
    return y


def _require_cuda(x):
    if not x.is_cuda:
        raise ops._lib.B200Error("B200 modules run on CUDA tensors only (no CPU fallback)")


@torch.no_grad()
def central_cnn_forward(mod, x):
    """CentralUnimodalImage / CentralUnimodalAudio headless forward -> [N, FLAT] (models/unimodal.py:127-143, 185-211)."""
    _require_cuda(x)
    k = 1
    while hasattr(mod, f"conv{k}"):
        x = _conv_block(x, getattr(mod, f"conv{k}"), getattr(mod, f"bn{k}"), mod.training)
        k += 1
    x = x.flatten(1)
    if mod.with_head:
        x = _linear(_linear(x, mod.fc1, act=1), mod.fc2)        # dropout is the identity in inference form
    return x


@torch.no_grad()
def sequential_cnn_forward(seq, x):
    """image_encoder()/audio_encoder()-style nn.Sequential: [conv, bn, relu, pool]* -> AdaptiveAvgPool -> Flatten -> Linear."""
    _require_cuda(x)
    mods = list(seq)
    i = 0
    while i < len(mods) and isinstance(mods[i], torch.nn.Conv2d):
        x = _conv_block(x, mods[i], mods[i + 1], seq.training)
        i += 4
    pooled = torch.empty(x.shape[0], x.shape[1], device=x.device)
    ops.avgpool_fwd(x.contiguous(), pooled)
    return _linear(pooled, mods[-1])


@torch.no_grad()
def projection_head_forward(head, x):
    """ProjectionHead in inference form (models/dino.py:1251-1254); dropout is the identity."""
    _require_cuda(x)
    lin0, bn, _, _, lin1 = list(head.mlp)
    h = _linear(x, lin0)
    M, C = h.shape
    stats = torch.zeros(C, 2, dtype=torch.float64, device=h.device)
    if head.training:
        ops.colstats(h, stats)
    scale, shift, mean, invstd = (torch.empty(1, C, device=h.device) for _ in range(4))
    ops.bn_finalize(stats, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.num_batches_tracked, scale, shift,
                    mean, invstd, 1, M, train=head.training, momentum=bn.momentum, eps=bn.eps)
    g = torch.empty_like(h)
    ops.bn1d_gelu_drop_fwd(h, scale, shift, None, 0.0, g)
    return _linear(g, lin1)


@torch.no_grad()
def fusion_forward(fusion, feats):
    lin0, _, _, lin1 = list(fusion)
    return _linear(_linear(feats, lin0, act=1), lin1)
