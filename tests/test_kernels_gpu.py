"""GPU parity tests, kernel by kernel, through the C ABI (ops.py -> libavmnist_b200.so) against the CPU oracle.

Tolerances (stated per test): bit-exact for the EMA and for integer gather maps; 1e-6 absolute for the fp32
augmentation arithmetic; 1e-5..1e-4 relative for fp32 reductions whose summation order differs from ATen's.
"""
import os
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F
import yaml

pytestmark = pytest.mark.gpu

from oracle import augment_ref as AR
from oracle import dino_ref as R
from multimodal_ssl_avmnist_b200 import augment as A
from multimodal_ssl_avmnist_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "multimodal_ssl_avmnist_b200", "AVMNIST_Experiments", "configs")
DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _close(got, want, rtol, atol, what=""):
    got = got.detach().cpu().double()
    want = want.detach().cpu().double()
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    assert not bad.any(), f"{what}: max err {err.max().item():.3e} (|ref| max {want.abs().max().item():.3e}), {int(bad.sum())} bad"


# ------------------------------------------------------------------------------------------------------------
def test_ema_bit_exact(golden):
    gen = torch.Generator().manual_seed(11)
    t = torch.randn(1000, generator=gen)
    s = torch.randn(1000, generator=gen)
    td = t.to(DEV)
    ops.ema_flat(td, s.to(DEV), 0.996)
    assert td.cpu()[:8].tolist() == golden["losses"]["ema_first8"]          # reference arithmetic, bit for bit
    assert float(td.cpu().double().sum()) == golden["losses"]["ema_sum64"]
    # large ragged arena + multi-tensor table, against the oracle
    n = 6_600_580 + 3
    t, s = _rand(n, seed=1), _rand(n, seed=2)
    want = {"w": t.clone()}
    R.ema_update(want, {"w": s}, 0.996)
    td = t.to(DEV)
    ops.ema_flat(td, s.to(DEV), 0.996)
    assert torch.equal(td.cpu(), want["w"])
    sizes = [5, 800, 1, 51200, 64, 1638400, 33]
    ts = [_rand(k, seed=10 + i).to(DEV) for i, k in enumerate(sizes)]
    ss = [_rand(k, seed=30 + i).to(DEV) for i, k in enumerate(sizes)]
    want = [0.996 * a.cpu() + (1 - 0.996) * b.cpu() for a, b in zip(ts, ss)]
    ops.ema_multi(ops.MultiTensorTable(ts, ss), 0.996)
    for a, w in zip(ts, want):
        assert torch.equal(a.cpu(), w)


def _dino_gpu(S, T, center, variant=0):
    Vs, B, D = S.shape
    s, t, c = S.to(DEV), T.to(DEV), center.to(DEV)
    parts = ops.dino_loss_parts(B)
    grad = torch.empty_like(s)
    pl = torch.empty(parts, device=DEV)
    pc = torch.empty(parts, D, device=DEV)
    cm = None
    if variant == 1:
        cm = torch.empty(T.shape[0], D, device=DEV)
        ops.teacher_norm_colmean(t, c, cm)
    ops.dino_loss_fwd_bwd(s, t, c, 0.1, 0.04, grad, pl, pc, variant=variant, t_colmean=cm)
    loss = torch.empty(1, device=DEV)
    cen = c.clone()
    ops.center_update(cen, pc, pl, T.shape[0] * B, 0.9, loss)
    return loss.cpu(), grad.cpu(), cen.cpu()


def test_dino_loss_known_answer_and_grad(golden):
    torch.manual_seed(7)
    S, T = torch.randn(6, 4, 128), torch.randn(2, 4, 128)
    zero = torch.zeros(128)
    loss, grad, cen = _dino_gpu(S, T, zero)
    assert abs(float(loss) - golden["losses"]["dino_multimodal"]) < 2e-6
    assert abs(float(grad.abs().sum()) - golden["losses"]["dino_multimodal_grad_abs_sum"]) < 1e-5
    np.testing.assert_allclose(grad.flatten()[:8].numpy(), golden["losses"]["dino_multimodal_grad_first8"], rtol=1e-4, atol=1e-8)
    loss_u, _, _ = _dino_gpu(S, T, zero, variant=1)
    assert abs(float(loss_u) - golden["losses"]["dino_unimodal"]) < 2e-6
    _close(cen, R.center_update(zero[None], T.reshape(-1, 128), 0.9)[0], 1e-6, 1e-7, "center")


@pytest.mark.parametrize("B,D,Vs,Vt", [(1, 128, 6, 2), (37, 128, 6, 2), (1000, 256, 6, 2), (64, 32, 3, 1), (5000, 128, 6, 2)])
def test_dino_loss_vs_oracle(B, D, Vs, Vt):
    S, T = _rand(Vs, B, D, seed=1), _rand(Vt, B, D, seed=2)
    center = _rand(D, seed=3, scale=0.1)
    for variant, fn in ((0, R.dino_loss_multimodal), (1, R.dino_loss_unimodal)):
        Sg = S.clone().requires_grad_(True)
        want = fn(Sg, T - center)
        want.backward()
        loss, grad, cen = _dino_gpu(S, T, center, variant)
        assert abs(float(loss) - float(want)) < 5e-6 * max(1.0, abs(float(want))), (variant, float(loss), float(want))
        _close(grad, Sg.grad, 2e-4, 1e-6 * float(Sg.grad.abs().max()), f"dino grad v{variant}")
        _close(cen, R.center_update(center[None], T.reshape(-1, D), 0.9)[0], 1e-5, 1e-7, "center")


def test_aux_losses_vs_oracle(golden):
    torch.manual_seed(7)
    S = torch.randn(6, 4, 128)
    a, b = S[0].contiguous(), S[1].contiguous()
    for B, D in ((4, 128), (300, 128), (1000, 64)):
        if B != 4:
            a, b = _rand(B, D, seed=5), _rand(B, D, seed=6)
        ag, bg = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        want = R.mse_align_loss(ag, bg)
        want.backward()
        ga, gb, lo = torch.empty(B, D, device=DEV), torch.empty(B, D, device=DEV), torch.empty(1, device=DEV)
        ops.mse_align_fwd_bwd(a.to(DEV), b.to(DEV), ga, gb, lo)
        assert abs(float(lo) - float(want)) < 1e-6 * max(1.0, float(want)) + 1e-9
        _close(ga, ag.grad, 1e-4, 1e-10, "mse grad a")
        _close(gb, bg.grad, 1e-4, 1e-10, "mse grad b")
        if B == 4:
            assert abs(float(lo) - golden["losses"]["mse"]) < 1e-8
        # InfoNCE
        ag, bg = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        want = R.infonce_loss(ag, bg)
        want.backward()
        work = torch.empty(ops.infonce_work_floats(B, D), device=DEV)
        ops.infonce_fwd_bwd(a.to(DEV), b.to(DEV), ga, gb, lo, work)
        assert abs(float(lo) - float(want)) < 5e-6 * max(1.0, float(want)), (float(lo), float(want))
        _close(ga, ag.grad, 5e-4, 1e-8, "infonce grad a")
        _close(gb, bg.grad, 5e-4, 1e-8, "infonce grad b")
        if B == 4:
            assert abs(float(lo) - golden["losses"]["infonce"]) < 2e-6
    # 10-way CE
    for B in (4, 1000):
        logits, labels = _rand(B, 10, seed=8), torch.randint(0, 10, (B,), generator=torch.Generator().manual_seed(9))
        lg = logits.clone().requires_grad_(True)
        want = F.cross_entropy(lg, labels)
        want.backward()
        g, lo = torch.empty(B, 10, device=DEV), torch.empty(1, device=DEV)
        ops.ce_fwd_bwd(logits.to(DEV), labels.to(DEV), g, lo)
        assert abs(float(lo) - float(want)) < 2e-6 * max(1.0, float(want))
        _close(g, lg.grad, 1e-4, 1e-9, "ce grad")
    # cosine consistency
    E = _rand(6, 50, 256, seed=12)
    Eg = E.clone().requires_grad_(True)
    want = R.cosine_consistency_loss(Eg)
    want.backward()
    g, lo = torch.empty_like(E, device=DEV), torch.empty(1, device=DEV)
    ops.cosine_consistency_fwd_bwd(E.to(DEV), g, lo)
    assert abs(float(lo) - float(want)) < 5e-6
    _close(g, Eg.grad, 2e-4, 1e-9, "cosine grad")


# ------------------------------------------------------------------------------------------------------------
def _chains(tag):
    ig, il = A.image_chains()
    if tag == "default":
        ag, al = A.default_audio_chains()
    else:
        name = {"tuned": "config_multimodal_dino.yaml", "old": "config_multimodal_dino_old_augments.yaml"}[tag]
        ag, al = A.audio_chains_from_values(A.values_from_config(yaml.safe_load(open(os.path.join(CFG, name)))))
    return ig, il, ag, al


def _host_views(tag, B, data_seed, rng_seed, Vg=2, Vl=4):
    """Host-sampled parameters for B samples + the oracle's outputs."""
    ig, il, ag, al = _chains(tag)
    V = Vg + Vl
    g = torch.Generator().manual_seed(data_seed)
    img = torch.rand(B, 28, 28, generator=g)
    aud_u8 = torch.randint(0, 256, (B, 112, 112), generator=g, dtype=torch.uint8)
    aud = (aud_u8.double() / 255.0).float()
    torch.manual_seed(rng_seed)
    random.seed(rng_seed)
    hs = A.HostSampler()
    img_ops = np.zeros((B, V, A.MAX_OPS, A.OP_WORDS), dtype=np.int32)
    aud_ops = np.zeros_like(img_ops)
    bits = np.zeros((B, V, A.GROUP_WORDS), dtype=np.uint32)
    noise = torch.zeros(B, V, 112, 112)
    want_img = np.zeros((V, B, 28, 28), dtype=np.float32)
    want_aud = np.zeros((V, B, 112, 112), dtype=np.float32)
    for b in range(B):
        for v in range(V):
            ci, ca = (ig, ag) if v < Vg else (il, al)
            o, gb, nz = hs.sample_view(ci, 28, 28)
            A.pack_ops(o, img_ops[b, v])
            want_img[v, b] = AR.apply_chain(img[b].numpy(), o, gb, None)
            o, gb, nz = hs.sample_view(ca, 112, 112)
            A.pack_ops(o, aud_ops[b, v])
            bits[b, v] = A.pack_group_bits(gb)
            if nz is not None:
                noise[b, v] = nz
            want_aud[v, b] = AR.apply_chain(aud[b].numpy(), o, gb, None if nz is None else nz.numpy())
    return img, aud_u8, aud, img_ops, aud_ops, bits, noise, want_img, want_aud


@pytest.mark.parametrize("tag", ["tuned", "old", "default"])
def test_augment_apply_vs_oracle(tag):
    B = 6
    img, aud_u8, aud, img_ops, aud_ops, bits, noise, want_img, want_aud = _host_views(tag, B, 77, 5)
    out_i = torch.empty(6, B, 28, 28, device=DEV)
    ops.aug_apply_image(img.to(DEV), torch.from_numpy(img_ops).to(DEV), out_i)
    # image chain: crop/resize (fp32 FMA sums, identical order) -> two integer gathers -> erase: bit-exact
    assert np.array_equal(out_i.cpu().numpy(), want_img), np.abs(out_i.cpu().numpy() - want_img).max()
    for src in (aud_u8, aud):            # uint8 source (device-side /255) and fp32 source
        out_a = torch.empty(6, B, 112, 112, device=DEV)
        ops.aug_apply_audio(src.to(DEV), torch.from_numpy(aud_ops).to(DEV), torch.from_numpy(bits.view(np.int32)).to(DEV), out_a,
                            noise=noise.to(DEV))
        got = out_a.cpu().numpy()
        np.testing.assert_allclose(got, want_aud, rtol=0, atol=1e-6)      # fp32 tolerance (time-warp |polar| rounding)
        assert (np.abs(got - want_aud) > 0).mean() < 0.02


def test_augment_matches_reference_fixture(golden, golden_aug):
    """End to end against the outputs of the imported reference (tests/golden/augment.npz)."""
    for tag in ("tuned", "old", "default"):
        ig, il, ag, al = _chains(tag)
        for s in range(2):
            g = torch.Generator().manual_seed(1000 + s)
            img = torch.rand(1, 28, 28, generator=g)
            aud = torch.rand(1, 112, 112, generator=g)
            torch.manual_seed(s)
            random.seed(s)
            hs = A.HostSampler()
            img_ops = np.zeros((1, 6, A.MAX_OPS, A.OP_WORDS), dtype=np.int32)
            aud_ops = np.zeros_like(img_ops)
            bits = np.zeros((1, 6, A.GROUP_WORDS), dtype=np.uint32)
            noise = torch.zeros(1, 6, 112, 112)
            for v in range(6):
                ci, ca = (ig, ag) if v < 2 else (il, al)
                o, gb, nz = hs.sample_view(ci, 28, 28)
                A.pack_ops(o, img_ops[0, v])
                o, gb, nz = hs.sample_view(ca, 112, 112)
                A.pack_ops(o, aud_ops[0, v])
                bits[0, v] = A.pack_group_bits(gb)
                if nz is not None:
                    noise[0, v] = nz
            out_i = torch.empty(6, 1, 28, 28, device=DEV)
            out_a = torch.empty(6, 1, 112, 112, device=DEV)
            ops.aug_apply_image(img.to(DEV), torch.from_numpy(img_ops).to(DEV), out_i)
            ops.aug_apply_audio(aud.to(DEV), torch.from_numpy(aud_ops).to(DEV), torch.from_numpy(bits.view(np.int32)).to(DEV), out_a,
                                noise=noise.to(DEV))
            oi, oa = out_i.cpu().numpy(), out_a.cpu().numpy()
            np.testing.assert_allclose(oi[:2].transpose(1, 0, 2, 3), golden_aug[f"{tag}_{s}_gi"][None, :, 0], atol=1e-6, rtol=0)
            np.testing.assert_allclose(oi[2:].transpose(1, 0, 2, 3), golden_aug[f"{tag}_{s}_li"][None, :, 0], atol=1e-6, rtol=0)
            for nm, sl in (("ga", slice(0, 2)), ("la", slice(2, 6))):
                a = oa[sl, 0]
                np.testing.assert_allclose(a[:, ::4, 1::4], golden_aug[f"{tag}_{s}_{nm}_dec"][:, 0], atol=1e-6, rtol=0)
                np.testing.assert_allclose(a.astype(np.float64).sum(-1), golden_aug[f"{tag}_{s}_{nm}_rows"][:, 0], atol=2e-3)


def test_augment_device_sampler_distributions():
    ig, il, ag, al = _chains("tuned")
    spec = np.stack([A.pack_spec(ig), A.pack_spec(il), A.pack_spec(ag), A.pack_spec(al)])
    B, Vg, Vl = 2048, 2, 4
    V = Vg + Vl
    io = torch.zeros(B, V, A.MAX_OPS, A.OP_WORDS, dtype=torch.int32, device=DEV)
    ao = torch.zeros_like(io)
    gb = torch.zeros(B, V, A.GROUP_WORDS, dtype=torch.int32, device=DEV)
    ops.aug_sample(torch.from_numpy(spec).to(DEV), B, Vg, Vl, 1234, 0, io, ao, gb)
    io2, ao2 = torch.zeros_like(io), torch.zeros_like(ao)
    ops.aug_sample(torch.from_numpy(spec).to(DEV), B, Vg, Vl, 1234, 0, io2, ao2, torch.zeros_like(gb))
    assert torch.equal(io, io2) and torch.equal(ao, ao2)                      # deterministic in (seed, step)
    ops.aug_sample(torch.from_numpy(spec).to(DEV), B, Vg, Vl, 1234, 1, io2, ao2, torch.zeros_like(gb))
    assert not torch.equal(io, io2)
    io, ao, gbn = io.cpu().numpy(), ao.cpu().numpy(), gb.cpu().numpy().view(np.uint32)
    # image: every view starts with a crop whose area fraction lies in the configured scale range
    for vs, (lo, hi) in ((slice(0, Vg), (0.75, 1.0)), (slice(Vg, V), (0.3, 0.75))):
        rec = io[:, vs, 0]
        assert (rec[..., 0] == A.OP_CROP_RESIZE).all()
        h, w, i, j = rec[..., 3], rec[..., 4], rec[..., 1], rec[..., 2]
        area = h * w / 784.0
        assert (area > lo - 0.06).all() and (area < hi + 0.06).all()
        assert (i >= 0).all() and (j >= 0).all() and (i + h <= 28).all() and (j + w <= 28).all()
        # same acceptance-biased mean as the host sampler (boxes with w > 28 or h > 28 are re-drawn)
        torch.manual_seed(0)
        hs, chain = A.HostSampler(), (ig if lo > 0.5 else il)
        host = np.array([(lambda r: r[2] * r[3] / 784.0)(hs._rrc(chain[0], 28, 28)) for _ in range(3000)])
        assert abs(area.mean() - host.mean()) < 0.015, (area.mean(), host.mean())
        assert (io[:, vs, 1, 0] == A.OP_AFFINE).all() and (io[:, vs, 2, 0] == A.OP_AFFINE).all()
    erased = (io[:, Vg:, 3, 0] == A.OP_ERASE).mean()
    assert abs(erased - 0.3) < 0.03
    # audio local views: application rates follow the YAML probabilities
    kinds = ao[:, Vg:, :, 0]
    probs = {A.OP_FREQ_MASK: 0.9357, A.OP_NOISE: 0.8428, A.OP_GROUP_MASK: 0.9764, A.OP_TIME_MASK: 0.9702, A.OP_TIME_WARP: 0.8198,
             A.OP_CROP_RESIZE: 0.5, A.OP_AFFINE: 0.5}
    for k, p in probs.items():
        rate = (kinds == k).any(-1).mean()
        assert abs(rate - p) < 0.03, (k, rate, p)
    # grouped masking selects exactly int(ratio*784) groups whenever it is applied
    n_mask = int(0.6483441701034119 * 784)
    applied = (kinds == A.OP_GROUP_MASK).any(-1)
    pop = np.unpackbits(gbn[:, Vg:].view(np.uint8), axis=-1).sum(-1)
    assert (pop[applied] == n_mask).all() and (pop[~applied] == 0).all()
    # ... uniformly: every one of the 784 groups is masked with frequency n_mask / 784 (5 sigma), none of the padding bits ever
    sel = np.unpackbits(gbn[:, Vg:].view(np.uint8), axis=-1, bitorder="little")[applied]          # [records, 28 * 32], bit g = group g
    assert sel[:, 784:].sum() == 0
    pr = n_mask / 784.0
    assert np.abs(sel[:, :784].mean(0) - pr).max() < 5.0 * (pr * (1 - pr) / sel.shape[0]) ** 0.5 + 1e-3
    both = (sel[:, :783] & sel[:, 1:784]).mean(0)                    # neighbouring groups: no correlation beyond the fixed-size constraint
    assert np.abs(both - pr * (n_mask - 1) / 783.0).max() < 5.0 * (0.25 / sel.shape[0]) ** 0.5 + 1e-3
    # the sampled records drive the apply kernels without faults and produce finite, bounded views
    src_i = torch.rand(B, 28, 28, device=DEV)
    src_a = torch.randint(0, 256, (B, 112, 112), dtype=torch.uint8, device=DEV)
    out_i = torch.empty(V, B, 28, 28, device=DEV)
    out_a = torch.empty(V, B, 112, 112, device=DEV)
    ops.aug_apply_image(src_i, torch.from_numpy(io).to(DEV), out_i)
    ops.aug_apply_audio(src_a, torch.from_numpy(ao).to(DEV), gb, out_a, seed=99)
    torch.cuda.synchronize()
    assert torch.isfinite(out_i).all() and torch.isfinite(out_a).all()
    assert float(out_i.min()) >= 0 and float(out_i.max()) <= 1.0 + 1e-6
    assert 0.15 < float(out_a.abs().mean()) < 0.6


# ------------------------------------------------------------------------------------------------------------
CONV_SHAPES = [(1, 32, 28, 28, 5, 2), (32, 64, 14, 14, 5, 0), (1, 8, 112, 112, 5, 2), (8, 16, 56, 56, 5, 2), (16, 32, 28, 28, 5, 2),
               (32, 64, 14, 14, 5, 2), (1, 32, 28, 28, 3, 1), (32, 64, 14, 14, 3, 1), (64, 128, 7, 7, 3, 1),
               # multi_simple audio stack (models/dino.py:43-72)
               (1, 32, 112, 112, 3, 1), (32, 64, 56, 56, 3, 1), (64, 128, 28, 28, 3, 1), (128, 256, 14, 14, 3, 1)]


@pytest.mark.parametrize("shape", CONV_SHAPES)
@pytest.mark.parametrize("N,npv", [(6, 2), (35, 7)])
def test_conv_block_fwd_bwd(shape, N, npv):
    """conv -> BN(train, per view-call) -> ReLU -> maxpool2, forward and all gradients, vs torch CPU autograd."""
    Cin, Cout, H, W, K, pad = shape
    n_views = N // npv
    x = _rand(N, Cin, H, W, seed=1)
    w = _rand(Cout, Cin, K, K, seed=2, scale=1.0 / (Cin * K * K) ** 0.5)
    b = _rand(Cout, seed=3, scale=0.1)
    gamma = 1 + _rand(Cout, seed=4, scale=0.1)
    beta = _rand(Cout, seed=5, scale=0.1)
    HO, WO = H + 2 * pad - K + 1, W + 2 * pad - K + 1
    # ---- oracle ----
    xr, wr, br, gr, ber = (t.clone().requires_grad_(True) for t in (x, w, b, gamma, beta))
    rm, rv = torch.zeros(Cout), torch.ones(Cout)
    outs = []
    for v in range(n_views):
        z = F.conv2d(xr[v * npv:(v + 1) * npv], wr, br, padding=pad)
        y = F.batch_norm(z, rm, rv, gr, ber, training=True, momentum=0.1, eps=1e-5)
        outs.append(F.max_pool2d(F.relu(y), 2))
    out_ref = torch.cat(outs)
    dout = _rand(*out_ref.shape, seed=6)
    (out_ref * dout).sum().backward()
    # ---- CUDA path ----
    xd, wd, bd, gd, bed = (t.to(DEV) for t in (x, w, b, gamma, beta))
    z = torch.empty(N, Cout, HO, WO, device=DEV)
    stats = torch.zeros(n_views, Cout, 2, dtype=torch.float64, device=DEV)
    ops.conv_fwd(xd, wd, bd, z, stats, npv, pad)
    rmd, rvd = torch.zeros(Cout, device=DEV), torch.ones(Cout, device=DEV)
    nbt = torch.zeros(1, dtype=torch.int64, device=DEV)
    scale, shift, mean, invstd = (torch.empty(n_views, Cout, device=DEV) for _ in range(4))
    ops.bn_finalize(stats, gd, bed, rmd, rvd, nbt, scale, shift, mean, invstd, n_views, npv * HO * WO)
    out = torch.empty(N, Cout, HO // 2, WO // 2, device=DEV)
    ops.bn_relu_pool_fwd(z, scale, shift, out, npv)
    zref = torch.cat([F.conv2d(x[v * npv:(v + 1) * npv], w, b, padding=pad) for v in range(n_views)])
    _close(z, zref, 1e-5, 2e-6 * max(1.0, (Cin * K * K / 200) ** 0.5), "conv z")        # rounding grows with the reduction length
    _close(out, out_ref, 2e-5, 1e-5, "block out")
    _close(rmd, rm, 1e-5, 1e-6, "running_mean")
    _close(rvd, rv, 1e-5, 1e-6, "running_var")
    assert int(nbt) == n_views
    # backward
    doutd = dout.to(DEV)
    sums = torch.zeros(n_views, Cout, 2, dtype=torch.float64, device=DEV)
    ops.bn_relu_pool_bwd_reduce(z, doutd, scale, shift, mean, invstd, sums, npv)
    dz = torch.empty_like(z)
    ops.bn_relu_pool_bwd_apply(z, doutd, scale, shift, mean, invstd, sums, dz, npv)
    dgam, dbet = torch.empty(Cout, device=DEV), torch.empty(Cout, device=DEV)
    ops.bn_param_grads(sums, dgam, dbet, n_views)
    gscale = float(gr.grad.abs().max())
    _close(dgam, gr.grad, 1e-4, 1e-5 * gscale, "dgamma")
    _close(dbet, ber.grad, 1e-4, 1e-5 * float(ber.grad.abs().max()), "dbeta")
    dw, db = torch.empty_like(wd), torch.empty(Cout, device=DEV)
    work = torch.empty(ops.conv_bwd_weight_work_floats(N, Cin, Cout, H, W, K, pad), device=DEV)
    ops.conv_bwd_weight(xd, dz, dw, db, work, pad)
    # the kernel itself, with the pooling decisions fixed: against the fp64 weight gradient of OUR dz (a near-tie arg-max that flips
    # between the CUDA and the ATen forward moves a whole dout term, which is not the convolution kernels' business)
    dz64 = dz.detach().cpu().double()
    dw_iso = torch.nn.grad.conv2d_weight(x.double(), w.shape, dz64, padding=pad)
    _close(dw, dw_iso, 2e-4, 2e-5 * float(dw_iso.abs().max()), "dw (given dz)")
    big = N * Cout * HO * WO > 2_000_000          # > 0.5 M pooling windows: end-to-end flips are expected, see above
    if not big:
        _close(dw, wr.grad, 2e-4, 2e-5 * float(wr.grad.abs().max()), "dw")
    assert float(db.abs().max()) < 1e-3 * max(1.0, float(dout.abs().sum()) ** 0.5)       # bias feeds BN: exact gradient is 0
    if Cin > 1:
        dx = torch.empty_like(xd)
        ops.conv_bwd_data(dz, wd, dx, pad)
        dx_iso = torch.nn.grad.conv2d_input(x.shape, w.double(), dz64, padding=pad)
        _close(dx, dx_iso, 2e-4, 2e-5 * float(dx_iso.abs().max()), "dx (given dz)")
        if not big:
            _close(dx, xr.grad, 2e-4, 2e-5 * float(xr.grad.abs().max()), "dx")


def test_conv_odd_output_and_avgpool():
    """image_simple conv3: 7x7 output is pooled with floor (-> 3x3), then AdaptiveAvgPool2d(1)."""
    N, npv = 8, 4
    x = _rand(N, 64, 7, 7, seed=1)
    w = _rand(128, 64, 3, 3, seed=2, scale=0.05)
    b = _rand(128, seed=3, scale=0.1)
    z = torch.empty(N, 128, 7, 7, device=DEV)
    stats = torch.zeros(2, 128, 2, dtype=torch.float64, device=DEV)
    ops.conv_fwd(x.to(DEV), w.to(DEV), b.to(DEV), z, stats, npv, 1)
    _close(z, F.conv2d(x, w, b, padding=1), 1e-5, 2e-6, "conv z 7x7")
    zc = z.cpu()
    _close(stats[:, :, 0], torch.stack([zc[:4].sum((0, 2, 3)), zc[4:].sum((0, 2, 3))]), 1e-5, 1e-4, "stats sum")
    p = _rand(N, 128, 3, 3, seed=4)
    o = torch.empty(N, 128, device=DEV)
    ops.avgpool_fwd(p.to(DEV), o)
    _close(o, p.mean((2, 3)), 1e-6, 1e-7, "avgpool")
    dx = torch.empty(N, 128, 3, 3, device=DEV)
    do = _rand(N, 128, seed=5)
    ops.avgpool_bwd(do.to(DEV), dx)
    _close(dx, (do / 9)[:, :, None, None].expand(N, 128, 3, 3), 1e-6, 1e-8, "avgpool bwd")


@pytest.mark.parametrize("M,N,K", [(24, 256, 1600), (24, 256, 3136), (100, 256, 512), (77, 512, 256), (300, 128, 512), (33, 10, 512),
                                   (5, 256, 256), (1000, 512, 128)])
def test_linear_fwd_bwd(M, N, K):
    x, w, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3, scale=0.1)
    keep = torch.rand(M, N, generator=torch.Generator().manual_seed(4)) >= 0.3
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    y = torch.empty(M, N, device=DEV)
    ops.linear_fwd(xd, wd, bd, y)
    _close(y, F.linear(x, w, b), 1e-5, 1e-5, "linear")
    ops.linear_fwd(xd, wd, bd, y, act=1)
    _close(y, F.relu(F.linear(x, w, b)), 1e-5, 1e-5, "linear relu")
    ops.linear_fwd(xd, wd, bd, y, act=2, mask=keep.to(torch.uint8).to(DEV), drop_p=0.3)
    want = F.relu(F.linear(x, w, b)) * keep / 0.7
    _close(y, want, 1e-5, 1e-5, "linear relu dropout")
    dy = _rand(M, N, seed=5)
    dyd = dy.to(DEV)
    ops.act_bwd(dyd, y, drop_p=0.3)
    dh = dy * (want > 0) / 0.7
    _close(dyd, dh, 1e-6, 1e-7, "act bwd")
    dx = torch.empty(M, K, device=DEV)
    ops.linear_bwd_data(dyd, wd, dx)
    _close(dx, dh @ w, 1e-4, 1e-5, "linear dx")
    dw, db = torch.empty(N, K, device=DEV), torch.empty(N, device=DEV)
    ops.linear_bwd_weight(dyd, xd, dw, db)
    _close(dw, dh.T @ x, 1e-4, 2e-5 * float((dh.T @ x).abs().max()), "linear dw")
    _close(db, dh.sum(0), 1e-4, 1e-5, "linear db")
    ops.linear_bwd_weight(dyd, xd, dw, db, accumulate=True)
    _close(dw, 2 * (dh.T @ x), 1e-4, 4e-5 * float((dh.T @ x).abs().max()), "linear dw accumulate")
    # strided views (the fusion layer reads/writes halves of a concatenated feature matrix)
    wide = torch.zeros(M, 2 * N, device=DEV)
    ops.linear_fwd(xd, wd, bd, wide[:, N:])
    _close(wide[:, N:], F.linear(x, w, b), 1e-5, 1e-5, "linear strided out")
    assert float(wide[:, :N].abs().max()) == 0.0


@pytest.mark.parametrize("M,C,p", [(24, 512, 0.3), (1000, 512, 0.0), (37, 96, 0.3)])
def test_head_middle_bn1d_gelu_dropout(M, C, p):
    h = _rand(M, C, seed=1)
    gamma, beta = 1 + _rand(C, seed=2, scale=0.1), _rand(C, seed=3, scale=0.1)
    keep = torch.rand(M, C, generator=torch.Generator().manual_seed(4)) >= p
    hr, gr, br = (t.clone().requires_grad_(True) for t in (h, gamma, beta))
    rm, rv = torch.zeros(C), torch.ones(C)
    y = F.gelu(F.batch_norm(hr, rm, rv, gr, br, training=True, momentum=0.1, eps=1e-5))
    out_ref = y * keep / (1 - p) if p > 0 else y
    dg = _rand(M, C, seed=5)
    (out_ref * dg).sum().backward()
    hd = h.to(DEV)
    stats = torch.zeros(C, 2, dtype=torch.float64, device=DEV)
    ops.colstats(hd, stats)
    rmd, rvd, nbt = torch.zeros(C, device=DEV), torch.ones(C, device=DEV), torch.zeros(1, dtype=torch.int64, device=DEV)
    scale, shift, mean, invstd = (torch.empty(1, C, device=DEV) for _ in range(4))
    ops.bn_finalize(stats, gamma.to(DEV), beta.to(DEV), rmd, rvd, nbt, scale, shift, mean, invstd, 1, M)
    g = torch.empty(M, C, device=DEV)
    mask = keep.to(torch.uint8).to(DEV)
    ops.bn1d_gelu_drop_fwd(hd, scale, shift, mask, p, g)
    _close(g, out_ref, 2e-5, 2e-6, "head middle fwd")
    _close(rmd, rm, 1e-5, 1e-6, "bn1d running_mean")
    _close(rvd, rv, 1e-5, 1e-6, "bn1d running_var")
    sums = torch.zeros(C, 2, dtype=torch.float64, device=DEV)
    dgd = dg.to(DEV)
    ops.bn1d_gelu_drop_bwd_reduce(hd, dgd, scale, shift, mean, invstd, mask, p, sums)
    dh = torch.empty(M, C, device=DEV)
    ops.bn1d_gelu_drop_bwd_apply(hd, dgd, scale, shift, mean, invstd, mask, p, sums, dh)
    _close(dh, hr.grad, 2e-4, 2e-5 * float(hr.grad.abs().max()), "head middle dh")
    dgam, dbet = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    ops.bn_param_grads(sums, dgam, dbet, 1)
    _close(dgam, gr.grad, 1e-4, 1e-5 * float(gr.grad.abs().max()), "bn1d dgamma")
    _close(dbet, br.grad, 1e-4, 1e-5 * float(br.grad.abs().max()), "bn1d dbeta")


def test_adam_and_dropout_mask():
    n = 100_003
    p, g = _rand(n, seed=1), _rand(n, seed=2, scale=0.01)
    params, state = {"w": p.clone()}, {}
    pd, gd = p.to(DEV), g.to(DEV)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in (1, 2, 3):
        R.adam_step(params, {"w": g}, state, lr=1e-4, weight_decay=1e-6)
        ops.adam_flat(pd, gd, m, v, step, 1e-4, weight_decay=1e-6)
    _close(pd, params["w"], 1e-6, 1e-7, "adam")
    mask = torch.empty(1_000_001, dtype=torch.uint8, device=DEV)
    ops.dropout_mask(mask, 0.3, 42, 7)
    assert abs(float(mask.float().mean()) - 0.7) < 2e-3
    m2 = torch.empty_like(mask)
    ops.dropout_mask(m2, 0.3, 42, 7)
    assert torch.equal(mask, m2)
    ops.dropout_mask(m2, 0.3, 42, 8)
    assert not torch.equal(mask, m2)


@pytest.mark.parametrize("B,D", [(4, 128), (37, 128), (300, 64), (1000, 256)])
def test_ntxent_vs_oracle(B, D, golden):
    """b200_ntxent_fwd_bwd (SimCLR NT-Xent on the InfoNCE tile kernels, self-similarity masked, positives B rows apart)
    against the oracle's restatement of the reference; B = 4 is the reference's own known answer."""
    if B == 4:
        torch.manual_seed(7)
        S = torch.randn(6, 4, 128)
        reps = torch.cat([S[0], S[1]])
    else:
        reps = torch.randn(2 * B, D, generator=torch.Generator().manual_seed(B))
    want_in = reps.clone().double().requires_grad_(True)
    want = R.ntxent_loss(want_in)
    want.backward()
    x = reps.to(DEV)
    grad, loss = torch.empty_like(x), torch.empty(1, device=DEV)
    work = torch.empty(ops.ntxent_work_floats(2 * B, D), device=DEV)
    ops.ntxent_fwd_bwd(x, grad, loss, work)
    torch.cuda.synchronize()
    if B == 4:
        assert abs(float(loss) - golden["losses"]["ntxent"]) < 2e-6 * golden["losses"]["ntxent"] + 1e-6
    assert abs(float(loss) - float(want)) < 1e-5 * abs(float(want)) + 1e-6
    ref = want_in.grad.float()
    assert float((grad.cpu() - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 1e-8


def test_knn_matches_sklearn_and_fp64_brute_force():
    """train_knn_classifier (training_structures/dino_train.py:349-368) = sklearn KNeighborsClassifier(5).  Device path:
    exact-fp32 score GEMM + top-k + vote; checked against sklearn's predictions and an fp64 brute-force neighbour search."""
    from sklearn.neighbors import KNeighborsClassifier
    g = torch.Generator().manual_seed(3)
    N, M, D, k = 6000, 1500, 256, 5
    centers = torch.randn(10, D, generator=g) * 1.5
    ytr = torch.randint(0, 10, (N,), generator=g)
    yte = torch.randint(0, 10, (M,), generator=g)
    xtr = centers[ytr] + torch.randn(N, D, generator=g) * 2.0
    xte = centers[yte] + torch.randn(M, D, generator=g) * 2.0
    pred, nbr = ops.knn_predict(xtr.to(DEV), ytr.to(DEV), xte.to(DEV), k=k, n_classes=10, return_neighbours=True, chunk=640)
    torch.cuda.synchronize()
    pred, nbr = pred.cpu(), nbr.cpu().long()
    d = torch.cdist(xte.double(), xtr.double())
    want_nbr = d.topk(k, dim=1, largest=False).indices
    same_sets = (nbr.sort(1).values == want_nbr.sort(1).values).all(1).float().mean()
    assert float(same_sets) > 0.998, float(same_sets)
    assert (nbr[:, 0] == want_nbr[:, 0]).float().mean() > 0.998          # nearest first
    sk = KNeighborsClassifier(n_neighbors=k).fit(xtr.numpy(), ytr.numpy())
    want = torch.from_numpy(sk.predict(xte.numpy()))
    assert float((pred == want).float().mean()) > 0.998, float((pred == want).float().mean())
    # vote ties go to the smallest label, equal scores to the smaller train index: a hand-made case
    xt = torch.tensor([[0.0, 0.0], [2.0, 0.0], [0.0, 2.0], [-2.0, 0.0], [9.0, 9.0]])
    yt = torch.tensor([3, 1, 3, 1, 0])
    q = torch.tensor([[0.0, 0.0]])
    p4, n4 = ops.knn_predict(xt.to(DEV), yt.to(DEV), q.to(DEV), k=4, n_classes=10, return_neighbours=True)
    assert int(p4[0]) == 1 and n4[0].tolist() == [0, 1, 2, 3]               # 2 votes each for labels 1 and 3 -> 1


def test_device_generated_noise_is_standard_normal():
    """GaussianNoise (utils/get_data.py:21-31) on the throughput path: Philox + Box-Muller inside the augmentation kernel.  Only the
    distribution is specified (the reference draws torch.randn): mean 0, variance std^2, normal tails, no repeats across views / seeds."""
    B, V, std = 64, 6, 0.25
    rec = np.zeros((B, V, A.MAX_OPS, A.OP_WORDS), dtype=np.int32)
    A.pack_ops([(A.OP_NOISE, (std,))], rec[0, 0])
    rec[:] = rec[0, 0]
    src = torch.zeros(B, 112, 112, device=DEV)
    bits = torch.zeros(B, V, A.GROUP_WORDS, dtype=torch.int32, device=DEV)
    outs = []
    for seed in (11, 12):
        out = torch.empty(V, B, 112, 112, device=DEV)
        ops.aug_apply_audio(src, torch.from_numpy(rec).to(DEV), bits, out, noise=None, seed=seed)
        outs.append(out)
    x = (outs[0] / std).double().flatten()
    n = x.numel()
    assert abs(float(x.mean())) < 5 / n ** 0.5 and abs(float(x.var()) - 1.0) < 5 * (2 / n) ** 0.5
    assert abs(float((x ** 4).mean()) - 3.0) < 0.02 and abs(float((x ** 3).mean())) < 0.01        # kurtosis / skewness of N(0, 1)
    assert abs(float((x.abs() > 3).double().mean()) - 0.0027) < 3e-4 and float(x.abs().max()) < 6.5
    assert float((outs[0][0, 0] - outs[0][1, 0]).abs().max()) > 0.1 and float((outs[0] - outs[1]).abs().max()) > 0.1
    a, b = outs[0][0].flatten(), outs[0][1].flatten()
    assert abs(float((a * b).mean() / (a.std() * b.std()))) < 0.01                                  # views are uncorrelated


def test_simclr_augmentation_kernels_vs_oracle_and_reference(golden):
    """the SimCLR chains (utils/get_data.py:299-408) on the CUDA augmentation kernels.  (1) host-sampled records incl.
    ElasticTransform grids and GaussianBlur taps -> kernel == numpy oracle (same arithmetic: <= 1e-6, image ops mostly bit-equal);
    (2) the API mirror's SimCLRMultiModalAugmentation, seeded like the reference, against the imported reference's outputs
    (tests/golden/simclr_aug.npz) at 2e-6; (3) the device-sampled variant (Philox elastic field + blur sigma) is finite, differs per
    view and applies each op at its configured rate."""
    import sys
    MIRROR = os.path.join(os.path.dirname(CFG))
    sys.path.insert(0, MIRROR)
    import utils.get_data as gd
    img_chain, aud_chain = A.simclr_chains()
    # (1) kernel vs oracle on host-sampled records, B = 5, 2 views
    B, V = 5, 2
    g = torch.Generator().manual_seed(31)
    img = torch.rand(B, 28, 28, generator=g)
    n_el = n_bl = 0
    for seed in range(12):
        torch.manual_seed(seed)
        random.seed(seed)
        hs = A.HostSampler()
        rec = np.zeros((B, V, A.MAX_OPS, A.OP_WORDS), dtype=np.int32)
        grids = np.zeros((B, V, 2, 28, 28), dtype=np.float32)
        want = np.zeros((V, B, 28, 28), dtype=np.float32)
        for v in range(V):
            o, _, _ = hs.sample_view(img_chain, 28, 28, batch=B)
            A.pack_ops(o, rec[0, v])
            rec[:, v] = rec[0, v]
            if hs.last_grid is not None:
                grids[:, v] = hs.last_grid
            n_el += any(k == A.OP_ELASTIC for k, _ in o)
            n_bl += any(k == A.OP_BLUR3 for k, _ in o)
            for b in range(B):
                want[v, b] = AR.apply_chain(img[b].numpy(), o, None, None, grid=hs.last_grid)
        out = torch.full((V, B, 28, 28), float("nan"), device=DEV)
        ops.aug_apply_image(img.to(DEV), torch.from_numpy(rec).to(DEV), out, elastic_grid=torch.from_numpy(grids).to(DEV))
        got = out.cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
        assert (got != want).mean() < 0.05
    assert n_el >= 3 and n_bl >= 3
    # (2) the API mirror against the imported reference
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simclr_aug.npz"))
    aug = gd.SimCLRMultiModalAugmentation()
    for s in range(8):
        gg = torch.Generator().manual_seed(2000 + s)
        im = torch.rand(3, 1, 28, 28, generator=gg)
        au = torch.rand(3, 1, 112, 112, generator=gg)
        torch.manual_seed(s)
        random.seed(s)
        i1, a1, i2, a2 = aug(im, au)
        assert i1.shape == (3, 1, 28, 28) and a2.shape == (3, 1, 112, 112) and i1.is_cuda
        np.testing.assert_allclose(i1.cpu().numpy(), fx[f"s{s}_i1"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(i2.cpu().numpy(), fx[f"s{s}_i2"], rtol=0, atol=2e-6)
        for nm, a in (("a1", a1), ("a2", a2)):
            a = a.cpu().numpy()
            np.testing.assert_allclose(a[:, :, ::4, 1::4], fx[f"s{s}_{nm}_dec"], rtol=0, atol=2e-6)
            np.testing.assert_allclose(a.astype(np.float64).sum(-1), fx[f"s{s}_{nm}_rows"], rtol=0, atol=2e-3)
    # (3) device-sampled throughput variant
    B = 512
    spec = torch.from_numpy(np.stack([A.pack_spec(c) for c in (img_chain, img_chain, aud_chain, aud_chain)])).to(DEV)
    io = torch.zeros(B, 2, A.MAX_OPS, A.OP_WORDS, dtype=torch.int32, device=DEV)
    ao = torch.zeros_like(io)
    gb = torch.zeros(B, 2, A.GROUP_WORDS, dtype=torch.int32, device=DEV)
    ops.aug_sample(spec, B, 2, 0, 77, 3, io, ao, gb)
    src = torch.rand(B, 28, 28, generator=g).to(DEV)
    out = torch.full((2, B, 28, 28), float("nan"), device=DEV)
    ops.aug_apply_image(src, io, out, seed=77)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all() and float(out.min()) >= -1e-6 and float(out.max()) <= 1.0 + 1e-6
    kinds = io.cpu().numpy()[:, :, :, 0]
    for k, p in ((A.OP_ELASTIC, 0.3), (A.OP_BLUR3, 0.3)):
        rate = (kinds == k).any(-1).mean()
        assert abs(rate - p) < 0.05, (k, rate)
    taps = io.cpu().numpy()[kinds == A.OP_BLUR3][:, 1:4].copy().view(np.float32)
    assert np.allclose(taps.sum(-1), 1.0, atol=1e-6) and (taps[:, 1] > taps[:, 0]).all() and np.allclose(taps[:, 0], taps[:, 2])
    el = torch.from_numpy((kinds == A.OP_ELASTIC).any(-1)).to(DEV)          # [B, 2]
    moved = (out - torch.stack([src, src])).abs().amax(dim=(2, 3)).t()      # every view is at least cropped / rotated
    assert float(moved.min()) > 1e-3 and float((out[0] - out[1]).abs().max()) > 1e-3 and bool(el.any())


def test_mix_kernels_gate_softmax_add2d():
    """csrc/mix.cu against plain torch fp32/fp64: sigmoid gates (models/dino.py:249-256), the attention row softmax and its
    backward (models/dino.py:400-401), the strided accumulate; strided (column-slice) operands as the engine passes them."""
    g = torch.Generator().manual_seed(3)
    M, E = 37, 24
    cat = torch.randn(M, 2 * E, generator=g).to(DEV)
    dy = torch.randn(M, 2 * E, generator=g).to(DEV)
    gate = torch.tensor(0.37, device=DEV)
    out = torch.full((M, 2 * E), float("nan"), device=DEV)
    ops.gate_apply(cat[:, E:], gate, out[:, E:])
    torch.cuda.synchronize()
    assert torch.allclose(out[:, E:], torch.sigmoid(gate) * cat[:, E:], rtol=1e-6, atol=1e-7) and bool(torch.isnan(out[:, :E]).all())
    work = torch.zeros(ops.gate_grad_work_floats(), device=DEV)
    dgate = torch.full((), float("nan"), device=DEV)
    s = torch.sigmoid(gate.double())
    want = float((dy[:, E:].double() * cat[:, E:].double()).sum() * s * (1 - s))
    for _ in range(2):          # the ticket is left zero: a second call gives the same answer
        ops.gate_grad(dy[:, E:], cat[:, E:], gate, dgate, work)
        torch.cuda.synchronize()
        assert abs(float(dgate) - want) < 1e-5 * max(1.0, abs(want)), (float(dgate), want)
    ops.gate_grad(dy[:, E:], cat[:, E:], gate, dgate, work, accumulate=True)
    torch.cuda.synchronize()
    assert abs(float(dgate) - 2 * want) < 2e-5 * max(1.0, abs(want))
    for B in (4, 300, 1024):
        S = torch.randn(B, B, generator=g).to(DEV) * 3
        A = S.clone()
        scale = 256 ** -0.5
        ops.softmax_rows(A, scale)
        ref = torch.softmax(S.double() * scale, dim=-1)
        assert float((A.double() - ref).abs().max()) < 1e-6
        dA = torch.randn(B, B, generator=g).to(DEV)
        d = dA.clone()
        ops.softmax_rows_bwd(d, A, scale)
        dref = scale * ref * (dA.double() - (dA.double() * ref).sum(-1, keepdim=True))
        assert float((d.double() - dref).abs().max()) < 1e-6 * max(1.0, float(dref.abs().max()))
    dst = cat.clone()
    ops.add2d(dst[:, :E], dy[:, E:])
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :E], cat[:, :E] + dy[:, E:]) and torch.equal(dst[:, E:], cat[:, E:])
