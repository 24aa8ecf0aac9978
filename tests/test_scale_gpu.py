"""GPU parity AT BENCHMARK SCALE (BASELINE configs 1/2: per-GPU batch 256 / 1024).  The kernel tests elsewhere run N <= 300
view-samples; the bench runs 6144 per launch with a different CTA partition (items*g/G bands, issuer split, grid-stride loops).
These tests put the same launch geometry under the oracle:

  * the whole step at B = 256 in both precisions -- loss, projections, EVERY layer's pre-BatchNorm z for all six view-calls,
    BatchNorm running statistics, gradients -- against the CPU oracle on identical inputs / masks / weights;
  * the device-sampled augmentation chain at B = 1024: the sampled op records are read back and replayed by the numpy oracle
    (image bit-exact, audio 1e-6), and the quad8 first-layer images the product path consumes are checked for every sample.
"""
import os

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

from oracle import augment_ref as AR
from oracle import dino_ref as R
from oracle.fixtures import make_masks, synth_views, views_to_vb
from multimodal_ssl_avmnist_b200 import augment as A
from multimodal_ssl_avmnist_b200 import ops
from multimodal_ssl_avmnist_b200.engine import DinoStepEngine

DEV = "cuda"
CFG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multimodal_ssl_avmnist_b200", "AVMNIST_Experiments", "configs")


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _l2(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _unact8(t):
    """act8 [N, C/8, H, W, 8] -> NCHW fp32"""
    N, P, H, W, _ = t.shape
    return t.float().permute(0, 1, 4, 2, 3).reshape(N, P * 8, H, W)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_step_vs_oracle_at_benchmark_batch(precision):
    B, V = 256, 6
    st = R.CentralDinoState(seed=21, mode="default")
    eng = DinoStepEngine(kind="multi_central", device=DEV, precision=precision)
    eng.load_named(student=st.student, teacher=st.teacher, student_head=st.student_head, teacher_head=st.teacher_head)
    img, aud = views_to_vb(*synth_views(B, seed=400))
    masks = make_masks(seed=401, V=V, Vg=2, B=B, E=256, hidden=512)
    R.TRACE = {}
    try:
        want = R.central_dino_step(st, img, aud, masks)
        trace = R.TRACE
    finally:
        R.TRACE = None
    loss = eng.forward_backward(img[:, :, 0].to(DEV).contiguous(), aud[:, :, 0].to(DEV).contiguous(),
                                masks={k: v.to(torch.uint8).to(DEV) for k, v in masks.items()})
    torch.cuda.synchronize()
    exact = precision == "fp32"
    w = eng._ws[B]
    # loss and projections
    assert abs(float(loss[3]) - float(want["loss"])) < (1e-5 if exact else 2e-3) * abs(float(want["loss"]))
    assert _rel(w["s.proj"].view(V, B, -1), want["student_out"]) < (5e-5 if exact else 2e-2)
    # every layer's pre-BatchNorm z, all six student view-calls (the oracle traces student calls first, then the teacher's)
    worst = {}
    for mod, layers, pre in (("img", eng.img_layers, "image_encoder.0."), ("aud", eng.aud_layers, "audio_encoder.0.")):
        for li, (conv, *_rest) in enumerate(layers):
            z = w[f"s.{mod}.z{li}"]
            z = _unact8(z) if z.dim() == 5 else z
            ref = torch.cat(trace[conv][:V])
            assert z.shape == ref.shape, (conv, z.shape, ref.shape)
            worst[conv] = (_rel(z, ref), _l2(z, ref))
            zt_ref = torch.cat(trace[conv][V:V + 2])
            if f"t.{mod}.z{li}" in w:
                zt = w[f"t.{mod}.z{li}"]
                zt = _unact8(zt) if zt.dim() == 5 else zt
            else:       # fused max-pool epilogue: the teacher never writes z, only the 2x2 window extreme (max / min by sign(gamma))
                zt = _unact8(w[f"t.{mod}.e{li}"])
                sgn = torch.where(st.teacher[conv.replace("conv", "bn") + ".weight"] < 0, -1.0, 1.0).view(1, -1, 1, 1)
                zt_ref = torch.nn.functional.max_pool2d(zt_ref * sgn, 2) * sgn
            worst[conv + "(teacher)"] = (_rel(zt, zt_ref), _l2(zt, zt_ref))
    print(precision, "per-layer z (max-rel, l2-rel):", {k: (round(a, 6), round(b, 6)) for k, (a, b) in worst.items()})
    for k, (mx, l2) in worst.items():
        assert mx < (2e-5 if exact else 4e-2) and l2 < (1e-5 if exact else 1.5e-2), (k, mx, l2)
    # BatchNorm running statistics after the six (student) / two (teacher) sequential updates
    btol = 2e-5 if exact else 1e-2
    for k in ("image_encoder.0.bn1", "image_encoder.0.bn2", "audio_encoder.0.bn1", "audio_encoder.0.bn2", "audio_encoder.0.bn3", "audio_encoder.0.bn4"):
        assert _rel(eng.bn_s["enc." + k].running_mean, st.student_buf[k + ".running_mean"]) < btol, k
        assert _rel(eng.bn_s["enc." + k].running_var, st.student_buf[k + ".running_var"]) < btol, k
        assert _rel(eng.bn_t["enc." + k].running_var, st.teacher_buf[k + ".running_var"]) < btol, k
    assert _rel(eng.center, st.center) < (1e-5 if exact else 1e-2)
    # gradients
    if exact:
        for prefix, gd in (("enc.", want["grads"]["student"]), ("head.", want["grads"]["student_head"])):
            for k, g in gd.items():
                if g.dim() >= 2:      # sums over 1536 x H x W terms that BatchNorm makes cancel: measured 1.0e-3 .. 4.0e-3 (B = 4: 3e-4)
                    assert _rel(eng.G[prefix + k], g) < 1e-2, (k, _rel(eng.G[prefix + k], g))
    else:
        fm, fw = [], []
        for prefix, gd in (("enc.", want["grads"]["student"]), ("head.", want["grads"]["student_head"])):
            for k, g in gd.items():
                if g.dim() >= 2:
                    fm.append(eng.G[prefix + k].detach().cpu().double().flatten())
                    fw.append(g.double().flatten())
        fm, fw = torch.cat(fm), torch.cat(fw)
        cos = float((fm @ fw) / (fm.norm() * fw.norm()))
        print("bf16 B=256: cosine(grad, oracle grad)", cos)
        assert cos > 0.97, cos


def _unpack_ops(rec):
    """inverse of augment.pack_ops: int32 [MAX_OPS, OP_WORDS] -> [(kind, params)]"""
    out = []
    for k in range(rec.shape[0]):
        kind = int(rec[k, 0])
        if kind == A.OP_NOP:
            continue
        if kind in (A.OP_CROP_RESIZE, A.OP_ERASE, A.OP_FREQ_MASK, A.OP_TIME_MASK):
            out.append((kind, tuple(int(v) for v in rec[k, 1:5])))
        elif kind in (A.OP_AFFINE, A.OP_NOISE):
            out.append((kind, tuple(float(v) for v in rec[k, 1:7].copy().view(np.float32))))
        elif kind == A.OP_TIME_WARP:
            out.append((kind, (float(rec[k, 1:3].copy().view(np.float64)[0]),)))
        elif kind == A.OP_GROUP_MASK:
            out.append((kind, ()))
        else:
            raise ValueError(kind)
    return out


def _quad8_of(views, pad):
    """fp32 [..., H, W] -> the quad8 image [..., H, ceil((W + 2 pad) / 4), 8] in bf16: unit xq = padded-row pixels 4 xq .. 4 xq + 7"""
    W = views.shape[-1]
    wq = ops.quad8_width(W, pad)
    padded = torch.nn.functional.pad(views, (pad, 4 * wq + 8 - W - pad))
    return padded.unfold(-1, 8, 4)[..., :wq, :].to(torch.bfloat16)


def test_device_sampled_augmentation_chain_at_bench_batch():
    B, Vg, Vl = 1024, 2, 4
    V = Vg + Vl
    ig, il = A.image_chains()
    ag, al = A.audio_chains_from_values(A.values_from_config(yaml.safe_load(open(os.path.join(CFG, "config_multimodal_dino.yaml")))))
    spec = torch.from_numpy(np.stack([A.pack_spec(c) for c in (ig, il, ag, al)])).to(DEV)
    g = torch.Generator().manual_seed(9)
    img = torch.rand(B, 28, 28, generator=g)
    aud_u8 = torch.randint(0, 256, (B, 112, 112), generator=g, dtype=torch.uint8)
    noise = torch.randn(B, V, 112, 112, generator=g)
    io = torch.zeros(B, V, A.MAX_OPS, A.OP_WORDS, dtype=torch.int32, device=DEV)
    ao = torch.zeros_like(io)
    gb = torch.zeros(B, V, A.GROUP_WORDS, dtype=torch.int32, device=DEV)
    ops.aug_sample(spec, B, Vg, Vl, 4321, 7, io, ao, gb)
    out_i = torch.empty(V, B, 28, 28, device=DEV)
    out_a = torch.empty(V, B, 112, 112, device=DEV)
    q_i = torch.empty(V, B, 28, ops.quad8_width(28, 2), 8, dtype=torch.bfloat16, device=DEV)
    q_a = torch.empty(V, B, 112, ops.quad8_width(112, 2), 8, dtype=torch.bfloat16, device=DEV)
    ops.aug_apply_image(img.to(DEV), io, out_i, out8=q_i, pad=2)
    ops.aug_apply_audio(aud_u8.to(DEV), ao, gb, out_a, noise=noise.to(DEV), out8=q_a, pad=2)
    torch.cuda.synchronize()
    # (1) the quad8 operand images == the fp32 views rounded to bf16 in the quad8 arrangement, for EVERY sample and view
    assert torch.equal(q_i, _quad8_of(out_i, 2))
    assert torch.equal(q_a, _quad8_of(out_a, 2))
    # (2) replay the SAMPLED records through the numpy oracle on a strided subset (every 23rd sample incl. the last CTA's)
    io_h, ao_h, gb_h = io.cpu().numpy(), ao.cpu().numpy(), gb.cpu().numpy().view(np.uint32)
    got_i, got_a = out_i.cpu().numpy(), out_a.cpu().numpy()
    aud = (aud_u8.double() / 255.0).float().numpy()
    n_diff = n_tot = 0
    for b in list(range(0, B, 23)) + [B - 1]:
        for v in range(V):
            want = AR.apply_chain(img[b].numpy(), _unpack_ops(io_h[b, v]), None, None)
            assert np.array_equal(got_i[v, b], want), (b, v, np.abs(got_i[v, b] - want).max())
            bits = np.unpackbits(gb_h[b, v].view(np.uint8), bitorder="little")[:784].astype(bool)
            want = AR.apply_chain(aud[b], _unpack_ops(ao_h[b, v]), bits, noise[b, v].numpy())
            np.testing.assert_allclose(got_a[v, b], want, rtol=0, atol=1e-6, err_msg=f"audio b={b} v={v}")
            n_diff += int((np.abs(got_a[v, b] - want) > 0).sum())
            n_tot += want.size
    assert n_diff / n_tot < 0.02


@pytest.mark.parametrize("mode", ["mse", "infonce", "semi_supervised"])
def test_mode_steps_on_the_product_path_at_benchmark_batch(mode):
    """BASELINE configs 3-5 (mse / infonce / semi_supervised: DINO + side loss on a 7th, un-augmented view-call) at B = 256 on the
    product path (bf16 tensor-core kernels, tf32 linears; InfoNCE through the tcgen05 3xTF32 similarity GEMM) against the fp32 CPU
    oracle: total loss 2e-3, side loss 5e-3, student projections 2e-2 of scale, mode-head outputs 3e-2, gradient direction > 0.97,
    BatchNorm running statistics after SEVEN sequential updates 1e-2."""
    from oracle.fixtures import synth_raw
    B, V = 256, 6
    st = R.CentralDinoState(seed=23, mode=mode)
    eng = DinoStepEngine(kind="multi_central", mode=mode, device=DEV, precision="bf16")
    eng.load_named(student=st.student, teacher=st.teacher, student_head=st.student_head, teacher_head=st.teacher_head,
                   aux_image=st.aux.get("image"), aux_audio=st.aux.get("audio"))
    img, aud = views_to_vb(*synth_views(B, seed=500))
    masks = make_masks(seed=501, V=V, Vg=2, B=B, E=256, hidden=512)
    image, audio, labels = synth_raw(B, seed=502)
    want = R.central_dino_step(st, img, aud, masks, raw=(image, audio), labels=labels)
    loss = eng.forward_backward(img[:, :, 0].to(DEV).contiguous(), aud[:, :, 0].to(DEV).contiguous(),
                                masks={k: v.to(torch.uint8).to(DEV) for k, v in masks.items()},
                                raw=(image[:, 0].to(DEV).contiguous(), audio[:, 0].to(DEV).contiguous()), labels=labels.to(DEV))
    torch.cuda.synchronize()
    w = eng._ws[B]
    assert abs(float(loss[3]) - float(want["loss"])) < 2e-3 * abs(float(want["loss"])), (float(loss[3]), float(want["loss"]))
    assert abs(float(loss[1]) - float(want["aux"])) < 5e-3 * max(abs(float(want["aux"])), 0.1), (float(loss[1]), float(want["aux"]))
    assert _rel(w["s.proj"].view(V, B, -1), want["student_out"]) < 2e-2
    fm, fw = [], []
    groups = [("enc.", want["grads"]["student"]), ("head.", want["grads"]["student_head"]), ("aux_image.", want["grads"]["image"]),
              ("aux_audio.", want["grads"]["audio"])]
    for prefix, gd in groups:
        for k, g in gd.items():
            if g.dim() >= 2:
                fm.append(eng.G[prefix + k].detach().cpu().double().flatten())
                fw.append(g.double().flatten())
    fm, fw = torch.cat(fm), torch.cat(fw)
    cos = float((fm @ fw) / (fm.norm() * fw.norm()))
    print(mode, "B=256 bf16: loss", float(loss[3]), float(want["loss"]), "aux", float(loss[1]), float(want["aux"]), "cos", cos)
    assert cos > 0.97, cos
    for k in ("image_encoder.0.bn1", "image_encoder.0.bn2", "audio_encoder.0.bn1", "audio_encoder.0.bn4"):
        assert _rel(eng.bn_s["enc." + k].running_mean, st.student_buf[k + ".running_mean"]) < 1e-2, k
        assert _rel(eng.bn_s["enc." + k].running_var, st.student_buf[k + ".running_var"]) < 1e-2, k
        assert int(eng.bn_s["enc." + k].num_batches_tracked) == int(st.student_buf[k + ".num_batches_tracked"]) == 7
