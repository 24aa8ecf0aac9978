"""GPU parity of the tcgen05 / TMEM / TMA convolutions (csrc/conv_tc.cu) against torch conv2d on the SAME bf16-rounded
operands (fp32 accumulate on both sides): reference in fp64; tolerance 2e-5 of the output scale for fp32 outputs (accumulation order only), 8e-3 for bf16 outputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from multimodal_ssl_avmnist_b200 import ops

DEV = "cuda"
# (Cin, Cout, H, W, K, pad) forward geometries of the encoders (models/unimodal.py:129-140, 186-208; dino.py:20-30)
FWD = [(8, 16, 56, 56, 5, 2), (16, 32, 28, 28, 5, 2), (32, 64, 14, 14, 5, 2), (32, 64, 14, 14, 5, 0), (32, 64, 14, 14, 3, 1)]


def _bf(x):
    return x.to(torch.bfloat16).float()


def _unpack8(y8):
    N, P, H, W, _ = y8.shape
    return y8.float().permute(0, 1, 4, 2, 3).reshape(N, P * 8, H, W)


def _prep(w, flip=False):
    Co, Ci, K, _ = w.shape
    cin, cout = (Co, Ci) if flip else (Ci, Co)
    buf = torch.empty(ops.conv_tc_weight_bytes(cin, cout, K), dtype=torch.uint8, device=DEV)
    ops.conv_tc_prep_weights(w, buf, flip=flip)
    return buf


@pytest.mark.parametrize("geom", FWD)
@pytest.mark.parametrize("views,B", [(1, 3), (3, 5), (7, 40)])
def test_conv_tc_forward(geom, views, B):
    Cin, Cout, H, W, K, pad = geom
    assert ops.conv_tc_supported(*geom)
    g = torch.Generator().manual_seed(Cin * 131 + Cout + views)
    N = views * B
    x = torch.randn(N, Cin, H, W, generator=g).to(DEV)
    w = (torch.randn(Cout, Cin, K, K, generator=g) / (Cin * K * K) ** 0.5).to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    x8 = torch.empty(N, Cin // 8, H, W, 8, dtype=torch.bfloat16, device=DEV)
    ops.pack_act8(x, x8)
    assert torch.equal(_unpack8(x8), _bf(x))
    want = F.conv2d(_bf(x).double(), _bf(w).double(), b.double(), padding=pad).float()   # fp64: no TF32 / FFT algorithms
    Ho = want.shape[-1]
    wp = _prep(w)
    for out_bf16 in (False, True):
        stats = torch.zeros(views, Cout, 2, dtype=torch.float64, device=DEV)
        if out_bf16:
            out = torch.full((N, Cout // 8, Ho, Ho, 8), float("nan"), dtype=torch.bfloat16, device=DEV)
        else:
            out = torch.full((N, Cout, Ho, Ho), float("nan"), device=DEV)
        ops.conv_tc(x8, wp, b, out, stats, B, Cout, K, pad)
        torch.cuda.synchronize()
        got = _unpack8(out) if out_bf16 else out
        tol = (8e-3 if out_bf16 else 2e-5) * float(want.abs().max())
        err = float((got - want).abs().max())
        assert err <= tol, (geom, out_bf16, err, tol)
        wv = want.view(views, B, Cout, Ho, Ho).double()
        s_want = torch.stack([wv.sum(dim=(1, 3, 4)), (wv * wv).sum(dim=(1, 3, 4))], dim=-1)
        rel = float(((stats - s_want).abs() / (s_want.abs() + 1.0)).max())
        assert rel < 1e-5, (geom, "stats", rel)


@pytest.mark.parametrize("geom", FWD)
def test_conv_tc_data_gradient(geom):
    """dx of conv(x, w): the same kernel with swapped channels, pad' = K-1-pad and flipped weights."""
    Cin, Cout, H, W, K, pad = geom
    Ho = H + 2 * pad - K + 1
    g = torch.Generator().manual_seed(7 + Cin)
    N = 9
    dz = torch.randn(N, Cout, Ho, Ho, generator=g).to(DEV)
    w = (torch.randn(Cout, Cin, K, K, generator=g) / (Cout * K * K) ** 0.5).to(DEV)
    want = F.conv_transpose2d(_bf(dz).double(), _bf(w).double(), padding=pad).float()
    assert want.shape == (N, Cin, H, W)
    dz8 = torch.empty(N, Cout // 8, Ho, Ho, 8, dtype=torch.bfloat16, device=DEV)
    ops.pack_act8(dz, dz8)
    wp = _prep(w, flip=True)
    assert ops.conv_tc_supported(Cout, Cin, Ho, Ho, K, K - 1 - pad)
    out = torch.full((N, Cin, H, W), float("nan"), device=DEV)
    ops.conv_tc(dz8, wp, None, out, None, N, Cin, K, K - 1 - pad)
    torch.cuda.synchronize()
    err = float((out - want).abs().max())
    assert err <= 2e-5 * float(want.abs().max()), (geom, err)


@pytest.mark.parametrize("geom", FWD)
@pytest.mark.parametrize("N", [1, 7, 300])
def test_conv_tc_weight_gradient(geom, N):
    """dW, db against autograd in fp64 on the same bf16-rounded x and dz (K = N*pixels, fp32 accumulate in TMEM)."""
    Cin, Cout, H, W, K, pad = geom
    Ho = H + 2 * pad - K + 1
    g = torch.Generator().manual_seed(11 + Cin + N)
    x = torch.randn(N, Cin, H, W, generator=g).to(DEV)
    dz = torch.randn(N, Cout, Ho, Ho, generator=g).to(DEV)
    xd = _bf(x).double().requires_grad_(False)
    wd = torch.zeros(Cout, Cin, K, K, dtype=torch.float64, device=DEV, requires_grad=True)
    bd = torch.zeros(Cout, dtype=torch.float64, device=DEV, requires_grad=True)
    F.conv2d(xd, wd, bd, padding=pad).backward(_bf(dz).double())
    x8 = torch.empty(N, Cin // 8, H, W, 8, dtype=torch.bfloat16, device=DEV)
    dz8 = torch.empty(N, Cout // 8, Ho, Ho, 8, dtype=torch.bfloat16, device=DEV)
    ops.pack_act8(x, x8)
    ops.pack_act8(dz, dz8)
    work = torch.empty(ops.conv_tc_wgrad_work_floats(N, Cin, Cout, H, W, K, pad), device=DEV)
    dw = torch.full((Cout, Cin, K, K), float("nan"), device=DEV)
    db = torch.full((Cout,), float("nan"), device=DEV)
    ops.conv_tc_wgrad(x8, dz8, dw, db, work, pad)
    torch.cuda.synchronize()
    scale = float(wd.grad.abs().max())
    err = float((dw.double() - wd.grad).abs().max())
    assert err <= 3e-5 * scale, (geom, N, err, scale)
    errb = float((db.double() - bd.grad).abs().max())
    assert errb <= 3e-5 * float(bd.grad.abs().max()), (geom, N, errb)
