"""Debug helper: error pattern of the x-phase-packed tensor-core convolution against an fp64 conv2d."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from multimodal_ssl_avmnist_b200 import ops

DEV = "cuda"
geom = tuple(int(a) for a in sys.argv[1:7]) if len(sys.argv) > 6 else (8, 16, 56, 56, 5, 2)
Cin, Cout, H, W, K, pad = geom
N = 2
g = torch.Generator().manual_seed(1)
mode = sys.argv[7] if len(sys.argv) > 7 else "rand"
x = torch.randn(N, Cin, H, W, generator=g).to(DEV)
w = (torch.randn(Cout, Cin, K, K, generator=g) / (Cin * K * K) ** 0.5).to(DEV)
if mode == "delta":
    x.zero_(); x[:, 0, 10, 10] = 1.0
b = torch.zeros(Cout, device=DEV)
if Cin == 1:
    x = x.abs()
    if mode.startswith("delta"):
        parts = mode.split(":")
        dy, dx = (int(parts[1]), int(parts[2])) if len(parts) > 2 else (10, 10)
        x.zero_(); x[:, 0, dy, dx] = 1.0
    if mode == "row":
        x.zero_(); x[:, 0, 10, :] = 1.0
    if mode == "col":
        x.zero_(); x[:, 0, :, 10] = 1.0
    x8 = torch.empty(N, H, ops.quad8_width(W, pad), 8, dtype=torch.bfloat16, device=DEV)
    ops.pack_quad8(x, x8, pad)
else:
    x8 = torch.empty(N, Cin // 8, H, W, 8, dtype=torch.bfloat16, device=DEV)
    ops.pack_act8(x, x8)
bf = lambda t: t.to(torch.bfloat16).float()
want = F.conv2d(bf(x).double(), bf(w).double(), b.double(), padding=pad).float()
Ho = want.shape[-1]
wp = torch.empty(ops.conv_tc_weight_bytes(Cin, Cout, K), dtype=torch.uint8, device=DEV)
ops.conv_tc_prep_weights(w, wp)
out = torch.full((N, Cout, Ho, Ho), float("nan"), device=DEV)
ops.conv_tc(x8, wp, b, out, None, N, Cout, K, pad)
torch.cuda.synchronize()
err = (out - want).abs()
print("max err", float(err.max()), "scale", float(want.abs().max()), "nan", int(torch.isnan(out).sum()))
e = err[0]
print("err by channel", [round(float(v), 3) for v in e.amax(dim=(1, 2))])
print("err by x (first 16)", [round(float(v), 3) for v in e.amax(dim=(0, 1))[:16]])
print("err by y (first 16)", [round(float(v), 3) for v in e.amax(dim=(0, 2))[:16]])
if mode != "rand":
    bad = (err[0, 0] > 1e-3).nonzero()
    print("bad positions ch0 (first 40):", bad[:40].tolist(), "count", len(bad))
    for (yy, xx) in bad[:6].tolist():
        print("   at", yy, xx, "got", float(out[0, 0, yy, xx]), "want", float(want[0, 0, yy, xx]))
if mode == "delta":
    nz = (out[0, 0].abs() > 1e-6).nonzero()
    print("nonzero got ch0:", nz[:30].tolist())
    nz = (want[0, 0].abs() > 1e-6).nonzero()
    print("nonzero want ch0:", nz[:30].tolist())
if mode == "col":
    torch.set_printoptions(linewidth=250, precision=3, sci_mode=False)
    print("diff[0, :, 60, 4:20]"); print((out - want)[0, :, 60, 4:20])
    print("diff[0, 0, 50:60, 4:20]"); print((out - want)[0, 0, 50:60, 4:20])
    print("diff[0, 0, 104:112, 4:20]"); print((out - want)[0, 0, 104:112, 4:20])
