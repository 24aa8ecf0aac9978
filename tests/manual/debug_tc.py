import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle.fixtures import make_masks, synth_views, views_to_vb
from multimodal_ssl_avmnist_b200.engine import DinoStepEngine
from multimodal_ssl_avmnist_b200 import ops
DEV = "cuda"
B = 8
def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))
def un8(t):
    N, P, H, W, _ = t.shape
    return t.float().permute(0, 1, 4, 2, 3).reshape(N, P * 8, H, W)
e16 = DinoStepEngine(kind="multi_central", device=DEV, precision="bf16", seed=3)
e32 = DinoStepEngine(kind="multi_central", device=DEV, precision="fp32", seed=3)
e32.student.flat.copy_(e16.student.flat); e32.sync_teacher()
img, aud = views_to_vb(*synth_views(B, seed=100))
masks = {k: v.to(torch.uint8).to(DEV) for k, v in make_masks(seed=200, V=6, Vg=2, B=B, E=256, hidden=512).items()}
xi, xa = img[:, :, 0].to(DEV).contiguous(), aud[:, :, 0].to(DEV).contiguous()
D = (torch.randn(6 * B, 128, generator=torch.Generator().manual_seed(5)) / (6 * B)).to(DEV)
w16 = e16.forward_pass(xi, xa, masks=masks)
w32 = e32.forward_pass(xi, xa, masks=masks)
for mod, nl in (("img", 2), ("aud", 4)):
    for li in range(nl):
        z16 = w16[f"s.{mod}.z{li}"]; z32 = w32[f"s.{mod}.z{li}"]
        z16 = un8(z16) if z16.dim() == 5 else z16
        print(mod, li, "z", rel(z16, z32), "mean", rel(w16[f"s.{mod}.mean{li}"], w32[f"s.{mod}.mean{li}"]), "invstd", rel(w16[f"s.{mod}.invstd{li}"], w32[f"s.{mod}.invstd{li}"]))
        k = f"s.{mod}.p{li}" if f"s.{mod}.p{li}" in w16 else None
        p16 = w16[k] if k else un8(w16[f"s.{mod}.p8{li}"])
        print(mod, li, "p", rel(p16, w32[f"s.{mod}.p{li}"]))
print("proj", rel(w16["s.proj"], w32["s.proj"]))
# hook: record dz / d_in by re-running pieces: monkeypatch ops to capture
cap = {}
def wrap(name, eng_tag):
    fn = getattr(ops, name)
    def f(*a, **k):
        r = fn(*a, **k)
        cap.setdefault((eng_tag, name), []).append([t.clone() if isinstance(t, torch.Tensor) else t for t in a])
        return r
    return fn, f
import multimodal_ssl_avmnist_b200.engine as E
for tag, eng, w in (("16", e16, w16), ("32", e32, w32)):
    saved = {}
    for name in ("conv_tc_wgrad", "conv_bwd_weight", "bn_relu_pool8_bwd_apply", "bn_relu_pool_bwd_apply", "conv_tc", "conv_bwd_data"):
        saved[name], f = wrap(name, tag)
        setattr(E.ops, name, f)
    eng.backward_pass(w, d_proj=D)
    torch.cuda.synchronize()
    for name, fn in saved.items():
        setattr(E.ops, name, fn)
# order of calls in backward: img stack (li=1, 0) then aud stack (3,2,1,0)
a16 = cap[("16", "bn_relu_pool8_bwd_apply")]; a32 = cap[("32", "bn_relu_pool_bwd_apply")]
tc_i = 0
for i, c32 in enumerate(a32):
    z32, dp32, dz32 = c32[0], c32[1], c32[7]
    # find the matching bf16 call by z shape
    m = [c for c in a16 if c[0].dim() == 5 and un8(c[0]).shape == z32.shape]
    if not m:
        continue
    c16 = m[0]
    dp16 = c16[1]
    dp16 = un8(dp16) if dp16.dim() == 5 else dp16.reshape(dp32.shape)
    print("layer z", tuple(z32.shape), "dp rel", rel(dp16, dp32.reshape(dp16.shape)), "sums rel", rel(c16[6], c32[6]), "dz rel", rel(un8(c16[7]), dz32))
n = e16.n_trainable_prefix
for k in ("enc.image_encoder.0.conv2.weight", "enc.image_encoder.0.conv1.weight", "enc.audio_encoder.0.conv4.weight", "enc.audio_encoder.0.conv3.weight", "enc.audio_encoder.0.conv2.weight", "enc.audio_encoder.0.conv1.weight", "enc.image_encoder.1.weight", "enc.fusion.0.weight", "head.mlp.0.weight"):
    print(k, rel(e16.G[k], e32.G[k]))
