"""Seeded synthetic inputs, dropout masks and tensor summaries shared by the golden generator, the oracle tests
and the GPU parity tests.  TEST INFRASTRUCTURE ONLY."""
import torch


def synth_views(B, seed, Vg=2, Vl=4):
    """AVMNIST-shaped multi-crop batch in the reference's collated layout [B,V,1,H,W]."""
    g = torch.Generator().manual_seed(seed)
    gi = torch.rand(B, Vg, 1, 28, 28, generator=g)
    ga = torch.rand(B, Vg, 1, 112, 112, generator=g)
    li = torch.rand(B, Vl, 1, 28, 28, generator=g)
    la = torch.rand(B, Vl, 1, 112, 112, generator=g)
    return gi, ga, li, la


def synth_raw(B, seed):
    """Un-augmented (image, audio, label) triple of the Extended dataset (utils/get_data.py:498-509)."""
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(B, 1, 28, 28, generator=g)
    audio = torch.randint(0, 256, (B, 1, 112, 112), generator=g).float() / 255.0
    labels = torch.randint(0, 10, (B,), generator=g)
    return image, audio, labels


def views_to_vb(gi, ga, li, la):
    """[B,V,1,H,W] x4 -> view-major ([V,B,1,28,28], [V,B,1,112,112]), global views first."""
    img = torch.cat([gi, li], dim=1).transpose(0, 1).contiguous()
    aud = torch.cat([ga, la], dim=1).transpose(0, 1).contiguous()
    return img, aud


def make_masks(seed, V, Vg, B, E, hidden, p_fusion=0.3, p_head=0.3):
    """Dropout keep-masks in the reference's call order: student fusion per view, teacher fusion per global
    view, student head (models/dino.py:225, 1247)."""
    g = torch.Generator().manual_seed(seed)
    m = {"student_fusion": [], "teacher_fusion": []}
    for _ in range(V):
        m["student_fusion"].append(torch.rand(B, E, generator=g) >= p_fusion)
    for _ in range(Vg):
        m["teacher_fusion"].append(torch.rand(B, E, generator=g) >= p_fusion)
    m["student_head"] = torch.rand(V * B, hidden, generator=g) >= p_head
    m["student_fusion"] = torch.stack(m["student_fusion"])
    m["teacher_fusion"] = torch.stack(m["teacher_fusion"])
    return m


def summarize(t):
    t = t.detach().double().flatten()
    return {"n": int(t.numel()), "sum": float(t.sum()), "abs_sum": float(t.abs().sum()),
            "first": t[:6].tolist(), "last": t[-2:].tolist()}


def summaries_close(a, b, rtol, atol):
    """Compare two `summarize` dicts; tolerances apply per element (first/last) and scale with n for the sums."""
    if a["n"] != b["n"]:
        return False, "numel"
    scale = max(a["abs_sum"], b["abs_sum"])
    if abs(a["sum"] - b["sum"]) > rtol * scale + atol * a["n"]:
        return False, f"sum {a['sum']} vs {b['sum']}"
    if abs(a["abs_sum"] - b["abs_sum"]) > rtol * scale + atol * a["n"]:
        return False, f"abs_sum {a['abs_sum']} vs {b['abs_sum']}"
    mean_abs = scale / max(a["n"], 1)
    for x, y in zip(a["first"] + a["last"], b["first"] + b["last"]):
        if abs(x - y) > rtol * max(abs(x), abs(y), mean_abs) + atol:
            return False, f"value {x} vs {y}"
    return True, ""


def contrastive_batch(B, it):
    """The seeded (img1, spec1, img2, spec2) batch of step `it` of tests/golden/make_golden.py::contrastive_fixture."""
    g = torch.Generator().manual_seed(400 + it)
    img1, spec1 = torch.rand(B, 1, 28, 28, generator=g), torch.rand(B, 1, 112, 112, generator=g)
    img2, spec2 = torch.rand(B, 1, 28, 28, generator=g), torch.rand(B, 1, 112, 112, generator=g)
    return img1, spec1, img2, spec2
