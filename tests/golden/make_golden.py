"""Generates tests/golden/*.npz|json by running the UNMODIFIED reference (imported from /root/reference through
oracle/refshim) on seeded inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The fixtures (not this script) are what the test-suite reads: /root/reference does not exist on the GPU box.
Weights are the deterministic scheme of oracle.dino_ref.make_params (loaded into the reference modules), and
dropout masks are drawn by oracle.dino_ref.make_masks through a patched torch.nn.functional.dropout, so the
oracle and the CUDA path can regenerate both from a seed instead of storing 26 MB of tensors.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
sys.path.insert(1, "/root/reference/AVMNIST_Experiments")
sys.path.insert(2, ROOT)

import numpy as np
import torch
import yaml

torch.set_num_threads(1)

import models.dino as md                      # noqa: E402  (reference)
import utils.get_data as gd                   # noqa: E402  (reference)
from hyperparameter_tuning.objective_augment import process_augment_config  # noqa: E402 (reference)

from oracle import dino_ref as R              # noqa: E402
from oracle.fixtures import make_masks, summarize, synth_views, synth_raw  # noqa: E402

CFG_DIR = "/root/reference/AVMNIST_Experiments/configs"


# ---------------------------------------------------------------------------------------------------------
def losses_kat():
    torch.manual_seed(7)
    S = torch.randn(6, 4, 128)
    T = torch.randn(2, 4, 128)

    class M(md.MultiModalDINOLightning):
        def __init__(self):
            md.pl.LightningModule.__init__(self)
            self.student_temperature, self.teacher_temperature = 0.1, 0.04

    class U(md.UniModalDINOLightning):
        def __init__(self):
            md.pl.LightningModule.__init__(self)
            self.student_temperature, self.teacher_temperature = 0.1, 0.04

    out = {
        "dino_multimodal": float(M().dino_loss(S, T)),
        "dino_unimodal": float(U().dino_loss(S, T)),
        "infonce": float(md.MultiModalDINOWithINFONCELightning.infoNCE_loss(None, S[0], S[1])),
        "mse": float(md.MultiModalDINOWithMSELightning.mse_loss(None, S[0], S[1])),
        "cosine_consistency": float(U()._cosine_consistency_loss(S)),
    }
    logits_i, logits_a = S[0][:, :10], S[1][:, :10]
    labels = torch.tensor([3, 1, 4, 1])
    ce = torch.nn.CrossEntropyLoss()
    out["supervised"] = float(ce(logits_i, labels) + ce(logits_a, labels))
    # SimCLR NT-Xent: the reference's own method on cat([z1, z2])
    import other_ssl.multimodal_simclr.multimodal_simclr as simclr
    reps = torch.cat([S[0], S[1]]).clone().requires_grad_(True)
    nt = simclr.MultiModalSimCLRLightning.nt_xent_loss(None, reps)
    nt.backward()
    out["ntxent"] = float(nt)
    out["ntxent_grad_abs_sum"] = float(reps.grad.abs().sum())
    out["ntxent_grad_first8"] = reps.grad.flatten()[:8].tolist()
    # gradient KAT for the fused DINO loss backward
    Sg = S.clone().requires_grad_(True)
    M().dino_loss(Sg, T).backward()
    out["dino_multimodal_grad_abs_sum"] = float(Sg.grad.abs().sum())
    out["dino_multimodal_grad_first8"] = Sg.grad.flatten()[:8].tolist()
    # EMA / centre KATs (update_teacher / update_center arithmetic)
    g = torch.Generator().manual_seed(11)
    t = torch.randn(1000, generator=g)
    s = torch.randn(1000, generator=g)
    ema = 0.996 * t + (1 - 0.996) * s
    out["ema_first8"] = ema[:8].tolist()
    out["ema_sum64"] = float(ema.double().sum())
    return out


# ---------------------------------------------------------------------------------------------------------
def augment_fixtures():
    res = {}
    arrays = {}
    for tag, cfgname in (("tuned", "config_multimodal_dino.yaml"), ("old", "config_multimodal_dino_old_augments.yaml"),
                         ("default", None)):
        av = None
        if cfgname is not None:
            av = process_augment_config(None, yaml.safe_load(open(os.path.join(CFG_DIR, cfgname))), False)
        aug = gd.MultiModalAugmentation(augment_values=av)
        sums = []
        for s in range(8):
            g = torch.Generator().manual_seed(1000 + s)
            img = torch.rand(1, 28, 28, generator=g)
            aud = torch.rand(1, 112, 112, generator=g)
            torch.manual_seed(s)
            random.seed(s)
            gi, ga, li, la = aug(img, aud)
            sums.append([float(x.double().sum()) for x in (gi, ga, li, la)])
            if s < 2:
                arrays[f"{tag}_{s}_gi"] = gi.numpy()
                arrays[f"{tag}_{s}_li"] = li.numpy()
                # audio views are large: keep a 4x-decimated grid + row/col sums
                for nm, a in (("ga", ga), ("la", la)):
                    a = a.numpy()
                    arrays[f"{tag}_{s}_{nm}_dec"] = a[:, :, ::4, 1::4].copy()
                    arrays[f"{tag}_{s}_{nm}_rows"] = a.astype(np.float64).sum(-1)
                    arrays[f"{tag}_{s}_{nm}_cols"] = a.astype(np.float64).sum(-2)
        res[tag] = {"config": cfgname, "sums": sums}
    # SURVEY §8c known answer (old_augments, seed 1 / data seed 1234)
    av = process_augment_config(None, yaml.safe_load(open(os.path.join(CFG_DIR, "config_multimodal_dino_old_augments.yaml"))), False)
    g = torch.Generator().manual_seed(1234)
    img = torch.rand(1, 28, 28, generator=g)
    aud = torch.rand(1, 112, 112, generator=g)
    torch.manual_seed(1)
    random.seed(1)
    outs = gd.MultiModalAugmentation(augment_values=av)(img, aud)
    res["survey_kat_sums"] = [float(x.sum()) for x in outs]
    np.savez_compressed(os.path.join(HERE, "augment.npz"), **arrays)
    return res


# ---------------------------------------------------------------------------------------------------------
class _DropoutPatch:
    """Replaces torch.nn.functional.dropout so that masks come from oracle.fixtures.make_masks (call order:
    student fusion x V, teacher fusion x Vg, student head; p == 0 calls pass through)."""

    def __init__(self, masks_in_order):
        self.queue = list(masks_in_order)
        self.orig = torch.nn.functional.dropout

    def __enter__(self):
        def patched(x, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return x
            keep = self.queue.pop(0)
            assert keep.shape == x.shape, (keep.shape, x.shape)
            return x * keep.to(x.dtype) / (1.0 - p)
        torch.nn.functional.dropout = patched
        torch.nn.modules.dropout.F.dropout = patched
        return self

    def __exit__(self, *a):
        torch.nn.functional.dropout = self.orig
        torch.nn.modules.dropout.F.dropout = self.orig
        assert not self.queue, "unused dropout masks"


def _load(module, prefix_params, buffers=None):
    sd = module.state_dict()
    for k, v in prefix_params.items():
        assert k in sd and sd[k].shape == v.shape, k
        sd[k] = v.clone()
    module.load_state_dict(sd)


ENCODER_CLASSES = {"multi_central": "CentralMultiModalEncoder", "multi_simple": "SimpleMultiModalEncoder",
                   "multi_simple_gated": "GatedMultiModalEncoder", "multi_cross_attention": "CrossAttentionMultiModalEncoder"}


def step_fixture(mode, B=4, seed=5, n_steps=2, kind="multi_central"):
    st = R.CentralDinoState(seed=seed, mode=mode, kind=kind)
    cls = {"default": md.MultiModalDINO, "semi_supervised": md.MultiModalDINOSemiSupervised,
           "infonce": md.MultiModalDINOWithINFONCE, "mse": md.MultiModalDINOWithMSE}[mode]
    lcls = {"default": md.MultiModalDINOLightning, "semi_supervised": md.MultiModalDINOSemiSupervisedLightning,
            "infonce": md.MultiModalDINOWithINFONCELightning, "mse": md.MultiModalDINOWithMSELightning}[mode]
    torch.manual_seed(0)
    model = cls(encoder_class=getattr(md, ENCODER_CLASSES[kind]), output_dim=256, encoder_output_dim=256, projection_dim=128,
                momentum=0.996, center_momentum=0.9, dropout=0.3)
    _load(model.student, st.student)
    _load(model.teacher, st.teacher)
    _load(model.student_projection, st.student_head)
    _load(model.teacher_projection, st.teacher_head)
    if mode == "semi_supervised":
        _load(model.image_classifier, st.aux["image"])
        _load(model.audio_classifier, st.aux["audio"])
    elif mode != "default":
        _load(model.image_projection_head, st.aux["image"])
        _load(model.audio_projection_head, st.aux["audio"])

    class L(lcls):
        def __init__(self, m):
            md.pl.LightningModule.__init__(self)
            self.model = m
            self.student_temperature, self.teacher_temperature = 0.1, 0.04
            self.alpha = 1
            self.ce_loss = torch.nn.CrossEntropyLoss()
            self.learning_rate, self.weight_decay, self.num_epochs = 1e-4, 1e-6, 100

    lit = L(model)
    lit.train()
    opt = torch.optim.Adam(lit.parameters(), lr=1e-4, weight_decay=1e-6)
    out = {"mode": mode, "kind": kind, "B": B, "seed": seed, "steps": []}
    for it in range(n_steps):
        gi, ga, li, la = synth_views(B, seed=100 + it)
        m = make_masks(seed=200 + it, V=6, Vg=2, B=B, E=256, hidden=512)
        order = [m["student_fusion"][v] for v in range(6)] + [m["teacher_fusion"][v] for v in range(2)] + [m["student_head"]]
        views = (gi, ga, li, la)
        if mode == "default":
            batch = views
        else:
            image, audio, labels = synth_raw(B, seed=300 + it)
            batch = (image, audio, labels, views)
        opt.zero_grad()
        with _DropoutPatch(order):
            loss = lit.training_step(batch, it)
        loss.backward()
        rec = {"loss": float(loss), "center": summarize(model.center)}
        grads = {}
        for n, p in lit.named_parameters():
            if p.grad is not None:
                grads[n] = summarize(p.grad)
        rec["grads"] = grads
        opt.step()
        rec["teacher"] = {n: summarize(p) for n, p in model.teacher.named_parameters() if "fc" not in n}
        rec["teacher_head"] = {n: summarize(p) for n, p in model.teacher_projection.named_parameters()}
        rec["student_after_adam"] = {n: summarize(p) for n, p in model.student.named_parameters() if "fc" not in n}
        rec["student_bn"] = {n: summarize(b) for n, b in model.student.named_buffers()}
        rec["teacher_bn"] = {n: summarize(b) for n, b in model.teacher.named_buffers()}
        rec["student_head_bn"] = {n: summarize(b) for n, b in model.student_projection.named_buffers()}
        out["steps"].append(rec)
    return out


def unimodal_fixture(B=4, seed=9, alpha=0):
    spec = R.image_simple_spec(256)
    hs = R.head_spec(256, 128)
    sp = R.make_params(spec, seed)
    hp = R.make_params(hs, seed + 1)
    torch.manual_seed(0)
    model = md.UniModalDINO(encoder_class=md.ImageEncoder, output_dim=256, projection_dim=128, dropout=0.3)
    _load(model.student, sp)
    _load(model.teacher, sp)
    _load(model.student_projection, hp)
    _load(model.teacher_projection, hp)

    class L(md.UniModalDINOLightning):
        def __init__(self, m):
            md.pl.LightningModule.__init__(self)
            self.model = m
            self.student_temperature, self.teacher_temperature = 0.1, 0.04
            self.cosine_loss_alpha = alpha

    lit = L(model)
    lit.train()
    opt = torch.optim.Adam(lit.parameters(), lr=1e-4, weight_decay=1e-6)
    gi, ga, li, la = synth_views(B, seed=100)
    m = make_masks(seed=200, V=6, Vg=2, B=B, E=256, hidden=512)
    with _DropoutPatch([m["student_head"]]):
        loss = lit.training_step((gi, ga, li, la), 0)
    loss.backward()
    rec = {"B": B, "seed": seed, "loss": float(loss), "center": summarize(model.center),
           "grads": {n: summarize(p.grad) for n, p in lit.named_parameters() if p.grad is not None}}
    opt.step()
    rec["teacher"] = {n: summarize(p) for n, p in model.teacher.named_parameters()}
    rec["student_after_adam"] = {n: summarize(p) for n, p in model.student.named_parameters()}
    return rec


def contrastive_fixture(kind, B=4, seed=21, modes=(2, 0, 1, 3)):
    """training_step -> backward -> Adam.step of the imported MultiModalInfoNCELightning (other_ssl/info_nce/info_nce.py) or
    MultiModalSimCLRLightning (other_ssl/multimodal_simclr/multimodal_simclr.py) with the oracle's deterministic weights.  SimCLR draws its
    modality pairing with torch.randint(0, 4, (1,)): patched here to the recorded `modes`, one per step."""
    st = R.ContrastiveState(seed=seed)
    if kind == "infonce":
        import other_ssl.info_nce.info_nce as mod
        lit = mod.MultiModalInfoNCELightning(projection_dim=256, output_dim=256)
        n_steps = 2
    else:
        import other_ssl.multimodal_simclr.multimodal_simclr as mod
        lit = mod.MultiModalSimCLRLightning(projection_dim=256, output_dim=256)
        n_steps = len(modes)
    for m in R.CONTRASTIVE_MODULES:
        _load(getattr(lit.model, m), st.params[m])
    lit.train()
    opt = torch.optim.Adam(lit.parameters(), lr=1e-4)
    out = {"kind": kind, "B": B, "seed": seed, "modes": list(modes), "steps": [],
           "state_dict": {k: list(v.shape) for k, v in lit.state_dict().items()}}          # API surface: keys / shapes of the Lightning module
    orig_randint = torch.randint
    for it in range(n_steps):
        g = torch.Generator().manual_seed(400 + it)
        img1, spec1 = torch.rand(B, 1, 28, 28, generator=g), torch.rand(B, 1, 112, 112, generator=g)
        img2, spec2 = torch.rand(B, 1, 28, 28, generator=g), torch.rand(B, 1, 112, 112, generator=g)
        opt.zero_grad(set_to_none=True)
        if kind == "infonce":
            loss = lit.training_step((img1, spec1, torch.zeros(B, dtype=torch.long)), it)
        else:
            torch.randint = lambda *a, **k: torch.tensor([modes[it]])
            try:
                loss = lit.training_step((img1, spec1, img2, spec2), it)
            finally:
                torch.randint = orig_randint
        loss.backward()
        rec = {"loss": float(loss.detach()), "grads": {n: summarize(p.grad) for n, p in lit.model.named_parameters() if p.grad is not None}}
        opt.step()
        rec["params_after_adam"] = {n: summarize(p) for n, p in lit.model.named_parameters()}
        rec["bn"] = {n: summarize(b) for n, b in lit.model.named_buffers()}
        out["steps"].append(rec)
    return out


def api_surface():
    """state_dict keys/shapes of the reference modules and seeded-initialisation checksums (API-compatibility fixtures)."""
    out = {}
    specs = {"default": (md.MultiModalDINO, {}), "semi_supervised": (md.MultiModalDINOSemiSupervised, {}),
             "infonce": (md.MultiModalDINOWithINFONCE, {}), "mse": (md.MultiModalDINOWithMSE, {})}
    for name, (cls, kw) in specs.items():
        torch.manual_seed(123)
        m = cls(encoder_class=md.CentralMultiModalEncoder, output_dim=256, encoder_output_dim=256, projection_dim=128, **kw)
        out[name] = {"keys": {k: list(v.shape) for k, v in m.state_dict().items()},
                     "init": {k: summarize(v) for k, v in m.state_dict().items() if v.dtype == torch.float32 and ("conv2.weight" in k or "mlp.0.weight" in k or "fusion.3.bias" in k)}}
    torch.manual_seed(123)
    u = md.UniModalDINO(encoder_class=md.ImageEncoder, output_dim=256, projection_dim=128)
    out["unimodal"] = {"keys": {k: list(v.shape) for k, v in u.state_dict().items()},
                       "init": {k: summarize(v) for k, v in u.state_dict().items() if "encoder.4.weight" in k or "mlp.4.weight" in k}}
    return out


def simclr_fixtures(n_seeds=8, B=3):
    """SimCLRMultiModalAugmentation of the reference (utils/get_data.py:299-408) on seeded batches -> tests/golden/simclr_aug.npz.
    The transforms are applied to the whole [B,1,H,W] batch at once: one parameter set per call, GaussianNoise per element."""
    arrays = {}
    aug = gd.SimCLRMultiModalAugmentation()
    for s in range(n_seeds):
        g = torch.Generator().manual_seed(2000 + s)
        img = torch.rand(B, 1, 28, 28, generator=g)
        aud = torch.rand(B, 1, 112, 112, generator=g)
        torch.manual_seed(s)
        random.seed(s)
        i1, a1, i2, a2 = aug(img, aud)
        arrays[f"s{s}_i1"], arrays[f"s{s}_i2"] = i1.numpy(), i2.numpy()
        for nm, a in (("a1", a1), ("a2", a2)):
            a = a.numpy()
            arrays[f"s{s}_{nm}_dec"] = a[:, :, ::4, 1::4].copy()
            arrays[f"s{s}_{nm}_rows"] = a.astype(np.float64).sum(-1)
    np.savez_compressed(os.path.join(HERE, "simclr_aug.npz"), **arrays)
    return sorted(arrays)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "simclr":
        print("wrote simclr_aug.npz:", len(simclr_fixtures()), "arrays")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "simple":
        # the other conv encoders (models/dino.py:214-263, 385-452) -> tests/golden/golden_simple.json
        fx = {f"{k}/{m}": step_fixture(m, kind=k) for k, m in (("multi_simple", "default"), ("multi_simple", "mse"), ("multi_simple_gated", "default"),
                                                              ("multi_simple_gated", "semi_supervised"), ("multi_cross_attention", "default"),
                                                              ("multi_cross_attention", "infonce"))}
        fx["versions"] = {"torch": torch.__version__}
        with open(os.path.join(HERE, "golden_simple.json"), "w") as f:
            json.dump(fx, f, indent=1)
        print("wrote golden_simple.json:", sorted(fx))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "contrastive":
        # the stand-alone contrastive steps (other_ssl/info_nce, other_ssl/multimodal_simclr) -> tests/golden/golden_contrastive.json
        fx = {"infonce": contrastive_fixture("infonce"), "simclr": contrastive_fixture("simclr"), "versions": {"torch": torch.__version__}}
        with open(os.path.join(HERE, "golden_contrastive.json"), "w") as f:
            json.dump(fx, f, indent=1)
        print("wrote golden_contrastive.json")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "api":
        with open(os.path.join(HERE, "api_surface.json"), "w") as f:
            json.dump(api_surface(), f, indent=1)
        sys.exit(0)
    fx = {"losses": losses_kat(), "augment": augment_fixtures()}
    fx["steps"] = {m: step_fixture(m) for m in ("default", "semi_supervised", "infonce", "mse")}
    fx["unimodal_image_simple"] = unimodal_fixture()
    fx["unimodal_image_simple_cosine"] = unimodal_fixture(alpha=0.3)          # + cosine consistency term (models/dino.py:1651-1656)
    fx["versions"] = {"torch": torch.__version__}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(fx, f, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))
