"""Oracle (CPU, plain torch ops) restatement of the DINO training step.  TEST INFRASTRUCTURE ONLY.

Functional style: parameters and buffers live in flat dicts keyed by the reference's state_dict names
(`student.image_encoder.0.conv1.weight`, ...), so weights can be exchanged with the imported reference.
Gradients come from torch autograd on CPU.  Dropout masks are injected (dict of keep-masks) so the CUDA path
and the oracle can share them.

Reference citations are relative to /root/reference/AVMNIST_Experiments/.
"""
import math

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------------------
# parameter inventories (names + shapes), in the reference's `.parameters()` order
# ------------------------------------------------------------------------------------------------------------

def _conv(name, cout, cin, k):
    return [(f"{name}.weight", (cout, cin, k, k)), (f"{name}.bias", (cout,))]


def _bn(name, c):
    return [(f"{name}.weight", (c,)), (f"{name}.bias", (c,))]


def _lin(name, out_f, in_f):
    return [(f"{name}.weight", (out_f, in_f)), (f"{name}.bias", (out_f,))]


def central_encoder_spec(E=256, O=256):
    """CentralMultiModalEncoder (models/dino.py:454-468) with CentralUnimodalImage/Audio
    (models/unimodal.py:105-125,155-183); includes the unused fc1/fc2 heads (they are parameters)."""
    s = []
    p = "image_encoder.0"
    s += _conv(f"{p}.conv1", 32, 1, 5) + _bn(f"{p}.bn1", 32) + _conv(f"{p}.conv2", 64, 32, 5) + _bn(f"{p}.bn2", 64)
    s += _lin(f"{p}.fc1", 1024, 1600) + _lin(f"{p}.fc2", 10, 1024)
    s += _lin("image_encoder.1", E, 1600)
    p = "audio_encoder.0"
    s += _conv(f"{p}.conv1", 8, 1, 5) + _bn(f"{p}.bn1", 8) + _conv(f"{p}.conv2", 16, 8, 5) + _bn(f"{p}.bn2", 16)
    s += _conv(f"{p}.conv3", 32, 16, 5) + _bn(f"{p}.bn3", 32) + _conv(f"{p}.conv4", 64, 32, 5) + _bn(f"{p}.bn4", 64)
    s += _lin(f"{p}.fc1", 1024, 3136) + _lin(f"{p}.fc2", 10, 1024)
    s += _lin("audio_encoder.1", E, 3136)
    s += _lin("fusion.0", E, 2 * E) + _lin("fusion.3", O, E)
    return s


def central_bn_names():
    return ["image_encoder.0.bn1", "image_encoder.0.bn2", "audio_encoder.0.bn1", "audio_encoder.0.bn2",
            "audio_encoder.0.bn3", "audio_encoder.0.bn4"]


def image_simple_spec(O=256):
    """ImageEncoder (models/dino.py:483-499) over image_encoder(512) (models/dino.py:18-41)."""
    s = []
    s += _conv("encoder.0", 32, 1, 3) + _bn("encoder.1", 32)
    s += _conv("encoder.4", 64, 32, 3) + _bn("encoder.5", 64)
    s += _conv("encoder.8", 128, 64, 3) + _bn("encoder.9", 128)
    s += _lin("encoder.14", 512, 128) + _lin("projection.0", O, 512)
    return s


def image_simple_bn_names():
    return ["encoder.1", "encoder.5", "encoder.9"]


def simple_multi_spec(E=256, O=256, mix=None):
    """SimpleMultiModalEncoder (models/dino.py:214-234) over image_encoder(E) / audio_encoder(E) (models/dino.py:18-73);
    mix="gated": GatedMultiModalEncoder (:237-263); mix="cross": CrossAttentionMultiModalEncoder (:407-452, CrossModalAttention :385-405)."""
    s = []
    for i, (ci, co) in zip((0, 4, 8), ((1, 32), (32, 64), (64, 128))):
        s += _conv(f"image_encoder.{i}", co, ci, 3) + _bn(f"image_encoder.{i + 1}", co)
    s += _lin("image_encoder.14", E, 128)
    for i, (ci, co) in zip((0, 4, 8, 12), ((1, 32), (32, 64), (64, 128), (128, 256))):
        s += _conv(f"audio_encoder.{i}", co, ci, 3) + _bn(f"audio_encoder.{i + 1}", co)
    s += _lin("audio_encoder.18", E, 256)
    s += _lin("fusion.0", E, 2 * E) + _lin("fusion.3", O, E)
    if mix == "gated":
        s += [("gate_image", ()), ("gate_audio", ())]
    elif mix == "cross":
        for a in ("image_to_audio_attention", "audio_to_image_attention"):
            s += _lin(f"{a}.q_proj", E, E) + _lin(f"{a}.kv_proj", 2 * E, E)
    return s


def simple_multi_bn_names():
    return ["image_encoder.1", "image_encoder.5", "image_encoder.9", "audio_encoder.1", "audio_encoder.5", "audio_encoder.9",
            "audio_encoder.13"]


KIND_MIX = {"multi_central": None, "multi_simple": None, "multi_simple_gated": "gated", "multi_cross_attention": "cross"}


def encoder_spec(kind, E=256, O=256):
    return central_encoder_spec(E, O) if kind == "multi_central" else simple_multi_spec(E, O, KIND_MIX[kind])


def encoder_bn_names(kind):
    return central_bn_names() if kind == "multi_central" else simple_multi_bn_names()


def head_spec(in_dim, out_dim, hidden=512):
    """ProjectionHead (models/dino.py:1240-1254)."""
    return _lin("mlp.0", hidden, in_dim) + _bn("mlp.1", hidden) + _lin("mlp.4", out_dim, hidden)


def bn_channels(spec, bn):
    return dict(spec)[f"{bn}.weight"][0]


def make_params(spec, seed, dtype=torch.float32):
    """Deterministic test weights (NOT the reference's init): U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv /
    linear weights and biases, 1+0.1*N(0,1) for BN weight, 0.1*N(0,1) for BN bias.  Same scheme is loaded into
    the imported reference by tests/golden/make_golden.py."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    fan = None
    for name, shape in spec:
        if len(shape) > 1:
            fan = math.prod(shape[1:])
            b = 1.0 / math.sqrt(fan)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b
        elif name.endswith(".bias") and fan is not None and f"{name[:-5]}.weight" in out and out[f"{name[:-5]}.weight"].dim() > 1:
            b = 1.0 / math.sqrt(fan)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b
        elif name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif len(shape) == 0:           # the scalar gates of GatedMultiModalEncoder
            t = 0.5 + 0.3 * torch.randn(shape, generator=g, dtype=torch.float64)
        else:
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        out[name] = t.to(dtype)
    return out


def make_bn_buffers(spec, bn_names, dtype=torch.float32):
    buf = {}
    for bn in bn_names:
        c = bn_channels(spec, bn)
        buf[f"{bn}.running_mean"] = torch.zeros(c, dtype=dtype)
        buf[f"{bn}.running_var"] = torch.ones(c, dtype=dtype)
        buf[f"{bn}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return buf


# ------------------------------------------------------------------------------------------------------------
# network forward (train-mode BatchNorm: batch statistics per call + running-stat update)
# ------------------------------------------------------------------------------------------------------------

def _bn_train(z, p, buf, bn):
    out = F.batch_norm(z, buf[f"{bn}.running_mean"], buf[f"{bn}.running_var"], p[f"{bn}.weight"], p[f"{bn}.bias"],
                       training=True, momentum=0.1, eps=1e-5)
    buf[f"{bn}.num_batches_tracked"] += 1
    return out


def _bn_eval(z, p, buf, bn):
    return F.batch_norm(z, buf[f"{bn}.running_mean"], buf[f"{bn}.running_var"], p[f"{bn}.weight"], p[f"{bn}.bias"],
                        training=False, eps=1e-5)


TRACE = None        # tests may set this to a dict: conv name -> list of pre-BatchNorm outputs z, one per call (view-call order)


def _block(x, p, buf, conv, bn, pad, train=True):
    z = F.conv2d(x, p[f"{conv}.weight"], p[f"{conv}.bias"], padding=pad)
    if TRACE is not None:
        TRACE.setdefault(conv, []).append(z.detach())
    y = _bn_train(z, p, buf, bn) if train else _bn_eval(z, p, buf, bn)
    return F.max_pool2d(F.relu(y), 2)


def central_image_features(x, p, buf, pre="image_encoder", train=True):
    """models/unimodal.py:127-143 (headless) + Linear(1600,E) (models/dino.py:459-462)."""
    x = _block(x, p, buf, f"{pre}.0.conv1", f"{pre}.0.bn1", 2, train)
    x = _block(x, p, buf, f"{pre}.0.conv2", f"{pre}.0.bn2", 0, train)
    return F.linear(x.flatten(1), p[f"{pre}.1.weight"], p[f"{pre}.1.bias"])


def central_audio_features(x, p, buf, pre="audio_encoder", train=True):
    """models/unimodal.py:185-211 (headless) + Linear(3136,E) (models/dino.py:465-468)."""
    for k in (1, 2, 3, 4):
        x = _block(x, p, buf, f"{pre}.0.conv{k}", f"{pre}.0.bn{k}", 2, train)
    return F.linear(x.flatten(1), p[f"{pre}.1.weight"], p[f"{pre}.1.bias"])


def _dropout(x, keep, prob):
    if keep is None or prob == 0.0:
        return x
    return x * keep.to(x.dtype) / (1.0 - prob)


def central_encoder(img, aud, p, buf, fusion_keep=None, fusion_p=0.3):
    """SimpleMultiModalEncoder.forward (models/dino.py:229-234) with the Central encoders."""
    fi = central_image_features(img, p, buf)
    fa = central_audio_features(aud, p, buf)
    h = F.relu(F.linear(torch.cat([fi, fa], 1), p["fusion.0.weight"], p["fusion.0.bias"]))
    h = _dropout(h, fusion_keep, fusion_p)
    return F.linear(h, p["fusion.3.weight"], p["fusion.3.bias"])


def simple_image_features(x, p, buf, pre="image_encoder", train=True):
    """image_encoder(E) (models/dino.py:18-41): 3 x [conv3x3 pad 1, BN, ReLU, maxpool 2] -> global average pool -> Linear(128, E)."""
    for i in (0, 4, 8):
        x = _block(x, p, buf, f"{pre}.{i}", f"{pre}.{i + 1}", 1, train)
    return F.linear(x.mean(dim=(2, 3)), p[f"{pre}.14.weight"], p[f"{pre}.14.bias"])


def simple_audio_features(x, p, buf, pre="audio_encoder", train=True):
    """audio_encoder(E) (models/dino.py:43-73): 4 x [conv3x3 pad 1, BN, ReLU, maxpool 2] -> global average pool -> Linear(256, E)."""
    for i in (0, 4, 8, 12):
        x = _block(x, p, buf, f"{pre}.{i}", f"{pre}.{i + 1}", 1, train)
    return F.linear(x.mean(dim=(2, 3)), p[f"{pre}.18.weight"], p[f"{pre}.18.bias"])


def cross_modal_attention(x1, x2, p, name):
    """CrossModalAttention.forward (models/dino.py:393-405): attention over the batch of one encoder call, plus the residual."""
    q = F.linear(x1, p[f"{name}.q_proj.weight"], p[f"{name}.q_proj.bias"])
    k, v = F.linear(x2, p[f"{name}.kv_proj.weight"], p[f"{name}.kv_proj.bias"]).chunk(2, dim=-1)
    attn = ((q @ k.transpose(-2, -1)) * (x1.shape[-1] ** -0.5)).softmax(dim=-1)
    return x1 + attn @ v


def image_features(kind, x, p, buf, train=True):
    return (central_image_features if kind == "multi_central" else simple_image_features)(x, p, buf, train=train)


def audio_features(kind, x, p, buf, train=True):
    return (central_audio_features if kind == "multi_central" else simple_audio_features)(x, p, buf, train=train)


def multimodal_encoder(kind, img, aud, p, buf, fusion_keep=None, fusion_p=0.3, train=True):
    """forward() of SimpleMultiModalEncoder (models/dino.py:229-234; inherited by CentralMultiModalEncoder), GatedMultiModalEncoder
    (:249-263) and CrossAttentionMultiModalEncoder (:431-452)."""
    fi = image_features(kind, img, p, buf, train)
    fa = audio_features(kind, aud, p, buf, train)
    mix = KIND_MIX[kind]
    if mix == "gated":
        fi, fa = torch.sigmoid(p["gate_image"]) * fi, torch.sigmoid(p["gate_audio"]) * fa
    elif mix == "cross":
        fi, fa = (cross_modal_attention(fi, fa, p, "image_to_audio_attention"),
                  cross_modal_attention(fa, fi, p, "audio_to_image_attention"))
    h = F.relu(F.linear(torch.cat([fi, fa], 1), p["fusion.0.weight"], p["fusion.0.bias"]))
    h = _dropout(h, fusion_keep, fusion_p)
    return F.linear(h, p["fusion.3.weight"], p["fusion.3.bias"])


def image_simple_encoder(img, p, buf):
    """ImageEncoder.forward (models/dino.py:493-499)."""
    x = _block(img, p, buf, "encoder.0", "encoder.1", 1)
    x = _block(x, p, buf, "encoder.4", "encoder.5", 1)
    x = _block(x, p, buf, "encoder.8", "encoder.9", 1)
    x = x.mean(dim=(2, 3))
    x = F.linear(x, p["encoder.14.weight"], p["encoder.14.bias"])
    return F.linear(x, p["projection.0.weight"], p["projection.0.bias"])


def projection_head(x, p, buf, keep=None, drop_p=0.0, train=True):
    """ProjectionHead.forward (models/dino.py:1251-1254): Linear -> BN1d -> GELU(erf) -> Dropout -> Linear."""
    h = F.linear(x, p["mlp.0.weight"], p["mlp.0.bias"])
    h = _bn_train(h, p, buf, "mlp.1") if train else _bn_eval(h, p, buf, "mlp.1")
    h = _dropout(F.gelu(h), keep, drop_p)
    return F.linear(h, p["mlp.4.weight"], p["mlp.4.bias"])


# ------------------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------------------

def dino_loss_multimodal(s, t, tau_s=0.1, tau_t=0.04):
    """MultiModalDINOLightning.dino_loss (models/dino.py:822-854); s [Vs,B,D], t [Vt,B,D] (already centred).
    Uses the factorised form  -(1/(Vs*Vt*B)) sum_b <sum_u p_u , sum_v q_v>  (SURVEY Appendix A8)."""
    q = F.log_softmax(F.normalize(s, p=2, dim=-1) / tau_s, dim=-1).sum(0)
    pr = F.softmax(F.normalize(t, p=2, dim=-1) / tau_t, dim=-1).sum(0)
    return -(pr * q).sum() / (s.shape[0] * t.shape[0] * s.shape[1])


def dino_loss_unimodal(s, t, tau_s=0.1, tau_t=0.04):
    """UniModalDINOLightning.dino_loss (models/dino.py:1596-1635): extra per-view batch-mean subtraction of the
    *normalised* teacher outputs (:1613-1614)."""
    q = F.log_softmax(F.normalize(s, p=2, dim=-1) / tau_s, dim=-1).sum(0)
    tn = F.normalize(t, p=2, dim=-1)
    tn = tn - tn.mean(dim=1, keepdim=True)
    pr = F.softmax(tn / tau_t, dim=-1).sum(0)
    return -(pr * q).sum() / (s.shape[0] * t.shape[0] * s.shape[1])


def cosine_consistency_loss(emb):
    """UniModalDINOLightning._cosine_consistency_loss (models/dino.py:1575-1594)."""
    e = F.normalize(emb, p=2, dim=-1)
    V = e.shape[0]
    tot, cnt = 0.0, 0
    for i in range(V):
        for j in range(i + 1, V):
            tot = tot + ((1 - (e[i] * e[j]).sum(-1)) ** 2).mean()
            cnt += 1
    return tot / cnt


def infonce_loss(a, b, temperature=0.07):
    """MultiModalDINOWithINFONCELightning.infoNCE_loss (models/dino.py:1091-1128)."""
    sim = F.normalize(a, p=2, dim=1) @ F.normalize(b, p=2, dim=1).T / temperature
    lab = torch.arange(a.shape[0])
    return 0.5 * (F.cross_entropy(sim, lab) + F.cross_entropy(sim.T, lab))


def ntxent_loss(reps, temperature=0.07):
    """MultiModalSimCLRLightning.nt_xent_loss (other_ssl/multimodal_simclr/multimodal_simclr.py:74-89): reps = cat([z1, z2]),
    self-similarities masked, positive of row i is row (i + B) mod 2B."""
    r = F.normalize(reps, dim=1)
    n = r.shape[0]
    sim = (r @ r.T) / temperature
    sim = sim.masked_fill(torch.eye(n, dtype=torch.bool, device=r.device), float("-inf"))
    lab = (torch.arange(n, device=r.device) + n // 2) % n
    return F.cross_entropy(sim, lab)


def mse_align_loss(a, b):
    """MultiModalDINOWithMSELightning.mse_loss (models/dino.py:1193-1211)."""
    return ((F.normalize(a, p=2, dim=1) - F.normalize(b, p=2, dim=1)) ** 2).mean()


def supervised_loss(image_logits, audio_logits, labels):
    """MultiModalDINOSemiSupervisedLightning.supervised_loss (models/dino.py:1001-1025)."""
    return F.cross_entropy(image_logits, labels) + F.cross_entropy(audio_logits, labels)


# ------------------------------------------------------------------------------------------------------------
# EMA / center / Adam
# ------------------------------------------------------------------------------------------------------------

def ema_update(teacher, student, momentum=0.996):
    """update_teacher (models/dino.py:635-646): t = m*t + (1-m)*s, as two products and a sum; `1-m` is formed
    in Python double and rounded when it meets the fp32 tensor."""
    for k in teacher:
        teacher[k] = momentum * teacher[k] + (1 - momentum) * student[k]


def center_update(center, teacher_projs, center_momentum=0.9):
    """update_center (models/dino.py:648-653); teacher_projs are the UNcentred projections [Vt*B, D]."""
    return center * center_momentum + teacher_projs.mean(dim=0, keepdim=True) * (1 - center_momentum)


def adam_step(params, grads, state, lr=1e-4, weight_decay=1e-6, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam as configured at models/dino.py:953-962 (L2 added to the gradient, bias correction)."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    for k, p in params.items():
        g = grads.get(k)
        if g is None:
            continue
        g = g + weight_decay * p
        m = state.setdefault(("m", k), torch.zeros_like(p))
        v = state.setdefault(("v", k), torch.zeros_like(p))
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        bc1 = 1 - betas[0] ** t
        bc2 = 1 - betas[1] ** t
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        params[k] = p - (lr / bc1) * m / denom


# ------------------------------------------------------------------------------------------------------------
# whole step (multi_central, all four training modes)
# ------------------------------------------------------------------------------------------------------------

class CentralDinoState:
    """All state of MultiModalDINO(+mode heads) as flat dicts; kind selects the encoder: CentralMultiModalEncoder (default),
    SimpleMultiModalEncoder, GatedMultiModalEncoder or CrossAttentionMultiModalEncoder."""

    def __init__(self, seed=0, E=256, O=256, P=128, mode="default", dtype=torch.float32, kind="multi_central"):
        self.E, self.O, self.P, self.mode, self.dtype, self.kind = E, O, P, mode, dtype, kind
        self.enc_spec = encoder_spec(kind, E, O)
        self.head_spec = head_spec(O, P)
        self.student = make_params(self.enc_spec, seed, dtype)
        self.student_head = make_params(self.head_spec, seed + 1, dtype)
        self.teacher = {k: v.clone() for k, v in self.student.items()}
        self.teacher_head = {k: v.clone() for k, v in self.student_head.items()}
        self.student_buf = make_bn_buffers(self.enc_spec, encoder_bn_names(kind), dtype)
        self.teacher_buf = make_bn_buffers(self.enc_spec, encoder_bn_names(kind), dtype)
        self.student_head_buf = make_bn_buffers(self.head_spec, ["mlp.1"], dtype)
        self.teacher_head_buf = make_bn_buffers(self.head_spec, ["mlp.1"], dtype)
        self.center = torch.zeros(1, P, dtype=dtype)
        self.aux = {}
        if mode != "default":
            out = 10 if mode == "semi_supervised" else P
            self.aux_spec = head_spec(E, out)
            for i, mod in enumerate(("image", "audio")):
                self.aux[mod] = make_params(self.aux_spec, seed + 2 + i, dtype)
                self.aux[mod + "_buf"] = make_bn_buffers(self.aux_spec, ["mlp.1"], dtype)
        self.adam = {}


def central_dino_step(st, img_views, aud_views, masks, n_global=2, tau_s=0.1, tau_t=0.04, momentum=0.996,
                      center_momentum=0.9, lr=1e-4, weight_decay=1e-6, dropout=0.3, raw=None, labels=None,
                      alpha=1.0, do_adam=True, loss_scale=1.0):
    """One reference training step in the live Lightning order: forward (student on all views,
    teacher on the global views, heads, centre update), loss, teacher EMA, backward, Adam.

    img_views [V,B,1,28,28], aud_views [V,B,1,112,112] (global views first);
    masks: dict with keep-masks 'student_fusion' [V,B,E], 'teacher_fusion' [Vg,B,E], 'student_head' [V*B,512]
    (+ 'aux_image_head'/'aux_audio_head' are never needed: the mode heads use dropout 0).
    raw = (image [B,1,28,28], audio [B,1,112,112]) for the non-default modes.
    loss_scale: static GradScaler stand-in for runs under fp16 autocast (Lightning's precision='16-mixed', run_dino.py:360):
    the backward runs on loss * loss_scale, gradients are unscaled, and a step with non-finite gradients skips Adam.
    Returns dict(loss, grads{student,student_head,aux...}, student_out, teacher_out)."""
    V, B = img_views.shape[0], img_views.shape[1]
    S = {k: v.clone().requires_grad_(True) for k, v in st.student.items()}
    SH = {k: v.clone().requires_grad_(True) for k, v in st.student_head.items()}
    AUX = {m: {k: v.clone().requires_grad_(True) for k, v in st.aux[m].items()} for m in ("image", "audio") if m in st.aux}

    kind = getattr(st, "kind", "multi_central")
    feats = [multimodal_encoder(kind, img_views[v], aud_views[v], S, st.student_buf, masks["student_fusion"][v], 0.3)
             for v in range(V)]
    student_features = torch.cat(feats)
    with torch.no_grad():
        tf = [multimodal_encoder(kind, img_views[v], aud_views[v], st.teacher, st.teacher_buf, masks["teacher_fusion"][v], 0.3)
              for v in range(n_global)]
        teacher_features = torch.cat(tf)
    student_projs = projection_head(student_features, SH, st.student_head_buf, masks["student_head"], dropout)
    with torch.no_grad():
        teacher_projs = projection_head(teacher_features, st.teacher_head, st.teacher_head_buf, None, 0.0)
        teacher_c = teacher_projs - st.center
        st.center = center_update(st.center, teacher_projs, center_momentum)
    s_out = student_projs.view(V, B, -1)
    t_out = teacher_c.view(n_global, B, -1)
    loss = dino_loss_multimodal(s_out, t_out, tau_s, tau_t)
    aux_val = None
    if st.mode != "default":
        image, audio = raw
        fi = image_features(kind, image, S, st.student_buf)
        fa = audio_features(kind, audio, S, st.student_buf)
        zi = projection_head(fi, AUX["image"], st.aux["image_buf"])
        za = projection_head(fa, AUX["audio"], st.aux["audio_buf"])
        if st.mode == "semi_supervised":
            aux_val = supervised_loss(zi, za, labels)
        elif st.mode == "infonce":
            aux_val = infonce_loss(zi, za)
        elif st.mode == "mse":
            aux_val = mse_align_loss(zi, za)
        else:
            raise ValueError(st.mode)
        loss = loss + alpha * aux_val

    # EMA before backward/optimizer (models/dino.py:871)
    ema_update(st.teacher, st.student, momentum)
    ema_update(st.teacher_head, st.student_head, momentum)

    (loss * loss_scale if loss_scale != 1.0 else loss).backward()
    unscale = (lambda g: g / loss_scale) if loss_scale != 1.0 else (lambda g: g)
    grads = {"student": {k: unscale(v.grad) for k, v in S.items() if v.grad is not None},
             "student_head": {k: unscale(v.grad) for k, v in SH.items() if v.grad is not None}}
    for m in AUX:
        grads[m] = {k: unscale(v.grad) for k, v in AUX[m].items() if v.grad is not None}
    if loss_scale != 1.0 and not all(bool(torch.isfinite(g).all()) for gd in grads.values() for g in gd.values()):
        do_adam = False          # GradScaler: skip the optimizer step on overflow
    if do_adam:
        # one optimizer over all parameters: shared step counter
        step = st.adam.get("step", 0)
        groups = [("student", st.student), ("student_head", st.student_head)] + [(m, st.aux[m]) for m in AUX]
        for gname, pd in groups:
            sub = st.adam.setdefault(gname, {})
            sub["step"] = step
            adam_step(pd, grads[gname], sub, lr, weight_decay)
        st.adam["step"] = step + 1
    return {"loss": loss.detach(), "aux": None if aux_val is None else aux_val.detach(), "grads": grads,
            "student_out": s_out.detach(), "teacher_out": t_out.detach(),
            "student_features": student_features.detach()}


# ------------------------------------------------------------------------------------------------------------
# stand-alone contrastive steps (SURVEY 8f-4 / BASELINE config 4): other_ssl/info_nce/info_nce.py:15-36, 120-142 and
# other_ssl/multimodal_simclr/multimodal_simclr.py:12-46, 91-110 -- ImageEncoder + SpectrogramEncoder + two ProjectionHeads,
# InfoNCE between the modalities of the un-augmented batch / NT-Xent between two augmented views with a random modality pairing
# ------------------------------------------------------------------------------------------------------------

def spectrogram_spec(O=256):
    """SpectrogramEncoder (models/dino.py:502-513) = audio_encoder(O) (models/dino.py:43-73)."""
    s = []
    for i, (ci, co) in zip((0, 4, 8, 12), ((1, 32), (32, 64), (64, 128), (128, 256))):
        s += _conv(f"encoder.{i}", co, ci, 3) + _bn(f"encoder.{i + 1}", co)
    return s + _lin("encoder.18", O, 256)


def spectrogram_bn_names():
    return ["encoder.1", "encoder.5", "encoder.9", "encoder.13"]


CONTRASTIVE_MODULES = ("image_encoder", "audio_encoder", "image_projection_head", "audio_projection_head")


class ContrastiveState:
    """Parameters / BatchNorm buffers / Adam state of InfoNCEModel or MultiModalSimCLRModel, one flat dict per sub-module."""

    def __init__(self, seed=0, O=256, P=256, dtype=torch.float32):
        self.O, self.P, self.dtype = O, P, dtype
        self.spec = {"image_encoder": image_simple_spec(O), "audio_encoder": spectrogram_spec(O),
                     "image_projection_head": head_spec(O, P), "audio_projection_head": head_spec(O, P)}
        bn = {"image_encoder": image_simple_bn_names(), "audio_encoder": spectrogram_bn_names(),
              "image_projection_head": ["mlp.1"], "audio_projection_head": ["mlp.1"]}
        self.params = {m: make_params(self.spec[m], seed + i, dtype) for i, m in enumerate(CONTRASTIVE_MODULES)}
        self.buf = {m: make_bn_buffers(self.spec[m], bn[m], dtype) for m in CONTRASTIVE_MODULES}
        self.adam = {m: {} for m in CONTRASTIVE_MODULES}


def contrastive_step(st, kind, batch, mode=None, lr=1e-4, temperature=0.07, do_adam=True):
    """One training step of MultiModalInfoNCELightning (kind="infonce": batch = (images [B,1,28,28], spectrograms [B,1,112,112])) or
    MultiModalSimCLRLightning (kind="simclr": batch = (img1, spec1, img2, spec2), mode in 0..3 = the reference's torch.randint draw:
    0 image-image, 1 audio-audio, 2 image-audio, 3 audio-image).  Adam(lr), no weight decay; torch's Adam skips parameters without
    a gradient and keeps a step count per parameter, so a branch that was not used this step is left untouched."""
    P = {m: {k: v.clone().requires_grad_(True) for k, v in st.params[m].items()} for m in CONTRASTIVE_MODULES}

    def enc(which, x):
        if which == "image":
            f = image_simple_encoder(x, P["image_encoder"], st.buf["image_encoder"])
            return projection_head(f, P["image_projection_head"], st.buf["image_projection_head"])
        f = simple_audio_features(x, P["audio_encoder"], st.buf["audio_encoder"], pre="encoder")
        return projection_head(f, P["audio_projection_head"], st.buf["audio_projection_head"])

    if kind == "infonce":
        images, specs = batch
        z1, z2 = enc("image", images), enc("audio", specs)
        loss = infonce_loss(z1, z2, temperature)
    else:
        img1, spec1, img2, spec2 = batch
        first, second = (("image", "image"), ("audio", "audio"), ("image", "audio"), ("audio", "image"))[mode]
        z1 = enc(first, img1 if first == "image" else spec1)
        z2 = enc(second, img2 if second == "image" else spec2)
        loss = ntxent_loss(torch.cat([z1, z2]), temperature)
    loss.backward()
    grads = {m: {k: v.grad for k, v in P[m].items() if v.grad is not None} for m in CONTRASTIVE_MODULES}
    if do_adam:
        for m in CONTRASTIVE_MODULES:
            if grads[m]:
                adam_step(st.params[m], grads[m], st.adam[m], lr, 0.0)
    return {"loss": loss.detach(), "grads": grads, "z1": z1.detach(), "z2": z2.detach()}
