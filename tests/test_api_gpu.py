"""GPU: the reference-shaped API (AVMNIST_Experiments mirror) drives the CUDA step: training_step -> loss.backward() ->
optimizer.step() through the Lightning modules, all four training modes + the unimodal model, and a short Trainer.fit on
synthetic AVMNIST files; the module path must agree with the bare engine on the same weights and views."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MIRROR = os.path.join(ROOT, "multimodal_ssl_avmnist_b200", "AVMNIST_Experiments")
sys.path.insert(0, MIRROR)

import models.dino as md  # noqa: E402
import utils.get_data as gd  # noqa: E402
from oracle.fixtures import synth_raw, synth_views  # noqa: E402

DEV = "cuda"
KW = dict(data_dir="x/", data_augmentation="burst_noise", dino_model=None, encoder_class=md.CentralMultiModalEncoder, encoder_kwargs=None,
          projection_dim=128, output_dim=256, encoder_output_dim=256, momentum=0.996, center_momentum=0.9, student_temperature=0.1,
          teacher_temperature=0.04, learning_rate=1e-3, use_mixed_precision=True, num_epochs=100, weight_decay=1e-6, dropout=0.3)
WRAPPERS = {"default": md.MultiModalDINOLightning, "semi_supervised": md.MultiModalDINOSemiSupervisedLightning,
            "infonce": md.MultiModalDINOWithINFONCELightning, "mse": md.MultiModalDINOWithMSELightning}


def _batch(mode, B, seed):
    gi, ga, li, la = synth_views(B, seed=seed)          # [B, V, 1, H, W] like the reference's collated batch
    views = tuple(t.to(DEV) for t in (gi, ga, li, la))
    if mode == "default":
        return views
    image, audio, labels = synth_raw(B, seed=seed + 1)
    return image.to(DEV), audio.to(DEV), labels.to(DEV), views


@pytest.mark.parametrize("mode", list(WRAPPERS))
def test_training_step_backward_optimizer(mode):
    torch.manual_seed(0)
    lit = WRAPPERS[mode](**KW).to(DEV)
    opt = lit.configure_optimizers()["optimizer"]
    B = 8
    before = {k: v.detach().clone() for k, v in lit.model.student.state_dict().items() if v.dtype == torch.float32}
    teacher0 = lit.model.teacher.fusion[0].weight.detach().clone()
    losses = []
    for it in range(3):
        opt.zero_grad(set_to_none=True)
        loss = lit.training_step(_batch(mode, B, 10 + it), it)
        assert loss.requires_grad and loss.dim() == 0
        loss.backward()
        g = lit.model.student.audio_encoder[0].conv2.weight.grad
        assert g is not None and float(g.abs().sum()) > 0
        assert lit.model.student.image_encoder[0].fc1.weight.grad is None or float(lit.model.student.image_encoder[0].fc1.weight.grad.abs().sum()) == 0
        opt.step()
        losses.append(float(loss))
    torch.cuda.synchronize()
    assert all(l == l and 3.0 < l < 12.0 for l in losses), losses
    moved = max(float((lit.model.student.state_dict()[k] - v).abs().max()) for k, v in before.items() if "fc1" not in k and "fc2" not in k)
    assert moved > 1e-4
    assert float((lit.model.teacher.fusion[0].weight - teacher0).abs().max()) > 0          # EMA follows the student
    assert float(lit.model.center.abs().max()) > 0
    # parameters still alias the engine arenas and the state_dict has the reference's keys
    eng = lit.model.engine
    assert lit.model.student.fusion[0].weight.data_ptr() == eng.S["enc.fusion.0.weight"].data_ptr()
    sd = lit.state_dict()
    assert "model.student.audio_encoder.0.conv4.weight" in sd and "model.center" in sd and "model.teacher_projection.mlp.4.bias" in sd


def test_module_path_matches_the_bare_engine():
    """Same weights, same views, dropout disabled: MultiModalDINOLightning.training_step == DinoStepEngine loss / gradients."""
    from multimodal_ssl_avmnist_b200.engine import DinoStepEngine
    torch.manual_seed(1)
    kw = dict(KW, dropout=0.0)
    lit = md.MultiModalDINOLightning(**kw).to(DEV)
    B = 8
    views = _batch("default", B, 50)
    loss = lit.training_step(views, 0)
    loss.backward()
    meng = lit.model.engine
    eng = DinoStepEngine(kind="multi_central", device=DEV, dropout=0.0, fusion_dropout=meng.fusion_dropout, seed=meng.seed,
                         precision=meng.precision)
    # the module ran update_teacher() already; copy the PRE-step weights: student is unchanged by training_step
    eng.student.flat.copy_(meng.student.flat)
    eng.sync_teacher()
    gi, ga, li, la = views
    xi = torch.cat([gi, li], 1).permute(1, 0, 2, 3, 4)[:, :, 0].contiguous()
    xa = torch.cat([ga, la], 1).permute(1, 0, 2, 3, 4)[:, :, 0].contiguous()
    eng.rng_step = 0
    if meng.fusion_dropout == 0:
        l2 = eng.forward_backward(xi, xa)
        torch.cuda.synchronize()
        assert abs(float(loss) - float(l2[0])) < 1e-5 * abs(float(l2[0]))
        n = eng.n_trainable_prefix
        assert float((eng.grad[:n] - meng.grad[:n]).abs().max()) <= 1e-6 * float(eng.grad[:n].abs().max()) + 1e-12
    else:       # fusion dropout is active in student and teacher (reference semantics): same Philox stream -> same masks
        l2 = eng.forward_backward(xi, xa)
        torch.cuda.synchronize()
        assert abs(float(loss) - float(l2[0])) < 1e-5 * abs(float(l2[0]))


def test_unimodal_training_step():
    torch.manual_seed(2)
    lit = md.UniModalDINOLightning(encoder_class=md.ImageEncoder, data_dir="x/", dropout=0.3, learning_rate=1e-3, projection_dim=128,
                                   output_dim=256, momentum=0.996, center_momentum=0.9, teacher_temperature=0.04, weight_decay=1e-6,
                                   cosine_loss_alpha=0, num_epochs=10, data_augmentation="burst_noise").to(DEV)
    opt = lit.configure_optimizers()["optimizer"]
    for it in range(2):
        opt.zero_grad(set_to_none=True)
        loss = lit.training_step(_batch("default", 8, 70 + it), it)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    assert 3.0 < float(loss) < 6.0


def test_unimodal_training_step_with_cosine_consistency_matches_engine():
    """cosine_loss_alpha > 0 (the reference's constructor default is 0.3): training_step returns dino + alpha * cosine and
    backward() carries the consistency gradient into the encoder - identical to the bare engine with the same weights."""
    from multimodal_ssl_avmnist_b200.engine import DinoStepEngine
    torch.manual_seed(4)
    lit = md.UniModalDINOLightning(encoder_class=md.ImageEncoder, data_dir="x/", dropout=0.0, learning_rate=1e-3, projection_dim=128,
                                   output_dim=256, cosine_loss_alpha=0.3, num_epochs=10, use_mixed_precision=False).to(DEV)
    views = _batch("default", 8, 90)
    loss = lit.training_step(views, 0)
    loss.backward()
    meng = lit.model.engine
    assert meng.precision == "fp32" and meng.cosine_loss_alpha == 0.3
    eng = DinoStepEngine(kind="image_simple", device=DEV, dropout=0.0, seed=meng.seed, precision="fp32", cosine_loss_alpha=0.3)
    eng.student.flat.copy_(meng.student.flat)
    eng.sync_teacher()
    gi, ga, li, la = views
    xi = torch.cat([gi, li], 1).permute(1, 0, 2, 3, 4)[:, :, 0].contiguous()
    l2 = eng.forward_backward(xi, None)
    torch.cuda.synchronize()
    assert float(l2[2]) > 0 and abs(float(l2[3]) - (float(l2[0]) + 0.3 * float(l2[2]))) < 1e-6
    assert abs(float(loss) - float(l2[3])) < 1e-5 * abs(float(l2[3]))
    n = eng.n_trainable_prefix
    # (fp32 SIMT path: float atomics in the statistics / bias-gradient sums give last-bit run-to-run differences)
    assert float((eng.grad[:n] - meng.grad[:n]).abs().max()) <= 1e-5 * float(eng.grad[:n].abs().max()) + 1e-12
    # and the term matters: without it the gradient differs
    eng0 = DinoStepEngine(kind="image_simple", device=DEV, dropout=0.0, seed=meng.seed, precision="fp32", cosine_loss_alpha=0.0)
    eng0.student.flat.copy_(meng.student.flat)
    eng0.sync_teacher()
    eng0.forward_backward(xi, None)
    assert float((eng0.grad[:n] - eng.grad[:n]).abs().max()) > 0


def test_trainer_fit_on_synthetic_files(tmp_path):
    """The run_dino.py flow in miniature: data module on synthetic files, Trainer(max_epochs) from the Lightning-surface
    shim (or real Lightning when installed), CSV logger, checkpoint, reload."""
    from _compat import pl, ModelCheckpoint, CSVLogger
    d = str(tmp_path) + "/"
    gd.write_synthetic_avmnist(d, n_train=64, n_test=16)
    dm = gd.AVMNISTDinoDataModule(data_dir=d, batch_size=16, num_workers=0, type="burst_noise")
    lit = md.MultiModalDINOLightning(**dict(KW, data_dir=d))
    ckpt = ModelCheckpoint(dirpath=str(tmp_path), monitor="train_loss_epoch", mode="min")
    tr = pl.Trainer(max_epochs=2, logger=CSVLogger(str(tmp_path), name="logs"), callbacks=[ckpt], log_every_n_steps=1, devices=1,
                    accelerator="gpu")
    tr.fit(lit, datamodule=dm)
    assert tr.global_step >= 4 and "train_loss_epoch" in tr.callback_metrics
    assert 3.0 < float(tr.callback_metrics["train_loss_epoch"]) < 6.0
    assert os.path.exists(ckpt.best_model_path)
    again = md.MultiModalDINOLightning.load_from_checkpoint(ckpt.best_model_path)
    a, b = again.state_dict(), lit.state_dict()
    assert set(a) == set(b) and all(a[k].shape == b[k].shape for k in a)


def test_device_resident_loader_feeds_the_step(tmp_path):
    """the training split lives in HBM (audio as the on-disk uint8), batches are device gathers and the /255
    normalisation happens inside the augmentation kernel; same samples as the DataLoader path."""
    d = str(tmp_path) + "/"
    gd.write_synthetic_avmnist(d, n_train=96, n_test=16)
    dm = gd.AVMNISTDinoDataModuleExtended(data_dir=d, batch_size=16, num_workers=0, type="burst_noise", device_resident=True)
    dm.setup("fit")
    dl = dm.train_dataloader()
    assert isinstance(dl, gd.DeviceResidentLoader) and len(dl) == len(dm.train_dataset) // 16
    image, audio, label = next(iter(dl))
    assert image.is_cuda and audio.dtype == torch.uint8 and image.shape == (16, 1, 28, 28) and audio.shape == (16, 1, 112, 112)
    # the resident tensors hold exactly the samples of the subset (the host dataset divides the audio by 255)
    sub = dm.train_dataset
    k = int(sorted(sub.indices)[3])
    himg, haud, hlab = sub.dataset[k]
    assert torch.allclose(dl.image[3].cpu(), himg.float(), atol=1e-7) and int(dl.labels[3]) == int(hlab)
    assert torch.allclose(dl.audio[3].cpu().float() / 255.0, haud.float(), atol=1e-7)
    lit = md.MultiModalDINOSemiSupervisedLightning(**dict(KW, data_dir=d)).to(DEV)
    opt = lit.configure_optimizers()["optimizer"]
    for it, batch in enumerate(dl):
        opt.zero_grad(set_to_none=True)
        loss = lit.training_step(batch, it)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    assert 3.0 < float(loss) < 12.0


def test_feature_path_and_linear_probe(tmp_path):
    """eval-mode encoder features through the CUDA kernels match the fp32 oracle encoder on the same weights
    (running statistics), train-mode features use batch statistics, and the per-epoch probe logs mlp_acc through Trainer.fit."""
    from oracle import dino_ref as R
    from _compat import pl, CSVLogger
    torch.manual_seed(3)
    lit = md.MultiModalDINOLightning(**KW).to(DEV)
    B = 16
    loss = lit.training_step(_batch("default", B, 90), 0)        # one step so that the running statistics are not the initial ones
    fx = md.FeatureExtractor(lit.model).eval()
    image, audio, _ = synth_raw(B, seed=91)
    feats = fx(image.to(DEV), audio.to(DEV))
    assert feats.shape == (B, 256) and torch.isfinite(feats).all()
    # oracle: eval-mode CentralMultiModalEncoder with the same parameters and buffers
    sd = {k: v.detach().cpu().float() for k, v in lit.model.student.state_dict().items()}
    p = {k: v for k, v in sd.items() if "running" not in k and "num_batches" not in k}
    buf = {k: v for k, v in sd.items() if "running" in k or "num_batches" in k}
    fi = R.central_image_features(image, p, buf, train=False)
    fa = R.central_audio_features(audio, p, buf, train=False)
    h = torch.relu(torch.cat([fi, fa], 1) @ p["fusion.0.weight"].t() + p["fusion.0.bias"])
    want = h @ p["fusion.3.weight"].t() + p["fusion.3.bias"]
    assert float((feats.cpu() - want).abs().max()) < 3e-2 * float(want.abs().max())
    ftrain = md.FeatureExtractor(lit.model).train()(image.to(DEV), audio.to(DEV))
    assert float((ftrain - feats).abs().max()) > 1e-3            # batch statistics + dropout differ from the eval path
    sd2 = lit.model.student.state_dict()
    assert all(torch.equal(sd2[k].cpu().float(), sd[k]) for k in sd if "running" in k), "feature passes must not touch the running statistics"
    # like the reference's deep-copied encoder: train-mode passes adapt the EXTRACTOR's running statistics, its eval-mode passes read them
    fx3 = md.FeatureExtractor(lit.model)
    fx3.train()
    a1, a2 = fx3(image.to(DEV), audio.to(DEV)), fx3(image.to(DEV), audio.to(DEV))
    assert float((a1 - a2).abs().max()) > 1e-4                   # a fresh dropout mask per probe batch
    adapted = fx3.eval()(image.to(DEV), audio.to(DEV))
    assert float((adapted - feats).abs().max()) > 1e-3           # the copy's statistics moved towards the un-augmented data
    assert torch.equal(md.FeatureExtractor(lit.model).eval()(image.to(DEV), audio.to(DEV)), feats)      # a new extractor starts from the live ones
    # the probe through the trainer
    d = str(tmp_path) + "/"
    gd.write_synthetic_avmnist(d, n_train=64, n_test=16)
    dm = gd.AVMNISTDinoDataModule(data_dir=d, batch_size=16, num_workers=0, type="burst_noise")
    lit2 = md.MultiModalDINOLightning(**dict(KW, data_dir=d))
    tr = pl.Trainer(max_epochs=1, logger=CSVLogger(str(tmp_path), name="logs"), log_every_n_steps=1, devices=1, accelerator="gpu")
    tr.fit(lit2, datamodule=dm)
    assert "mlp_acc" in tr.callback_metrics and 0.0 <= float(tr.callback_metrics["mlp_acc"]) <= 100.0
    # kNN evaluation of the frozen encoder (training_structures/dino_train.py:349-368) on the CUDA feature + kNN kernels
    from training_structures.dino_train import train_knn_classifier
    from sklearn.neighbors import KNeighborsClassifier
    train_loader, val_loader = dm.probe_dataloaders()
    knn, acc = train_knn_classifier(lit2.model, train_loader, val_loader, n_neighbors=5, device=DEV)
    assert 0.0 <= acc <= 100.0 and knn.train_features.shape[1] == 256 and knn.train_features.is_cuda
    fx2 = md.FeatureExtractor(lit2.model).eval()
    xs, ys = [], []
    for b in val_loader:
        xs.append(fx2(b[0].to(DEV), b[1].to(DEV)).cpu())
        ys.append(b[2])
    xs, ys = torch.cat(xs), torch.cat(ys)
    sk = KNeighborsClassifier(n_neighbors=5).fit(knn.train_features.cpu().numpy(), knn.train_labels.cpu().numpy())
    assert abs(100.0 * sk.score(xs.numpy(), ys.numpy()) - acc) <= 100.0 / len(ys) + 1e-6      # same features -> same accuracy (one near-tie allowed)


def test_standalone_ntxent_is_a_drop_in_for_the_reference_loss():
    """binding.standalone_ntxent_loss: forward value and autograd gradient of the fused kernel == torch restatement of
    MultiModalSimCLRLightning.nt_xent_loss on CUDA tensors that require grad."""
    import torch.nn.functional as F
    from multimodal_ssl_avmnist_b200 import binding as B
    torch.manual_seed(3)
    z = torch.randn(64, 128, device=DEV, requires_grad=True)
    loss = B.standalone_ntxent_loss(z)
    loss.backward()
    z2 = z.detach().clone().double().requires_grad_(True)
    r = F.normalize(z2, dim=1)
    sim = (r @ r.T / 0.07).masked_fill(torch.eye(64, dtype=torch.bool, device=DEV), float("-inf"))
    want = F.cross_entropy(sim, (torch.arange(64, device=DEV) + 32) % 64)
    want.backward()
    assert abs(float(loss) - float(want)) < 1e-5 * float(want)
    assert float((z.grad - z2.grad.float()).abs().max()) <= 2e-5 * float(z2.grad.abs().max())


@pytest.mark.parametrize("mode", ["default", "semi_supervised"])
def test_fused_graph_step_through_the_trainer_equals_the_step_by_step_path(tmp_path, mode):
    """Small batches through `Trainer.fit`: the fused path (the whole step = one CUDA-graph replay; loss.backward() and
    optimizer.step() become no-ops for that batch) must train exactly like training_step -> backward -> B200Adam.step():
    same seeds, same data order -> bit-identical student, teacher, Adam moments, centre after two epochs; the cosine LR schedule
    reaches the captured Adam through the device-side learning rate."""
    from _compat import pl
    d = str(tmp_path) + "/"
    gd.write_synthetic_avmnist(d, n_train=96, n_test=16)
    cls = WRAPPERS[mode]
    results = []
    for fused in (False, True):
        torch.manual_seed(7)
        dm_cls = gd.AVMNISTDinoDataModuleExtended if mode != "default" else gd.AVMNISTDinoDataModule
        dm = dm_cls(data_dir=d, batch_size=16, num_workers=0, type="burst_noise", device_resident=True)
        dm.probe_dataloaders = None
        lit = cls(**dict(KW, data_dir=d, num_epochs=2))
        lit.b200_fused_step = fused
        tr = pl.Trainer(max_epochs=2, logger=None, log_every_n_steps=1, devices=1, accelerator="gpu")
        tr.fit(lit, datamodule=dm)
        torch.cuda.synchronize()
        eng = lit.model.engine
        assert (eng._graph is not None) == fused
        assert tr.global_step == 2 * (len(dm.train_dataset) // 16) and eng.step_count == tr.global_step
        results.append((eng.student.flat.clone(), eng.teacher.flat.clone(), eng.exp_avg.clone(), eng.exp_avg_sq.clone(), lit.model.center.clone(),
                        float(tr.callback_metrics["train_loss_epoch"])))
    for a, b in zip(results[0][:5], results[1][:5]):
        assert torch.equal(a, b), float((a - b).abs().max())
    assert abs(results[0][5] - results[1][5]) < 1e-6


def test_adam_state_belongs_to_the_optimizer_object():
    """run_dino.py builds a fresh optimizer per fit / per seed on the same model (run_dino.py:94-104): a new B200Adam must start from
    zero moments and step 0 like a new torch.optim.Adam, and state_dict() / load_state_dict() must carry the arena moments."""
    torch.manual_seed(0)
    lit = md.MultiModalDINOLightning(**KW).to(DEV)
    B = 8

    def steps(opt, n, seed0):
        for it in range(n):
            opt.zero_grad(set_to_none=True)
            lit.training_step(_batch("default", B, seed0 + it), it).backward()
            opt.step()

    opt1 = lit.configure_optimizers()["optimizer"]
    steps(opt1, 2, 40)
    eng = lit.model.engine
    assert eng.step_count == 2 and float(eng.exp_avg.abs().max()) > 0
    sd = opt1.state_dict()
    assert sd["b200_arena_state"]["step"] == 2 and torch.equal(sd["b200_arena_state"]["exp_avg"], eng.exp_avg.cpu())
    m_after_2 = eng.exp_avg.clone()
    # a second fit: fresh optimizer -> fresh Adam (bias correction restarts at step 1, moments from zero)
    opt2 = lit.configure_optimizers()["optimizer"]
    steps(opt2, 1, 50)
    assert eng.step_count == 1
    n = eng.n_trainable_prefix          # first step from zero moments: exp_avg = (1 - beta1) * (grad + weight_decay * param), wd = 1e-6
    assert torch.allclose(eng.exp_avg[:n], 0.1 * eng.grad[:n], rtol=1e-2, atol=1e-7)
    # resume: a new optimizer that loads the checkpointed state continues from it
    opt3 = lit.configure_optimizers()["optimizer"]
    opt3.load_state_dict(sd)
    assert sd["b200_arena_state"]["step"] == 2                      # the caller's dict is not consumed
    steps(opt3, 1, 60)
    assert eng.step_count == 3
    assert not torch.equal(eng.exp_avg, m_after_2) and float((eng.exp_avg - 0.9 * m_after_2).abs().max()) < float(m_after_2.abs().max())


SIMPLE_ENCODERS = [("SimpleMultiModalEncoder", "default"), ("GatedMultiModalEncoder", "mse"), ("CrossAttentionMultiModalEncoder", "semi_supervised")]


@pytest.mark.parametrize("enc,mode", SIMPLE_ENCODERS)
def test_simple_family_through_the_lightning_modules(enc, mode):
    """SURVEY 8f-4 through the reference-shaped API: `--model multi_simple | multi_simple_gated | multi_cross_attention` of run_dino.py
    (models/dino.py:214-263, 385-452): training_step -> backward -> B200Adam.step moves every encoder parameter (gates and attention
    projections included), the teacher follows by EMA, and the containers' own inference forward (FeatureExtractor-style call of
    `student(images, spectrograms)`) agrees with the engine's evaluation forward on the same weights."""
    torch.manual_seed(0)
    lit = WRAPPERS[mode](**dict(KW, encoder_class=getattr(md, enc))).to(DEV)
    opt = lit.configure_optimizers()["optimizer"]
    B = 8
    before = {k: v.detach().clone() for k, v in lit.model.student.named_parameters()}
    losses = []
    for it in range(3):
        opt.zero_grad(set_to_none=True)
        loss = lit.training_step(_batch(mode, B, 10 + it), it)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    torch.cuda.synchronize()
    assert all(l == l and 3.0 < l < 14.0 for l in losses), losses
    for k, v in lit.model.student.named_parameters():
        assert float((v - before[k]).abs().max()) > 0, f"{k} did not move"
    eng = lit.model.engine
    assert eng.kind == getattr(md, enc).B200_KIND and all(eng.tc["aud"]) and all(eng.tc["img"])
    if enc == "GatedMultiModalEncoder":
        assert lit.model.student.gate_image.data_ptr() == eng.S["enc.gate_image"].data_ptr()
        assert float((lit.model.teacher.gate_image - 0.5).abs()) > 0          # EMA'd like every other parameter
    # inference forward of the containers (fp32 kernels) vs the engine's evaluation forward (bf16 product path)
    image, audio, _ = synth_raw(B, seed=77)
    lit.model.eval()
    with torch.no_grad():
        feats_mod = lit.model.student(image.to(DEV), audio.to(DEV))
        feats_eng = eng.encode_features(image[:, 0].to(DEV).contiguous(), audio[:, 0].to(DEV).contiguous(), train=False)
    torch.cuda.synchronize()
    rel = float((feats_mod - feats_eng).norm() / feats_eng.norm())
    assert rel < 3e-2, rel


@pytest.mark.parametrize("which", ["infonce", "simclr"])
def test_contrastive_lightning_modules_train(which):
    """other_ssl/info_nce and other_ssl/multimodal_simclr through their Lightning-shaped modules: training_step -> backward ->
    optimizer.step; parameters alias the engine arena, the branch a SimCLR step did not use has .grad None and does not move."""
    import other_ssl.info_nce.info_nce as nce
    import other_ssl.multimodal_simclr.multimodal_simclr as simclr
    torch.manual_seed(3)
    lit = (nce.MultiModalInfoNCELightning if which == "infonce" else simclr.MultiModalSimCLRLightning)(learning_rate=1e-3).to(DEV)
    opt = lit.configure_optimizers()["optimizer"]
    B = 16
    losses = []
    for it in range(6):
        g = torch.Generator().manual_seed(it)
        i1, s1 = torch.rand(B, 1, 28, 28, generator=g).to(DEV), torch.rand(B, 1, 112, 112, generator=g).to(DEV)
        i2, s2 = torch.rand(B, 1, 28, 28, generator=g).to(DEV), torch.rand(B, 1, 112, 112, generator=g).to(DEV)
        batch = (i1, s1, torch.zeros(B, dtype=torch.long, device=DEV)) if which == "infonce" else (i1, s1, i2, s2)
        before = {k: v.detach().clone() for k, v in lit.model.named_parameters()}
        opt.zero_grad(set_to_none=True)
        loss = lit.training_step(batch, it)
        assert loss.requires_grad and loss.dim() == 0
        loss.backward()
        opt.step()
        losses.append(float(loss))
        eng = lit._b200.engine
        for k, p in lit.model.named_parameters():
            branch = "img" if k.startswith("image") else "aud"
            moved = float((p.detach() - before[k]).abs().max())
            if branch in eng._used:
                assert p.grad is not None and p.grad.data_ptr() == eng.G["enc." + k].data_ptr()
            else:
                assert p.grad is None and moved == 0.0, k
    torch.cuda.synchronize()
    assert all(l == l and 0.0 < l < 20.0 for l in losses), losses
    eng = lit._b200.engine
    assert lit.model.image_encoder.encoder[0].weight.data_ptr() == eng.S["enc.image_encoder.encoder.0.weight"].data_ptr()
    assert sum(eng.step_counts.values()) == 12 if which == "infonce" else 6 <= sum(eng.step_counts.values()) <= 12
    sd = lit.state_dict()
    assert "model.audio_encoder.encoder.18.weight" in sd and "model.image_projection_head.mlp.1.running_mean" in sd
    # the containers' own inference forward runs on the CUDA kernels too
    lit.model.eval()
    with torch.no_grad():
        out = lit.model(batch)
    assert out[0].shape == (B, 256) and torch.isfinite(out[0]).all() and torch.isfinite(out[1]).all()


def test_infonce_trainer_fit_on_synthetic_files(tmp_path):
    """The flow of other_ssl/info_nce/info_nce.ipynb in miniature: AVMNISTDataModule batches (image, spectrogram, label) ->
    Trainer.fit(MultiModalInfoNCELightning) with the CSV logger and a checkpoint; the loss decreases, the checkpoint reloads with the
    reference's state_dict keys."""
    import other_ssl.info_nce.info_nce as nce
    from _compat import pl, ModelCheckpoint, CSVLogger
    d = str(tmp_path) + "/"
    gd.write_synthetic_avmnist(d, n_train=96, n_test=16)
    dm = gd.AVMNISTDataModule(data_dir=d, batch_size=16, num_workers=0, type="burst_noise")
    torch.manual_seed(0)
    lit = nce.MultiModalInfoNCELightning(projection_dim=256, output_dim=256, learning_rate=1e-3, num_epochs=3)
    ckpt = ModelCheckpoint(dirpath=str(tmp_path), monitor="train_loss_epoch", mode="min")
    tr = pl.Trainer(max_epochs=3, logger=CSVLogger(str(tmp_path), name="logs"), callbacks=[ckpt], log_every_n_steps=1, devices=1, accelerator="gpu")
    tr.fit(lit, datamodule=dm)
    assert tr.global_step >= 9 and "train_loss_epoch" in tr.callback_metrics
    assert 0.0 < float(tr.callback_metrics["train_loss_epoch"]) < 4.0
    assert sum(lit._b200.engine.step_counts.values()) == 2 * tr.global_step
    assert os.path.exists(ckpt.best_model_path)
    again = nce.MultiModalInfoNCELightning.load_from_checkpoint(ckpt.best_model_path)
    a, b = again.state_dict(), lit.state_dict()
    assert set(a) == set(b) and all(a[k].shape == b[k].shape for k in a)
