"""Lightning-free pre-training loop with the reference's `pretrain_dino` signature (training_structures/dino_train.py:104-186).

Differences that are documented, not hidden: the reference's legacy loop expects a 2-tuple model output and a free
`dino_loss(s, t, tau_s, tau_t)` that no longer exists in its own models/dino.py (stale code); here `dino_loss` may be None
(the fused CUDA loss is used) or any callable of (student_out, teacher_out, tau_s=..., tau_t=...).  Like the reference's
loop -- and unlike its Lightning path -- the teacher EMA runs AFTER the optimizer step.

Downstream evaluation: `feature_extraction_loop` / `train_knn_classifier` (reference :327-369) run on the CUDA feature path +
the device kNN kernels; `train_downstream` / `compute_classification_metrics` (an MLP probe trained for 10 epochs +
sklearn reports) stay outside the hot-path scope -- the per-epoch probe the checkpoint metric needs is
`_DinoLightningBase.probe_accuracy` in models/dino.py."""
import csv
import json
import os
import sys
from datetime import datetime

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _compat  # noqa: F401,E402
from multimodal_ssl_avmnist_b200 import binding as B  # noqa: E402


def pretrain_dino(model, trainloader, dino_loss=None, align=False, num_epochs=100, learning_rate=0.0001, save_path="pretrained_dino.pt",
                  log_path="pretrain_log.csv"):
    for path in (save_path, log_path):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    stamp = datetime.now().strftime("%Y-%m-%d %H-%M-%S")
    save_path, log_path = save_path.replace(".pt", f"_{stamp}.pt"), log_path.replace(".csv", f"_{stamp}.csv")
    info = {"start_time": stamp, "learning_rate": learning_rate, "batch_size": getattr(trainloader, "batch_size", None), "epochs": num_epochs,
            "model_name": type(model).__name__}
    with open(log_path, "w", newline="") as f:
        csv.writer(f).writerow(["epoch", "train_loss", f"# {json.dumps(info)}"])
    opt = B.B200Adam(model.parameters(), model._b200, lr=learning_rate, weight_decay=0.01)     # the legacy loop uses AdamW's default decay
    best = float("inf")
    for epoch in range(num_epochs):
        model.train()
        total, n = 0.0, 0
        for batch in trainloader:
            opt.zero_grad(set_to_none=True)
            if len(batch) == 4:
                student_out, teacher_out, _ = model(tuple(batch))
            else:
                student_out, teacher_out, _ = model.forward_raw(batch[0], batch[1])
            if dino_loss is None:
                loss = model._b200.dino_loss(student_out, teacher_out, 0.1, 0.04, 0)
            else:
                loss = dino_loss(student_out, teacher_out, tau_s=0.1, tau_t=0.04)
            loss.backward()
            opt.step()
            model.update_teacher()
            total += float(loss)
            n += 1
        avg = total / max(n, 1)
        with open(log_path, "a", newline="") as f:
            csv.writer(f).writerow([epoch + 1, avg])
        if avg < best:
            best = avg
            torch.save({"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": opt.state_dict(), "loss": best}, save_path)
    return model


def feature_extraction_loop(device, model, dataloader):
    """Frozen-encoder features of every (image, audio, label) batch (reference :327-347), kept ON THE DEVICE: returns
    (features [N, O] fp32 CUDA tensor, labels [N] int64 CUDA tensor) instead of numpy arrays."""
    model.eval()
    feats, labels = [], []
    with torch.no_grad():
        for batch in dataloader:
            images, spectrograms, lab = batch[0], batch[1], batch[2]
            feats.append(model(images.to(device), spectrograms.to(device)).float())
            labels.append(lab.to(device).long())
    return torch.cat(feats), torch.cat(labels)


class B200KNN:
    """What the reference keeps of its fitted sklearn KNeighborsClassifier: predict / score on new features."""

    def __init__(self, train_features, train_labels, n_neighbors=5, n_classes=10):
        self.train_features, self.train_labels = train_features.contiguous(), train_labels.contiguous()
        self.n_neighbors, self.n_classes = n_neighbors, n_classes

    def predict(self, features):
        from multimodal_ssl_avmnist_b200 import ops
        return ops.knn_predict(self.train_features, self.train_labels, features.to(self.train_features.device).float().contiguous(),
                               k=self.n_neighbors, n_classes=self.n_classes)

    def score(self, features, labels):
        pred = self.predict(features)
        return float((pred == labels.to(pred.device)).float().mean())


def train_knn_classifier(pretrained_dino, train_dataloader, test_dataloader, n_neighbors=5, device="cuda", is_dino_based=True):
    """kNN accuracy of the frozen student encoder (reference :349-369): features through the CUDA encoder forward, Euclidean
    k-nearest neighbours + majority vote in CUDA kernels (csrc/knn.cu).  Returns (knn, accuracy in %)."""
    from models.dino import FeatureExtractor
    extractor = FeatureExtractor(pretrained_dino, is_dino_based=is_dino_based)
    train_features, train_labels = feature_extraction_loop(device, extractor, train_dataloader)
    test_features, test_labels = feature_extraction_loop(device, extractor, test_dataloader)
    n_classes = int(max(int(train_labels.max()), int(test_labels.max()))) + 1
    knn = B200KNN(train_features, train_labels, n_neighbors=n_neighbors, n_classes=max(n_classes, 10))
    accuracy = 100.0 * knn.score(test_features, test_labels)
    print(f"KNN Accuracy (k={n_neighbors}): {accuracy:.4f}%")
    return knn, accuracy


def _evaluation_out_of_scope(*a, **k):
    raise NotImplementedError("the 10-epoch MLP probe / sklearn metric reports are outside the B200 hot-path scope (SURVEY 8f-3); "
                              "see _DinoLightningBase.probe_accuracy for the per-epoch probe")


train_downstream = compute_classification_metrics = _evaluation_out_of_scope
