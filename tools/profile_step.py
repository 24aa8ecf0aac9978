"""Small fixed workload for ncu: W warm-up steps + K steps of the DINO step at per-GPU batch B.
    python tools/profile_step.py [B] [warmup] [steps] [kind = multi_central]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import augment_values
from multimodal_ssl_avmnist_b200 import ops
from multimodal_ssl_avmnist_b200.engine import DinoStepEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = int(sys.argv[3]) if len(sys.argv) > 3 else 1
KIND = sys.argv[4] if len(sys.argv) > 4 else "multi_central"
eng = DinoStepEngine(kind=KIND, augment_values=augment_values(), seed=1, device="cuda:0")
g = torch.Generator().manual_seed(1)
img = torch.rand(B, 28, 28, generator=g).cuda()
aud = torch.randint(0, 256, (B, 112, 112), generator=g, dtype=torch.uint8).cuda()
for _ in range(W):
    eng.train_step(img, aud)
torch.cuda.synchronize()
l0 = ops.launch_count()
for _ in range(K):
    loss = eng.train_step(img, aud)
torch.cuda.synchronize()
print("launches per step", (ops.launch_count() - l0) // K, "loss", float(loss[3]))
