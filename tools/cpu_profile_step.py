"""Host-side profile of the step loop at a tiny batch (GPU work negligible): where do the ~16 us per launch go?"""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import augment_values
from multimodal_ssl_avmnist_b200.engine import DinoStepEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
eng = DinoStepEngine(kind="multi_central", augment_values=augment_values(), seed=1, device="cuda:0")
g = torch.Generator().manual_seed(1)
img = torch.rand(B, 28, 28, generator=g).cuda()
aud = torch.randint(0, 256, (B, 112, 112), generator=g, dtype=torch.uint8).cuda()
for _ in range(5):
    eng.train_step(img, aud)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    eng.train_step(img, aud)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
