import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_ssl_avmnist_b200 import ops
DEV = "cuda"
M, N, K = 256, 128, 64
g = torch.Generator().manual_seed(0)
w = (torch.randn(N, K, generator=g)).to(DEV)
dy = torch.randn(M, N, generator=g).to(DEV)
x = torch.randn(M, K, generator=g).to(DEV)
dx = torch.full((M, K), float("nan"), device=DEV)
ops.linear_bwd_data(dy, w, dx, tc=True)
ref = dy.double() @ w.double()
torch.cuda.synchronize()
print("bwd_data: max|dx|", float(dx.abs().max()), "max err", float((dx.double() - ref).abs().max()), "ref max", float(ref.abs().max()))
print(dx[:2, :8]); print(ref[:2, :8].float())
dw = torch.full((N, K), float("nan"), device=DEV); db = torch.zeros(N, device=DEV)
ops.linear_bwd_weight(dy, x, dw, db, tc=True)
refw = dy.double().t() @ x.double()
torch.cuda.synchronize()
print("wgrad: max|dw|", float(dw.abs().max()), "max err", float((dw.double() - refw).abs().max()), "ref max", float(refw.abs().max()))
print(dw[:2, :8]); print(refw[:2, :8].float())
