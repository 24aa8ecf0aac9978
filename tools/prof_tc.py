"""Small driver for ncu: one launch of each tensor-core conv kernel + the act8 BN kernels at bench shapes (N = 6*B)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_ssl_avmnist_b200 import ops

DEV = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
views = 6
N = views * B
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
BF = torch.bfloat16


def layer(Cin, Cout, H, K, pad):
    Ho = H + 2 * pad - K + 1
    if Cin == 1:
        x8 = torch.rand(N, H, ops.quad8_width(H, pad), 8, device=DEV).to(BF)
    else:
        x8 = torch.randn(N, Cin // 8, H, H, 8, device=DEV).to(BF)
    dz8 = torch.randn(N, Cout // 8, Ho, Ho, 8, device=DEV).to(BF)
    w = torch.randn(Cout, Cin, K, K, device=DEV) * 0.05
    b = torch.zeros(Cout, device=DEV)
    wp = torch.empty(ops.conv_tc_weight_bytes(Cin, Cout, K), dtype=torch.uint8, device=DEV)
    ops.conv_tc_prep_weights(w, wp)
    stats = torch.zeros(views, Cout, 2, dtype=torch.float64, device=DEV)
    z8 = torch.empty(N, Cout // 8, Ho, Ho, 8, dtype=torch.float16, device=DEV)
    dw = torch.empty(Cout, Cin, K, K, device=DEV)
    work = torch.empty(ops.conv_tc_wgrad_work_floats(N, Cin, Cout, H, H, K, pad), device=DEV) if Cin > 1 else None
    for _ in range(reps):
        ops.conv_tc(x8, wp, b, z8, stats, B, Cout, K, pad)
        if Cin > 1:
            ops.conv_tc_wgrad(x8, dz8, dw, work, pad)
    if Cin > 1:
        wpf = torch.empty(ops.conv_tc_weight_bytes(Cout, Cin, K), dtype=torch.uint8, device=DEV)
        ops.conv_tc_prep_weights(w, wpf, flip=True)
        dx8 = torch.empty(N, Cin // 8, H, H, 8, dtype=BF, device=DEV)
        for _ in range(reps):
            ops.conv_tc(dz8, wpf, None, dx8, None, N, Cin, K, K - 1 - pad)
    # BN / ReLU / pool on this layer's z
    sc = torch.ones(views, Cout, device=DEV); sh = torch.zeros(views, Cout, device=DEV)
    mu = torch.zeros(views, Cout, device=DEV); inv = torch.ones(views, Cout, device=DEV)
    p8 = torch.empty(N, Cout // 8, Ho // 2, Ho // 2, 8, dtype=BF, device=DEV)
    dp8 = torch.randn(N, Cout // 8, Ho // 2, Ho // 2, 8, device=DEV).to(BF)
    sums = torch.zeros(views, Cout, 2, dtype=torch.float64, device=DEV)
    for _ in range(reps):
        ops.bn_relu_pool8_fwd(z8, sc, sh, p8, B)
        ops.bn_relu_pool8_bwd_reduce(z8, dp8, sc, sh, mu, inv, sums, B)
        ops.bn_relu_pool8_bwd_apply(z8, dp8, sc, sh, mu, inv, sums, dz8, B)
        ops.bn_pool8_bwd_reduce_p(p8, dp8, sc[0].contiguous(), sh[0].contiguous(), sums, B)
    if Cin == 1:        # fused first-layer backward (BatchNorm-apply + weight gradient)
        work2 = torch.empty(ops.conv_tc_wgrad_l0_fused_work_floats(N, B, Cout, H, H, K, pad), device=DEV)
        for _ in range(reps):
            ops.conv_tc_wgrad_l0_fused(x8, z8, dp8, sc, sh, mu, inv, sums, dw, None, work2, B, pad)


layer(1, 8, 112, 5, 2)
layer(1, 32, 28, 5, 2)
layer(8, 16, 56, 5, 2)
layer(16, 32, 28, 5, 2)
layer(32, 64, 14, 5, 2)
torch.cuda.synchronize()
print("ok")
