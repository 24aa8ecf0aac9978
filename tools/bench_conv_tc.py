"""Times the tensor-core conv kernels (fwd / data gradient / weight gradient) per encoder layer at bench shapes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_ssl_avmnist_b200 import ops

DEV = "cuda"
LAYERS = [("aud2", 8, 16, 56, 5, 2), ("aud3", 16, 32, 28, 5, 2), ("aud4", 32, 64, 14, 5, 2), ("img2", 32, 64, 14, 5, 0)]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    views = 6
    N = views * B
    rows = []
    for name, Cin, Cout, H, K, pad in LAYERS:
        Ho = H + 2 * pad - K + 1
        x8 = torch.randn(N, Cin // 8, H, H, 8, device=DEV).to(torch.bfloat16)
        dz8 = torch.randn(N, Cout // 8, Ho, Ho, 8, device=DEV).to(torch.bfloat16)
        w = torch.randn(Cout, Cin, K, K, device=DEV) * 0.05
        b = torch.zeros(Cout, device=DEV)
        wp = torch.empty(ops.conv_tc_weight_bytes(Cin, Cout, K), dtype=torch.uint8, device=DEV)
        wpf = torch.empty(ops.conv_tc_weight_bytes(Cout, Cin, K), dtype=torch.uint8, device=DEV)
        ops.conv_tc_prep_weights(w, wp)
        ops.conv_tc_prep_weights(w, wpf, flip=True)
        stats = torch.zeros(views, Cout, 2, dtype=torch.float64, device=DEV)
        z8 = torch.empty(N, Cout // 8, Ho, Ho, 8, dtype=torch.bfloat16, device=DEV)
        zf = torch.empty(N, Cout, Ho, Ho, device=DEV)
        dx8 = torch.empty(N, Cin // 8, H, H, 8, dtype=torch.bfloat16, device=DEV)
        dw = torch.empty(Cout, Cin, K, K, device=DEV)
        work = torch.empty(ops.conv_tc_wgrad_work_floats(N, Cin, Cout, H, H, K, pad), device=DEV)
        fl = 2.0 * N * Cout * Ho * Ho * Cin * K * K
        t_f8 = timeit(lambda: ops.conv_tc(x8, wp, b, z8, stats, B, Cout, K, pad))
        t_ff = timeit(lambda: ops.conv_tc(x8, wp, b, zf, stats, B, Cout, K, pad))
        t_d = timeit(lambda: ops.conv_tc(dz8, wpf, None, dx8, None, N, Cin, K, K - 1 - pad))
        t_w = timeit(lambda: ops.conv_tc_wgrad(x8, dz8, dw, work, pad))
        by_f = x8.numel() * 2 + z8.numel() * 2
        rows.append({"layer": name, "N": N, "gflop": fl / 1e9, "fwd_bf16_ms": t_f8, "fwd_f32out_ms": t_ff, "dgrad_ms": t_d, "wgrad_ms": t_w,
                     "fwd_tflops": fl / t_f8 / 1e9, "fwd_gbs": by_f / t_f8 / 1e6, "dgrad_tflops": fl / t_d / 1e9, "wgrad_tflops": fl / t_w / 1e9,
                     "wgrad_gbs": (x8.numel() + dz8.numel()) * 2 / t_w / 1e6})
        print(json.dumps(rows[-1]))


if __name__ == "__main__":
    main()
