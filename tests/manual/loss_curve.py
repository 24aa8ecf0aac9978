"""Loss-curve comparison over many steps: the CPU oracle (== the reference's arithmetic, tests/test_oracle_golden.py) against
the engine in both precisions, on identical per-step inputs (pre-augmented synthetic views, shared dropout masks) from
identical initial weights.  Writes a JSON with the three curves and summary statistics.

    python tests/manual/loss_curve.py --steps 1000 --batch 16 --out profiles/r1b_loss_curve_1k.json
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import dino_ref as R
from oracle.fixtures import make_masks, synth_views, views_to_vb
from multimodal_ssl_avmnist_b200.engine import DinoStepEngine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--out", default="gpurun_out/loss_curve.json")
    ap.add_argument("--pool", type=int, default=64, help="number of distinct synthetic batches cycled through")
    args = ap.parse_args()
    dev = "cuda:0"
    B = args.batch
    torch.set_num_threads(os.cpu_count() or 1)
    st = R.CentralDinoState(seed=11, mode="default")
    engines = {}
    for prec in ("fp32", "bf16"):
        e = DinoStepEngine(kind="multi_central", device=dev, precision=prec, learning_rate=args.lr)
        e.load_named(student=st.student, teacher=st.teacher, student_head=st.student_head, teacher_head=st.teacher_head)
        engines[prec] = e
    curves = {"oracle": [], "fp32": [], "bf16": []}
    t0 = time.time()
    for it in range(args.steps):
        k = it % args.pool
        img, aud = views_to_vb(*synth_views(B, seed=1000 + k))
        masks = make_masks(seed=5000 + it, V=6, Vg=2, B=B, E=256, hidden=512)
        want = R.central_dino_step(st, img, aud, masks, lr=args.lr)
        curves["oracle"].append(float(want["loss"]))
        gi, ga = img[:, :, 0].to(dev).contiguous(), aud[:, :, 0].to(dev).contiguous()
        gm = {k_: v.to(torch.uint8).to(dev) for k_, v in masks.items()}
        for prec, e in engines.items():
            loss = e.train_step_views(gi, ga, masks=gm)
            curves[prec].append(float(loss[3]))
        if it % 100 == 0:
            print(it, [round(curves[k_][-1], 5) for k_ in curves], f"{time.time() - t0:.0f}s", flush=True)

    def stats(a, b):
        d = [abs(x - y) for x, y in zip(a, b)]
        n = len(d)
        return {"max_abs": max(d), "mean_abs": sum(d) / n, "mean_abs_last_100": sum(d[-100:]) / min(100, n),
                "final": [a[-1], b[-1]], "mean_last_100": [sum(a[-100:]) / min(100, n), sum(b[-100:]) / min(100, n)]}

    out = {"steps": args.steps, "batch": B, "lr": args.lr, "pool": args.pool,
           "note": "default mode multi_central, identical inputs / masks / initial weights; oracle = CPU fp32 restatement of the reference",
           "fp32_vs_oracle": stats(curves["fp32"], curves["oracle"]), "bf16_vs_oracle": stats(curves["bf16"], curves["oracle"]),
           "curves_every_10": {k_: v[::10] for k_, v in curves.items()}}
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps({k_: out[k_] for k_ in ("fp32_vs_oracle", "bf16_vs_oracle")}))


if __name__ == "__main__":
    main()
