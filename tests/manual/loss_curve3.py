"""1k-step loss curves, four arms on identical per-step inputs (pre-augmented synthetic views, shared dropout masks) from
identical initial weights:

  oracle_fp32      the oracle's torch calls on cuda:0 in fp32 (TF32 off): the reference's arithmetic, stock ATen / cuDNN kernels
  oracle_fp16ac    the same under torch.autocast(fp16): the reference's REAL GPU precision (run_dino.py:360, precision='16-mixed')
  engine_fp32      this repo's exact-fp32 CUDA path
  engine_bf16      this repo's product path (tcgen05 convolutions, bf16 / fp16 activations, tf32 linears)

Acceptance bound (asserted at the end, exit code 1 if violated): the product path's deviation from the fp32 oracle is no larger
than 1.5 x the deviation of the reference's own fp16-autocast arithmetic (mean |dloss| over all steps and over the last 100).
Also reports, at step 0, the cosine between the fp16-autocast gradient and the fp32 gradient per weight matrix -- the measured
size of the routing-flip effect that the bf16 step test's direction tolerance (cos > 0.97) rests on.

    python tests/manual/loss_curve3.py --steps 1000 --batch 64 --out gpurun_out/r2_loss_curve_1k_b64.json
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


def to_cuda(obj):
    if torch.is_tensor(obj):
        return obj.cuda()
    if isinstance(obj, dict):
        return {k: to_cuda(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_cuda(v) for v in obj)
    return obj


def cuda_state(seed, kind="multi_central"):
    st = R.CentralDinoState(seed=seed, mode="default", kind=kind)
    for name, val in list(vars(st).items()):
        setattr(st, name, to_cuda(val))
    return st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--pool", type=int, default=64, help="number of distinct synthetic batches cycled through")
    ap.add_argument("--kind", default="multi_central", help="multi_central | multi_simple | multi_simple_gated | multi_cross_attention")
    ap.add_argument("--out", default="gpurun_out/r2_loss_curve.json")
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev, B = "cuda:0", args.batch
    cpu0 = R.CentralDinoState(seed=11, mode="default", kind=args.kind)
    arms = {"oracle_fp32": cuda_state(11, args.kind), "oracle_fp16ac": cuda_state(11, args.kind)}
    engines = {}
    for prec in ("fp32", "bf16"):
        e = DinoStepEngine(kind=args.kind, device=dev, precision=prec, learning_rate=args.lr)
        e.load_named(student=cpu0.student, teacher=cpu0.teacher, student_head=cpu0.student_head, teacher_head=cpu0.teacher_head)
        engines["engine_" + prec] = e
    curves = {k: [] for k in list(arms) + list(engines)}
    grad_cos = None
    t0 = time.time()
    for it in range(args.steps):
        k = it % args.pool
        img, aud = views_to_vb(*synth_views(B, seed=1000 + k))
        masks = make_masks(seed=5000 + it, V=6, Vg=2, B=B, E=256, hidden=512)
        gimg, gaud, gmask = img.cuda(), aud.cuda(), to_cuda(masks)
        outs = {}
        for name, st in arms.items():
            with torch.autocast("cuda", dtype=torch.float16, enabled=name.endswith("fp16ac")):
                outs[name] = R.central_dino_step(st, gimg, gaud, gmask, lr=args.lr, loss_scale=65536.0 if name.endswith("fp16ac") else 1.0)
            curves[name].append(float(outs[name]["loss"]))
        if it == 0:        # the reference's own fp16 routing-flip effect on the gradient, per weight matrix
            grad_cos = {}
            for grp in ("student", "student_head"):
                for name, g32 in outs["oracle_fp32"]["grads"][grp].items():
                    g16 = outs["oracle_fp16ac"]["grads"][grp].get(name)
                    if g16 is not None and g32.dim() >= 2:
                        a, b = g16.double().flatten(), g32.double().flatten()
                        grad_cos[f"{grp}.{name}"] = float((a @ b) / (a.norm() * b.norm() + 1e-300))
        gi, ga = gimg[:, :, 0].contiguous(), gaud[:, :, 0].contiguous()
        gm = {k_: v.to(torch.uint8) for k_, v in gmask.items()}
        for name, e in engines.items():
            curves[name].append(float(e.train_step_views(gi, ga, masks=gm)[3]))
        if it % 100 == 0:
            print(it, {k_: round(v[-1], 5) for k_, v in curves.items()}, f"{time.time() - t0:.0f}s", flush=True)

    ref = curves["oracle_fp32"]

    def stats(a):
        d = [abs(x - y) for x, y in zip(a, ref)]
        n = len(d)
        return {"max_abs": max(d), "mean_abs": sum(d) / n, "mean_abs_last_100": sum(d[-100:]) / min(100, n), "final": a[-1],
                "mean_last_100": sum(a[-100:]) / min(100, n)}

    summary = {k_: stats(v) for k_, v in curves.items()}
    bound = {"rule": "engine_bf16 deviation from oracle_fp32 <= 1.5 x oracle_fp16ac deviation (mean_abs and mean_abs_last_100)",
             "mean_abs": [summary["engine_bf16"]["mean_abs"], 1.5 * summary["oracle_fp16ac"]["mean_abs"]],
             "mean_abs_last_100": [summary["engine_bf16"]["mean_abs_last_100"], 1.5 * summary["oracle_fp16ac"]["mean_abs_last_100"]]}
    bound["ok"] = all(a <= b for a, b in (bound["mean_abs"], bound["mean_abs_last_100"]))
    out = {"steps": args.steps, "batch": B, "lr": args.lr, "pool": args.pool,
           "kind": args.kind,
           "note": "default mode; identical inputs / masks / initial weights in all four arms; deviations are |loss - oracle_fp32 loss|",
           "summary": summary, "bound": bound,
           "fp16_autocast_vs_fp32_gradient_cosine_step0": {"min": min(grad_cos.values()), "per_matrix": grad_cos},
           "curves_every_10": {k_: v[::10] for k_, v in curves.items()}}
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps({"summary": summary, "bound": bound, "fp16ac_grad_cos_min": min(grad_cos.values())}))
    sys.exit(0 if bound["ok"] else 1)


if __name__ == "__main__":
    main()
