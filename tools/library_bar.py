"""Library-kernel bar: the reference's training step run with STOCK PyTorch kernels on the B200.

/root/reference does not exist on the GPU box, so the step that is timed is the oracle's restatement of the reference
(`oracle/dino_ref.central_dino_step`: the same torch calls as models/dino.py / models/unimodal.py) moved to cuda:0 and run
the way the reference runs on a GPU: eager ATen / cuDNN / cuBLAS, fp16 autocast (run_dino.py:360).  The inputs are
ALREADY-AUGMENTED views resident in HBM: the reference has no GPU augmentation (its torchvision chains run in DataLoader
workers on the CPU), so this bar EXCLUDES the augmentation that our step includes.

A measurement tool, not part of the product path: nothing in multimodal_ssl_avmnist_b200/ imports it.

    python tools/library_bar.py --batch 1024 --steps 10 --warmup 3 [--fp32] > gpurun_out/library_bar.json
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dino_ref, fixtures  # noqa: E402


def to_cuda(obj):
    if torch.is_tensor(obj):
        return obj.cuda()
    if isinstance(obj, dict):
        return {k: to_cuda(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_cuda(v) for v in obj)
    return obj


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--fp32", action="store_true", help="no autocast (TF32 convolutions / matmuls allowed, as torch defaults)")
    ap.add_argument("--channels-last", action="store_true")
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    B, V, Vg = args.batch, 6, 2
    st = dino_ref.CentralDinoState(seed=0)
    for name, val in list(vars(st).items()):
        setattr(st, name, to_cuda(val))
    g = torch.Generator(device="cuda").manual_seed(1)
    img = torch.rand(V, B, 1, 28, 28, device="cuda", generator=g)
    aud = torch.rand(V, B, 1, 112, 112, device="cuda", generator=g)
    if args.channels_last:
        pass        # C = 1 inputs: channels_last is the same memory; the conv stacks pick their own cuDNN layouts
    masks = to_cuda(fixtures.make_masks(3, V, Vg, B, st.E, 512))

    def step():
        with torch.autocast("cuda", dtype=torch.float16, enabled=not args.fp32):
            return dino_ref.central_dino_step(st, img, aud, masks)["loss"]

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"what": "stock PyTorch (cuDNN / cuBLAS / ATen eager) DINO step on B200, multi_central default mode, "
                              "views pre-augmented in HBM (no augmentation in the timed region)",
                      "precision": "fp32 (torch defaults)" if args.fp32 else "fp16 autocast (run_dino.py:360)",
                      "per_gpu_batch": B, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                      "samples_per_s": B / ms * 1e3, "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                      "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}))


if __name__ == "__main__":
    main()
