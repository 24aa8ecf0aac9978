"""Per-stream timeline of ONE overlapped training step from CUDA events (no nsys in this image): every op wrapper records a start / end
event on the stream it launches on; positions are relative to an event recorded on the main stream at the start of the step.
    python tools/timeline_step.py [B] [kind] > gpurun_out/timeline.txt
Prints the ops in start order with their stream, start, end, and for the MAIN stream the idle gaps between consecutive ops."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import augment_values
from multimodal_ssl_avmnist_b200 import ops
from multimodal_ssl_avmnist_b200.engine import DinoStepEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
KIND = sys.argv[2] if len(sys.argv) > 2 else "multi_central"
eng = DinoStepEngine(kind=KIND, augment_values=augment_values(), seed=1, device="cuda:0")
g = torch.Generator().manual_seed(1)
img = torch.rand(B, 28, 28, generator=g).cuda()
aud = torch.randint(0, 256, (B, 112, 112), generator=g, dtype=torch.uint8).cuda()
for _ in range(4):
    eng.train_step(img, aud)
    eng.prefetch_augment(img, aud)
torch.cuda.synchronize()
# steady state: the host runs ahead of the GPU (no synchronisation between steps); only the LAST of a few back-to-back steps is recorded
for _ in range(3):
    eng.train_step(img, aud)
    eng.prefetch_augment(img, aud)
t0 = torch.cuda.Event(enable_timing=True)
t0.record()
rec = ops.start_profile()
eng.train_step(img, aud)
eng.prefetch_augment(img, aud)
ops.stop_profile()
t1 = torch.cuda.Event(enable_timing=True)
t1.record()
torch.cuda.synchronize()
main = torch.cuda.current_stream().cuda_stream
streams = {}
rows = []
for name, a, b, meta, st in rec:
    sid = streams.setdefault(st, len(streams))
    rows.append((t0.elapsed_time(a), t0.elapsed_time(b), sid, name, meta[0] if meta else ()))
rows.sort()
print(f"step {t0.elapsed_time(t1):.3f} ms (with event overhead); streams: {len(streams)} (0 = first used = main)")
busy = {}
last_end = {}
for s, e, sid, name, shp in rows:
    gap = s - last_end.get(sid, 0.0)
    last_end[sid] = e
    busy[sid] = busy.get(sid, 0.0) + (e - s)
    print(f"{s:8.3f} {e:8.3f}  s{sid}  {e - s:7.3f} ms  gap {gap:7.3f}  {name:28s} {str(shp)[:60]}")
print("busy per stream (ms):", {k: round(v, 3) for k, v in sorted(busy.items())})
# union coverage of all ops
iv = sorted((s, e) for s, e, *_ in rows)
cov, cur_s, cur_e = 0.0, None, None
for s, e in iv:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            cov += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
cov += (cur_e - cur_s) if cur_e is not None else 0.0
print(f"time covered by at least one op: {cov:.3f} ms; first start {iv[0][0]:.3f}, last end {max(e for _, e in iv):.3f}")
