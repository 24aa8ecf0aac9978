"""Multi-GPU plumbing check (run under torchrun, N >= 2): every rank feeds the SAME views and dropout masks, so after the
gradient / centre all-reduces each rank must reproduce the single-process step exactly (sum of N identical gradients
times 1/N; mean over N identical row sets).  Also checks that ranks stay bit-identical with different per-rank data.
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/manual/dp_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from oracle.fixtures import make_masks, synth_views, views_to_vb
from multimodal_ssl_avmnist_b200.engine import DinoStepEngine

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
B = 8
ok = True
for mode in ("default", "infonce"):
    solo = DinoStepEngine(kind="multi_central", mode=mode, seed=3, device=dev, data_parallel=False)
    par = DinoStepEngine(kind="multi_central", mode=mode, seed=3, device=dev)
    assert par.world == world and solo.world == 1
    for it in range(2):
        img, aud = views_to_vb(*synth_views(B, seed=50 + it))
        masks = {k: v.to(torch.uint8).to(dev) for k, v in make_masks(seed=60 + it, V=6, Vg=2, B=B, E=256, hidden=512).items()}
        xi, xa = img[:, :, 0].to(dev).contiguous(), aud[:, :, 0].to(dev).contiguous()
        raw = labels = None
        if mode != "default":
            raw = (xi[0].clone(), xa[0].clone())
        l1 = solo.forward_backward(xi, xa, masks=masks, raw=raw).clone()
        l2 = par.forward_backward(xi, xa, masks=masks, raw=raw).clone()
        torch.cuda.synchronize()
        # the all-reduced gradient is the SUM over ranks (1/N is applied inside Adam): N identical gradients -> N * g, exact
        d_grad = float((solo.grad * world - par.grad).abs().max() / (solo.grad.abs().max() * world))
        for e in (solo, par):
            e.update_teacher()
            e.optimizer_step()
            e.rng_step += 1
        torch.cuda.synchronize()
        d_loss = float((l1 - l2).abs().max())
        d_par = float((solo.student.flat - par.student.flat).abs().max())
        d_tea = float((solo.teacher.flat - par.teacher.flat).abs().max())
        d_cen = float((solo.center - par.center).abs().max())
        # tolerances: fp64 atomics inside the BatchNorm reductions make gradients reproducible only to ~1 ulp; Adam turns a
        # 1-ulp difference on a rounding-noise gradient into at most a 2*lr parameter difference
        good = d_loss <= 5e-6 and d_grad <= 2e-5 and d_par <= 3e-4 and d_tea <= 2e-6 and d_cen <= 1e-6
        ok &= good
        if rank == 0:
            print(f"mode={mode} step={it} |dloss|={d_loss:.1e} |dgrad|/max={d_grad:.1e} |dparam|={d_par:.1e} |dteacher|={d_tea:.1e} "
                  f"|dcenter|={d_cen:.1e} {'OK' if good else 'FAIL'}")
    # different data per rank: replicas must remain identical (same all-reduced gradients, same centre)
    eng = DinoStepEngine(kind="multi_central", mode=mode, seed=3, device=dev)
    for it in range(3):
        g = torch.Generator().manual_seed(100 * rank + it)
        img = torch.rand(B, 28, 28, generator=g).to(dev)
        aud = torch.randint(0, 256, (B, 112, 112), generator=g, dtype=torch.uint8).to(dev)
        eng.seed = 7 + rank
        eng.train_step(img, aud)
    chk = torch.stack([eng.student.flat.double().sum(), eng.teacher.flat.double().sum(), eng.center.double().sum()])
    gathered = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    same = all(torch.equal(gathered[0], g_) for g_ in gathered)
    ok &= same
    if rank == 0:
        print(f"mode={mode} replicas identical after 3 steps on different shards: {same}")
# CUDA-graph replay of the data-parallel step: the NCCL all-reduces (C ABI, communication stream) are captured with the kernels
eager = DinoStepEngine(kind="multi_central", mode="default", seed=11, device=dev)      # same initial weights on every rank ...
graph = DinoStepEngine(kind="multi_central", mode="default", seed=11, device=dev)
eager.seed = graph.seed = 11 + rank                                                    # ... different augmentation / dropout streams
graph.student.flat.copy_(eager.student.flat)
graph.sync_teacher()
eager.sync_teacher()
gen = torch.Generator().manual_seed(1000 + rank)
batches = [(torch.rand(B, 28, 28, generator=gen).to(dev), torch.randint(0, 256, (B, 112, 112), generator=gen, dtype=torch.uint8).to(dev))
           for _ in range(4)]
eager.train_step(*batches[0])
graph.train_step(*batches[0])
graph.capture_train_step(B)
for img, aud in batches[1:]:
    eager.train_step(img, aud)
    graph.graph_step(img, aud)
torch.cuda.synchronize()
same_graph = torch.equal(eager.student.flat, graph.student.flat) and torch.equal(eager.teacher.flat, graph.teacher.flat) \
    and torch.equal(eager.center, graph.center)
chk = eager.student.flat.double().sum().reshape(1)
gathered = [torch.empty_like(chk) for _ in range(world)]
dist.all_gather(gathered, chk)
same_replicas = all(torch.equal(gathered[0], g_) for g_ in gathered)
ok &= same_graph and same_replicas
if rank == 0:
    print(f"data-parallel CUDA-graph replay == eager data-parallel steps (bit for bit): {same_graph}; replicas identical: {same_replicas}; "
          f"NCCL {eager.comm.version} through the C ABI")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DP_CHECK", "PASS" if int(flag) == 1 else "FAIL")
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
