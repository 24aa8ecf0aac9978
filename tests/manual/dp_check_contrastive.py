"""torchrun --nproc-per-node 2 tests/manual/dp_check_contrastive.py -- data-parallel plumbing of ContrastiveStepEngine on real GPUs:
(1) every rank steps on the SAME shard: the all-reduced, 1/world-scaled gradient equals the local one exactly, so the data-parallel engine
must reproduce a single-process engine bit for bit (loss, parameters, Adam moments); (2) on DIFFERENT shards the replicas stay identical."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
from multimodal_ssl_avmnist_b200.contrastive import ContrastiveStepEngine

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
ok = True
B = 16
for kind in ("infonce", "simclr"):
    torch.manual_seed(5)
    random.seed(5)                                         # same host-sampled SimCLR parameters on every rank and in both engines
    e_dp = ContrastiveStepEngine(kind=kind, device=dev, seed=3, learning_rate=1e-3)
    e_one = ContrastiveStepEngine(kind=kind, device=dev, seed=3, learning_rate=1e-3, data_parallel=False)
    assert e_dp.world == dist.get_world_size() and e_one.world == 1
    g = torch.Generator().manual_seed(11)
    for it in range(3):
        img = torch.rand(B, 28, 28, generator=g).to(dev)
        aud = torch.randint(0, 256, (B, 112, 112), generator=g, dtype=torch.uint8).to(dev)
        st, pst = torch.get_rng_state(), random.getstate()
        l1 = e_dp.train_step(img, aud).clone()
        torch.set_rng_state(st)                            # the single-process engine draws the same augmentation parameters
        random.setstate(pst)                               # (the time-warp rate comes from Python's generator, like in the reference)
        l2 = e_one.train_step(img, aud).clone()
        same = torch.equal(l1, l2) and torch.equal(e_dp.student.flat, e_one.student.flat) and torch.equal(e_dp.exp_avg, e_one.exp_avg)
        ok &= same
        if rank == 0:
            print(f"kind={kind} step={it} same-shard DP == single process (bit for bit): {same}  loss {float(l1[3]):.6f}  used {e_dp._used}")
    # different shards per rank: replicas stay identical
    g2 = torch.Generator().manual_seed(100 + rank)
    for it in range(3):
        img = torch.rand(B, 28, 28, generator=g2).to(dev)
        aud = torch.randint(0, 256, (B, 112, 112), generator=g2, dtype=torch.uint8).to(dev)
        e_dp.train_step(img, aud)
    flat = e_dp.student.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = torch.equal(flat, ref)
    ok &= same
    if rank == 0:
        print(f"kind={kind} replicas identical after 3 steps on different shards: {same}")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DP_CHECK_CONTRASTIVE", "PASS" if int(flag) else "FAIL")
dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
