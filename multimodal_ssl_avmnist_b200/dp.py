"""Data-parallel host logic (one process per GPU, torch.distributed for the plumbing).

The step shards by samples: global batch G -> contiguous G/N samples per rank; weights replicated;
BatchNorm statistics and InfoNCE negatives stay rank-local.  Exactly two exchanges cross GPUs per step, both
`all_reduce(sum)`:
  * the trainable prefix of the flat gradient arena (plus the mode-head range) -- the 1/N average is folded into Adam's
    grad_scale instead of a separate scaling pass;
  * the [D] column sums of the un-centred teacher projections, from which every rank forms the same centre EMA
    (mean over ALL ranks' rows == the single-process large-batch result).
On GPUs the exchange runs INSIDE the C ABI (`AbiComm`: b200_dp_* of libavmnist_b200.so, NCCL over NVLink, asynchronous on a
communication stream and capturable into the step's CUDA graph); torch.distributed only carries the 128-byte NCCL unique id
between the ranks.  The torch.distributed functions below are the same logic on CPU tensors with gloo (tests/test_dp_gloo.py).
"""
import ctypes

import torch
import torch.distributed as dist


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None):
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def shard_range(global_batch, rank_, world):
    """Contiguous sample range [lo, hi) of `rank_`; the first (global_batch % world) ranks take one extra sample."""
    if not (0 <= rank_ < world) or global_batch < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(global_batch, world)
    lo = rank_ * base + min(rank_, extra)
    return lo, lo + base + (1 if rank_ < extra else 0)


class GradientPlan:
    """Which slices of the flat gradient arena are exchanged: [(lo, hi), ...] in floats."""

    def __init__(self, ranges):
        self.ranges = [(int(lo), int(hi)) for lo, hi in ranges if hi > lo]

    def bytes(self):
        return 4 * sum(hi - lo for lo, hi in self.ranges)

    def allreduce_(self, flat_grad, group=None):
        """In-place sum over ranks of every planned slice.  Returns the grad_scale (1/world) Adam must apply."""
        w = world_size(group)
        if w > 1:
            for lo, hi in self.ranges:
                dist.all_reduce(flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=group)
        return 1.0 / w


def allreduce_colsum_(colsum, n_rows_local, group=None):
    """Sum the teacher column sums over ranks; returns the global row count the centre mean must divide by."""
    w = world_size(group)
    if w > 1:
        dist.all_reduce(colsum, op=dist.ReduceOp.SUM, group=group)
    return n_rows_local * w


def center_ema_reference(center, colsum, n_rows, m_c):
    """Plain-torch statement of what b200_center_apply computes (used by the CPU tests of this module only)."""
    return center * m_c + (colsum / n_rows) * (1 - m_c)


class AbiComm:
    """The process-wide NCCL communicator inside libavmnist_b200.so (b200_dp_init / _allreduce_* / _destroy).
    Bootstrap: rank 0 draws the NCCL unique id (b200_dp_unique_id), torch.distributed broadcasts its 128 bytes."""

    _instance = None

    @classmethod
    def get(cls, group=None):
        if cls._instance is None:
            cls._instance = cls(group)
        return cls._instance

    def __init__(self, group=None):
        from . import _lib
        self.lib = _lib.load()
        self._check = _lib.check
        self.rank, self.world = rank(group), world_size(group)
        if self.world < 2:
            raise _lib.B200Error("AbiComm needs an initialised torch.distributed process group with world_size > 1")
        buf = ctypes.create_string_buffer(128)
        if self.rank == 0:
            self._check(self.lib.b200_dp_unique_id(buf), "dp_unique_id")
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = bytes(t.cpu().tolist())
        self._check(self.lib.b200_dp_init(uid, self.rank, self.world), "dp_init")
        self.version = int(self.lib.b200_dp_nccl_version())

    def allreduce_grads_(self, t, stream):
        """in-place sum over ranks of a contiguous fp32 CUDA slice, asynchronous on `stream` (a raw cudaStream_t)"""
        self._check(self.lib.b200_dp_allreduce_grads(t.data_ptr(), t.numel(), stream), "dp_allreduce_grads")

    def allreduce_center_(self, colsum, stream):
        self._check(self.lib.b200_dp_allreduce_center(colsum.data_ptr(), colsum.numel(), stream), "dp_allreduce_center")

    def destroy(self):
        self._check(self.lib.b200_dp_destroy(), "dp_destroy")
        type(self)._instance = None
