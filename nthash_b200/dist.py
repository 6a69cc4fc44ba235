"""One process per GPU: how a read batch is split over ranks and how per-rank timings are combined.

The hash path has no exchange step (reads are independent: SURVEY.md §8e), so the only collective
is the barrier + MAX reduction of the timing; data never crosses GPUs.  torch.distributed is the
plumbing: backend "nccl" on the GPU box, "gloo" in the CPU tests (tests/test_dist.py)."""
import os

import numpy as np
import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment; no-op for a single process."""
    rank, world, _ = env_rank()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), **kw)
    return rank, world


def shard_uniform(n_reads, rank, world):
    """Contiguous read range [lo, hi) of `rank` when all reads have the same length."""
    return n_reads * rank // world, n_reads * (rank + 1) // world


def shard_by_bases(read_off, world):
    """Cut ragged reads into `world` contiguous ranges of (nearly) equal total bases.

    read_off: int64/uint64 array of n+1 byte offsets.  Returns a list of (lo, hi) read ranges that
    partition [0, n)."""
    off = np.asarray(read_off, dtype=np.int64)
    n = len(off) - 1
    total = int(off[-1] - off[0])
    cuts = [0]
    for r in range(1, world):
        target = off[0] + total * r // world
        cuts.append(int(np.searchsorted(off, target, side="left")))
    cuts.append(n)
    cuts = [min(max(c, 0), n) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def barrier():
    if dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(values):
    """Element-wise MAX of a list of floats over all ranks (the slowest rank defines the step time)."""
    if not dist.is_initialized():
        return list(values)
    dev = "cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu"
    t = torch.tensor(list(values), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(value):
    """64-bit wrap-around sum of one unsigned checksum per rank (a checksum of checksums)."""
    if not dist.is_initialized():
        return int(value) & (2**64 - 1)
    dev = "cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu"
    v = int(value) & (2**64 - 1)
    t = torch.tensor([v - 2**64 if v >= 2**63 else v], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)  # int64 addition wraps like uint64 addition
    return int(t[0]) & (2**64 - 1)


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()
