#!/usr/bin/env python
"""Raw PCIe floor of bench.py's end-to-end step at N GPUs: every rank moves what one C2 step moves (1.5 GB up from pinned host
memory, 9.75 GB + bitmap down) on two streams at once, nothing else — all ranks concurrently, max over ranks, like the bench.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        profiles/microbench/pcie_floor.py [--affinity]

--affinity binds each rank to the CPUs `nvidia-smi topo -m` lists next to its GPU before the pinned buffers are allocated
(first touch then lands on that NUMA node).  Prints one JSON line per variant on rank 0."""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def topo_cpus(gpu):
    """CPU affinity column of `nvidia-smi topo -m` for one GPU, as a set of CPU ids (None when it cannot be parsed)."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        hdr = None
        for line in out.splitlines():
            cols = line.split("\t")
            if hdr is None and "CPU Affinity" in line:
                hdr = [c.strip() for c in cols]
                continue
            if hdr and cols and cols[0].strip() == f"GPU{gpu}":
                spec = cols[hdr.index("CPU Affinity")].strip()
                cpus = set()
                for part in spec.split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
                return cpus
    except Exception:
        return None
    return None


note = {"affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."], "n_cpus": os.cpu_count()}
if "--affinity" in sys.argv:
    cpus = topo_cpus(local)
    if cpus:
        os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
    note["bound_to"] = len(os.sched_getaffinity(0))

UP, DOWN = 1_500_000_000, 9_750_000_000 + 37_500_000
h_up = torch.empty(UP, dtype=torch.uint8).pin_memory()
h_up.fill_(65)
h_down = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
h_down.fill_(0)
d_up = torch.empty(UP, dtype=torch.uint8, device="cuda")
d_down = torch.zeros(DOWN, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def step(up=True, down=True):
    if up:
        with torch.cuda.stream(s1):
            d_up.copy_(h_up, non_blocking=True)
    if down:
        with torch.cuda.stream(s2):
            h_down.copy_(d_down, non_blocking=True)


def timed(up, down, reps=3):
    step(up, down)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        step(up, down)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


res = {"n_gpus": world, "variant": "affinity from nvidia-smi topo" if "--affinity" in sys.argv else "default placement", **note}
for name, up, down, nbytes in (("both", True, True, UP + DOWN), ("down_only", False, True, DOWN), ("up_only", True, False, UP)):
    dt = timed(up, down)
    res[name] = {"ms_per_step": dt * 1e3, "aggregate_GBps": world * nbytes / dt / 1e9, "per_gpu_GBps": nbytes / dt / 1e9}
res["e2e_floor_kmers_per_s"] = world * 1_200_000_000 / (res["both"]["ms_per_step"] * 1e-3)
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
