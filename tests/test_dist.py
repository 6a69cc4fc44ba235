"""Host-side multi-GPU logic on CPU: two gloo processes shard a ragged batch with no data-path
collective, each hashes its own shard (the oracle stands in for the GPU kernel here), and the only
communication is the barrier / MAX / checksum reduction bench.py also uses."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from nthash_b200 import dist as nd
    from oracle_lib import ORACLE
    r, w = nd.init("gloo")
    assert (r, w) == (rank, world)
    rng = np.random.default_rng(7)  # same batch on every rank
    lens = rng.integers(0, 400, 5000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    bases = rng.choice(np.frombuffer(b"ACGTN", np.uint8), int(off[-1]), p=[0.2499, 0.2499, 0.2499, 0.2499, 0.0004])
    lo, hi = nd.shard_by_bases(off, world)[rank]
    mine = ORACLE.kmer_batch(bases, off[lo:hi + 1], 31, 1, want=())       # absolute offsets: no copy, no exchange
    nd.barrier()
    total_sum = nd.sum_over_ranks(mine["sum"])
    step_ms = nd.max_over_ranks([10.0 + rank, 1.0])
    whole = ORACLE.kmer_batch(bases, off, 31, 1, want=())
    q.put((rank, lo, hi, int(off[hi] - off[lo]), total_sum == whole["sum"], step_ms))
    nd.finalize()


def test_two_rank_sharding_with_gloo():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, b0, ok0, ms0), (r1, lo1, hi1, b1, ok1, ms1) = res
    assert lo0 == 0 and hi0 == lo1 and hi1 == 5000          # shards partition the batch
    assert abs(b0 - b1) <= 400                               # balanced by bases (within one read)
    assert ok0 and ok1                                       # checksum of checksums == single-process checksum
    assert ms0 == ms1 == [11.0, 1.0]                         # slowest rank defines the step


def test_shard_helpers():
    sys.path.insert(0, ROOT)
    from nthash_b200 import dist as nd
    assert [nd.shard_uniform(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    off = np.array([0, 10, 10, 500, 510, 1000])
    shards = nd.shard_by_bases(off, 3)
    assert shards[0][0] == 0 and shards[-1][1] == 5 and all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
    assert nd.shard_by_bases(np.array([0]), 2) == [(0, 0), (0, 0)]
    assert nd.max_over_ranks([1.5]) == [1.5] and nd.sum_over_ranks(2**64 + 5) == 5
