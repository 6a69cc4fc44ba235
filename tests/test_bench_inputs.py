"""bench.py's synthetic reads are the generator SURVEY.md 8(d) / Appendix C pins (splitmix64, 32 bases per draw), in
both arms, and both arms print the same `config` object (CPU only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle_lib import ORACLE  # noqa: E402


def test_splitmix_generators_match_the_oracle_generator():
    want = ORACLE.gen_bases(200_000, 42)
    assert bytes(want[:64]) == b"CCCGGTGCTGGTTTGAGCGAGATATCCTCTTGTAAACATTGCGCGATGTATATAGTTTGTAGGA"
    assert (bench.splitmix_bases_numpy(200_000, 42) == want).all()
    t = bench.splitmix_bases_torch(torch, 200_000, 42, device="cpu", slack=64)
    assert t.numel() == 200_064 and (t[:200_000].numpy() == want).all() and int(t[200_000:].sum()) == 0


def test_rank_shards_are_slices_of_one_stream():
    full = ORACLE.gen_bases(4 * 6400, 44)
    for r in range(4):
        assert (bench.splitmix_bases_numpy(6400, 44, first_base=r * 6400) == full[r * 6400:(r + 1) * 6400]).all()
        assert (bench.splitmix_bases_torch(torch, 6400, 44, first_base=r * 6400, device="cpu")[:6400].numpy() == full[r * 6400:(r + 1) * 6400]).all()
    for name, cfg in bench.CONFIGS.items():  # every rank's shard starts on a draw boundary
        assert (cfg["n_reads"] * cfg["read_len"]) % 32 == 0, name


def test_c5_is_config4_as_stated_at_eight_ranks():
    c5 = bench.CONFIGS["c5"]
    assert 8 * c5["n_reads"] == 100_000 and c5["read_len"] == 50_000 and c5["k"] == 63 and c5["h"] == 1


def test_cpu_reference_measure_small():
    cfg = dict(bench.CONFIGS["c2"], n_reads=2000)
    m = bench.cpu_reference_measure(cfg, 2000, passes=2)
    assert m["emitted_per_pass"] == 2000 * 120 and m["passes"] == 2 and m["value"] > 0
    ora = ORACLE.kmer_batch(bench.splitmix_bases_numpy(2000 * 150, 42), np.arange(2001, dtype=np.uint64) * 150, 31, 1, want=())
    assert ora["sum"] == m["sum"]
    assert bench.workload_config(cfg, 2000)["reads_per_gpu"] == 2000
