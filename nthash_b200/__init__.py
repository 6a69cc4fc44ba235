"""nthash_b200 — B200-native (sm_100a) batch engine for the ntHash v2 rolling k-mer hash.

Layout: csrc/ holds the hand-written CUDA kernels and the C ABI (include/nthash_b200.h);
_lib.py binds that ABI with ctypes; api.py is a thin device-resident front end that uses
PyTorch only for HBM buffers and streams; dist.py shards read batches over one process per GPU.
"""
from ._lib import LIB, LIB_PATH, NtHashError  # noqa: F401
from .api import (HashBatch, RaggedPlan, SeedPlan, kmer_hashes_planned, seed_hashes_planned, blind_peek4, blind_roll, blind_seed_roll, kmer_hashes, kmer_hashes_uniform, kmer_hashes_packed2bit_uniform,  # noqa: F401
                  bloom_filter, compact, fastq_extract, kmer_bloom, kmer_bloom_uniform, kmer_reduce, kmer_reduce_uniform,
                  seed_hashes, seed_hashes_uniform, seed_reduce_uniform, kmer_sketch, kmer_sketch_uniform, kmer_minimizers_uniform)

FN_NAME = LIB.nthash_fn_name().decode()
