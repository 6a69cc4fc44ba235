#!/usr/bin/env python
"""Generates tests/golden/ref_vectors.npz from the UNMODIFIED reference (bcgsc/ntHash 2.4.0 compiled from
/root/reference into oracle/_ref by oracle/Makefile).  Run in the build container (the reference tree does not
exist on the GPU box):   python tests/golden/make_golden.py
Every case stores its inputs and the reference's outputs in the engine's dense layout (row = koff[read] + pos)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import ORACLE, REF  # noqa: E402

assert REF is not None, "oracle/_ref is missing: run `make -C oracle` where /root/reference exists"
BAD = np.frombuffer(b"NnRYKMSWBDHV-*.", np.uint8)


def reads(seed, lens, p_bad, lower):
    rng = np.random.default_rng(seed)
    n = int(sum(lens))
    a = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
    m = rng.random(n) < lower
    a[m] |= 0x20
    a[(rng.random(n) < lower / 2) & (a == ord("T"))] = ord("U")
    bad = rng.random(n) < p_bad
    a[bad] = rng.choice(BAD, int(bad.sum()))
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    return a, off


out = {}
cases = []
# NtHash: BASELINE configs[0] (1 kb, seed 42 of the survey's generator), ragged dirty reads for several k / h
c1 = ORACLE.gen_bases(1000, 42)
kmer_cases = [("c1_k31_h1", c1, np.array([0, 1000], np.uint64), 31, 1), ("c1_k31_h4", c1, np.array([0, 1000], np.uint64), 31, 4),
              ("c1_k63_h1", c1, np.array([0, 1000], np.uint64), 63, 1)]
for i, (k, h) in enumerate([(3, 1), (5, 3), (31, 2), (32, 1), (33, 4), (64, 1), (127, 2), (255, 1)]):
    lens = np.random.default_rng(100 + i).integers(0, 2 * k + 90, 24)
    lens[:5] = [0, k - 1, k, k + 1, 2 * k]
    a, off = reads(200 + i, lens, 0.01, 0.1)
    kmer_cases.append((f"ragged_k{k}_h{h}", a, off, k, h))
a, off = reads(7, [150] * 24, 0.002, 0.0)
kmer_cases.append(("uniform150_k31_h1", a, off, 31, 1))
for name, a, off, k, h in kmer_cases:
    r = REF.kmer_batch(a, off, k, h)
    out[f"kmer/{name}/bases"], out[f"kmer/{name}/off"], out[f"kmer/{name}/kh"] = a, off, np.array([k, h])
    for key in ("out", "valid", "fwd", "rev"):
        out[f"kmer/{name}/{key}"] = r[key]
    cases.append(name)
# SeedNtHash: the survey's C4 seeds, the reference tests' seed, an asymmetric seed set, dirty bytes incl. NUL
seed_cases = [("c4_seeds_h3", ["1010101010101010101010101010101", "1101101101101101011011011011011"], 3, 0.0),
              ("tests_cpp_seed_h3", ["11100111"], 3, 0.01), ("asymmetric_h2", ["1101001110", "1011100011", "1111100000"], 2, 0.01),
              ("k21_three_seeds_h1", ["111011101110111011101", "101010111101111010101", "110000111111111000011"], 1, 0.005)]
for i, (name, seeds, h, p_bad) in enumerate(seed_cases):
    k = len(seeds[0])
    lens = np.random.default_rng(300 + i).integers(0, 3 * k + 100, 24)
    lens[:4] = [k, k + 1, 2 * k, 150]
    a, off = reads(400 + i, lens, p_bad, 0.05)
    if p_bad:
        a[int(off[3]) + k // 2] = 0  # a NUL byte inside a window (seed.cpp:151 tests against it)
    r = REF.seed_batch(a, off, seeds, h)
    out[f"seed/{name}/bases"], out[f"seed/{name}/off"], out[f"seed/{name}/h"] = a, off, np.array([h])
    out[f"seed/{name}/seeds"] = np.array(seeds)
    for key in ("out", "valid", "fwd", "rev"):
        out[f"seed/{name}/{key}"] = r[key]
# BlindNtHash: states fed with caller-supplied bases
rng = np.random.default_rng(9)
for k, h in ((5, 3), (31, 1), (64, 2)):
    kmer = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), k))
    feed = bytes(rng.choice(np.frombuffer(b"ACGTacgtU", np.uint8), 50))
    h0, hv, fw, rv = REF.blind_read(kmer, h, feed)
    name = f"blind/k{k}_h{h}"
    out[f"{name}/kmer"], out[f"{name}/feed"] = np.frombuffer(kmer, np.uint8), np.frombuffer(feed, np.uint8)
    out[f"{name}/h0"], out[f"{name}/hashes"], out[f"{name}/fwd"], out[f"{name}/rev"] = h0, hv, fw, rv
np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
print("wrote", os.path.join(HERE, "ref_vectors.npz"), os.path.getsize(os.path.join(HERE, "ref_vectors.npz")), "bytes;", len(out), "arrays")
