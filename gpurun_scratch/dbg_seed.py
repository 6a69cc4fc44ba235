import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import nthash_b200
from gpu_util import synth, to_dev, u64
from oracle_lib import ORACLE
seeds, h, n, L, p_bad = ["1" * 40 + "0" * 23 + "1" * 40], 1, 500, 300, 0.003
rng = np.random.default_rng(n + L)
bases = synth(rng, n * L, p_bad=p_bad, lower=0.05)
bases[rng.integers(0, n * L, 5)] = 0
bases[rng.integers(0, n * L, 20)] = rng.choice(np.array([1, 3, 4, 5, 7], np.uint8), 20)
d_b, _keep = to_dev(bases)
plan = nthash_b200.SeedPlan(seeds, h)
off = np.arange(n + 1, dtype=np.uint64) * L
ora = ORACLE.seed_batch(bases, off, seeds, h, threads=8)
for env in (None, "NTHASH_B200_DISABLE_SEED_JIT"):
    if env: os.environ[env] = "1"
    for strands in (False, True):
        res = nthash_b200.seed_hashes_uniform(plan, d_b, n, L, want_strands=strands)
        torch.cuda.synchronize()
        out = u64(res.out).reshape(-1, 1)
        vm = res.valid_mask().cpu().numpy()
        bad_v = np.argwhere(vm != ora["valid"].astype(bool)).reshape(-1)
        bad_o = np.argwhere((out != ora["out"]).any(axis=1)).reshape(-1)
        print("env", env, "strands", strands, "valid mismatches", len(bad_v), bad_v[:8], "hash mismatches", len(bad_o), bad_o[:8])
        nk = L - 103 + 1
        for w in list(bad_v[:3]) + list(bad_o[:3]):
            r, p = divmod(int(w), nk)
            seq = bases[r * L:(r + 1) * L]
            badpos = [(i, int(c)) for i, c in enumerate(seq) if chr(c) not in "ACGTUacgtu"]
            print("  row", w, "read", r, "pos", p, "gpu valid", bool(vm[w]), "ora valid", bool(ora["valid"][w]), "gpu", hex(int(out[w, 0])), "ora", hex(int(ora["out"][w, 0])), "non-ACGT at", badpos)
