"""Golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py -> ref_vectors.npz).

`-m "not gpu"`: the CPU oracle (oracle/nthash_oracle.c) must reproduce every vector - this is what pins the oracle
on a machine where /root/reference and oracle/_ref do not exist.  `-m gpu`: the CUDA path, through the C ABI, must
reproduce them too."""
import os

import numpy as np
import pytest

from oracle_lib import ORACLE

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))
KMER = sorted({k.split("/")[1] for k in G.files if k.startswith("kmer/")})
SEED = sorted({k.split("/")[1] for k in G.files if k.startswith("seed/")})
BLIND = sorted({k.split("/")[1] for k in G.files if k.startswith("blind/")})


def _case(kind, name):
    pre = f"{kind}/{name}/"
    return {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}


@pytest.mark.parametrize("name", KMER)
def test_oracle_matches_reference_kmer(name):
    c = _case("kmer", name)
    k, h = (int(x) for x in c["kh"])
    r = ORACLE.kmer_batch(c["bases"], c["off"], k, h)
    for key in ("out", "valid", "fwd", "rev"):
        assert (r[key] == c[key]).all(), key


@pytest.mark.parametrize("name", SEED)
def test_oracle_matches_reference_seed(name):
    c = _case("seed", name)
    r = ORACLE.seed_batch(c["bases"], c["off"], [str(s) for s in c["seeds"]], int(c["h"][0]))
    for key in ("out", "valid", "fwd", "rev"):
        assert (r[key] == c[key]).all(), key


@pytest.mark.parametrize("name", BLIND)
def test_oracle_matches_reference_blind(name):
    c = _case("blind", name)
    h0, hv, fw, rv = ORACLE.blind_read(c["kmer"].tobytes(), c["hashes"].shape[1], c["feed"].tobytes())
    assert (h0 == c["h0"]).all() and (hv == c["hashes"]).all() and (fw == c["fwd"]).all() and (rv == c["rev"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", KMER)
def test_gpu_matches_reference_kmer(name):
    import torch

    import nthash_b200
    from gpu_util import to_dev, u64
    c = _case("kmer", name)
    k, h = (int(x) for x in c["kh"])
    d_b, _keep = to_dev(c["bases"])
    for strands in (False, True):   # without strands: the fast kernel; with: the general kernel
        res = nthash_b200.kmer_hashes(d_b, torch.from_numpy(c["off"].astype(np.int64)).cuda(), k, h, want_valid=True, want_strands=strands)
        torch.cuda.synchronize()
        assert (u64(res.out).reshape(c["out"].shape) == c["out"]).all()
        assert (res.valid_mask().cpu().numpy() == c["valid"].astype(bool)).all()
        if strands:
            assert (u64(res.fwd) == c["fwd"]).all() and (u64(res.rev) == c["rev"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", SEED)
def test_gpu_matches_reference_seed(name):
    import torch

    import nthash_b200
    from gpu_util import to_dev, u64
    c = _case("seed", name)
    plan = nthash_b200.SeedPlan([str(s) for s in c["seeds"]], int(c["h"][0]))
    d_b, _keep = to_dev(c["bases"])
    res = nthash_b200.seed_hashes(plan, d_b, torch.from_numpy(c["off"].astype(np.int64)).cuda(), want_valid=True, want_strands=True)
    torch.cuda.synchronize()
    assert (u64(res.out).reshape(c["out"].shape) == c["out"]).all()
    assert (res.valid_mask().cpu().numpy() == c["valid"].astype(bool)).all()
    assert (u64(res.fwd).reshape(c["fwd"].shape) == c["fwd"]).all() and (u64(res.rev).reshape(c["rev"].shape) == c["rev"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", BLIND)
def test_gpu_matches_reference_blind(name):
    import torch

    import nthash_b200
    from gpu_util import to_dev, u64
    c = _case("blind", name)
    kmer, feed, h = c["kmer"], c["feed"], c["hashes"].shape[1]
    k = len(kmer)
    d_k, _keep = to_dev(kmer)
    init = nthash_b200.kmer_hashes_uniform(d_k, 1, k, k, h, want_strands=True)
    assert (u64(init.out)[0] == c["h0"]).all()
    fwd, rev = init.fwd.clone(), init.rev.clone()
    window = list(kmer)
    for i, ch in enumerate(feed):   # BlindNtHash::roll(char): the caller supplies the incoming base (kmer.cpp:355-364)
        out_b = torch.tensor([window[0]], dtype=torch.uint8, device="cuda")
        in_b = torch.tensor([ch], dtype=torch.uint8, device="cuda")
        hv = nthash_b200.blind_roll(fwd, rev, out_b, in_b, k, h)
        window = window[1:] + [ch]
        assert (u64(hv)[0] == c["hashes"][i]).all(), i
        assert int(u64(fwd)[0]) == int(c["fwd"][i]) and int(u64(rev)[0]) == int(c["rev"][i])
