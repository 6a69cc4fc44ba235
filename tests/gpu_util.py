"""Helpers shared by the `-m gpu` parity tests: run the CUDA path through the C ABI and compare
with the CPU oracle (oracle/) on identical inputs.  The oracle is only ever the checker."""
import numpy as np
import torch

BAD_BYTES = np.frombuffer(b"NnRYKMSWBDHV-*.", np.uint8)


def synth(rng, n, p_bad=0.0, lower=0.0):
    a = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
    if lower:
        m = rng.random(n) < lower
        a[m] = a[m] | 0x20
        u = rng.random(n) < lower / 2
        a[u & (a == ord("T"))] = ord("U")
    if p_bad:
        bad = rng.random(n) < p_bad
        a[bad] = rng.choice(BAD_BYTES, int(bad.sum()))
    return a


def ragged_offsets(lens):
    return np.concatenate([[0], np.cumsum(np.asarray(lens, np.int64))]).astype(np.int64)


def to_dev(a, pad=32):
    """uint8 numpy -> CUDA tensor whose allocation is readable `pad` bytes past the data."""
    t = torch.zeros(len(a) + pad, dtype=torch.uint8, device="cuda")
    t[: len(a)] = torch.from_numpy(np.array(a, dtype=np.uint8, copy=True))
    return t[: len(a)], t


def u64(t):
    return t.detach().cpu().numpy().view(np.uint64)


def assert_batch_equal(res, ora, H, check_strands=False):
    """res: nthash_b200.HashBatch, ora: dict from oracle_lib (dense layout)."""
    out = u64(res.out).reshape(-1, H)
    assert out.shape == ora["out"].shape, (out.shape, ora["out"].shape)
    if res.valid_bits is not None:
        vm = res.valid_mask().cpu().numpy()
        assert (vm == ora["valid"].astype(bool)).all(), "validity bitmap differs from the reference's emitted positions"
    # rows the reference does not emit read back as zero; emitted rows are bit-exact
    if not (out == ora["out"]).all():
        bad = np.argwhere(out != ora["out"])[:5]
        raise AssertionError(f"hash mismatch at rows/cols {bad.tolist()}")
    if check_strands:
        assert (u64(res.fwd).reshape(ora["fwd"].shape) == ora["fwd"]).all()
        assert (u64(res.rev).reshape(ora["rev"].shape) == ora["rev"]).all()
