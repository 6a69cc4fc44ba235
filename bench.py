#!/usr/bin/env python
"""bench.py — k-mers hashed per second on N B200s (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3                 # this repo's CUDA engine
    torchrun ... bench.py --gpus N ...                             # N ranks, weak scaling (one shard per GPU)
    python bench.py --impl reference ...                           # the reference's own CPU roll() loop

A "step" is one pass of the hot path (NtHash roll over every read) over one batch of synthetic
reads: BASELINE.json configs[1], 10 M x 150 bp, k=31, h=1 per GPU.  `value` is device-resident
throughput (inputs in HBM, CUDA events on the launching stream, max over ranks); `e2e` is the same
metric through the host-buffer C ABI call (pinned host buffers, H2D + kernel + D2H inside the timed
region); `roofline` compares the kernel's algorithmic bytes/s with the measured HBM copy peak;
`cpu_baseline` is the compiled reference (oracle/_ref) timed on this box's host cores.
The oracle/ directory is executed here only for cpu_baseline / --impl reference and a checksum check.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # name: (reads per GPU, read length, k, hashes)
    "c2": dict(n_reads=10_000_000, read_len=150, k=31, h=1, desc="10M x 150bp reads, k=31, h=1 canonical (BASELINE.json configs[1])"),
    "c3": dict(n_reads=10_000_000, read_len=150, k=31, h=4, desc="10M x 150bp reads, k=31, h=4 (configs[2])"),
    "c4": dict(n_reads=10_000_000, read_len=150, k=31, h=3, seeds=["1010101010101010101010101010101", "1101101101101101011011011011011"],
               desc="SeedNtHash: 10M x 150bp reads, two spaced seeds, k=31, h=3 per seed (configs[3])"),
    "c5": dict(n_reads=12_500, read_len=50_000, k=63, h=1, desc="12.5k x 50kb reads per GPU, k=63, h=1 (configs[4] shard)"),
}
METRIC = "kmers_hashed_per_sec"
UNIT = "kmers/s"


def algorithmic_bytes(n_reads, read_len, k, H):
    """every base read once + every hash written once."""
    return n_reads * read_len + n_reads * max(read_len - k + 1, 0) * H * 8


def load_traffic(config):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(config)
    except Exception:
        return None


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region: NVML every 2 ms (nvidia_ml_py), or
    `nvidia-smi` every 100 ms when NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self._stop, self._t, self.source = index, [], [], set(), threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; NVML's does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nv = pynvml
            self.source = "nvml"
        except Exception:
            self._nv = None
            self.source = "nvidia-smi"

    def _poll_nvml(self):
        nv = self._nv
        bits = [(getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]
        try:
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)))
        except Exception:
            pass
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = get_reasons(self._h)
                self.reasons.update(name for bit, name in bits if r & bit)
            except Exception:
                pass
            self._stop.wait(0.002)

    def _poll_smi(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [x.strip() for x in out.split(",")]
                if len(r) >= 7:
                    self.sm.append(float(r[0]))
                    self.mx.append(float(r[1]))
                    self.reasons.update(self.NAMES[i] for i in range(4) if r[3 + i].lower().startswith("active"))
            except Exception:
                pass
            self._stop.wait(0.1)

    def sample_now(self):
        """One sample from the calling thread (work is queued on the GPU at this point)."""
        if self._nv:
            try:
                self.sm.append(float(self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM)))
            except Exception:
                pass

    def __enter__(self):
        self._t = threading.Thread(target=self._poll_nvml if self._nv else self._poll_smi, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def synth_reads_device(torch, n_bases, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    buf = torch.zeros(n_bases + 64, dtype=torch.uint8, device="cuda")  # readable slack past the last base
    step = 1 << 28
    for o in range(0, n_bases, step):
        m = min(step, n_bases - o)
        buf[o:o + m] = lut[torch.randint(0, 4, (m,), dtype=torch.uint8, device="cuda", generator=g).long()]
    return buf


def cpu_reference_pass(lib, bases_np, n_reads, read_len, k, h, threads, seeds=None):
    """One pass of the reference's own loop over `n_reads` reads -> (windows/s, emitted, sum)."""
    import numpy as np
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    t0 = time.perf_counter()
    if seeds:
        r = lib.seed_batch(bases_np, off, seeds, h, want=(), threads=threads)
    else:
        r = lib.kmer_batch(bases_np, off, k, h, want=(), threads=threads)
    dt = time.perf_counter() - t0
    return r["n_emit"] / dt, r["n_emit"], r["sum"], dt


def run_reference(args, cfg):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle_lib import ORACLE, REF
    lib = REF if REF is not None else ORACLE
    threads = os.cpu_count() or 1
    sample_reads = min(cfg["n_reads"], args.ref_sample_reads)
    bases = ORACLE.gen_bases(sample_reads * cfg["read_len"], 42)
    for _ in range(args.warmup):
        cpu_reference_pass(lib, bases, sample_reads, cfg["read_len"], cfg["k"], cfg["h"], threads, cfg.get("seeds"))
    t0 = time.perf_counter()
    emitted = 0
    for _ in range(args.steps):
        _, ne, _, _ = cpu_reference_pass(lib, bases, sample_reads, cfg["read_len"], cfg["k"], cfg["h"], threads, cfg.get("seeds"))
        emitted += ne
    dt = time.perf_counter() - t0
    value = emitted / dt
    sample = f"{sample_reads} reads x {cfg['read_len']} bp per step (a bounded sample of the {cfg['n_reads']}-read batch)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": cfg["desc"], "reads_per_step": sample_reads, "read_len": cfg["read_len"], "k": cfg["k"], "h": cfg["h"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": lib.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU (debugging only; invalidates the number)")
    ap.add_argument("--ref-sample-reads", type=int, default=2_000_000)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.reads:
        cfg["n_reads"] = args.reads
    if args.impl == "reference":
        return run_reference(args, cfg)

    import numpy as np
    import torch

    import nthash_b200
    from nthash_b200 import dist as nd
    from nthash_b200._lib import LIB, check

    rank, world, local = nd.env_rank()
    torch.cuda.set_device(local)
    try:  # pin this rank to the CPUs next to its GPU, so that the pinned host buffers land on that NUMA node (best effort)
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
    except Exception:
        pass
    nd.init("nccl", torch.device("cuda", local))
    barrier = nd.barrier

    n_reads, L, k, h = cfg["n_reads"], cfg["read_len"], cfg["k"], cfg["h"]
    seeds = cfg.get("seeds")
    H = h * (len(seeds) if seeds else 1)  # u64 values per window
    plan = nthash_b200.SeedPlan(seeds, h) if seeds else None
    nk = L - k + 1
    rows = n_reads * nk
    n_bases = n_reads * L
    bases_buf = synth_reads_device(torch, n_bases, 1234 + rank)
    bases = bases_buf[:n_bases]
    out = torch.empty((rows, H), dtype=torch.int64, device="cuda")
    valid = torch.empty(int(LIB.nthash_valid_words(rows)), dtype=torch.int32, device="cuda")

    def step(with_valid=True):
        if plan is not None:
            nthash_b200.seed_hashes_uniform(plan, bases, n_reads, L, want_valid=with_valid, out=out,
                                            valid_bits=valid if with_valid else None)
        else:
            nthash_b200.kmer_hashes_uniform(bases, n_reads, L, k, h, want_valid=with_valid, out=out,
                                            valid_bits=valid if with_valid else None)

    # ---- device-resident throughput (value) -------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:  # clocks are sampled (NVML, ~5 ms per sample) over both timed loops below
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        clk.sample_now()
        torch.cuda.synchronize()
        barrier()
        # ---- dominant kernel alone (roofline): same launch without the bitmap memset ----
        step(False)
        torch.cuda.synchronize()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        k0.record()
        marks[0].record()
        for i in range(args.steps):
            step(False)
            marks[i + 1].record()
        k1.record()
        clk.sample_now()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    kernel_ms = k0.elapsed_time(k1) / args.steps
    each = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps))  # per-launch times (best and median)
    ms, kernel_ms = nd.max_over_ranks([ms, kernel_ms])

    # ---- end to end through the host-buffer C ABI call ----------------------------------------
    e2e = None
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    e2e_reads = n_reads
    need = (n_bases + rows * H * 8) * world
    while e2e_reads > 100_000 and avail and need * e2e_reads / n_reads > 0.5 * avail:
        e2e_reads //= 2
    e_rows = e2e_reads * nk
    h_bases = torch.empty(e2e_reads * L, dtype=torch.uint8).pin_memory()
    h_bases.copy_(bases[: e2e_reads * L])
    h_off = (torch.arange(e2e_reads + 1, dtype=torch.int64) * L)
    h_out = torch.empty((e_rows, H), dtype=torch.int64).pin_memory()
    h_valid = torch.empty(int(LIB.nthash_valid_words(e_rows)), dtype=torch.int32).pin_memory()

    import ctypes
    seed_arr = (ctypes.c_char_p * len(seeds))(*[s.encode() for s in seeds]) if seeds else None

    def e2e_step():
        if seeds:
            check(LIB.nthash_seed_batch(h_bases.data_ptr(), h_off.data_ptr(), e2e_reads, seed_arr, len(seeds), k, h,
                                        h_out.data_ptr(), h_valid.data_ptr(), None, None, local))
        else:
            check(LIB.nthash_kmer_batch_uniform(h_bases.data_ptr(), e2e_reads, L, k, h, h_out.data_ptr(), h_valid.data_ptr(), None, None, local))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e_dt = (time.perf_counter() - t0) / args.e2e_steps
    e_dt = nd.max_over_ranks([e_dt])[0]
    # parity of the e2e result with the device-resident one (same reads): checksum of checksums
    same = bool((h_out[: 1000 * nk].cuda() == out[: 1000 * nk]).all())
    e2e = {"value": world * e_rows / e_dt, "unit": UNIT, "h2d_bytes_per_step": int(h_bases.numel()),
           "d2h_bytes_per_step": int(h_out.numel() * 8 + h_valid.numel() * 4), "reads_per_step": e2e_reads,
           "ms_per_step": e_dt * 1e3, "matches_device_path": same}

    # ---- fused consumer (count/sum/xor on the device: what the reference's own benchmark loop computes) ----
    consumer = None
    if not seeds:
        red = nthash_b200.kmer_reduce_uniform(bases, n_reads, L, k, h)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(args.steps):
            red = nthash_b200.kmer_reduce_uniform(bases, n_reads, L, k, h)
        r1.record()
        torch.cuda.synchronize()
        red_ms = r0.elapsed_time(r1) / args.steps
        h_res = torch.zeros(3, dtype=torch.int64)

        def e2e_reduce_step():
            check(LIB.nthash_kmer_reduce(h_bases.data_ptr(), h_off.data_ptr(), e2e_reads, k, h, h_res.data_ptr(), local))

        e2e_reduce_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_reduce_step()
        er_dt = (time.perf_counter() - t0) / args.e2e_steps
        # the same consumer fed with 2-bit packed bases from the host (nthash_kmer_reduce_packed2bit): a quarter of the H2D bytes
        lut2 = torch.zeros(256, dtype=torch.uint8, device="cuda")
        lut2[torch.tensor(list(b"ACGT"), device="cuda").long()] = torch.arange(4, dtype=torch.uint8, device="cuda")
        codes = lut2[bases[: e2e_reads * L].long()]
        if codes.numel() % 4:
            codes = torch.cat([codes, torch.zeros(4 - codes.numel() % 4, dtype=torch.uint8, device="cuda")])
        c4 = codes.view(-1, 4)
        h_packed = (c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).cpu().pin_memory()
        del codes, c4, lut2
        h_res2 = torch.zeros(3, dtype=torch.int64)

        def e2e_packed_step():
            check(LIB.nthash_kmer_reduce_packed2bit(h_packed.data_ptr(), None, None, e2e_reads, L, k, h, h_res2.data_ptr(), local))

        e2e_packed_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_packed_step()
        ep_dt = (time.perf_counter() - t0) / args.e2e_steps
        red_ms, er_dt, ep_dt = nd.max_over_ranks([red_ms, er_dt, ep_dt])
        consumer = {"kind": "count/sum/xor of all hashes (nthash_kmer_reduce*), no hash leaves the GPU",
                    "value": world * rows / (red_ms * 1e-3), "unit": UNIT, "ms_per_step": red_ms,
                    "e2e": {"value": world * e_rows / er_dt, "unit": UNIT, "ms_per_step": er_dt * 1e3,
                            "h2d_bytes_per_step": int(h_bases.numel()), "d2h_bytes_per_step": 24},
                    "e2e_packed2bit": {"value": world * e_rows / ep_dt, "unit": UNIT, "ms_per_step": ep_dt * 1e3,
                                       "h2d_bytes_per_step": int(h_packed.numel()), "d2h_bytes_per_step": 24,
                                       "matches_ascii_path": bool((h_res2 == h_res).all())},
                    "windows": int(red[0]), "sum": int(red[1]) & (2**64 - 1)}

        # second fused consumer: Bloom filter insert / query (the caller nthash.hpp:14-17 names), 3 hashes per
        # k-mer into a 1 GiB (2^33 bit) device-resident filter; bound by random 32-byte atomics, not by streaming
        try:
            bits = 1 << 33
            filt = nthash_b200.bloom_filter(bits)
            b0, b1, b2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            nthash_b200.kmer_bloom_uniform(bases, min(n_reads, 100_000), L, k, 3, filt, bits)  # warm-up
            filt.zero_()
            torch.cuda.synchronize()
            b0.record()
            ins = nthash_b200.kmer_bloom_uniform(bases, n_reads, L, k, 3, filt, bits)
            b1.record()
            qry = nthash_b200.kmer_bloom_uniform(bases, n_reads, L, k, 3, filt, bits, query=True)
            b2.record()
            torch.cuda.synchronize()
            ins_ms, qry_ms = nd.max_over_ranks([b0.elapsed_time(b1), b1.elapsed_time(b2)])
            consumer["bloom"] = {"filter_bits": bits, "hashes_per_kmer": 3,
                                 "insert": {"value": world * rows / (ins_ms * 1e-3), "unit": UNIT, "ms_per_step": ins_ms},
                                 "query": {"value": world * rows / (qry_ms * 1e-3), "unit": UNIT, "ms_per_step": qry_ms},
                                 "windows": int(ins[0]), "query_hits": int(qry[1])}
            del filt
        except Exception as e:  # the filter did not fit next to the batch
            consumer["bloom"] = {"skipped": str(e)[:80]}

    if seeds:
        # SeedNtHash consumer: count / sum / xor of every visited window's hashes, two passes on the device (rows to scratch
        # memory, then a reduction): no hash crosses PCIe
        red = nthash_b200.seed_reduce_uniform(plan, bases, n_reads, L)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(args.steps):
            red = nthash_b200.seed_reduce_uniform(plan, bases, n_reads, L)
        r1.record()
        torch.cuda.synchronize()
        red_ms = r0.elapsed_time(r1) / args.steps
        h_res = torch.zeros(3, dtype=torch.int64)

        def e2e_seed_reduce_step():
            check(LIB.nthash_seed_reduce(h_bases.data_ptr(), h_off.data_ptr(), e2e_reads, seed_arr, len(seeds), k, h, h_res.data_ptr(), local))

        e2e_seed_reduce_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_seed_reduce_step()
        er_dt = (time.perf_counter() - t0) / args.e2e_steps
        red_ms, er_dt = nd.max_over_ranks([red_ms, er_dt])
        consumer = {"kind": "count/sum/xor of all SeedNtHash hashes (nthash_seed_reduce*: rows to device scratch, then a reduction; not fused)",
                    "value": world * rows / (red_ms * 1e-3), "unit": UNIT, "ms_per_step": red_ms,
                    "e2e": {"value": world * e_rows / er_dt, "unit": UNIT, "ms_per_step": er_dt * 1e3,
                            "h2d_bytes_per_step": int(h_bases.numel()), "d2h_bytes_per_step": 24},
                    "windows": int(red[0]), "sum": int(red[1]) & (2**64 - 1),
                    "e2e_matches_device": bool(e2e_reads != n_reads or (h_res.cuda() == red).all())}

    if rank != 0:
        nd.finalize()
        return

    # ---- CPU baseline: the compiled reference on this box's host cores -------------------------
    cpu = None
    checksum_ok = None
    if not args.no_cpu_baseline:
        from oracle_lib import ORACLE, REF
        lib = REF if REF is not None else ORACLE
        threads = os.cpu_count() or 1
        sample_reads = min(n_reads, args.ref_sample_reads)
        h_sample = bases[: sample_reads * L].cpu().numpy()
        v, ne, s, dt = cpu_reference_pass(lib, h_sample, sample_reads, L, k, h, threads, seeds)
        # grow the sample until it is a meaningful amount of CPU work (about 10-30 core-seconds)
        if dt * threads < 10 and sample_reads < n_reads:
            sample_reads = min(n_reads, int(sample_reads * 20 / max(dt * threads, 0.5)))
            h_sample = bases[: sample_reads * L].cpu().numpy()
            v, ne, s, dt = cpu_reference_pass(lib, h_sample, sample_reads, L, k, h, threads, seeds)
        got = int(out[: sample_reads * nk].cpu().numpy().view(np.uint64).sum(dtype=np.uint64))
        checksum_ok = (got == s) and ne == sample_reads * nk
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": lib.kind,
               "sample": f"first {sample_reads} of {n_reads} reads, {dt:.2f} s wall on {threads} threads", "checksum_matches_gpu": checksum_ok}

    peak, peak_src = load_peak()
    abytes = algorithmic_bytes(n_reads, L, k, H)
    achieved = abytes / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": world * rows / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": cfg["desc"], "reads_per_gpu": n_reads, "read_len": L, "k": k, "h": h, "seeds": seeds,
                   "l2_policy": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2; no explicit flush" % (abytes / 1e9),
                   "sharding": "independent read shards per GPU, no collective"},
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (load_traffic(args.config) or {}).get("dram_bytes_per_launch") if not args.reads else None,
                     "traffic_source": (load_traffic(args.config) or {}).get("source"), "peak_source": peak_src, "kernel": ("seed_jit_kernel (NVRTC-specialised)" if seeds else "kmer_fast_kernel<H=%d>" % h), "kernel_ms": kernel_ms,
                     "kernel_ms_best": each[0], "kernel_ms_median": statistics.median(each), "algorithmic_bytes_per_launch": abytes},
        "cpu_baseline": cpu, "fused_consumer": consumer,
    }
    if consumer and cpu and sample_reads == n_reads:
        consumer["matches_cpu_reference"] = consumer["sum"] == s and consumer["windows"] == ne
    print(json.dumps(line))
    nd.finalize()


if __name__ == "__main__":
    main()
