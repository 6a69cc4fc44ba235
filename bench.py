#!/usr/bin/env python
"""bench.py — k-mers hashed per second on N B200s (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3                 # this repo's CUDA engine
    torchrun ... bench.py --gpus N ...                             # N ranks, weak scaling (one shard per GPU)
    python bench.py --impl reference ...                           # the reference's own CPU roll() loop

A "step" is one pass of the hot path (NtHash / SeedNtHash roll over every read) over one batch of synthetic reads.
The headline line is BASELINE.json configs[1] (C2: 10 M x 150 bp, k=31, h=1 per GPU); `value` is device-resident
throughput (inputs in HBM, CUDA events on the launching stream, max over ranks); `e2e` is the same metric through the
host-buffer C ABI call (pinned host buffers, H2D + kernel + D2H inside the timed region); `roofline` compares the
kernel's algorithmic bytes/s with the measured HBM copy peak; `cpu_baseline` is the compiled reference (oracle/_ref)
timed on this box's host cores.  The other BASELINE configs ride along as compact sub-records under `configs`:
c3 (h=4), c4 (SeedNtHash, two seeds x 3) and c5 (50 kb reads, k=63: 100 k reads over 8 GPUs = 12.5 k per rank, so a
`--gpus 8` run IS configs[4] as stated), each with its device-resident time, roofline fraction and a checksum against
the compiled reference on a stated prefix of the same reads.

Reads are the generator SURVEY.md §8(d) pins, in BOTH arms: one splitmix64 stream (seed 42; 44 for c5), 32 bases per
draw; rank r hashes reads [r*n, (r+1)*n) of that stream, the reference arm a prefix of it.
The oracle/ directory is executed here only for cpu_baseline / --impl reference and the checksum checks.
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

SEED_A = "1010101010101010101010101010101"
SEED_B = "1101101101101101011011011011011"
CONFIGS = {
    # reads per GPU, read length, k, hashes (per seed), splitmix seed, reads of the CPU sample
    "c2": dict(n_reads=10_000_000, read_len=150, k=31, h=1, seed=42, cpu_reads=2_000_000,
               desc="10M x 150bp reads, k=31, h=1 canonical (BASELINE.json configs[1])"),
    "c3": dict(n_reads=10_000_000, read_len=150, k=31, h=4, seed=42, cpu_reads=1_000_000,
               desc="10M x 150bp reads, k=31, h=4 (configs[2])"),
    "c4": dict(n_reads=10_000_000, read_len=150, k=31, h=3, seed=42, cpu_reads=300_000, seeds=[SEED_A, SEED_B],
               desc="SeedNtHash: 10M x 150bp reads, two spaced seeds, k=31, h=3 per seed (configs[3])"),
    "c5": dict(n_reads=12_500, read_len=50_000, k=63, h=1, seed=44, cpu_reads=4_000,
               desc="50kb reads, k=63, h=1: 100k reads sharded over 8 GPUs = 12.5k reads per GPU (configs[4])"),
}
METRIC = "kmers_hashed_per_sec"
UNIT = "kmers/s"
M64 = (1 << 64) - 1


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


# ---- the synthetic reads of SURVEY.md §8(d) / Appendix C: one splitmix64 stream, 32 bases per draw -------------------
def _s64(v):
    v &= M64
    return v - (1 << 64) if v >= (1 << 63) else v


def splitmix_bases_torch(torch, n_bases, seed, first_base=0, device="cuda", slack=64):
    """Bases [first_base, first_base + n_bases) of the stream, as a uint8 tensor with `slack` readable bytes after
    them.  int64 arithmetic wraps like uint64; logical shifts are emulated with a mask."""
    assert first_base % 32 == 0
    buf = torch.zeros(n_bases + slack, dtype=torch.uint8, device=device)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    shifts = torch.arange(0, 64, 2, dtype=torch.int64, device=device)
    d0, n_draws = first_base // 32, (n_bases + 31) // 32
    step = 1 << 22
    for o in range(0, n_draws, step):
        m = min(step, n_draws - o)
        i = torch.arange(d0 + o + 1, d0 + o + 1 + m, dtype=torch.int64, device=device)
        z = i * _s64(0x9E3779B97F4A7C15) + _s64(seed)
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _s64(0xBF58476D1CE4E5B9)
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _s64(0x94D049BB133111EB)
        z = z ^ ((z >> 31) & ((1 << 33) - 1))
        b = lut[((z.unsqueeze(1) >> shifts) & 3)].reshape(-1)
        lo, hi = o * 32, min(n_bases, (o + m) * 32)
        buf[lo:hi] = b[: hi - lo]
    return buf


def splitmix_bases_numpy(n_bases, seed, first_base=0):
    import numpy as np
    assert first_base % 32 == 0
    d0, n_draws = first_base // 32, (n_bases + 31) // 32
    out = np.empty(n_draws * 32, np.uint8)
    lut = np.frombuffer(b"ACGT", np.uint8)
    shifts = np.arange(0, 64, 2, dtype=np.uint64)
    step = 1 << 20
    with np.errstate(over="ignore"):
        for o in range(0, n_draws, step):
            m = min(step, n_draws - o)
            z = np.arange(d0 + o + 1, d0 + o + 1 + m, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[o * 32:(o + m) * 32] = lut[((z[:, None] >> shifts) & np.uint64(3)).astype(np.intp)].reshape(-1)
    return out[:n_bases]


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


# ---- the reference's own loop on the host cores (oracle/_ref = the unmodified reference compiled here) -----------------
def cpu_reference_pass(lib, bases_np, n_reads, read_len, k, h, threads, seeds=None):
    """One pass of the reference's own loop over `n_reads` reads -> (windows/s, emitted, sum, seconds)."""
    import numpy as np
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    t0 = time.perf_counter()
    if seeds:
        r = lib.seed_batch(bases_np, off, seeds, h, want=(), threads=threads)
    else:
        r = lib.kmer_batch(bases_np, off, k, h, want=(), threads=threads)
    dt = time.perf_counter() - t0
    return r["n_emit"] / dt, r["n_emit"], r["sum"], dt


def cpu_reference_measure(cfg, sample_reads, passes, warm=1):
    """The policy BOTH CPU numbers (`--impl reference` and the in-arm `cpu_baseline`) follow: the first `sample_reads`
    reads of the config's splitmix stream, all host threads, `warm` untimed passes, then `passes` timed ones."""
    from oracle_lib import ORACLE, REF
    lib = REF if REF is not None else ORACLE
    threads = os.cpu_count() or 1
    L, k, h, seeds = cfg["read_len"], cfg["k"], cfg["h"], cfg.get("seeds")
    bases = splitmix_bases_numpy(sample_reads * L, cfg["seed"])
    for _ in range(warm):
        cpu_reference_pass(lib, bases, sample_reads, L, k, h, threads, seeds)
    emitted, ssum, each = 0, 0, []
    t0 = time.perf_counter()
    for _ in range(passes):
        _, ne, s, dt1 = cpu_reference_pass(lib, bases, sample_reads, L, k, h, threads, seeds)
        emitted += ne
        ssum = s
        each.append(dt1)
    dt = time.perf_counter() - t0
    return dict(value=emitted / dt, emitted_per_pass=emitted // max(passes, 1), sum=ssum, seconds=dt, threads=threads, kind=lib.kind,
                sample_reads=sample_reads, passes=passes, best_pass_s=min(each) if each else None)


def cpu_sample_label(cfg, m):
    H = cfg["h"] * (len(cfg["seeds"]) if cfg.get("seeds") else 1)
    return (f"first {m['sample_reads']} reads x {cfg['read_len']} bp of the splitmix64(seed={cfg['seed']}) stream "
            f"(a bounded sample of the {cfg['n_reads']}-read batch), {m['passes']} timed passes after 1 warm-up, "
            f"{m['seconds']:.2f} s wall on {m['threads']} threads, {H} hash value(s) per window")


def run_reference(args, cfg):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_reads = min(cfg["n_reads"], args.ref_sample_reads or cfg["cpu_reads"])
    m = cpu_reference_measure(cfg, sample_reads, args.steps, warm=max(1, min(args.warmup, 2)))
    value = m["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["seconds"] / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(cfg, cfg["n_reads"]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": m["threads"], "kind": m["kind"], "sample": cpu_sample_label(cfg, m)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(cfg, n_reads):
    """The `config` object both arms print (identical: the driver compares them)."""
    abytes = algorithmic_bytes(n_reads, cfg["read_len"], cfg["k"], cfg["h"] * (len(cfg["seeds"]) if cfg.get("seeds") else 1))
    return {"workload": cfg["desc"], "reads_per_gpu": n_reads, "read_len": cfg["read_len"], "k": cfg["k"], "h": cfg["h"],
            "seeds": cfg.get("seeds"), "generator": "splitmix64 seed %d, 32 bases per draw (SURVEY.md 8d)" % cfg["seed"],
            "l2_policy": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2; no explicit flush" % (abytes / 1e9),
            "sharding": "independent read shards per GPU, no collective"}


class Workload:
    """One BASELINE config resident on this rank's GPU: bases, output rows, validity bitmap, and its step()."""

    def __init__(self, torch, nthash_b200, LIB, cfg, rank):
        self.torch, self.nb, self.cfg = torch, nthash_b200, cfg
        self.n, self.L, self.k, self.h = cfg["n_reads"], cfg["read_len"], cfg["k"], cfg["h"]
        self.seeds = cfg.get("seeds")
        self.H = self.h * (len(self.seeds) if self.seeds else 1)
        self.plan = nthash_b200.SeedPlan(self.seeds, self.h) if self.seeds else None
        self.nk = self.L - self.k + 1
        self.rows = self.n * self.nk
        self.n_bases = self.n * self.L
        self.buf = splitmix_bases_torch(torch, self.n_bases, cfg["seed"], first_base=rank * self.n_bases)
        self.bases = self.buf[: self.n_bases]
        self.out = torch.empty((self.rows, self.H), dtype=torch.int64, device="cuda")
        self.valid = torch.empty(int(LIB.nthash_valid_words(self.rows)), dtype=torch.int32, device="cuda")

    def step(self, with_valid=True):
        if self.plan is not None:
            self.nb.seed_hashes_uniform(self.plan, self.bases, self.n, self.L, want_valid=with_valid, out=self.out,
                                        valid_bits=self.valid if with_valid else None)
        else:
            self.nb.kmer_hashes_uniform(self.bases, self.n, self.L, self.k, self.h, want_valid=with_valid, out=self.out,
                                        valid_bits=self.valid if with_valid else None)

    def time_steps(self, steps, with_valid, warm):
        """CUDA events on the launching stream around `steps` back-to-back calls -> (mean ms, sorted per-launch ms)."""
        torch = self.torch
        for _ in range(warm):
            self.step(with_valid)
        torch.cuda.synchronize()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        marks[0].record()
        for i in range(steps):
            self.step(with_valid)
            marks[i + 1].record()
        torch.cuda.synchronize()
        return marks[0].elapsed_time(marks[steps]) / steps, sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(steps))

    def checksum_prefix(self, reads):
        """(windows, 64-bit sum of every hash value) of the first `reads` reads' rows, from the device-resident output."""
        v = self.out[: reads * self.nk]
        return reads * self.nk, int(v.sum()) & M64

    def kernel_name(self):
        return "seed_jit_kernel (NVRTC-specialised)" if self.seeds else "kmer_fast_kernel<H=%d>" % self.h


def sub_record(torch, nthash_b200, LIB, nd, name, args, rank, world, peak, want_cpu):
    """Compact record of one of the other BASELINE configs: device-resident time, roofline fraction, checksum vs the
    compiled reference on a stated prefix (rank 0).  Everything is freed before the next config is built."""
    cfg = dict(CONFIGS[name])
    w = Workload(torch, nthash_b200, LIB, cfg, rank)
    steps = max(3, min(args.steps, args.sub_steps))
    with ClockSampler(torch.cuda.current_device()) as clk:  # these launches run for 50-100 ms back to back: note what the clocks did
        ms_valid, _ = w.time_steps(steps, True, 2)
        nd.barrier()
        ms, each = w.time_steps(steps, False, 1)
    ms_valid, ms = nd.max_over_ranks([ms_valid, ms])
    abytes = algorithmic_bytes(w.n, w.L, w.k, w.H)
    achieved = abytes / (ms * 1e-3) / 1e9
    rec = {"workload": cfg["desc"], "reads_per_gpu": w.n, "n_gpus": world, "value": world * w.rows / (ms_valid * 1e-3), "unit": UNIT,
           "ms_per_step": ms_valid, "steps": steps, "clocks": clk.summary(),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "kernel": w.kernel_name(), "kernel_ms": ms, "kernel_ms_best": each[0], "kernel_ms_median": statistics.median(each),
                        "algorithmic_bytes_per_launch": abytes,
                        "traffic": (load_traffic(name) or {}).get("dram_bytes_per_launch")}}
    if name == "c5":
        # end to end through the host-buffer C ABI for the long-read shard too (this is configs[4]'s per-GPU share)
        h_bases = torch.empty(w.n_bases, dtype=torch.uint8).pin_memory()
        h_bases.copy_(w.bases)
        h_out = torch.empty((w.rows, w.H), dtype=torch.int64).pin_memory()
        h_valid = torch.empty(int(LIB.nthash_valid_words(w.rows)), dtype=torch.int32).pin_memory()
        from nthash_b200._lib import check
        local = torch.cuda.current_device()

        def e2e_step():
            check(LIB.nthash_kmer_batch_uniform(h_bases.data_ptr(), w.n, w.L, w.k, w.h, h_out.data_ptr(), h_valid.data_ptr(), None, None, local))

        e2e_step()
        nd.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_step()
        e_dt = nd.max_over_ranks([(time.perf_counter() - t0) / 3])[0]
        rec["e2e"] = {"value": world * w.rows / e_dt, "unit": UNIT, "ms_per_step": e_dt * 1e3, "h2d_bytes_per_step": int(h_bases.numel()),
                      "d2h_bytes_per_step": int(h_out.numel() * 8 + h_valid.numel() * 4),
                      "matches_device_path": bool((h_out[: 100 * w.nk].cuda() == w.out[: 100 * w.nk]).all())}
        del h_bases, h_out, h_valid
    # whole-job checksum of checksums (every rank's full output) — the size-independent property the multi-GPU run reports
    rec["sum_all_ranks"] = nd.sum_over_ranks(int(w.out.sum()) & M64)
    if rank == 0 and want_cpu:
        m = cpu_reference_measure(cfg, min(cfg["cpu_reads"], w.n), 1, warm=0)
        windows, gsum = w.checksum_prefix(m["sample_reads"])
        rec["cpu_reference"] = {"value": m["value"], "unit": UNIT, "cores": m["threads"], "kind": m["kind"], "sample": cpu_sample_label(cfg, m)}
        rec["checksum_matches_reference"] = bool(gsum == m["sum"] and windows == m["emitted_per_pass"])
    del w
    torch.cuda.empty_cache()
    return rec


RAGGED_CASES = {
    # trimmed-read shaped batches (no BASELINE config of their own: what short-read data looks like after adapter / quality
    # trimming, and what the ragged entry points of the C ABI are for): lengths uniform in [lo, hi]
    "kmer_100_150": dict(n_reads=10_000_000, lo=100, hi=150, k=31, h=1, seeds=None, seed=61),
    "kmer_36_150": dict(n_reads=10_000_000, lo=36, hi=150, k=31, h=1, seeds=None, seed=62),
    "seed_100_150": dict(n_reads=5_000_000, lo=100, hi=150, k=31, h=3, seeds=[SEED_A, SEED_B], seed=63),
}


def ragged_records(torch, nthash_b200, nd, args, rank, world, peak, want_cpu):
    """Ragged batches through the planned entry points (nthash_ragged_plan_create once, then nthash_kmer_batch_planned_dev /
    nthash_seed_batch_planned_dev: kernels only): device-resident time, roofline fraction on the batch's own algorithmic
    bytes (bases + rows x H x 8), and a checksum of the first reads' rows against the compiled reference (rank 0)."""
    import numpy as np
    out_recs = {}
    steps = max(3, min(args.steps, args.sub_steps))
    for name, c in RAGGED_CASES.items():
        time.sleep(args.sub_pause)
        try:
            n, k, h, seeds = c["n_reads"], c["k"], c["h"], c["seeds"]
            H = h * (len(seeds) if seeds else 1)
            g = torch.Generator(device="cuda")
            g.manual_seed(c["seed"] * 1000 + rank)
            lens = torch.randint(c["lo"], c["hi"] + 1, (n,), device="cuda", generator=g, dtype=torch.int64)
            off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
            off[1:] = torch.cumsum(lens, 0)
            nb = int(off[-1])
            bases = splitmix_bases_torch(torch, (nb + 31) // 32 * 32, c["seed"], first_base=rank * ((nb + 31) // 32 * 32))[:nb]
            plan = nthash_b200.RaggedPlan(off, k)
            rows = plan.rows
            out = torch.empty((rows, H), dtype=torch.int64, device="cuda")
            splan = nthash_b200.SeedPlan(seeds, h) if seeds else None

            def step():
                if splan is not None:
                    nthash_b200.seed_hashes_planned(splan, plan, bases, want_valid=False, out=out)
                else:
                    nthash_b200.kmer_hashes_planned(plan, bases, h, want_valid=False, out=out)

            for _ in range(2):
                step()
            torch.cuda.synchronize()
            nd.barrier()
            with ClockSampler(torch.cuda.current_device()) as clk:
                marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
                marks[0].record()
                for i in range(steps):
                    step()
                    marks[i + 1].record()
                torch.cuda.synchronize()
            each = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(steps))
            ms = nd.max_over_ranks([marks[0].elapsed_time(marks[steps]) / steps])[0]
            abytes = nb + rows * H * 8
            rec = {"workload": f"{n} reads of {c['lo']}-{c['hi']} bp (uniform), k={k}, " + (f"{len(seeds)} spaced seeds x {h}" if seeds else f"h={h}")
                               + ", planned ragged entry points", "reads_per_gpu": n, "rows_per_gpu": rows, "n_gpus": world,
                   "value": world * rows / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "clocks": clk.summary(),
                   "roofline": {"bound": "hbm", "achieved": abytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": abytes / (ms * 1e-3) / 1e9 / peak, "kernel_ms": ms, "kernel_ms_best": each[0],
                                "kernel_ms_median": statistics.median(each), "algorithmic_bytes_per_launch": abytes,
                                "kernel": "seed_jit_kernel (ragged, direct form)" if seeds else "kmer_fast_kernel<H=%d> (ragged, direct stores)" % h}}
            if rank == 0 and want_cpu:
                from oracle_lib import ORACLE, REF
                lib = REF if REF is not None else ORACLE
                m_reads = min(n, 200_000)
                off_np = off[: m_reads + 1].cpu().numpy().astype(np.uint64)
                bases_np = bases[: int(off_np[-1])].cpu().numpy()
                koff = plan.koff()
                pre_rows = int(koff[m_reads])
                t0 = time.perf_counter()
                r = (lib.seed_batch(bases_np, off_np, seeds, h, want=(), threads=os.cpu_count() or 1) if seeds
                     else lib.kmer_batch(bases_np, off_np, k, h, want=(), threads=os.cpu_count() or 1))
                dt = time.perf_counter() - t0
                gsum = int(out[:pre_rows].sum()) & M64
                rec["cpu_reference"] = {"value": r["n_emit"] / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": lib.kind,
                                        "sample": f"first {m_reads} reads of the batch, one pass, {dt:.2f} s"}
                rec["checksum_matches_reference"] = bool(gsum == r["sum"] and pre_rows == r["n_emit"])
                del koff
            out_recs[name] = rec
            del plan, splan, out, bases, off, lens
        except Exception as e:  # never take the headline line down
            out_recs[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    return out_recs



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU (debugging only; invalidates the number)")
    ap.add_argument("--ref-sample-reads", type=int, default=0, help="reads of the CPU sample (default: the config's cpu_reads)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--sub-steps", type=int, default=5, help="timed launches per sub-record config")
    ap.add_argument("--sub-pause", type=float, default=1.5, help="idle seconds in front of every sub-record (lets the power cap of the previous one clear)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true", help="only the headline config (skip configs.c3/c4/c5)")
    ap.add_argument("--no-consumers", action="store_true")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.reads:
        cfg["n_reads"] = args.reads
        cfg["cpu_reads"] = min(cfg["cpu_reads"], args.reads)
    if args.impl == "reference":
        return run_reference(args, cfg)

    import numpy as np  # noqa: F401
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
    peak, peak_src = load_peak()

    w = Workload(torch, nthash_b200, LIB, cfg, rank)
    n_reads, L, k, h, seeds, H, nk, rows, n_bases, plan = w.n, w.L, w.k, w.h, w.seeds, w.H, w.nk, w.rows, w.n_bases, w.plan
    bases, out = w.bases, w.out
    step = w.step

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
    step()  # leave the full result (with its bitmap) in `out` for the checks below
    torch.cuda.synchronize()
    sum_all = nd.sum_over_ranks(int(out.sum()) & M64)

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
    same = bool((h_out[: 1000 * nk].cuda() == out[: 1000 * nk]).all()) and (e2e_reads != n_reads or (int(h_out.sum()) & M64) == (int(out.sum()) & M64))
    e2e = {"value": world * e_rows / e_dt, "unit": UNIT, "h2d_bytes_per_step": int(h_bases.numel()),
           "d2h_bytes_per_step": int(h_out.numel() * 8 + h_valid.numel() * 4), "reads_per_step": e2e_reads,
           "ms_per_step": e_dt * 1e3, "steps": args.e2e_steps, "matches_device_path": same}

    # ---- fused consumer (count/sum/xor on the device: what the reference's own benchmark loop computes) ----
    consumer = None
    if not seeds and not args.no_consumers:
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
                    "windows": int(red[0]), "sum": int(red[1]) & M64,
                    "matches_stored_hashes": bool((int(red[1]) & M64) == (int(out.sum()) & M64))}
        del h_packed

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
        consumer.update(extra_consumers(torch, nthash_b200, nd, w, args, world))

    if seeds and not args.no_consumers:
        # SeedNtHash consumer: count / sum / xor of every visited window's hashes; no hash crosses PCIe
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
        consumer = {"kind": "count/sum/xor of all SeedNtHash hashes (nthash_seed_reduce*)",
                    "value": world * rows / (red_ms * 1e-3), "unit": UNIT, "ms_per_step": red_ms,
                    "e2e": {"value": world * e_rows / er_dt, "unit": UNIT, "ms_per_step": er_dt * 1e3,
                            "h2d_bytes_per_step": int(h_bases.numel()), "d2h_bytes_per_step": 24},
                    "windows": int(red[0]), "sum": int(red[1]) & M64,
                    "e2e_matches_device": bool(e2e_reads != n_reads or (h_res.cuda() == red).all())}

    # ---- CPU baseline: the compiled reference on this box's host cores (rank 0; same policy as --impl reference) ----
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        m = cpu_reference_measure(cfg, min(n_reads, args.ref_sample_reads or cfg["cpu_reads"]), 3)
        windows, gsum = w.checksum_prefix(m["sample_reads"])
        cpu = {"value": m["value"], "unit": UNIT, "cores": m["threads"], "kind": m["kind"], "sample": cpu_sample_label(cfg, m),
               "checksum_matches_gpu": bool(gsum == m["sum"] and windows == m["emitted_per_pass"])}
        try:  # the one-core number next to the all-core one (one pass over the first 200 k reads)
            from oracle_lib import ORACLE, REF
            one_reads = min(n_reads, 200_000)
            v1, _, _, dt1 = cpu_reference_pass(REF if REF is not None else ORACLE, splitmix_bases_numpy(one_reads * L, cfg["seed"]), one_reads, L, k, h, 1,
                                               cfg.get("seeds"))
            cpu["one_core"] = {"value": v1, "unit": UNIT, "sample": f"first {one_reads} reads, one thread, one pass, {dt1:.2f} s"}
        except Exception as e:
            cpu["one_core"] = {"error": str(e)[:200]}

    abytes = algorithmic_bytes(n_reads, L, k, H)
    achieved = abytes / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": world * rows / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": workload_config(cfg, n_reads),
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (load_traffic(args.config) or {}).get("dram_bytes_per_launch") if not args.reads else None,
                     "traffic_source": (load_traffic(args.config) or {}).get("source"), "peak_source": peak_src, "kernel": w.kernel_name(), "kernel_ms": kernel_ms,
                     "kernel_ms_best": each[0], "kernel_ms_median": statistics.median(each), "algorithmic_bytes_per_launch": abytes},
        "cpu_baseline": cpu, "fused_consumer": consumer, "sum_all_ranks": sum_all,
    }
    # ---- the other BASELINE configs as sub-records (every rank takes part; rank 0 checks against the reference) ----
    del h_bases, h_out, h_valid, w, bases, out, step
    torch.cuda.empty_cache()
    if not args.no_sub_records and not args.reads:
        subs = {}
        # c4 last among the BASELINE configs and a short pause in front of every sub-record: C4's launches run the GPU into its
        # power cap (sw_power_cap, SM clock down to ~1750 MHz for a few hundred ms afterwards), which otherwise lands on
        # whichever config is measured next (C5 read 0.73 of the peak right after C4 and 0.90 on its own, same kernel)
        for name in ("c3", "c5", "c4"):
            if name == args.config:
                continue
            time.sleep(args.sub_pause)
            try:
                subs[name] = sub_record(torch, nthash_b200, LIB, nd, name, args, rank, world, peak, not args.no_cpu_baseline)
            except Exception as e:  # a sub-record must never take the headline line down with it
                subs[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                torch.cuda.empty_cache()
        subs["ragged"] = ragged_records(torch, nthash_b200, nd, args, rank, world, peak, not args.no_cpu_baseline)
        line["configs"] = subs
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["class_api"] = class_api_record()
    if rank == 0:
        print(json.dumps(line))
    nd.finalize()


def class_api_record():
    """The reference's C++ class loop (`NtHash h(read, ...); while (h.roll()) ...`, examples/benchmark.cpp:31-39) timed
    through this repo's drop-in header against the unmodified reference, both compiled from tests/cpp/class_bench.cpp
    (oracle/Makefile).  Short reads are rolled on the host by the header (no GPU round trip per object); one long sequence
    goes through the CUDA engine in bounded chunks.  Single host thread in both builds."""
    rec = {}
    cases = {"short_reads": ["kmer", "300000", "100", "64", "3"],       # the shape of the reference's own benchmark
             "long_sequence": ["kmer", "1", "100000000", "31", "1"],    # one 100 Mbp sequence
             "seed_long_sequence": ["seed", "1", "20000000", "3"]}
    for name, argv in cases.items():
        r = {}
        for which in ("ref", "shim"):
            exe = os.path.join(ROOT, "oracle", "_ref", "class_bench_" + which)
            if not os.path.exists(exe):
                r[which] = {"unavailable": "oracle/_ref/class_bench_%s not built" % which}
                continue
            try:
                env = dict(os.environ, NTHASH_BENCH_PASSES="2", NTHASH_BENCH_NO_STRANDS="1")
                out = subprocess.run([exe] + argv, capture_output=True, text=True, timeout=300, env=env)
                j = json.loads(out.stdout.strip().splitlines()[-1])
                # first pass (pays the CUDA context once per process) and the better of two passes
                r[which] = {"windows_per_sec": j["windows_per_sec_best"], "windows_per_sec_first_pass": j["windows_per_sec"],
                            "seconds": j["seconds_best"], "max_rss_mb": j["max_rss_mb"], "sum": j["sum"], "windows": j["windows"]}
            except Exception as e:
                r[which] = {"error": f"{type(e).__name__}: {e}"[:200]}
        if "sum" in r.get("ref", {}) and "sum" in r.get("shim", {}):
            r["bit_exact"] = r["ref"]["sum"] == r["shim"]["sum"] and r["ref"]["windows"] == r["shim"]["windows"]
            r["shim_over_ref"] = r["shim"]["windows_per_sec"] / r["ref"]["windows_per_sec"]
        r["workload"] = " ".join(argv)
        rec[name] = r
    return rec


def extra_consumers(torch, nthash_b200, nd, w, args, world):
    """Consumers added in round 2 on the headline batch: window minimizers (w = 10) and the ntCard-style cardinality sketch.
    Device-resident time (CUDA events, max over ranks) + what leaves the device per k-mer."""
    rec = {}
    steps = max(3, min(args.steps, 5))

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return nd.max_over_ranks([e0.elapsed_time(e1) / steps])[0], out

    try:
        window = 10
        cap = w.rows // 4  # ~2/(w+1) of the rows are selected on random sequence
        ms, (bits, mh, mr, n_sel) = timed(lambda: nthash_b200.kmer_minimizers_uniform(w.bases, w.n, w.L, w.k, window, capacity=cap))
        # size-independent properties: every selected row carries the hash the row path stored; selection density ~ 2/(w+1)
        sample = mr[:100000].long()
        same = bool((w.out.view(-1)[sample * w.H] == mh[:100000]).all())
        rec["minimizer"] = {"window": window, "value": world * w.rows / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "selected": int(n_sel),
                            "selected_fraction": n_sel / w.rows, "expected_fraction_random_sequence": 2 / (window + 1),
                            "bytes_out_per_kmer": (16 * n_sel + w.rows / 8) / w.rows, "selected_hashes_match_row_path": same}
        del bits, mh, mr
    except Exception as e:
        rec["minimizer"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    try:
        ms, (counters, res) = timed(lambda: nthash_b200.kmer_sketch_uniform(w.bases, w.n, w.L, w.k, sample_bits=7, index_bits=20))
        total = int(counters.sum())
        rec["cardinality_sketch"] = {"sample_bits": 7, "index_bits": 20, "value": world * w.rows / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                     "windows": int(res[0]), "sampled_per_step": int(res[1]),
                                     "counters_sum_equals_sampled": total == int(res[1]),
                                     "bytes_out_per_step": int(counters.numel() * 4)}
    except Exception as e:
        rec["cardinality_sketch"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    torch.cuda.empty_cache()
    return rec


if __name__ == "__main__":
    main()
