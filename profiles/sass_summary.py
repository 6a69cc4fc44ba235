#!/usr/bin/env python
"""Per-kernel SASS evidence for libnthash_b200.so: registers, shared memory and the counts of the mnemonics that prove the
TMA paths (UTMASTG = tensor store, UBLKCP = bulk copy, SYNCS = mbarrier) plus the ALU- / FMA-pipe instruction totals.

    python profiles/sass_summary.py > profiles/r02_sass_summary.txt

Needs only cuobjdump (no GPU).  The spaced-seed kernels are compiled per seed set at run time with NVRTC and are therefore
not in the library; nthash_seed_jit_selftest compiles one without a GPU (see __graft_entry__.build)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nthash_b200", "libnthash_b200.so")
ALU = ("LOP3", "SHF", "PRMT", "IADD3", "VIADD", "LEA", "ISETP", "PLOP3", "SEL", "VIMNMX", "POPC", "FLO", "BREV")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.split("\n"):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = {}, None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            counts[cur]["alu" if op in ALU else "fma" if op in ("IMAD", "FFMA") else op] += 1
    names = demangle(list(counts))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels (cuobjdump -sass / -res-usage, sm_100a)")
    print(f"# {'kernel':<84} {'regs':>4} {'smem':>6} {'instr':>6} {'ALU':>6} {'FMA':>6} {'LDS':>5} {'STS':>5} {'UTMASTG':>7} {'UBLKCP':>6} {'SYNCS':>5}")
    tot = collections.Counter()
    for fn in sorted(counts, key=lambda f: names[f]):
        c = counts[fn]
        nm = re.sub(r"nthb::\(anonymous namespace\)::|\(nthb::\w+, CUtensorMap_st\)|\(nthb::\w+\)|void ", "", names[fn])[:84]
        r = usage.get(fn, (0, 0))
        print(f"  {nm:<84} {r[0]:>4} {r[1]:>6} {c['total']:>6} {c['alu']:>6} {c['fma']:>6} {c['LDS']:>5} {c['STS']:>5} {c['UTMASTG']:>7} {c['UBLKCP']:>6} {c['SYNCS']:>5}")
        for k in ("UTMASTG", "UBLKCP", "SYNCS"):
            tot[k] += c[k]
    print(f"# totals: UTMASTG {tot['UTMASTG']}, UBLKCP {tot['UBLKCP']}, SYNCS {tot['SYNCS']}; no UTC*MMA / tcgen05: there is no contraction on this path")


if __name__ == "__main__":
    sys.exit(main())
