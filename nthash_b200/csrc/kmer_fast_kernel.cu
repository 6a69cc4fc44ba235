// kmer_fast_kernel.cu — the NtHash batch kernel: uniform and ragged batches, 1..8 hashes per k-mer, optional strand
// hashes, and the fused consumers (count/sum/xor, Bloom filter).  Same arithmetic and the same reference lines as
// kmer_kernel.cu (NtHash::roll src/kmer.cpp:246-264, extend_hashes src/internal.hpp:104-118); what differs is how
// bytes get in and hashes get out.
//
// What the profiles (profiles/r01_ncu_kmer_c2_v*.txt) and the microbenchmarks (profiles/r01_microbench_*.txt)
// showed, and what this kernel does about it:
//  * L1TEX data pipe 99.5 % busy in the first kernel: two 16-byte byte-indexed table lookups (4 wavefronts each)
//    and two bank-conflicting LDS.U8 per window.  Here each lane streams its row as 32-bit words (one LDS.32 per
//    four windows per stream, requested one group ahead, realigned with PRMT) and ONE lookup per window and strand
//    fetches the combined in/out contribution from a 16-entry pair table indexed by the 2-bit codes
//    (byte >> 1) & 3 of the incoming and outgoing base.  Codes of non-ACGTU bytes are garbage but self-consistent
//    (the same byte enters and leaves with the same code), so windows free of such bytes stay exact; a SWAR test on
//    the incoming words flags rows that need the exact scrub pass.
//  * the HBM write path wants long contiguous pieces (128-byte pieces of 960-byte rows: 5.1-5.4 TB/s, 192-byte 5.6,
//    320-byte 6.3-7.0, lane-strided 32-byte sector stores 4.6).  Uniform batches of whole-read items: a warp tile
//    [blocks][32 rows][8 u64] under the 64-byte TMA swizzle leaves through ONE 3-D tensor store (UTMASTG) of
//    192-256 bytes per row.  Everything else (ragged, cut-up reads, odd rows, 5..8 hashes, strand outputs): per-lane
//    256-byte rows in shared memory, copied out with coalesced 16-byte stores, a half-warp per row.
#include "kmer_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

namespace nthb {

namespace {

// The pair table is kept as two 16 x 8 B halves: 128 bytes each = every entry in its own bank pair, so
// the per-lane LDS.64 are conflict-free whatever codes the lanes hold (one 16 x 16 B table put
// entries e and e+8 in the same banks: 205 M conflict wavefronts in profiles/r01_ncu_kmer_c2_v3_fast.txt).
constexpr int F_PAIR_OFF = 0;     // 16 x 8 B : [code_in][code_out] -> S[in]^Sk[out]    (forward strand)
constexpr int F_PAIR_R_OFF = 128; // 16 x 8 B : [code_in][code_out] -> Skc[in]^Sc[out]  (reverse strand)
constexpr int F_IN_OFF = 256;     //  4 x 16 B : [code_in] -> {S[in], Skc[in]}  (warm-up)
constexpr int F_LUT_OFF = 320;    // 256 x 1 B : 0 for ACGTUacgtu, 1 otherwise
constexpr int F_BAR_OFF = 576;    // mbarrier
constexpr int F_RANGE_OFF = 592;  // u64: first staged byte of the CTA (g0), written by thread 0
constexpr int F_TILE_OFF = 608;   // 16-byte pad + staged bases
constexpr int F_TILE_PAD = 16;
constexpr int T4_BYTES = 256 * 16; // tetramer warm-up table

NTH_D uint32_t lds_u8(uint32_t a)
{
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
NTH_D uint32_t lds_u32(uint32_t a)
{
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
NTH_D uint2 lds_v2(uint32_t a)
{
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
NTH_D uint4 lds_v4(uint32_t a)
{
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}

NTH_D void st_shared_u64(uint32_t saddr, uint64_t a)
{
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(saddr), "l"(a) : "memory");
}
NTH_D void tma_store_3d(const void* tmap, uint32_t saddr, int c0, int c1, int c2)
{
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tmap), "r"(c0),
               "r"(c1), "r"(c2), "r"(saddr)
               : "memory");
}
template<int N>
NTH_D void bulk_wait_read()
{
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template<int LUT>
NTH_D uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return d;
}
// LUT bytes: a = 0xF0, b = 0xCC, c = 0xAA

// Makes a value opaque to the compiler so that it stays in a register instead of being re-derived (the shared
// window base costs an S2R SR_CgaCtaId + LEA every time it is rematerialised; ncu showed one per four windows).
NTH_D uint32_t keep(uint32_t v)
{
  asm volatile("" : "+r"(v));
  return v;
}

// Per byte of x: non-zero iff the byte is not one of ACGTUacgtu (what NtHash hashes, src/internal.hpp:132-165).
// A byte is hashable iff bit7 = 0, bit6 = 1, bit3 = 0, bit4 == (bit2 & ~bit1) and (bit4 | bit0) = 1.
NTH_D uint32_t swar_bad(uint32_t x)
{
  const uint32_t t1 = lop3<(0xF0 & ~0xCC) & 0xFF>(x, x << 1, 0u);                 // x & ~(x << 1): bit2 = b2 & ~b1
  const uint32_t e1 = lop3<(0xF0 ^ 0xCC) & 0xAA>(t1, x >> 2, 0x04040404u);        // b4 != (b2 & ~b1)
  const uint32_t e2 = lop3<(~(0xF0 | 0xCC)) & 0xAA & 0xFF>(x, x >> 4, 0x01010101u); // !(b0 | b4)
  const uint32_t e3 = lop3<(0xF0 ^ 0xCC) & 0xAA>(x, 0x40404040u, 0xC8C8C8C8u);    // bit7, bit3 set or bit6 clear
  return lop3<0xF0 | 0xCC | 0xAA>(e1, e2, e3);
}

constexpr int LUT_SEL_C = (0xF0 & ~0xAA & 0xFF) | (0xCC & 0xAA); // (a & ~c) | (b & c)
constexpr int LUT_OR_AND = 0xF0 | (0xCC & 0xAA);                 // a | (b & c)
constexpr int LUT_XOR_AND = (0xF0 ^ 0xCC) & 0xAA;                // (a ^ b) & c
constexpr int LUT_XOR3 = 0xF0 ^ 0xCC ^ 0xAA;                     // a ^ b ^ c

// F <- srol(F) ^ e.xy ; R <- sror(R ^ e.zw) with one combined table entry: 10 ALU-pipe ops + 3 on the FMA pipe
// (IMAD.SHL / IMAD.HI stand in for the plain shifts).
// NTH_ROLL_V2 (round-2 experiment, off): 8 ALU + 7 FMA-pipe ops — "x + (x << 31)" as one IMAD, the bit that crosses from the
// low into the high word as the addend of an IMAD.HI whose multiplier 2 comes from the constant bank (`two`; a literal 2
// is strength-reduced back into an ALU-pipe LEA.HI).  ALU-pipe instructions per window fell 12.5 -> 10.9, the total rose
// 21 -> 22.5, and on the same box the store kernels got 1-1.5 % SLOWER (C2 1.809 -> 1.836 ms, C5 1.007 -> 1.019, C4 10.29 ->
// 10.38; only the store-free reduce consumer gained, 1.657 -> 1.595-1.648): issue slots and latency, not the ALU pipe, are
// what these kernels run out of (profiles/r02_ab_roll_step.txt).
NTH_D uint32_t mad_hi(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t d;
  asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// SH: plain shifts (ALU pipe) instead of shift-by-multiply (FMA pipe).  Measured on one box (profiles/r02_ab_kmer_shifts.txt): the
// consumers gain 2-4 % and C5's flat items 3 % with the shifts, C2 loses 0.5 %: chosen per kernel variant (ROLL_SHIFTS below).
template<bool SH = false>
NTH_D void roll_step(State& s, const uint4 e, const uint32_t two)
{
#ifndef NTH_ROLL_V2
  {
    const uint32_t lo = s.flo, hi = s.fhi;
    const uint32_t hi1 = __funnelshift_l(lo, hi, 1);
    const uint32_t nhi = lop3<LUT_SEL_C>(hi1, SH ? hi >> 30 : __umulhi(hi, 4u), 2u);
    const uint32_t nlo = lop3<LUT_OR_AND>(lo + lo, hi, 1u);
    s.flo = nlo ^ e.x;
    s.fhi = nhi ^ e.y;
  }
  {
    const uint32_t lo = s.rlo ^ e.z, hi = s.rhi ^ e.w;
    s.rlo = __funnelshift_r(lo, hi, 1);
    const uint32_t y = __funnelshift_r(hi, SH ? hi >> 1 : __umulhi(hi, 0x80000000u), 1);
    s.rhi = lop3<LUT_SEL_C>(y, lo, 1u);
  }
  (void)two;
  return;
#endif
  {
    const uint32_t lo = s.flo, hi = s.fhi;
    // 31-bit field (bits 31:1 of hi) rotated left by one, bit 0 clear: bits 31:2 <- hi << 1, bit 1 <- old bit 31
    const uint32_t ch = lop3<LUT_SEL_C>(hi * 2u, __umulhi(hi, 4u), 2u);
    const uint32_t nhi = mad_hi(lo, two, ch);                     // bit 0 (word bit 32) <- old bit 31 of lo
    const uint32_t nlo = lop3<LUT_OR_AND>(lo * 2u, hi, 1u);       // bit 0 <- old bit 32
    s.flo = nlo ^ e.x;
    s.fhi = nhi ^ e.y;
  }
  {
    const uint32_t lo = s.rlo ^ e.z, hi = s.rhi ^ e.w;
    s.rlo = __funnelshift_r(lo, hi, 1);                           // bit 31 <- old bit 32
    const uint32_t a = __umulhi(hi, 0x80000000u);                 // hi >> 1: bits 30:1 in place, bit 0 = old bit 33
    const uint32_t y = a * 0x80000001u;                           // a + (a << 31): bit 31 (word bit 63) <- old bit 33
    s.rhi = lop3<LUT_SEL_C>(y, lo, 1u);                           // bit 0 (word bit 32) <- old bit 0
  }
}

// Four in-only steps at once (warm-up): F <- srol^4(F) ^ t.xy ; R <- sror^4(R ^ t.zw), where t is the
// precombined contribution of four consecutive bases (the idea of the reference's TETRAMER_TAB init,
// src/kmer.cpp:43-73, src/internal.hpp:420-541, applied to the rolling form).  srol^d for d <= 31 is a
// plain 64-bit rotate followed by swapping the d bits that crossed the 33|31 split
// (src/internal.hpp:56-66); sror^d is the same swap followed by the opposite rotate.
NTH_D void roll4_in(State& s, const uint4 t)
{
  {
    const uint32_t vlo = __funnelshift_l(s.fhi, s.flo, 4), vhi = __funnelshift_l(s.flo, s.fhi, 4);
    const uint32_t y = lop3<LUT_XOR_AND>(vlo, vhi >> 1, 0xFu);
    s.flo = lop3<LUT_XOR3>(vlo, y, t.x);
    s.fhi = lop3<LUT_XOR3>(vhi, y << 1, t.y);
  }
  {
    const uint32_t xlo = s.rlo ^ t.z, xhi = s.rhi ^ t.w;
    const uint32_t y = lop3<LUT_XOR_AND>(xlo, xhi >> 1, 0xFu);
    const uint32_t zlo = xlo ^ y, zhi = xhi ^ (y << 1);
    s.rlo = __funnelshift_r(zlo, zhi, 4);
    s.rhi = __funnelshift_r(zhi, zlo, 4);
  }
}

NTH_D uint64_t canonical2(const State& s)
{
  uint32_t lo, hi;
  asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;" : "=r"(lo), "=r"(hi) : "r"(s.flo), "r"(s.rlo), "r"(s.fhi), "r"(s.rhi));
  return ((uint64_t)hi << 32) | lo;
}

// Geometry of item i (shared with kmer_kernel.cu's item_geom, restated here with the last item of a cut-up
// read allowed to be short).
// byte offset of dense window w of a uniform batch (flat geometry)
NTH_D uint64_t flat_byte(const KmerGeom& g, uint64_t w)
{
  const uint64_t r = (w >> 32) ? w / g.nk : (uint64_t)((uint32_t)w / g.nk);
  return r * g.read_len + (w - r * g.nk);
}

NTH_D void fast_item_geom(const KmerGeom& g, uint64_t i, uint64_t& byte, uint64_t& out, uint32_t& n)
{
  if (g.flat) { // a fixed number of dense windows per item; the bytes of two reads are adjacent, so a lane that crosses a
    out = i * g.seg; // read boundary simply keeps rolling (its rows past the boundary are redone by kmer_flat_fix_kernel)
    byte = flat_byte(g, out);
    n = g.seg;
  } else if (g.item_byte) { // ragged: arrays are read_off/koff themselves when every read is one item
    byte = g.item_byte[i];
    out = g.item_out[i];
    n = (uint32_t)(g.item_out[i + 1] - out);
  } else {
    uint64_t r = i;
    uint32_t s = 0;
    if (g.segs > 1) {
      r = (i >> 32) ? i / g.segs : (uint64_t)((uint32_t)i / g.segs); // (a 64-bit division is ~100 instructions)
      s = (uint32_t)(i - r * g.segs);
    }
    byte = r * g.read_len + (uint64_t)s * g.seg;
    out = r * g.nk + (uint64_t)s * g.seg;
    n = min(g.seg, g.nk - s * g.seg);
  }
}

// exact clean-up of one item: windows touching a non-ACGTU byte are not emitted (kmer.cpp:232-235, :255-258)
template<int H0>
__device__ __noinline__ void scrub_lane(const KmerParams& P, uint32_t lut, uint32_t ps, uint64_t my_out, uint32_t n)
{
  const uint32_t k = P.k, H = H0 ? (uint32_t)H0 : P.h;
  uint32_t run = 0;
  for (uint32_t j = 0; j < n + k - 1; ++j) {
    run = lds_u8(lut + lds_u8(ps + j)) ? 0 : run + 1;
    if (j >= k - 1 && run < k) {
      const uint64_t w = my_out + (j - (k - 1));
      for (uint32_t q = 0; q < H; ++q) P.out[w * H + q] = 0;
      if (P.out_fwd) {
        P.out_fwd[w] = 0;
        P.out_rev[w] = 0;
      }
      if (P.valid_bits) atomicAnd(&P.valid_bits[(P.valid_row0 + w) >> 5], ~(1u << ((P.valid_row0 + w) & 31)));
    }
  }
}

// General output path: one lane's shared-memory row = 256 bytes of hashes + 16 bytes of padding (an odd number of
// 16-byte chunks, so the per-lane STS.128 of a quarter-warp fall into distinct bank groups)
constexpr uint32_t ROW1_BYTES = 272;

// REDUCE: fused consumer (the loop of the reference's examples/benchmark.cpp:34-39): instead of storing
// the hashes, count the visited windows and accumulate the 64-bit sum and xor of all their hash values.
//
// Output (REDUCE == false): hashes leave the SM as contiguous pieces of 192-256 bytes per row, which is what the
// HBM write path wants (profiles/r01_microbench_tma_store_patterns_v2.txt: 128-byte pieces 5.1-5.4 TB/s, 192-byte
// 5.6, 320-byte 6.3-7.0), staged through shared memory; pieces start on 32-byte (sector) boundaries: the first
// (-row) mod (4/H) windows of an item go out as plain stores.  (A per-lane cp.async.bulk per piece was tried and
// dropped: UBLKCP is a uniform-datapath instruction, so the compiler serialises it over the 32 lanes at ~12
// instructions each; profiles/r01_sweep_kmer_fast_v5a_1d_bulk.txt.)
//
// BOX == true (uniform batches of whole-read items whose rows are multiples of 64 bytes): the 32 lanes of a warp
// share one tile laid out [WS*H/8 blocks][32 rows][8 u64] under the 64-byte TMA swizzle (conflict-free STS.128),
// and ONE elected lane issues a 3-D tensor store (cp.async.bulk.tensor.3d -> UTMASTG) of box 8 x 32 x blocks:
// the same WS*H*8 contiguous bytes per row, without the 32-iteration issue loop a per-lane bulk copy costs.
//
// CONS: 0 = store the hashes; 1 = REDUCE (count / sum / xor); 5 = cardinality sketch; 2 / 3 = Bloom-filter insert / query (runtime number
// of hashes P.h; position = hash % P.bloom_bits; counts the windows whose positions were all set already).
template<int H, int CONS, int WS, int NBUF, bool BOX>
__global__ void __launch_bounds__(256)
kmer_fast_kernel(const __grid_constant__ KmerParams P, const __grid_constant__ CUtensorMap omap)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + F_BAR_OFF);
  uint64_t* s_range = reinterpret_cast<uint64_t*>(smem + F_RANGE_OFF);
  uint8_t* tile = smem + F_TILE_OFF;

  constexpr bool DIRECT = !BOX && NBUF == 2;       // general output path without the shared-memory rows: 32-byte stores from registers
  constexpr bool STR = CONS == 4;                  // CONS == 4: store the hashes AND the strand hashes (general output path)
  // DIRECT with whole-sector stores at row boundaries too: the u64s between a row's start and the next 32-byte boundary
  // travel through shared memory to the lane that owns the END of the previous row, which stores the shared sector whole
  constexpr bool MERGE = DIRECT && (H == 1 || (H == 2 && !STR));
  constexpr bool REDUCE = CONS != 0 && CONS != 4;  // consumers proper: nothing is stored
  // roll_step with plain shifts: the consumers and the long-piece tensor-store variants (flat items of long reads, h <= 2)
  constexpr bool ROLL_SHIFTS = REDUCE || (BOX && H >= 1 && H <= 2 && WS == (H == 1 ? 40 : 20));
  const uint32_t HH = H ? (uint32_t)H : P.h; // H == 0: runtime number of hashes (5..255, general output path only)
  const uint32_t NT = blockDim.x;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t i0 = (uint64_t)blockIdx.x * NT;
  const uint64_t i1 = min(i0 + (uint64_t)NT, P.g.n_items);
  const uint32_t k = P.k;

  // ---- stage the CTA's byte range (TMA bulk copy): the very first thing thread 0 does --------------
  // The range = first byte of the CTA's first item .. last byte of its last item, in batch order (uniform batches: plain
  // arithmetic; ragged: four entries of the item tables).  Nothing else is in front of the copy — no CTA barrier, no
  // per-thread geometry — so that everything below (own item, tables, the ragged deal) hides behind its latency: the wait for
  // the tile was 17 % of the stall samples and the serial prologue in front of the copy another 10 % (profiles/r02_ncu_c2_head.txt).
  // row buffers follow the staged bases; the 4 KB tetramer table is parked in them for the warm-up phase
  const uint32_t rb_base = (sbase + F_TILE_OFF + F_TILE_PAD + P.tile_cap + 16 + 1023u) & ~1023u;
  if (P.g.flat) asm volatile("griddepcontrol.launch_dependents;"); // the fix-up launch may queue up behind this grid's last wave
  if (tid == 0) {
    uint64_t lo_byte, eb, eo;
    uint32_t en;
    fast_item_geom(P.g, i0, lo_byte, eo, en);
    fast_item_geom(P.g, i1 - 1, eb, eo, en);
    uint64_t g1 = P.g.flat ? flat_byte(P.g, eo + en - 1) + k : eb + (en ? en + k - 1 : 0);
    g1 = max(g1, lo_byte);
    const uint64_t g0 = (lo_byte ? lo_byte - 1 : 0) & ~15ull;
    if (g1 - g0 > P.tile_cap) __trap();
    const uint64_t bulk_end = min((g1 + 15) & ~15ull, P.n_bases & ~15ull);
    const uint32_t bulk_bytes = bulk_end > g0 ? (uint32_t)(bulk_end - g0) : 0u;
    mbar_init(bar, 1);
    fence_mbar_init();
    mbar_expect_tx(bar, bulk_bytes + T4_BYTES);
    if (bulk_bytes) bulk_g2s(tile + F_TILE_PAD, P.bases + g0, bulk_bytes, bar);
    bulk_g2s(smem + (rb_base - sbase), P.t4, T4_BYTES, bar);
    s_range[0] = g0;
    // the (at most 15) bytes of a range that ends in the buffer's unaligned tail
    for (uint64_t g = max(bulk_end, g0); g < g1; ++g) tile[F_TILE_PAD + (g - g0)] = P.bases[g];
    // pull the bases of the CTA that will take this SM slot next into L2 (same extent, one residency wave ahead),
    // so that its start-up wait is an L2 hit instead of a DRAM read queued behind the output stream
    if (P.prefetch_ctas) {
      const uint64_t nxt = g0 + (uint64_t)P.prefetch_ctas * ((g1 - g0) & ~15ull);
      const uint64_t len = (g1 - g0 + 15) & ~15ull;
      if (len && nxt + len <= (P.n_bases & ~15ull)) bulk_prefetch_l2(P.bases + nxt, (uint32_t)len);
    }
  }
  if (tid < F_TILE_PAD) tile[tid] = 'A'; // (visible after the barrier in front of the tile wait)

  uint64_t my_byte = 0, my_out = 0;
  uint32_t n = 0;
  uint32_t slot = tid; // position of this thread's item among the CTA's items (ragged batches: dealt out by length class)
  // ragged batch with a precomputed deal (KmerGeom::item_perm, CTAs of 256 threads): take the item straight away
  const bool presorted = P.g.item_perm != nullptr;
  if (presorted) slot = P.g.item_perm[i0 + tid];
  if (i0 + slot < i1) fast_item_geom(P.g, i0 + slot, my_byte, my_out, n);
  // (the loads are in flight from here on; the bookkeeping below hides behind them)
  if (P.g.item_byte && !presorted) {
    // Ragged batch: a warp runs as long as its longest item, so hand the CTA's items out by length class
    // (32 classes, longest first; counting sort through shared memory that the tables overwrite later).
    // 10 M reads of 100-150 bases: 3.10 -> 2.63 ms per call, of 36-150 bases: 3.55 -> 2.25 ms (profiles/r01_ragged_bench.txt);
    // nothing else depends on which lane owns which item.
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + F_PAIR_OFF); // 32 class counters + the CTA's longest item
    uint8_t* perm = smem + F_LUT_OFF;                               // slot -> thread that first held the item
    if (tid < 33) cnt[tid] = 0;
    __syncthreads();
    atomicMax(&cnt[32], n);
    __syncthreads();
    const uint32_t cls = 31u - (uint32_t)(((uint64_t)n * 32u) / ((uint64_t)cnt[32] + 1u));
    const uint32_t pos = atomicAdd(&cnt[cls], 1u);
    __syncthreads();
    if (tid < 32) { // exclusive scan of the class counters
      const uint32_t c = cnt[tid];
      uint32_t x = c;
      for (uint32_t o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      cnt[tid] = x - c;
    }
    __syncthreads();
    perm[cnt[cls] + pos] = (uint8_t)tid;
    __syncthreads();
    const uint32_t src = perm[tid];
    slot = src;
    my_byte = my_out = 0;
    n = 0;
    if (i0 + src < i1) fast_item_geom(P.g, i0 + src, my_byte, my_out, n);
    __syncthreads(); // the scratch is reused for the tables below
  }

  {
    // code (byte >> 1) & 3 : 0 = A, 1 = C, 2 = T/U, 3 = G ; complement = code ^ 2
    auto code2base = [](int c) { return c ^ (c >> 1); }; // code -> index into P.s / P.sk (A, C, G, T order)
    if (tid < 16) {
      const int ci = tid >> 2, co = tid & 3;
      const uint64_t f = P.s[code2base(ci)] ^ P.sk[code2base(co)];
      const uint64_t r = P.sk[code2base(ci ^ 2)] ^ P.s[code2base(co ^ 2)];
      reinterpret_cast<uint64_t*>(smem + F_PAIR_OFF)[tid] = f;
      reinterpret_cast<uint64_t*>(smem + F_PAIR_R_OFF)[tid] = r;
    } else if (tid < 20) {
      const int ci = tid - 16;
      const uint64_t f = P.s[code2base(ci)], r = P.sk[code2base(ci ^ 2)];
      reinterpret_cast<uint4*>(smem + F_IN_OFF)[ci] = make_uint4((uint32_t)f, (uint32_t)(f >> 32), (uint32_t)r, (uint32_t)(r >> 32));
    }
    for (uint32_t c = tid; c < 256; c += NT) smem[F_LUT_OFF + c] = is_acgtu(c) ? 0 : 1;
  }
  __syncthreads(); // tables, pad and tail bytes written; the barrier's initialisation and the range visible to everybody
  const uint64_t g0 = s_range[0];
  mbar_wait(bar, 0);

  const bool active = n != 0;
  if (BOX && !REDUCE && !active) { // warp-synchronous output path: idle lanes hash a dummy row that the TMA store clips
    my_byte = g0 + 1;
    n = P.g.seg;
  }
  const uint32_t ps = keep(sbase + F_TILE_OFF + F_TILE_PAD + (uint32_t)(my_byte - g0)); // shared address of base 0
  const uint32_t lut = keep(sbase + F_LUT_OFF);
  const uint32_t pair = keep(sbase + F_PAIR_OFF); // 256-byte aligned: a table offset (< 128) can be spliced in with one PRMT

  // ---- warm-up: k in-only steps over bases -1 .. k-2 (base -1 is cancelled by the first roll), ----
  // ---- four bases per step through the tetramer table, then k % 4 single steps               ----
  State s = { 0u, 0u, 0u, 0u };
  uint32_t bad = 0;
  uint32_t run = 0;                   // REDUCE: hashable bases in a row, ending at the newest one
  uint64_t acc_sum = 0, acc_xor = 0;  // REDUCE accumulators
  uint32_t acc_cnt = 0;
  if (n) {
    const uint32_t a_w = ps - 1;
    uint32_t wp = a_w & ~3u;
    const uint32_t sel = 0x3210u + 0x1111u * (a_w & 3u);
    uint32_t w0 = lds_u32(wp);
    const uint32_t nq = k >> 2;
    for (uint32_t q = 0; q < nq; ++q) {
      wp += 4;
      const uint32_t w1 = lds_u32(wp);
      const uint32_t x = __byte_perm(w0, w1, sel);
      w0 = w1;
      // base -1 is checked along with the rest: a false alarm only costs the (exact) scrub pass
      if constexpr (REDUCE) { // per-base flags for the run counter: LUT loads (the consumers are issue-bound, not shared-memory-bound)
        const uint32_t v0 = lds_u8(lut + __byte_perm(x, 0u, 0x4440u)), v1 = lds_u8(lut + __byte_perm(x, 0u, 0x4441u));
        const uint32_t v2 = lds_u8(lut + __byte_perm(x, 0u, 0x4442u)), v3 = lds_u8(lut + (x >> 24));
        bad |= v0 | v1 | v2 | v3;
        run = v0 ? 0 : run + 1;
        run = v1 ? 0 : run + 1;
        run = v2 ? 0 : run + 1;
        run = v3 ? 0 : run + 1;
      } else { // SWAR test, as in the main loop: the four byte-indexed loads were a fifth of the warm-up's shared-memory traffic
        bad |= swar_bad(x);
      }
      const uint32_t y2 = (x >> 1) & 0x03030303u;          // 2-bit codes, first base in byte 0
      const uint32_t off = ((y2 * 0x40100401u) >> 20) & 0xFF0u; // 16 * (c0<<6 | c1<<4 | c2<<2 | c3)
      roll4_in(s, lds_v4(rb_base + off));
    }
    for (uint32_t j = 4 * nq; j < k; ++j) {
      const uint32_t c = lds_u8(a_w + j);
      const uint32_t v = lds_u8(lut + c);
      bad |= v;
      if (REDUCE) run = v ? 0 : run + 1;
      roll_step<ROLL_SHIFTS>(s, lds_v4(sbase + F_IN_OFF + ((c & 6u) << 3)), P.two);
    }
  }
  if (!MERGE) __syncthreads(); // the tetramer table is dead from here on: its bytes become row buffers

  // one window through single-byte loads (alignment peel)
  auto roll1 = [&](uint32_t p) -> uint64_t {
    const uint32_t ci = lds_u8(ps + k - 1 + p), co = lds_u8(ps - 1 + p);
    const uint32_t v = lds_u8(lut + ci);
    bad |= v;
    if (REDUCE) run = v ? 0 : run + 1;
    const uint32_t ea = pair + (((ci & 6u) << 4) | ((co & 6u) << 2));
    const uint2 ef = lds_v2(ea), er = lds_v2(ea + (F_PAIR_R_OFF - F_PAIR_OFF));
    roll_step<ROLL_SHIFTS>(s, make_uint4(ef.x, ef.y, er.x, er.y), P.two);
    return canonical2(s);
  };
  auto consume = [&](uint64_t h0) { // one window the reference visits
    {
      {
      if (CONS == 1) {
        acc_sum += h0;
        acc_xor ^= h0;
        if (H) {
#pragma unroll
          for (int q = 1; q < H; ++q) {
            const uint64_t e = ext_hash(h0, P.mult[q]);
            acc_sum += e;
            acc_xor ^= e;
          }
        } else { // runtime number of hashes (5..255); mult[0] = k * MULTISEED
          for (uint32_t q = 1; q < HH; ++q) {
            const uint64_t e = ext_hash(h0, (uint64_t)q ^ P.mult[0]);
            acc_sum += e;
            acc_xor ^= e;
          }
        }
      } else if (CONS == 5) {
        // ntCard-style cardinality sketch: of the k-mers whose canonical hash has its top s bits clear (a 2^-s sample),
        // count the multiplicity of the next r bits in a table of 2^r counters (bloom_bits = s << 8 | r)
        const uint32_t sb = (uint32_t)(P.bloom_bits >> 8), rb = (uint32_t)(P.bloom_bits & 0xFF);
        if ((h0 >> (64 - sb)) == 0) {
          atomicAdd(P.bloom_words + (uint32_t)((h0 >> (64 - sb - rb)) & ((1ull << rb) - 1)), 1u);
          acc_sum += 1; // k-mers sampled
        }
      } else { // Bloom filter: P.h positions per window (extend_hashes, src/internal.hpp:104-118, with a runtime count)
        const uint64_t kmul = P.mult[0]; // k * MULTISEED
        const bool pow2 = (P.bloom_bits & (P.bloom_bits - 1)) == 0;
        bool all = true;
        for (uint32_t q = 0; q < P.h; ++q) {
          const uint64_t hq = q ? ext_hash(h0, (uint64_t)q ^ kmul) : h0;
          const uint64_t pos = pow2 ? hq & (P.bloom_bits - 1) : hq % P.bloom_bits;
          const uint32_t bit = 1u << (pos & 31);
          const uint32_t old = CONS == 2 ? atomicOr(P.bloom_words + (pos >> 5), bit) : __ldg(P.bloom_words + (pos >> 5));
          all = all && (old & bit);
        }
        acc_sum += all ? 1u : 0u;
      }
      }
    }
  };
  auto reduce_add = [&](uint64_t h0) {
    if (run >= k) { // the window is one the reference visits
      ++acc_cnt;
      consume(h0);
    }
  };

  // ---- alignment peel: the windows before the lane's next output u64 sits on a 32-byte boundary ----
  // MERGE: shared memory behind the (still live) tetramer table: [ok flag per item][32 B per item and output array]
  const uint32_t hd_ok = rb_base + T4_BYTES, hd_val = hd_ok + 256;
  bool okme = false;
  uint32_t p = 0;
  if (!REDUCE) {
    // windows per 32-byte boundary of the row: 4 / gcd(h, 4)
    const uint32_t ALIGN_W = STR ? 4 : (HH & 3) == 0 ? 1 : (HH & 1) == 0 ? 2 : 4; // strand rows are 8 bytes per window
    const uint32_t pw = (uint32_t)((0 - my_out) & (uint64_t)(ALIGN_W - 1));
    const uint32_t peel = min(n, pw);
    for (; p < peel; ++p) {
      const uint64_t h0 = roll1(p);
      if (MERGE) { // parked; who stores them is decided after the barrier below
        st_shared_u64(hd_val + slot * 32u + p * (uint32_t)H * 8u, h0);
        if (H == 2) st_shared_u64(hd_val + slot * 32u + 8u, ext_hash(h0, P.mult[1]));
        if (STR) {
          st_shared_u64(hd_val + (NT + slot) * 32u + p * 8u, ((uint64_t)s.fhi << 32) | s.flo);
          st_shared_u64(hd_val + (2u * NT + slot) * 32u + p * 8u, ((uint64_t)s.rhi << 32) | s.rlo);
        }
        continue;
      }
      uint64_t* o = P.out + (my_out + p) * HH;
      o[0] = h0;
      if (H) {
#pragma unroll
        for (int q = 1; q < H; ++q) o[q] = ext_hash(h0, P.mult[q]);
      } else {
        for (uint32_t q = 1; q < HH; ++q) o[q] = ext_hash(h0, (uint64_t)q ^ P.mult[0]);
      }
      if (STR) { // get_forward_hash() / get_reverse_hash() of this window (nthash.hpp:183-194)
        P.out_fwd[my_out + p] = ((uint64_t)s.fhi << 32) | s.flo;
        P.out_rev[my_out + p] = ((uint64_t)s.rhi << 32) | s.rlo;
      }
    }
    if (MERGE) {
      // mergeable: the row reaches its first sector boundary, and no byte seen so far (bases -1 .. peel+k-2) was flagged,
      // i.e. the parked values are final: the owner's scrub pass, which is not ordered against the neighbour's store of
      // them, will never have to zero them
      okme = n >= pw && n != 0 && bad == 0;
      asm volatile("st.shared.u8 [%0], %1;" ::"r"(hd_ok + slot), "r"(okme ? 1u : 0u) : "memory");
      __syncthreads();
      // the previous item's lane takes these u64s along with its own last ones iff both rows reach the shared sector's
      // boundaries and both items sit in this CTA; otherwise they go out from here as plain (partial-sector) stores
      const bool prev_takes = okme && slot > 0 && lds_u8(hd_ok + slot - 1) != 0;
      if (peel && !prev_takes) {
        for (uint32_t j = 0; j < peel * (uint32_t)H; ++j) {
          uint64_t v;
          asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(hd_val + slot * 32u + j * 8u));
          P.out[my_out * H + j] = v;
        }
        if (STR)
          for (uint32_t j = 0; j < peel; ++j) {
            uint64_t a, b;
            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(a) : "r"(hd_val + (NT + slot) * 32u + j * 8u));
            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(b) : "r"(hd_val + (2u * NT + slot) * 32u + j * 8u));
            P.out_fwd[my_out + j] = a;
            P.out_rev[my_out + j] = b;
          }
      }
    }
  }

  // ---- main loop: word streams + pair table ------------------------------------------------------
  // in-stream starts at base k-1+p, out-stream at base p-1; both are read as aligned words + PRMT realign
  const uint32_t a_in = ps + k - 1 + p, a_out = ps - 1 + p;
  uint32_t wp_in = a_in & ~3u, wp_out = a_out & ~3u;
  const uint32_t sel_in = 0x3210u + 0x1111u * (a_in & 3u), sel_out = 0x3210u + 0x1111u * (a_out & 3u);
  uint32_t w_in = n ? lds_u32(wp_in) : 0u, w_out = n ? lds_u32(wp_out) : 0u;

  uint32_t w_in_n = n ? lds_u32(wp_in + 4) : 0u, w_out_n = n ? lds_u32(wp_out + 4) : 0u; // one word ahead
  wp_in += 4;
  wp_out += 4;

  // Four windows (FULL) or the first cnt < 4 of them.  The loop is a two-stage pipeline: while a group is rolled, the words of
  // the group after the next are requested and the eight pair-table entries of the NEXT group are fetched (each entry right
  // after the roll that consumed its predecessor), so every shared-memory load has a whole group of rolls to land in.  (With
  // the loads next to their uses the DIRECT loop, which ptxas keeps at 32-40 registers, had a third of its stall samples on
  // them: profiles/r02_ncu_ragged_head.txt.)
  uint64_t fw4[4], rv4[4]; // STR: strand hashes of the group's four windows
  uint2 efv[4], erv[4];    // the current group's table entries
  uint32_t bw_cur = 0;     // ... and its per-byte validity flags
  auto fetch_codes = [&]() -> uint32_t { // consumes one realigned word of each stream: table offsets of the next group
    const uint32_t x_in = __byte_perm(w_in, w_in_n, sel_in), x_out = __byte_perm(w_out, w_out_n, sel_out);
    w_in = w_in_n;
    w_out = w_out_n;
    wp_in += 4;
    wp_out += 4;
    w_in_n = lds_u32(wp_in);
    w_out_n = lds_u32(wp_out);
    bw_cur = swar_bad(x_in); // bytes past the row's end may flag a false alarm: that only costs the exact scrub
    // per byte: code_in at bits 5-6, code_out at bits 3-4  =>  byte = 8 * (4*code_in + code_out) = table offset
    return lop3<LUT_SEL_C>(x_out << 2, x_in << 4, 0x60606060u) & 0x78787878u;
  };
  // (the runtime-h variant, H == 0, sits at ptxas' 128-register ceiling already: the 16 extra live registers spilled and cost
  // 2-6 % there, so it fetches a group's entries at the start of the group instead)
  constexpr bool PIPE = H != 0 && (BOX || DIRECT || REDUCE); // (the shared-memory-row variants are near the ceiling too)
  if (PIPE && n) { // prime: the first group's entries
    const uint32_t c4 = fetch_codes();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t ea = __byte_perm(c4, pair, 0x7650u | i); // (pair & ~0xFF) | byte i of c4
      efv[i] = lds_v2(ea);
      erv[i] = lds_v2(ea + (F_PAIR_R_OFF - F_PAIR_OFF));
    }
  }
  auto roll4 = [&](uint64_t (&hv)[4], auto full, uint32_t cnt) {
    constexpr bool FULL = decltype(full)::value;
    uint32_t bw = bw_cur;
    const uint32_t c4n = fetch_codes(); // PIPE: the group after this one (reads at most two words past the row: staged or padding)
    if (!PIPE) {
      bw = bw_cur;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t ea = __byte_perm(c4n, pair, 0x7650u | i);
        efv[i] = lds_v2(ea);
        erv[i] = lds_v2(ea + (F_PAIR_R_OFF - F_PAIR_OFF));
      }
    }
    if (!REDUCE) bad |= bw;
    if (REDUCE && FULL && bw == 0 && run + 1 >= k) { // consumer, steady state: all four windows are visited ones
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        roll_step<ROLL_SHIFTS>(s, make_uint4(efv[i].x, efv[i].y, erv[i].x, erv[i].y), P.two);
        if (PIPE) {
          const uint32_t ea = __byte_perm(c4n, pair, 0x7650u | i);
          efv[i] = lds_v2(ea);
          erv[i] = lds_v2(ea + (F_PAIR_R_OFF - F_PAIR_OFF));
        }
        consume(canonical2(s));
      }
      run += 4;
      acc_cnt += 4;
      return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!FULL && i && (uint32_t)i >= cnt) break; // (a partial group is the item's last: nothing is fetched after it)
      if (REDUCE) {
        const uint32_t v = __byte_perm(bw, 0u, 0x4440u | i);
        bad |= v;
        run = v ? 0 : run + 1;
      }
      roll_step<ROLL_SHIFTS>(s, make_uint4(efv[i].x, efv[i].y, erv[i].x, erv[i].y), P.two);
      if (PIPE && FULL) {
        const uint32_t ea = __byte_perm(c4n, pair, 0x7650u | i);
        efv[i] = lds_v2(ea);
        erv[i] = lds_v2(ea + (F_PAIR_R_OFF - F_PAIR_OFF));
      }
      hv[i] = canonical2(s);
      if (STR) {
        fw4[i] = ((uint64_t)s.fhi << 32) | s.flo;
        rv4[i] = ((uint64_t)s.rhi << 32) | s.rlo;
      }
      if (REDUCE) reduce_add(hv[i]);
    }
  };
  using full_t = std::true_type;
  using part_t = std::false_type;

  if constexpr (REDUCE) {
    for (; p + 4 <= n; p += 4) {
      uint64_t hv[4];
      roll4(hv, full_t(), 4u);
    }
    if (p < n) {
      uint64_t hv[4];
      roll4(hv, part_t(), n - p);
    }
    for (int o = 16; o; o >>= 1) {
      acc_cnt += __shfl_down_sync(0xffffffffu, acc_cnt, o);
      acc_sum += __shfl_down_sync(0xffffffffu, acc_sum, o);
      acc_xor ^= __shfl_down_sync(0xffffffffu, acc_xor, o);
    }
    if (lane == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(P.reduce_out), (unsigned long long)acc_cnt);
      atomicAdd(reinterpret_cast<unsigned long long*>(P.reduce_out) + 1, (unsigned long long)acc_sum);
      atomicXor(reinterpret_cast<unsigned long long*>(P.reduce_out) + 2, (unsigned long long)acc_xor);
    }
    return;
  } else if constexpr (BOX) {
    constexpr uint32_t TILE = (uint32_t)(WS * H / 8) * 2048u; // blocks x 32 rows x 64 bytes
    const uint32_t tb0 = rb_base + warp * (NBUF * TILE);
    // SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3; tiles are 2048-byte multiples, so the lane term can be OR-ed in
    const uint32_t lterm = keep(lane * 64 + (((lane >> 1) & 3) << 4));
    const int row0 = (int)(i0 + warp * 32);
    auto chunk_addr = [&](uint32_t tbl, uint32_t cc) { return (tbl ^ ((cc & 3u) << 4)) + (cc >> 2) * 2048u; };
    uint32_t t = 0, b = 0;
    for (; p < n; p += WS) { // n is the same for every lane here
      const uint32_t cnt = min((uint32_t)WS, n - p);
      const uint32_t tb_raw = tb0 + b * TILE, tb = tb_raw + lterm;
      // one group of four windows -> the tile (FULL: all four, no per-window branches)
      auto group = [&](uint32_t q, auto full, uint32_t c4n) {
        uint64_t hv[4];
        roll4(hv, full, c4n);
        if (q == 0 && t >= (uint32_t)NBUF) { // the tile stored NBUF steps ago must have left shared memory
          if (lane == 0) bulk_wait_read<NBUF - 1>();
          __syncwarp();
        }
        if (H == 1) {
          st_shared_v2_u64(chunk_addr(tb, 2 * q), hv[0], hv[1]);
          st_shared_v2_u64(chunk_addr(tb, 2 * q + 1), hv[2], hv[3]);
        } else if (H == 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i) st_shared_v2_u64(chunk_addr(tb, 4 * q + i), hv[i], ext_hash(hv[i], P.mult[1]));
        } else if (H == 3) { // 12 u64 per group = 6 chunks; chunks straddle windows
          uint64_t v[12];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[3 * i] = hv[i];
            v[3 * i + 1] = ext_hash(hv[i], P.mult[1]);
            v[3 * i + 2] = ext_hash(hv[i], P.mult[2]);
          }
#pragma unroll
          for (int c = 0; c < 6; ++c) st_shared_v2_u64(chunk_addr(tb, 6 * q + c), v[2 * c], v[2 * c + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            st_shared_v2_u64(chunk_addr(tb, 8 * q + 2 * i), hv[i], ext_hash(hv[i], P.mult[1]));
            st_shared_v2_u64(chunk_addr(tb, 8 * q + 2 * i + 1), ext_hash(hv[i], P.mult[2]), ext_hash(hv[i], P.mult[3]));
          }
        }
      };
      if (cnt == (uint32_t)WS) {
#pragma unroll
        for (uint32_t q = 0; q < (uint32_t)WS / 4; ++q) group(q, full_t(), 4u);
      } else { // the last, shorter tile of a row: kept rolled up (small code for the instruction cache)
#pragma unroll 1
        for (uint32_t q = 0; 4 * q < cnt; ++q) {
          if (4 * q + 4 <= cnt) group(q, full_t(), 4u);
          else group(q, part_t(), cnt - 4 * q);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&omap, tb_raw, 0, row0, (int)(p * H / 8));
        bulk_commit();
      }
      ++t;
      b = (b + 1 == (uint32_t)NBUF) ? 0 : b + 1;
    }
    const bool dirty = active && bad != 0;
    const bool any_dirty = __any_sync(0xffffffffu, dirty);
    if (lane == 0) {
      if (any_dirty) bulk_wait_all0(); // zeros must land after the tile they overwrite
      else bulk_wait_read0();          // shared memory must outlive the TMA read
    }
    __syncwarp();
    if (dirty) { // flat geometry: only the windows before the read boundary are this lane's to clean
      const uint32_t n_own = P.g.flat ? (uint32_t)min((uint64_t)n, (uint64_t)P.g.nk - my_out % P.g.nk) : n;
      scrub_lane<H>(P, lut, ps, my_out, n_own);
    }
  } else {
    // General output path (ragged batches, cut-up reads, rows that are not 64-byte multiples): every lane collects
    // 256 bytes of its row (32/H windows) in a private shared-memory row, then the warp copies the 32 rows out
    // with coalesced 16-byte stores, two rows per instruction (a half-warp per row).  Each row's global address
    // and byte count travel through a 16-byte descriptor, so lanes may differ in length and alignment.
    if constexpr (DIRECT && H != 0) {
      // DIRECT: no staging at all.  After the peel every group of four windows starts on a 32-byte boundary of each
      // output array, so it leaves as full-sector 32-byte stores straight from the registers (STG.E.ENL2.256; a 16-byte
      // store is a partial sector and ran 5x slower in profiles/r01_microbench_store_patterns.txt).  Lanes are
      // independent here: no warp-level hand-off, no row buffers, no descriptors.
      for (; p + 4 <= n; p += 4) {
        uint64_t hv[4];
        roll4(hv, full_t(), 4u);
        uint64_t* o = P.out + (my_out + p) * H;
        if (H == 1) {
          st_global_v4_u64(o, hv[0], hv[1], hv[2], hv[3]);
        } else if (H == 2) {
          st_global_v4_u64(o, hv[0], ext_hash(hv[0], P.mult[1]), hv[1], ext_hash(hv[1], P.mult[1]));
          st_global_v4_u64(o + 4, hv[2], ext_hash(hv[2], P.mult[1]), hv[3], ext_hash(hv[3], P.mult[1]));
        } else if (H == 3) {
          uint64_t v[12];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[3 * i] = hv[i];
            v[3 * i + 1] = ext_hash(hv[i], P.mult[1]);
            v[3 * i + 2] = ext_hash(hv[i], P.mult[2]);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) st_global_v4_u64(o + 4 * c, v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            st_global_v4_u64(o + 4 * i, hv[i], ext_hash(hv[i], P.mult[1]), ext_hash(hv[i], P.mult[2]), ext_hash(hv[i], P.mult[3]));
        }
        if (STR) {
          st_global_v4_u64(P.out_fwd + my_out + p, fw4[0], fw4[1], fw4[2], fw4[3]);
          st_global_v4_u64(P.out_rev + my_out + p, rv4[0], rv4[1], rv4[2], rv4[3]);
        }
      }
      if (p < n) { // the last one to three windows of the item
        uint64_t hv[4];
        const uint32_t cnt = n - p;
        roll4(hv, part_t(), cnt);
        bool done = false;
        if constexpr (MERGE) {
          // the sector this row ends in also holds the first u64s of the next row: if that row's lane parked them
          // (same CTA, both rows reach their boundaries), store the sector whole
          if (okme && slot + 1 < (uint32_t)(i1 - i0) && lds_u8(hd_ok + slot + 1) != 0) {
            auto nh = [&](uint32_t arr, uint32_t j) {
              uint64_t v;
              asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(hd_val + (arr * NT + slot + 1) * 32u + j * 8u));
              return v;
            };
            auto put4 = [&](uint64_t* o, uint32_t arr, const uint64_t (&x)[4]) { // x[0..cnt) then the neighbour's first 4 - cnt
              if (cnt == 1) st_global_v4_u64(o, x[0], nh(arr, 0), nh(arr, 1), nh(arr, 2));
              else if (cnt == 2) st_global_v4_u64(o, x[0], x[1], nh(arr, 0), nh(arr, 1));
              else st_global_v4_u64(o, x[0], x[1], x[2], nh(arr, 0));
            };
            if (H == 1) {
              put4(P.out + my_out + p, 0, hv);
              if (STR) {
                put4(P.out_fwd + my_out + p, 1, fw4);
                put4(P.out_rev + my_out + p, 2, rv4);
              }
            } else { // H == 2: whole sectors for window pairs, the odd last window shares its sector with the next row
              uint64_t* o = P.out + (my_out + p) * 2;
              if (cnt >= 2) st_global_v4_u64(o, hv[0], ext_hash(hv[0], P.mult[1]), hv[1], ext_hash(hv[1], P.mult[1]));
              if (cnt & 1) {
                const uint64_t hl = cnt == 1 ? hv[0] : hv[2];
                st_global_v4_u64(o + 2 * (cnt - 1), hl, ext_hash(hl, P.mult[1]), nh(0, 0), nh(0, 1));
              }
            }
            done = true;
          }
        }
        for (uint32_t i = 0; i < cnt && !done; ++i) {
          const uint64_t h0 = i == 0 ? hv[0] : i == 1 ? hv[1] : hv[2];
          uint64_t* o = P.out + (my_out + p + i) * H;
          o[0] = h0;
#pragma unroll
          for (int q = 1; q < H; ++q) o[q] = ext_hash(h0, P.mult[q]);
          if (STR) {
            P.out_fwd[my_out + p + i] = i == 0 ? fw4[0] : i == 1 ? fw4[1] : fw4[2];
            P.out_rev[my_out + p + i] = i == 0 ? rv4[0] : i == 1 ? rv4[1] : rv4[2];
          }
        }
        p = n;
      }
      __syncwarp();
      if (bad != 0) scrub_lane<H>(P, lut, ps, my_out, n);
      return;
    }
    constexpr uint32_t WS1 = H == 3 ? 8 : 32 / (H ? H : 1); // windows per row piece (compile-time h): <= 256 bytes
    constexpr uint32_t NARR = STR ? 3 : 1; // output arrays: hashes (+ forward and reverse strand hashes)
    const uint32_t wbase = rb_base + (tid & ~31u) * NARR * (ROW1_BYTES + 16); // this warp: NARR x [32 descriptors], NARR x [32 rows]
    const uint32_t desc0 = wbase, rows0 = wbase + NARR * 32 * 16;
    const uint32_t rb = rows0 + lane * ROW1_BYTES;
    const uint32_t rbf = rb + 32 * ROW1_BYTES, rbr = rb + 64 * ROW1_BYTES; // STR only
    const uint32_t hw = lane >> 4, c16 = (lane & 15) * 16;
    // the warp's 32 rows of one output array -> global memory: all loads of a batch of rows first (they do not depend
    // on each other), then the stores; a half-warp per row, each row's address and byte count from its descriptor
    auto copy_rows = [&](uint32_t descA, uint32_t rowsA) {
#pragma unroll
      for (uint32_t r0 = 0; r0 < 16; r0 += 8) {
        uint4 d[8], v[8];
#pragma unroll
        for (uint32_t rr = 0; rr < 8; ++rr) {
          const uint32_t row = 2 * (r0 + rr) + hw;
          d[rr] = lds_v4(descA + row * 16); // {address lo, address hi, bytes, 0}
          v[rr] = lds_v4(rowsA + row * ROW1_BYTES + c16);
        }
#pragma unroll
        for (uint32_t rr = 0; rr < 8; ++rr) {
          uint8_t* ga = reinterpret_cast<uint8_t*>(((uint64_t)d[rr].y << 32) | d[rr].x) + c16;
          if (c16 + 16 <= d[rr].z)
            asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(ga), "r"(v[rr].x), "r"(v[rr].y), "r"(v[rr].z), "r"(v[rr].w) : "memory");
          else if (c16 + 8 == d[rr].z) // odd last u64 of an item
            asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(ga), "r"(v[rr].x), "r"(v[rr].y) : "memory");
        }
      }
    };
    if constexpr (H == 0) {
      // Runtime number of hashes (5..255): a lane's output is one contiguous stream of u64 (window after window, h values
      // each), cut into 256-byte pieces wherever they fall.  Values are appended to the lane's row; when a running lane
      // has 32 of them the warp copies its rows out.  `ufill` counts as a lane that is still running does (warp-uniform),
      // `fill` what this lane really appended (lanes whose item has ended append nothing).
      uint32_t fill = 0, ufill = 0;
      uint64_t gaddr = (uint64_t)(uintptr_t)(P.out + (my_out + p) * HH);
      auto flush = [&]() {
        st_shared_v2_u64(desc0 + lane * 16, gaddr, (uint64_t)(fill * 8u));
        __syncwarp();
        copy_rows(desc0, rows0);
        __syncwarp();
        gaddr += fill * 8u;
        fill = 0;
        ufill = 0;
      };
      const uint64_t kmul = P.mult[0]; // k * MULTISEED
      while (__any_sync(0xffffffffu, p < n)) {
        const uint32_t cnt = p < n ? min(4u, n - p) : 0u;
        uint64_t hv[4] = { 0, 0, 0, 0 };
        if (cnt == 4) roll4(hv, full_t(), 4u);
        else if (cnt) roll4(hv, part_t(), cnt);
        if (STR) {
          st_shared_v2_u64(rbf, fw4[0], fw4[1]);
          st_shared_v2_u64(rbf + 16, fw4[2], fw4[3]);
          st_shared_v2_u64(rbr, rv4[0], rv4[1]);
          st_shared_v2_u64(rbr + 16, rv4[2], rv4[3]);
          st_shared_v2_u64(desc0 + (32 + lane) * 16, (uint64_t)(uintptr_t)(P.out_fwd + my_out + p), (uint64_t)(cnt * 8));
          st_shared_v2_u64(desc0 + (64 + lane) * 16, (uint64_t)(uintptr_t)(P.out_rev + my_out + p), (uint64_t)(cnt * 8));
          __syncwarp();
          copy_rows(desc0 + 32 * 16, rows0 + 32 * ROW1_BYTES);
          copy_rows(desc0 + 64 * 16, rows0 + 64 * ROW1_BYTES);
          __syncwarp();
        }
#pragma unroll 1
        for (uint32_t i = 0; i < 4; ++i) {
          const uint64_t h0 = hv[0];
          hv[0] = hv[1]; // rotate instead of indexing: hv stays in registers
          hv[1] = hv[2];
          hv[2] = hv[3];
          const bool on = i < cnt;
#pragma unroll 1
          for (uint32_t e = 0; e < HH; ++e) {
            if (on) {
              st_shared_u64(rb + fill * 8u, e ? ext_hash(h0, (uint64_t)e ^ kmul) : h0);
              ++fill;
            }
            if (++ufill == 32) flush();
          }
        }
        p += cnt;
      }
      if (ufill) flush();
    } else {
    while (__any_sync(0xffffffffu, p < n)) {
      const uint32_t cnt = p < n ? min(WS1, n - p) : 0u;
      auto group = [&](uint32_t q, auto full, uint32_t c4n) {
        uint64_t hv[4];
        roll4(hv, full, c4n);
        if (STR) {
          st_shared_v2_u64(rbf + (2 * q) * 16, fw4[0], fw4[1]);
          st_shared_v2_u64(rbf + (2 * q + 1) * 16, fw4[2], fw4[3]);
          st_shared_v2_u64(rbr + (2 * q) * 16, rv4[0], rv4[1]);
          st_shared_v2_u64(rbr + (2 * q + 1) * 16, rv4[2], rv4[3]);
        }
        if (H == 1) {
          st_shared_v2_u64(rb + (2 * q) * 16, hv[0], hv[1]);
          st_shared_v2_u64(rb + (2 * q + 1) * 16, hv[2], hv[3]);
        } else if (H == 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i) st_shared_v2_u64(rb + (4 * q + i) * 16, hv[i], ext_hash(hv[i], P.mult[1]));
        } else if (H == 3) {
          uint64_t v[12];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[3 * i] = hv[i];
            v[3 * i + 1] = ext_hash(hv[i], P.mult[1]);
            v[3 * i + 2] = ext_hash(hv[i], P.mult[2]);
          }
#pragma unroll
          for (int c = 0; c < 6; ++c) st_shared_v2_u64(rb + (6 * q + c) * 16, v[2 * c], v[2 * c + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            st_shared_v2_u64(rb + (8 * q + 2 * i) * 16, hv[i], ext_hash(hv[i], P.mult[1]));
            st_shared_v2_u64(rb + (8 * q + 2 * i + 1) * 16, ext_hash(hv[i], P.mult[2]), ext_hash(hv[i], P.mult[3]));
          }
        }
      };
      if (cnt == WS1) {
#pragma unroll
        for (uint32_t q = 0; q < WS1 / 4; ++q) group(q, full_t(), 4u);
      } else { // the last piece of an item: kept rolled up (once per item; the code stays small for the instruction cache)
#pragma unroll 1
        for (uint32_t q = 0; 4 * q < cnt; ++q) {
          if (4 * q + 4 <= cnt) group(q, full_t(), 4u);
          else group(q, part_t(), cnt - 4 * q);
        }
      }
      st_shared_v2_u64(desc0 + lane * 16, (uint64_t)(uintptr_t)(P.out + (my_out + p) * HH), (uint64_t)(cnt * HH * 8));
      if (STR) {
        st_shared_v2_u64(desc0 + (32 + lane) * 16, (uint64_t)(uintptr_t)(P.out_fwd + my_out + p), (uint64_t)(cnt * 8));
        st_shared_v2_u64(desc0 + (64 + lane) * 16, (uint64_t)(uintptr_t)(P.out_rev + my_out + p), (uint64_t)(cnt * 8));
      }
      __syncwarp();
#pragma unroll
      for (uint32_t arr = 0; arr < NARR; ++arr) copy_rows(desc0 + arr * 32 * 16, rows0 + arr * 32 * ROW1_BYTES);
      __syncwarp(); // rows are rewritten next; also orders these stores before the scrub's zeros
      p += cnt;
    }
    }
    if (bad != 0) scrub_lane<H>(P, lut, ps, my_out, n);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// kmer_pack_kernel — the BOX path (uniform batches, 1..4 hashes, tensor stores) with WARP-PRIVATE, NIBBLE-PACKED bases.
//
// What the profile of kmer_fast_kernel says at HEAD of round 2 (profiles/r02_ncu_c2_head.txt): 17.8 % of the stall samples
// sit on the CTA's wait for its base tile, sm__warps_active is 23.8 % (2 CTAs x 8 warps: 150 bytes of staged ASCII per
// thread next to the 192-byte output rows), and neither the ALU pipe (66 %) nor DRAM (74 %) is saturated: the kernel
// runs out of warps to hide latency with.  Here
//   * every warp stages the bytes of ITS 32 items itself: coalesced 16-byte global loads, converted in registers to one
//     nibble per base (2-bit code (byte >> 1) & 3 in the low bits) and stored to a warp-private strip — 1/2 byte of shared
//     memory per base instead of 1, no CTA-wide barrier after the tables, no TMA round trip: a warp that waits for its
//     bytes only stalls itself, and ~1.5x as many warps fit an SM;
//   * validity is decided once per 16-byte chunk during the conversion (SWAR test, one ballot per 32 chunks); a lane
//     whose row touches a flagged chunk takes the exact scrub pass over the ASCII bytes in global memory afterwards;
//   * the main loop reads one aligned word per 8 windows per stream (funnel shift by the lane's nibble offset), merges
//     the in- and out-stream into one word of 4-bit table indices with a single IMAD, and forms each pair-table address
//     with one rotate and one LOP3: ~2.4 ALU-pipe ops per window for the lookups against 5.25 in kmer_fast_kernel;
//   * hashes leave exactly as there: warp tile [blocks][32 rows][8 u64] under the 64-byte swizzle, one 3-D tensor store.
// Same arithmetic, same reference lines (NtHash::roll src/kmer.cpp:246-264, extend_hashes src/internal.hpp:104-118).
//
// MEASURED (profiles/r02_ab_pack.txt, profiles/r02_ncu_c2_pack.txt, same box as kmer_fast_kernel): sm__warps_active 23.8 -> 35.9 %,
// issue slots 57.6 -> 65 %, the tile wait is gone — but the conversion costs more than the leaner loop saves: 38 instructions
// per window overall against 31.8 (1.43 G vs 1.19 G warp instructions on C2; the SWAR validity test alone is 9 ALU ops per
// four staged bases, and 1.26 bases are staged per window of a 150 bp read), the ALU pipe goes from 66 to 75 % busy, and the
// kernel is SLOWER: C2 1.896 ms vs 1.809 (0.89 vs 0.94 of the HBM peak), C5 1.030 vs 0.986, C3 6.46 vs 6.24.  (With one
// chunk load in flight per lane instead of four it was 2.22 ms: the warp-private staging is latency-bound.)  It therefore
// stays OFF by default (NTHASH_B200_PACK=1 selects it; tests/test_gpu_kmer.py runs both).  What it would take to win: input
// that needs no validity scan and no byte -> code step, i.e. 2-bit packed bases handed in by the caller.
// shared memory: [output tiles: one per warp, 1024-byte aligned][tables at PK_TAB][strips: one per warp]
constexpr int PK_PAIR_OFF = 0;     // (from PK_TAB) 16 x 8 B forward pair table, 16 x 8 B reverse pair table (256-byte aligned)
constexpr int PK_IN_OFF = 256;     // 4 x 16 B in-only entries (warm-up tail)
constexpr int PK_T4_OFF = 512;     // 256 x 16 B tetramer table, re-indexed by c0 | c1 << 2 | c2 << 4 | c3 << 6
constexpr int PK_TAB_BYTES = 512 + T4_BYTES;
constexpr uint32_t PK_MAX_CHUNKS = 512; // 16-byte chunks of ASCII a warp may stage (16 bad-chunk words)

// PACKED: the input is the caller's 2-bit packed sequence (4 bases per byte, A C G T = 0 1 2 3, base j of the staged slice at
// bits 2 (j & 3) of byte j >> 2) plus an optional invalid-base bitmap (bit j set: base j is an N), P.packed / P.inv_bits, with
// the batch's base 0 at slice position P.packed_first.  Chunks are then 64 bases (one 16-byte load); the conversion is a
// 2-bit -> nibble spread and a code remap (no SWAR validity scan, no byte -> code step), validity is the bitmap's.
template<int H, int WS, bool PACKED>
__global__ void __launch_bounds__(256)
kmer_pack_kernel(const __grid_constant__ KmerParams P, const __grid_constant__ CUtensorMap omap, const uint32_t strip_bytes)
{
  constexpr uint32_t CSH = PACKED ? 6 : 4;      // log2(bases per chunk)
  constexpr uint32_t LEAD = PACKED ? 32u : 16u; // nibbles of zero padding in front of a strip (base -1 of the first read lives there); keeps the chunk stores aligned
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t TILE = (uint32_t)(WS * H / 8) * 2048u; // blocks x 32 rows x 64 bytes
  const uint32_t NT = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tab_off = (NT / 32) * TILE;
  const uint32_t sbase = smem_u32(smem) + tab_off; // the tables' base: PK_* offsets count from here
  uint8_t* const tabs = smem + tab_off;
  const uint32_t k = P.k;
  // ---- CTA-wide tables (the only barrier of the kernel) ----
  {
    auto code2base = [](int c) { return c ^ (c >> 1); }; // code 0 = A, 1 = C, 2 = T/U, 3 = G -> index into P.s / P.sk (A, C, G, T)
    if (tid < 16) {
      const int ci = tid >> 2, co = tid & 3;
      reinterpret_cast<uint64_t*>(tabs + PK_PAIR_OFF)[tid] = P.s[code2base(ci)] ^ P.sk[code2base(co)];
      reinterpret_cast<uint64_t*>(tabs + PK_PAIR_OFF + 128)[tid] = P.sk[code2base(ci ^ 2)] ^ P.s[code2base(co ^ 2)];
    } else if (tid < 20) {
      const int ci = tid - 16;
      const uint64_t f = P.s[code2base(ci)], r = P.sk[code2base(ci ^ 2)];
      reinterpret_cast<uint4*>(tabs + PK_IN_OFF)[ci] = make_uint4((uint32_t)f, (uint32_t)(f >> 32), (uint32_t)r, (uint32_t)(r >> 32));
    }
    for (uint32_t j = tid; j < 256; j += NT) { // j = c0 | c1 << 2 | c2 << 4 | c3 << 6  ->  P.t4's c0 << 6 | c1 << 4 | c2 << 2 | c3
      const uint32_t src = ((j & 3u) << 6) | ((j & 12u) << 2) | ((j & 48u) >> 2) | (j >> 6);
      reinterpret_cast<uint4*>(tabs + PK_T4_OFF)[j] = __ldg(P.t4 + src);
    }
  }
  __syncthreads();

  // ---- this warp's 32 items ----
  const uint64_t it0 = (uint64_t)blockIdx.x * NT + warp * 32;
  if (it0 >= P.g.n_items) return;
  const uint32_t n = P.g.seg; // windows per item: the same for every item of a BOX batch
  auto item_byte = [&](uint64_t i) -> uint64_t { return P.g.flat ? flat_byte(P.g, i * n) : i * (uint64_t)P.g.read_len; };
  const uint64_t it_last = min(it0 + 31, P.g.n_items - 1);
  const uint64_t xo = PACKED ? P.packed_first : 0; // position of the batch's base 0 in the staged slice
  const uint64_t b_first = item_byte(it0) + xo;
  const uint64_t b_end = (P.g.flat ? flat_byte(P.g, it_last * n + n - 1) : it_last * (uint64_t)P.g.read_len + n - 1) + k + xo; // one past the last base
  const uint64_t a0 = (b_first ? b_first - 1 : 0) & ~(uint64_t)((1u << CSH) - 1); // chunk-aligned start of the staged range
  const uint32_t n_chunks = (uint32_t)((b_end - a0 + (1u << CSH) - 1) >> CSH);
  const uint32_t strip = sbase + PK_TAB_BYTES + warp * (strip_bytes + 64u); // [nibble strip][16 bad-chunk words]
  const uint32_t badw = strip + strip_bytes;
  const uint32_t tb0 = smem_u32(smem) + warp * TILE;
  if ((n_chunks << (CSH - 1)) + LEAD / 2 + 32u > strip_bytes || n_chunks > PK_MAX_CHUNKS) __trap(); // the launcher sized the strip for this geometry

  // ---- stage + convert: chunk c = bases a0 + (c << CSH) .. -> nibbles at strip + 8 + (c << (CSH - 1)) ----
  if (lane < LEAD / 8) asm volatile("st.shared.u32 [%0], %1;" ::"r"(strip + lane * 4), "r"(0u) : "memory"); // the lead-in: code 0 ('A'), cancels out
  // four chunks per lane in flight (the loads do not depend on each other: one round trip to DRAM per 2 KB of ASCII)
  auto load_chunk = [&](uint32_t c) {
    if constexpr (PACKED) {
      uint4 x = make_uint4(0u, 0u, 0u, 0u);
      if (c < n_chunks) x = __ldg(reinterpret_cast<const uint4*>(P.packed + (a0 >> 2)) + c); // the slice is allocated past its end
      return x;
    } else {
      uint4 x = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u); // 'A': hashable, never stored past n_chunks
      const uint64_t gb = a0 + 16ull * c;
      if (c < n_chunks) {
        if (gb + 16 <= P.n_bases) {
          x = __ldg(reinterpret_cast<const uint4*>(P.bases + gb));
        } else { // the last bytes of the buffer
          uint32_t w[4] = { 0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u };
          for (uint32_t j = 0; j < 16 && gb + j < P.n_bases; ++j) w[j >> 2] = (w[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | ((uint32_t)P.bases[gb + j] << (8 * (j & 3)));
          x = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      return x;
    }
  };
  auto convert_chunk = [&](uint32_t c0, const uint4 x) {
    const uint32_t c = c0 + lane;
    if constexpr (PACKED) {
      uint32_t anybad = 0;
      if (P.inv_bits && c < n_chunks) {
        const uint2 iv = __ldg(reinterpret_cast<const uint2*>(P.inv_bits + (a0 >> 5)) + c);
        anybad = iv.x | iv.y;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, anybad != 0);
      if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(badw + (c0 >> 5) * 4), "r"(m) : "memory");
      // eight 2-bit codes (16 bits) -> eight nibbles; then A C G T = 0 1 2 3 -> the kernels' (byte >> 1) & 3 order A C T G: p ^ (p >> 1)
      auto spread = [](uint32_t h16) {
        uint32_t v = (h16 | (h16 << 8)) & 0x00FF00FFu;
        v = (v | (v << 4)) & 0x0F0F0F0Fu;
        v = (v | (v << 2)) & 0x33333333u;
        return v ^ ((v >> 1) & 0x11111111u);
      };
      if (c < n_chunks) {
        const uint32_t dst = strip + LEAD / 2 + 32u * c;
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(spread(x.x & 0xFFFFu)), "r"(spread(x.x >> 16)), "r"(spread(x.y & 0xFFFFu)),
                     "r"(spread(x.y >> 16))
                     : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst + 16u), "r"(spread(x.z & 0xFFFFu)), "r"(spread(x.z >> 16)), "r"(spread(x.w & 0xFFFFu)),
                     "r"(spread(x.w >> 16))
                     : "memory");
      }
    } else {
      const uint32_t anybad = swar_bad(x.x) | swar_bad(x.y) | swar_bad(x.z) | swar_bad(x.w);
      const uint32_t m = __ballot_sync(0xffffffffu, anybad != 0);
      if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(badw + (c0 >> 5) * 4), "r"(m) : "memory");
      // per word: codes (x >> 1) & 3 per byte, then byte pairs folded into nibbles: z = y | y >> 4 has c0 | c1 << 4 in byte 0 and
      // c2 | c3 << 4 in byte 2; PRMT gathers bytes 0, 2 of two words
      auto nib = [](uint32_t v) {
        const uint32_t y = (v >> 1) & 0x03030303u;
        return y | (y >> 4);
      };
      const uint32_t lo = __byte_perm(nib(x.x), nib(x.y), 0x6420u), hi = __byte_perm(nib(x.z), nib(x.w), 0x6420u);
      if (c < n_chunks) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(strip + LEAD / 2 + 8u * c), "r"(lo), "r"(hi) : "memory");
    }
  };
  for (uint32_t c0 = 0; c0 < n_chunks; c0 += 128) {
    uint4 x[4];
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u) x[u] = load_chunk(c0 + 32 * u + lane);
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u)
      if (c0 + 32 * u < n_chunks) convert_chunk(c0 + 32 * u, x[u]);
  }
  __syncwarp();

  // ---- this lane's item ----
  const uint64_t item = it0 + lane;
  const bool active = item < P.g.n_items;
  const uint64_t my_byte = active ? item_byte(item) + xo : b_first; // idle lanes hash the first item again; the tensor store clips their rows
  const uint64_t my_out = item * (uint64_t)n;
  const uint32_t q0 = LEAD + (uint32_t)(my_byte - a0); // nibble index of the item's base 0
  // any flagged chunk among the row's bases -1 .. n+k-2?
  bool bad = false;
  {
    const uint32_t c_lo = (q0 - 1 - LEAD) >> CSH, c_hi = (q0 + n + k - 2 - LEAD) >> CSH; // q0 - 1 - LEAD may wrap for the slice's first base
    for (uint32_t c = (q0 > LEAD ? c_lo : 0u); c <= c_hi; ++c) bad |= (lds_u32(badw + (c >> 5) * 4) >> (c & 31)) & 1u;
  }
  const uint32_t pair = keep(sbase + PK_PAIR_OFF);
  const uint32_t t4a = keep(sbase + PK_T4_OFF);

  // nibble stream reader: aligned word j of the stream that starts at nibble position `pos`
  auto stream_word = [&](uint32_t pos, uint32_t j) {
    const uint32_t wa = strip + ((pos >> 3) + j) * 4;
    return __funnelshift_r(lds_u32(wa), lds_u32(wa + 4), (pos & 7u) * 4u);
  };

  // ---- warm-up: k in-only steps over bases -1 .. k-2, four per step through the tetramer table ----
  State s = { 0u, 0u, 0u, 0u };
  {
    const uint32_t nq = k >> 2;
    uint32_t wv = 0;
    for (uint32_t q = 0; q < nq; ++q) {
      if ((q & 1u) == 0) wv = stream_word(q0 - 1, q >> 1);
      const uint32_t v = (q & 1u) ? (wv >> 16) : (wv & 0xFFFFu);      // c0 | c1 << 4 | c2 << 8 | c3 << 12
      const uint32_t t = (v | (v >> 2)) & 0x0F0Fu;
      const uint32_t j = (t | (t >> 4)) & 0xFFu;                       // c0 | c1 << 2 | c2 << 4 | c3 << 6
      roll4_in(s, lds_v4(t4a + j * 16u));
    }
    for (uint32_t j = 4 * nq; j < k; ++j) {
      const uint32_t pos = q0 - 1 + j;
      const uint32_t c = (lds_u32(strip + (pos >> 3) * 4) >> ((pos & 7u) * 4u)) & 3u;
      roll_step(s, lds_v4(sbase + PK_IN_OFF + c * 16u), P.two);
    }
  }

  // ---- main loop ----
  const uint32_t lterm = keep(lane * 64 + (((lane >> 1) & 3) << 4));
  const uint32_t tb = tb0 + lterm;
  const int row0 = (int)it0;
  auto chunk_addr = [&](uint32_t cc) { return (tb ^ ((cc & 3u) << 4)) + (cc >> 2) * 2048u; };
  const uint32_t pos_in = q0 + k - 1, pos_out = q0 - 1;
  const uint32_t sh_in = (pos_in & 7u) * 4u, sh_out = (pos_out & 7u) * 4u;
  uint32_t wa_in = strip + (pos_in >> 3) * 4, wa_out = strip + (pos_out >> 3) * 4;
  uint32_t in0 = lds_u32(wa_in), out0 = lds_u32(wa_out);
  // eight windows: one aligned word of each stream -> nibble i = code_in << 2 | code_out = pair-table index of window i
  auto next8 = [&]() {
    wa_in += 4;
    wa_out += 4;
    const uint32_t in1 = lds_u32(wa_in), out1 = lds_u32(wa_out);
    const uint32_t wi = __funnelshift_r(in0, in1, sh_in), wo = __funnelshift_r(out0, out1, sh_out);
    in0 = in1;
    out0 = out1;
    return wi * 4u + wo;
  };
  auto window = [&](uint32_t cw, int i) -> uint64_t { // window i (0..7) of the index word
    const uint32_t ea = lop3<LUT_OR_AND>(pair, __funnelshift_r(cw, cw, (4 * i + 29) & 31), 0x78u); // pair | (idx << 3)
    const uint2 ef = lds_v2(ea), er = lds_v2(ea + 128u);
    roll_step(s, make_uint4(ef.x, ef.y, er.x, er.y), P.two);
    return canonical2(s);
  };
  auto store4 = [&](uint32_t g, const uint64_t (&hv)[4]) { // group g of four windows of the tile -> shared memory
    if (H == 1) {
      st_shared_v2_u64(chunk_addr(2 * g), hv[0], hv[1]);
      st_shared_v2_u64(chunk_addr(2 * g + 1), hv[2], hv[3]);
    } else if (H == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) st_shared_v2_u64(chunk_addr(4 * g + i), hv[i], ext_hash(hv[i], P.mult[1]));
    } else if (H == 3) {
      uint64_t v[12];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[3 * i] = hv[i];
        v[3 * i + 1] = ext_hash(hv[i], P.mult[1]);
        v[3 * i + 2] = ext_hash(hv[i], P.mult[2]);
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) st_shared_v2_u64(chunk_addr(6 * g + c), v[2 * c], v[2 * c + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        st_shared_v2_u64(chunk_addr(8 * g + 2 * i), hv[i], ext_hash(hv[i], P.mult[1]));
        st_shared_v2_u64(chunk_addr(8 * g + 2 * i + 1), ext_hash(hv[i], P.mult[2]), ext_hash(hv[i], P.mult[3]));
      }
    }
  };
  static_assert(WS % 8 == 0, "tiles are whole index words");
  for (uint32_t p = 0; p < n; p += WS) {
    if (p) { // the previous tile must have left shared memory
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
    }
    if (p + WS <= n) {
#pragma unroll
      for (uint32_t w8 = 0; w8 < (uint32_t)WS / 8; ++w8) {
        const uint32_t cw = next8();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint64_t hv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) hv[i] = window(cw, 4 * half + i);
          store4(2 * w8 + half, hv);
        }
      }
    } else { // the last, shorter tile of a row (kept rolled up); windows past the row's end land in the tile but are clipped
#pragma unroll 1
      for (uint32_t w8 = 0; 8 * w8 < n - p; ++w8) {
        const uint32_t cw = next8();
#pragma unroll 1
        for (uint32_t half = 0; half < 2; ++half) {
          uint64_t hv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) hv[i] = window(half ? cw >> 16 : cw, i);
          store4(2 * w8 + half, hv);
        }
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(&omap, tb0, 0, row0, (int)(p * H / 8));
      bulk_commit();
    }
  }
  const bool dirty = active && bad;
  const bool any_dirty = __any_sync(0xffffffffu, dirty);
  if (lane == 0) {
    if (any_dirty) bulk_wait_all0(); // zeros must land after the tile they overwrite
    else bulk_wait_read0();          // shared memory must outlive the TMA read
  }
  __syncwarp();
  if (dirty) { // exact clean-up from the ASCII bytes: windows touching a non-ACGTU byte are not emitted (kmer.cpp:232-235, :255-258)
    const uint32_t n_own = P.g.flat ? (uint32_t)min((uint64_t)n, (uint64_t)P.g.nk - my_out % P.g.nk) : n;
    uint32_t run = 0;
    for (uint32_t j = 0; j < n_own + k - 1; ++j) {
      bool ok;
      if constexpr (PACKED) ok = !((P.inv_bits[(my_byte + j) >> 5] >> ((my_byte + j) & 31)) & 1u);
      else ok = is_acgtu(P.bases[my_byte + j]);
      run = ok ? run + 1 : 0;
      if (j >= k - 1 && run < k) {
        const uint64_t w = my_out + (j - (k - 1));
        for (uint32_t q = 0; q < (uint32_t)H; ++q) P.out[w * H + q] = 0;
        if (P.valid_bits) atomicAnd(&P.valid_bits[(P.valid_row0 + w) >> 5], ~(1u << ((P.valid_row0 + w) & 31)));
      }
    }
  }
}

// FLAT geometry, second launch: the rows the main kernel cannot produce.  Item f >= 1 = the windows of read f that
// share a flat item with the end of read f-1 (the main kernel rolled across the boundary there and stored junk);
// item 0 = the partial last item of the batch, which the tensor map leaves out.  At most seg windows each, FIX_LANES
// lanes per item: every lane hashes its own run of windows from scratch straight out of global memory, the way the main
// kernel does it — k in-only steps (four bases at a time through the tetramer table; base_forward_hash /
// base_reverse_hash, src/kmer.cpp:43-73, :123-152) starting one base early, then NtHash::roll (src/kmer.cpp:246-264).
constexpr uint32_t FIX_LANES = 32; // a warp per item: its bytes are staged once (coalesced), every lane hashes n / 32 windows
__global__ void __launch_bounds__(128) kmer_flat_fix_kernel(const __grid_constant__ KmerParams P, uint64_t n_reads, uint32_t warp_bytes)
{
  // [tetramer table 4 KB][seeds: S[4], srol^k S[4]][per warp: the item's bytes]; everything a step touches is in shared
  // memory — the kernel is a few thousand short dependent chains, so its duration is the latency of one chain
  extern __shared__ __align__(16) uint8_t fix_smem[];
  uint4* t4 = reinterpret_cast<uint4*>(fix_smem);
  uint64_t* seeds = reinterpret_cast<uint64_t*>(fix_smem + T4_BYTES);
  for (uint32_t j = threadIdx.x; j < 256; j += blockDim.x) t4[j] = __ldg(P.t4 + j);
  if (threadIdx.x < 4) seeds[threadIdx.x] = P.s[threadIdx.x];
  else if (threadIdx.x < 8) seeds[threadIdx.x] = P.sk[threadIdx.x - 4];
  __syncthreads();
  const uint32_t wpb = blockDim.x / FIX_LANES, warp = threadIdx.x / FIX_LANES, sub = threadIdx.x % FIX_LANES;
  const uint64_t f = (uint64_t)blockIdx.x * wpb + warp;
  if (f >= n_reads) return; // whole warps leave together
  const uint64_t nk = P.g.nk, seg = P.g.seg, w_tail = P.g.total / seg * seg;
  uint64_t w0;
  uint32_t n;
  if (f == 0) {
    w0 = w_tail;
    n = (uint32_t)(P.g.total - w_tail);
  } else {
    w0 = f * nk;
    const uint32_t rem = (uint32_t)(w0 % seg);
    n = (rem && w0 < w_tail) ? (uint32_t)min(seg - rem, nk) : 0u;
  }
  if (n == 0) return;
  const uint32_t k = P.k, h = P.h;
  // the item's bytes -1 .. n+k-2 (byte -1 exists: f >= 1 items start a read that is not the first, and the batch's tail
  // item never starts at base 0 of the batch because flat batches hold at least two full items)
  const uint64_t r = w0 / nk;
  const uint8_t* g0 = P.bases + r * P.g.read_len + (w0 - r * nk) - 1;
  uint8_t* sb = fix_smem + T4_BYTES + 64 + (size_t)warp * warp_bytes;
  for (uint32_t j = sub; j < n + k; j += FIX_LANES) sb[j] = g0[j];
  __syncwarp();
  const uint32_t per = (n + FIX_LANES - 1) / FIX_LANES, p_lo = min(n, sub * per), p_hi = min(n, p_lo + per);
  if (p_lo >= p_hi) return;
  const uint8_t* sq = sb + 1 + p_lo; // first base of this lane's first window; sq[-1] is staged
  w0 += p_lo;
  n = p_hi - p_lo;
  // 2-bit code (byte >> 1) & 3: 0 = A, 1 = C, 2 = T/U, 3 = G; index into S / srol^k S (A, C, G, T) = code ^ (code >> 1)
  auto sidx = [](uint32_t c) { const uint32_t x = (c >> 1) & 3u; return x ^ (x >> 1); };
  State st = { 0u, 0u, 0u, 0u };
  uint32_t run = 0; // hashable bases in a row, ending at the newest base
  const uint32_t nq = k >> 2;
  for (uint32_t q = 0; q < nq; ++q) { // bases 4q-1 .. 4q+2
    const uint32_t c0 = sq[(int)(4 * q) - 1], c1 = sq[4 * q], c2 = sq[4 * q + 1], c3 = sq[4 * q + 2];
    const uint32_t idx = ((c0 & 6u) << 5) | ((c1 & 6u) << 3) | ((c2 & 6u) << 1) | ((c3 & 6u) >> 1);
    roll4_in(st, t4[idx]);
    // validity of bases 4q .. 4q+3 (base -1 does not count): SWAR test on the four bytes at once
    if (4 * q + 3 < k - 1) {
      const uint32_t bw = swar_bad(c1 | (c2 << 8) | (c3 << 16) | ((uint32_t)sq[4 * q + 3] << 24));
      if (bw == 0) run += 4;
      else
        for (uint32_t j = 0; j < 4; ++j) run = ((bw >> (8 * j)) & 0xFFu) ? 0 : run + 1;
    }
  }
  for (uint32_t j = (k - 1) / 4 * 4; j + 1 < k; ++j) run = is_acgtu(sq[j]) ? run + 1 : 0; // the bases the whole words above left out
  for (uint32_t j = 4 * nq; j < k; ++j) { // bases j-1
    const uint32_t c = sq[j - 1];
    const uint64_t fi = seeds[sidx(c)], ri = seeds[4 + sidx(c ^ 4u)]; // c ^ 4 flips code bit 1: the complement
    roll_step(st, make_uint4((uint32_t)fi, (uint32_t)(fi >> 32), (uint32_t)ri, (uint32_t)(ri >> 32)), P.two);
  }
  // Launched with programmatic stream serialisation (launch_flat_fix): everything above — tables, the item's bytes, the k-base
  // warm-up — only reads the input and runs while the main kernel's last CTAs drain; the rows it overwrites must have
  // landed first, so the first store waits for the main kernel (a no-op when launched the ordinary way).  (Keeping the
  // lane's windows in registers across the wait as well was slower: 0.944 vs 0.936 ms on C5.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (uint32_t p = 0; p < n; ++p) {
    const uint32_t cin = sq[p + k - 1], cout = sq[(int)p - 1];
    run = is_acgtu(cin) ? run + 1 : 0;
    const uint64_t fe = seeds[sidx(cin)] ^ seeds[4 + sidx(cout)], re = seeds[4 + sidx(cin ^ 4u)] ^ seeds[sidx(cout ^ 4u)];
    roll_step(st, make_uint4((uint32_t)fe, (uint32_t)(fe >> 32), (uint32_t)re, (uint32_t)(re >> 32)), P.two);
    const uint64_t w = w0 + p, h0 = canonical2(st);
    uint64_t* o = P.out + w * h;
    if (run >= k) {
      o[0] = h0;
      for (uint32_t q = 1; q < h; ++q) o[q] = ext_hash(h0, P.mult[q]);
    } else { // a window the reference does not visit (kmer.cpp:232-235, :255-258)
      for (uint32_t q = 0; q < h; ++q) o[q] = 0;
      if (P.valid_bits) atomicAnd(&P.valid_bits[(P.valid_row0 + w) >> 5], ~(1u << ((P.valid_row0 + w) & 31)));
    }
  }
}

uint32_t fast_smem_bytes(uint32_t tile_cap, uint32_t buf_bytes)
{
  return F_TILE_OFF + F_TILE_PAD + tile_cap + 16 + 1024 + (buf_bytes > (uint32_t)T4_BYTES ? buf_bytes : (uint32_t)T4_BYTES);
}

// 256 x {TF, TR}: combined in-only contribution of four consecutive bases c0..c3 (codes, c0 first):
//   TF = srol^3 S[c0] ^ srol^2 S[c1] ^ srol S[c2] ^ S[c3]
//   TR = Skc[c0] ^ srol Skc[c1] ^ srol^2 Skc[c2] ^ srol^3 Skc[c3],  Skc[c] = srol^k S[complement c]
// One small device buffer per (device, k), created on first use and kept for the life of the process.
cudaError_t get_t4_table_impl(uint32_t k, const uint4** out)
{
  static std::mutex mu;
  static std::map<std::pair<int, uint32_t>, uint4*> cache;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find({ dev, k });
  if (it != cache.end()) {
    *out = it->second;
    return cudaSuccess;
  }
  const uint64_t seed_of_code[4] = { SEED_A, SEED_C, SEED_T, SEED_G };
  std::vector<uint4> tab(256);
  for (unsigned idx = 0; idx < 256; ++idx) {
    uint64_t tf = 0, tr = 0;
    for (unsigned j = 0; j < 4; ++j) {
      const unsigned c = (idx >> (6 - 2 * j)) & 3;
      tf ^= srol_n(seed_of_code[c], 3 - j);
      tr ^= srol_n(seed_of_code[c ^ 2], k + j);
    }
    tab[idx] = make_uint4((uint32_t)tf, (uint32_t)(tf >> 32), (uint32_t)tr, (uint32_t)(tr >> 32));
  }
  uint4* d = nullptr;
  e = cudaMalloc(&d, T4_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaMemcpy(d, tab.data(), T4_BYTES, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(d);
    return e;
  }
  cache[{ dev, k }] = d;
  *out = d;
  return cudaSuccess;
}

struct FastCfg
{
  uint32_t nt, ws, nbuf;
  bool box, direct;
};

// Tensor map of the output seen as [rows*H/8 blocks][n_items rows][8 u64] (blocks outermost), boxes of
// `blocks` x 32 rows x 8 u64 under the 64-byte swizzle.
cudaError_t make_out_map(const KmerParams& P, uint32_t blocks, CUtensorMap* map)
{
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static const EncodeFn encode = []() -> EncodeFn { // initialised once, thread-safe (function-local static)
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) != cudaSuccess) return nullptr;
    return (fp && qr == cudaDriverEntryPointSuccess) ? (EncodeFn)fp : nullptr;
  }();
  if (!encode) return cudaErrorNotSupported;
  const uint64_t row_u64 = (uint64_t)P.g.seg * P.h;
  const cuuint64_t dims[3] = { 8, P.g.n_items, row_u64 / 8 }; // flat: n_items = the FULL items only, so nothing lands past `out`
  const cuuint64_t strides[2] = { row_u64 * 8, 64 };
  const cuuint32_t box[3] = { 8, 32, blocks };
  const cuuint32_t estr[3] = { 1, 1, 1 };
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, P.out, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template<int H, int CONS, int WS, int NBUF, bool BOX>
cudaError_t launch_fast_t(const KmerParams& P, uint32_t nt, cudaStream_t st)
{
  constexpr bool REDUCE = CONS != 0 && CONS != 4; // consumers allocate no output buffers
  auto fn = kmer_fast_kernel<H, CONS, WS, NBUF, BOX>;
  const uint32_t buf = BOX ? (nt / 32) * NBUF * (uint32_t)(WS * H / 8) * 2048u
                       : (NBUF == 2 && H != 0) ? (uint32_t)T4_BYTES + 256u + nt * 32u * (CONS == 4 ? 3u : 1u) // DIRECT: table + parked row heads
                                               : nt * (CONS == 4 ? 3u : 1u) * (ROW1_BYTES + 16);
  uint32_t smem_bytes = fast_smem_bytes(P.tile_cap, REDUCE ? 0u : buf);
  if (const char* e = getenv("NTHASH_B200_SMEM_PAD")) smem_bytes += (uint32_t)atoi(e); // experiments: lower the occupancy
  if (smem_bytes > 227u * 1024u) return cudaErrorInvalidConfiguration;
  CUtensorMap map;
  memset(&map, 0, sizeof map);
  cudaError_t e = (BOX && !REDUCE) ? make_out_map(P, (uint32_t)(WS * H / 8), &map) : cudaSuccess;
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (P.g.n_items + nt - 1) / nt;
  if (ctas > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  fn<<<(unsigned)ctas, nt, smem_bytes, st>>>(P, map);
  return cudaGetLastError();
}

template<int H, int WS>
cudaError_t launch_fast_nbuf(const KmerParams& P, const FastCfg& c, cudaStream_t st)
{
  return c.nbuf >= 2 ? launch_fast_t<H, 0, WS, 2, true>(P, c.nt, st) : launch_fast_t<H, 0, WS, 1, true>(P, c.nt, st);
}

// windows per tensor store: index 0/1/2 = 192-256 / 256-384 / 320-512 bytes per row
template<int H>
constexpr int fast_ws(int idx)
{
  return H == 1   ? (idx == 0 ? 24 : idx == 1 ? 32 : 40)
         : H == 2 ? (idx == 0 ? 12 : idx == 1 ? 16 : 20)
         : H == 3 ? (idx == 0 ? 8 : idx == 1 ? 16 : 24)
                  : (idx == 0 ? 8 : idx == 1 ? 12 : 16);
}

int fast_ws_rt(uint32_t h, uint32_t idx)
{
  return h == 1 ? fast_ws<1>((int)idx) : h == 2 ? fast_ws<2>((int)idx) : h == 3 ? fast_ws<3>((int)idx) : fast_ws<4>((int)idx);
}

template<int H>
cudaError_t launch_fast_h(const KmerParams& P, const FastCfg& c, cudaStream_t st)
{
  if (P.reduce_out) return launch_fast_t<H, 1, fast_ws<H>(0), 1, false>(P, c.nt, st);
  if (P.out_fwd) // hashes + strand hashes
    return c.direct ? launch_fast_t<H, 4, fast_ws<H>(0), 2, false>(P, c.nt, st) : launch_fast_t<H, 4, fast_ws<H>(0), 1, false>(P, c.nt, st);
  if (!c.box) // general output path: WS is unused there, NBUF == 2 selects the DIRECT form
    return c.direct ? launch_fast_t<H, 0, fast_ws<H>(0), 2, false>(P, c.nt, st) : launch_fast_t<H, 0, fast_ws<H>(0), 1, false>(P, c.nt, st);
  switch (c.ws) {
    case 0: return launch_fast_nbuf<H, fast_ws<H>(0)>(P, c, st);
    case 1: return launch_fast_nbuf<H, fast_ws<H>(1)>(P, c, st);
    default: return launch_fast_nbuf<H, fast_ws<H>(2)>(P, c, st);
  }
}

// ---- kmer_pack_kernel launch: strip size from the geometry, CTA size by resident warps ----
template<int H>
constexpr int pack_ws()
{
  return H == 1 ? 24 : H == 2 ? 16 : 8; // windows per tensor store: 192 / 256 / 192 / 256 bytes per row, whole index words
}

// 16-byte chunks of ASCII one warp (32 consecutive items) stages at most
uint64_t pack_max_chunks(const KmerGeom& g, uint32_t k, bool packed)
{
  const uint64_t span = g.flat ? 32ull * g.seg + (32ull * g.seg / g.nk + 2) * (k - 1) : 31ull * g.read_len + g.seg + k - 1;
  const uint64_t cb = packed ? 64 : 16;  // bases per chunk
  return (span + 1 + 2 * (cb - 1)) / cb + 1; // base -1, the aligned start, the rounded-up end
}

template<int H, bool PACKED>
cudaError_t launch_pack_t(const KmerParams& P, uint32_t nt_env, cudaStream_t st)
{
  constexpr int WS = pack_ws<H>();
  constexpr uint32_t TILE = (uint32_t)(WS * H / 8) * 2048u;
  const uint32_t strip_bytes = (uint32_t)pack_max_chunks(P.g, P.k, PACKED) * (PACKED ? 32u : 8u) + 64u;
  auto smem_for = [&](uint32_t nt) { return (nt / 32) * (TILE + strip_bytes + 64u) + (uint32_t)PK_TAB_BYTES; };
  // the CTA size that leaves the most warps resident (shared memory: 227 KB per SM, 1 KB reserved per CTA; registers: <= 32
  // warps at the kernel's ~64 registers per thread); larger CTAs share the tables among more warps and win ties
  uint32_t nt = 0, best = 0;
  for (uint32_t cand : { 256u, 192u, 160u, 128u, 96u, 64u }) {
    const uint32_t sm = smem_for(cand) + 1024u;
    if (sm > 227u * 1024u) continue;
    const uint32_t warps = std::min(32u / (cand / 32) * (cand / 32), (227u * 1024u / sm) * (cand / 32));
    if (warps > best) {
      best = warps;
      nt = cand;
    }
  }
  if (nt_env) nt = nt_env;
  if (!nt || nt % 32 || nt > 256 || smem_for(nt) > 227u * 1024u) return cudaErrorInvalidConfiguration;
  CUtensorMap map;
  memset(&map, 0, sizeof map);
  cudaError_t e = make_out_map(P, (uint32_t)(WS * H / 8), &map);
  if (e != cudaSuccess) return e;
  auto fn = kmer_pack_kernel<H, WS, PACKED>;
  const uint32_t smem_bytes = smem_for(nt);
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (P.g.n_items + nt - 1) / nt;
  if (ctas > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  fn<<<(unsigned)ctas, nt, smem_bytes, st>>>(P, map, strip_bytes);
  return cudaGetLastError();
}

uint32_t env_u32(const char* name, uint32_t dflt)
{
  const char* e = getenv(name);
  return e ? (uint32_t)atoi(e) : dflt;
}

// FLAT geometry: rows behind every read boundary + the partial last item, after the main kernel's junk has landed
cudaError_t launch_flat_fix(const KmerParams& P, cudaStream_t st)
{
  const uint64_t n_reads = P.g.total / P.g.nk;
  const uint32_t warp_bytes = (P.g.seg + P.k + 16 + 15) & ~15u; // bytes -1 .. n+k-2 of one item, n <= seg
  const uint32_t fixed = (uint32_t)T4_BYTES + 64u;              // tetramer table + seeds
  if (warp_bytes > 200u * 1024u) return cudaErrorInvalidConfiguration;
  const uint32_t wpb = std::max(1u, std::min(4u, (200u * 1024u) / warp_bytes));
  const uint32_t smem = fixed + wpb * warp_bytes;
  if (smem > 48u * 1024u) {
    const cudaError_t e = cudaFuncSetAttribute(kmer_flat_fix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  // programmatic dependent launch: the fix-up's CTAs may start once every CTA of the main kernel has started (it signals
  // griddepcontrol.launch_dependents first thing) and take the slots its last wave frees; they wait before their first store
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((n_reads + wpb - 1) / wpb));
  cfg.blockDim = dim3(wpb * FIX_LANES);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = getenv("NTHASH_B200_NO_PDL") ? 0 : 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kmer_flat_fix_kernel, P, n_reads, warp_bytes);
}

} // namespace

cudaError_t get_t4_table(uint32_t k, const uint4** out) { return get_t4_table_impl(k, out); }

bool kmer_packed_direct_ok(uint64_t n_reads, uint32_t read_len, uint32_t k, uint32_t h)
{
  if (n_reads == 0 || read_len < k || h < 1 || h > 4 || getenv("NTHASH_B200_NO_PACKED_DIRECT") || getenv("NTHASH_B200_FAST_NO_BOX") ||
      getenv("NTHASH_B200_DISABLE_TMA_STORE"))
    return false;
  if (read_len > env_u32("NTHASH_B200_FAST_WHOLE_READ", 400)) return false; // longer reads are cut / flat: ASCII kernels
  KmerGeom g;
  g.read_len = read_len;
  g.nk = read_len - k + 1;
  g.seg = g.nk;
  g.segs = 1;
  return ((uint64_t)g.nk * h) % 8 == 0 && pack_max_chunks(g, k, true) <= PK_MAX_CHUNKS;
}

bool kmer_fast_ok(const KmerParams& P)
{
  const KmerGeom& g = P.g;
  if (P.out_fwd && (P.reduce_out || P.bloom_mode || (((uintptr_t)P.out_fwd | (uintptr_t)P.out_rev) & 31))) return false;
  return (P.h >= 1 && P.h <= 255) && g.n_items > 0 &&
         (P.reduce_out || ((uintptr_t)P.out & 31) == 0) && (g.item_byte || (g.seg && g.segs));
}

// Launch configuration.  Uniform batches get their own item geometry here (reads longer than FAST_WHOLE_READ
// bases are cut into ~FAST_SEG-window items, the last one short); ragged batches keep the caller's item arrays.
cudaError_t launch_kmer_fast(const KmerParams& Pin, cudaStream_t st)
{
  KmerParams P = Pin;
  FastCfg c;
  c.box = false;
  c.direct = false;
  c.nbuf = env_u32("NTHASH_B200_FAST_NBUF", 1);
  c.ws = env_u32("NTHASH_B200_FAST_WS", 0);
  const bool uniform = !P.g.item_byte;
  if (uniform) {
    KmerGeom& g = P.g;
    const uint64_t n_reads = g.n_items / g.segs;
    const uint32_t whole = env_u32("NTHASH_B200_FAST_WHOLE_READ", 400), seg_t = env_u32("NTHASH_B200_FAST_SEG", 240);
    // windows per flat item: a multiple of 8 (rows of whole 64-byte blocks for every h), 168 by default — the round-2 sweep
    // (profiles/r02_c5_flatseg_sweep.txt: 120 0.816, 144 0.851, 168 0.864, 216 0.837, 264 0.840, 360 0.812 of the HBM peak on
    // C5; 192 = 48 words between lanes, 16-way LDS conflicts: 0.636): shorter items leave shared memory for more resident
    // warps, longer ones amortise the k-base warm-up
    // h <= 2 (second sweep, profiles/r02_c5_flat_sweep2.txt, after the lookups were pipelined): 200-window items = 5 stores of 40
    // windows (320-byte pieces, 50 words between lanes' rows) in CTAs of 96 threads: C5 0.979 -> 0.956 ms on the same box
    const bool flat_long_pieces = P.h <= 2 && !getenv("NTHASH_B200_FAST_WS") && !getenv("NTHASH_B200_FLAT_SEG");
    uint32_t flat_seg = env_u32("NTHASH_B200_FLAT_SEG", flat_long_pieces ? 200 : 168);
    flat_seg = std::max(24u, flat_seg / 8u * 8u);
    const bool flat_ok = !P.reduce_out && !P.out_fwd && !P.bloom_mode && P.h <= 4 && g.nk >= 2 * flat_seg &&
                         (n_reads * (uint64_t)g.nk) / flat_seg > 0 && !getenv("NTHASH_B200_FAST_NO_BOX") && !getenv("NTHASH_B200_NO_FLAT");
    if (g.read_len <= whole) { // one item per read
      g.seg = g.nk;
      g.segs = 1;
      // tensor stores need rows that are whole 64-byte blocks (then every row is 64-byte aligned too)
      c.box = !P.reduce_out && !P.out_fwd && P.h <= 4 && ((uint64_t)g.nk * P.h) % 8 == 0 && !getenv("NTHASH_B200_FAST_NO_BOX");
    } else if (flat_ok) {
      // FLAT: item i = dense windows [i*seg, (i+1)*seg) of the whole batch.  Every row of the [items][seg] view is full
      // and seg*h*8-byte pitched, so long reads leave through the same 3-D tensor stores as short ones.  seg = 168 =
      // 7 tiles of 24 windows (12 x h=2, 8 x h=3/4) and 42 words between neighbouring lanes' rows (2-way LDS.32
      // conflicts; 48 or 64 words would be 16- or 32-way).  Rows past a read boundary are junk and are redone by the fix-up kernel.
      g.flat = 1;
      g.total = n_reads * g.nk;
      g.seg = flat_seg;
      g.segs = 1;
      c.box = true;
      if (flat_long_pieces) c.ws = 2;
    } else {
      // balanced items, the last one shorter.  seg = 4 (mod 8): an odd number of 32-bit words between the rows of
      // neighbouring lanes keeps their LDS.32 on distinct banks (seg = 256 ran 1.7x slower than 244), and a
      // multiple of 4 windows keeps every item's output on a 32-byte boundary when the read's is.
      const uint32_t segs = (g.nk + seg_t - 1) / seg_t, x = (g.nk + segs - 1) / segs;
      g.seg = (x & ~7u) + ((x & 7u) <= 4 ? 4u : 12u);
      g.segs = (g.nk + g.seg - 1) / g.seg;
    }
    g.n_items = g.flat ? g.total / g.seg : n_reads * g.segs; // flat: full items only; the tail belongs to the fix-up kernel
  }
  // General output path, two forms (profiles/r02_ragged_*.txt): DIRECT (32-byte stores from registers, row-boundary sectors
  // merged through shared memory; 32-48 registers, almost no shared memory) wins wherever items are short or ragged —
  // ragged 100-150 bp reads 0.57 -> 0.72 of the HBM peak, 36-150 bp 0.45 -> 0.64, ragged long reads 0.65 -> 0.80 — while
  // uniform reads cut into long items keep the shared-memory rows (0.76 vs 0.67).
  c.direct = env_u32("NTHASH_B200_FAST_DIRECT", (!uniform || P.g.segs == 1) ? 1u : 0u) != 0 && !c.box && P.h <= 4 && !P.reduce_out && !P.bloom_mode;
  auto tile_cap_for = [&](uint32_t nt) -> uint64_t {
    // measured at plan time (KmerParams::tile_cap_256): 10 M reads of 100-150 bp go from four to five resident CTAs, 0.77 -> 0.81 of
    // the HBM peak; with two hashes per window the extra CTA is a loss (0.82 -> 0.78: more stores in flight, see the C2 probe), so
    // h = 1 only (profiles/r02_ab_exact_tile.txt)
    if (!uniform && nt == 256 && Pin.tile_cap_256 && P.h == 1 && !P.out_fwd) return (uint64_t)Pin.tile_cap_256 + 2 * (uint64_t)P.k + 64;
    if (!uniform) return ((uint64_t)Pin.tile_cap * nt + KMER_NT - 1) / KMER_NT + 2 * (uint64_t)P.k + 64; // sized for KMER_NT items
    if (P.g.flat) return (uint64_t)nt * P.g.seg + ((uint64_t)nt * P.g.seg / P.g.nk + 2) * (P.k - 1) + 64;
    if (P.g.segs == 1) return (uint64_t)nt * P.g.read_len + 64;
    return (uint64_t)nt * P.g.seg + ((uint64_t)nt / P.g.segs + 2) * (P.k - 1) + 64;
  };
  // defaults from profiles/sweeps/ (round 1): what matters most is resident warps per SM (small pieces win over
  // long ones that cost occupancy), then CTA size (fewer prologues): the largest CTA that leaves >= 16 warps resident
  const uint32_t buf_per_warp = P.reduce_out ? 0u
                                : c.box    ? c.nbuf * (uint32_t)(fast_ws_rt(P.h, c.ws) * P.h / 8) * 2048u
                                : c.direct ? 32u * 32u * (P.out_fwd ? 3u : 1u) // DIRECT: parked row heads only
                                           : 32u * (P.out_fwd ? 3u : 1u) * (ROW1_BYTES + 16);
  c.nt = 32;
  for (uint32_t nt : { 256u, 192u, 128u, 96u, 64u, 32u }) { // the small ones only matter for huge k
    const uint64_t cap = tile_cap_for(nt);
    if (cap > 227u * 1024u) continue;
    const uint32_t smem = fast_smem_bytes((uint32_t)cap, (nt / 32) * buf_per_warp + (c.direct ? (uint32_t)T4_BYTES + 256u : 0u)) + 1024;
    c.nt = nt;
    if (smem <= 227u * 1024u && ((227u * 1024u / smem) * (nt / 32) >= 16 || nt <= 96)) break;
  }
  if (P.g.flat && c.ws == 2 && !getenv("NTHASH_B200_FAST_WS") && tile_cap_for(96) <= 227u * 1024u) c.nt = 96;
  c.nt = env_u32("NTHASH_B200_FAST_NT", c.nt);
  if (c.nt != 256) P.g.item_perm = nullptr; // the precomputed deal is per block of 256 items
  if (c.nt < 32 || c.nt > 256 || c.nt % 32) return cudaErrorInvalidValue;
  if (tile_cap_for(c.nt) > 227u * 1024u) return cudaErrorInvalidConfiguration;
  P.tile_cap = (uint32_t)tile_cap_for(c.nt);
  cudaError_t e = get_t4_table_impl(P.k, &P.t4);
  if (e != cudaSuccess) return e;
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t smem = fast_smem_bytes(P.tile_cap, (c.nt / 32) * buf_per_warp + (c.direct ? (uint32_t)T4_BYTES + 256u : 0u)) + 1024;
    const uint32_t resident = std::max(1u, 227u * 1024u / smem) * (uint32_t)sms;
    // measured (profiles/r01_prefetch_sweep.txt): the prefetch costs 1-10 % on every config, so it stays off
    P.prefetch_ctas = env_u32("NTHASH_B200_PREFETCH_CTAS", 0);
    if (P.prefetch_ctas == 1) P.prefetch_ctas = resident;
  }
  // BOX batches (whole-read items with 64-byte-multiple rows, or flat items), opt-in (NTHASH_B200_PACK=1; measured slower, see
  // kmer_pack_kernel): the warp-private nibble-strip kernel, when a warp's bytes fit its strip (reads up to ~250 bases, any
  // flat item); kmer_fast_kernel keeps everything else
  if (P.packed) { // 2-bit packed input: only kmer_pack_kernel reads it (whole-read items; the flat fix-up kernel wants ASCII)
    if (!c.box || P.g.flat || P.reduce_out || P.out_fwd || P.bloom_mode || P.h > 4 || pack_max_chunks(P.g, P.k, true) > PK_MAX_CHUNKS)
      return cudaErrorNotSupported;
    const uint32_t nt_env = env_u32("NTHASH_B200_PACK_NT", 0);
    return P.h == 1 ? launch_pack_t<1, true>(P, nt_env, st) : P.h == 2 ? launch_pack_t<2, true>(P, nt_env, st)
         : P.h == 3 ? launch_pack_t<3, true>(P, nt_env, st) : launch_pack_t<4, true>(P, nt_env, st);
  }
  if (c.box && !P.reduce_out && !P.out_fwd && !P.bloom_mode && P.h <= 4 && pack_max_chunks(P.g, P.k, false) <= PK_MAX_CHUNKS &&
      env_u32("NTHASH_B200_PACK", 0) != 0) {
    const uint32_t nt_env = env_u32("NTHASH_B200_PACK_NT", 0);
    e = P.h == 1 ? launch_pack_t<1, false>(P, nt_env, st) : P.h == 2 ? launch_pack_t<2, false>(P, nt_env, st)
      : P.h == 3 ? launch_pack_t<3, false>(P, nt_env, st) : launch_pack_t<4, false>(P, nt_env, st);
    if (e != cudaErrorInvalidConfiguration) {
      if (e == cudaSuccess && P.g.flat) e = launch_flat_fix(P, st);
      return e;
    }
    cudaGetLastError();
  }
  if (P.bloom_mode) { // runtime number of hashes; mult[0] carries k * MULTISEED
    P.mult[0] = (uint64_t)P.k * MULTISEED;
    return P.bloom_mode == 1   ? launch_fast_t<1, 2, fast_ws<1>(0), 1, false>(P, c.nt, st)
           : P.bloom_mode == 2 ? launch_fast_t<1, 3, fast_ws<1>(0), 1, false>(P, c.nt, st)
                               : launch_fast_t<1, 5, fast_ws<1>(0), 1, false>(P, c.nt, st); // 3: cardinality sketch
  }
  switch (P.h) {
    case 1: e = launch_fast_h<1>(P, c, st); break;
    case 2: e = launch_fast_h<2>(P, c, st); break;
    case 3: e = launch_fast_h<3>(P, c, st); break;
    case 4: e = launch_fast_h<4>(P, c, st); break;
    default: // 5..255 hashes: runtime count, general output path (a stream of 256-byte pieces per lane)
      return P.reduce_out ? launch_fast_t<0, 1, 8, 1, false>(P, c.nt, st)
             : P.out_fwd  ? launch_fast_t<0, 4, 8, 1, false>(P, c.nt, st)
                          : launch_fast_t<0, 0, 8, 1, false>(P, c.nt, st);
  }
  if (e == cudaSuccess && P.g.flat) e = launch_flat_fix(P, st);
  return e;
}

} // namespace nthb
