// kmer_fast_kernel.cu — NtHash batch kernel for uniform batches (fixed read length, every item full,
// rows that are 16-byte multiples): the BASELINE configs' shape.  Same arithmetic and the same
// reference lines as kmer_kernel.cu (NtHash::roll src/kmer.cpp:246-264, extend_hashes
// src/internal.hpp:104-118); what differs is how bytes get in and hashes get out.
//
// What the first profile (profiles/r01_ncu_kmer_c2_v1.txt) and the microbenchmarks
// (profiles/r01_microbench_*.txt) showed, and what this kernel does about it:
//  * L1TEX data pipe 99.5 % busy: two 16-byte byte-indexed table lookups (4 wavefronts each) and two
//    bank-conflicting LDS.U8 per window.  Here each lane streams its row as 32-bit words (one LDS.32
//    per four windows per stream, realigned with PRMT) and ONE 16-byte lookup per window fetches the
//    combined in/out contribution from a 16-entry pair table indexed by the 2-bit codes
//    (byte >> 1) & 3 of the incoming and outgoing base.  Codes of non-ACGTU bytes are garbage but
//    self-consistent (the same byte enters and leaves with the same code), so windows free of such
//    bytes stay exact; a 256-byte validity LUT flags rows that need the exact scrub pass.
//  * lane-strided 32-byte stores cap at 4.6 TB/s (every sector its own L2 request).  Here each warp
//    collects a [32 items] x [16 u64] tile in shared memory (128-byte rows, SWIZZLE_128B so the
//    per-lane STS.128 are conflict-free) and one elected lane issues cp.async.bulk.tensor.2d
//    (UTMASTG): the output leaves the SM as full 128-byte row segments (7.2 TB/s pattern).
#include "kmer_common.cuh"

#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <map>
#include <mutex>
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
constexpr int F_TILE_OFF = 592;   // 16-byte pad + staged bases
constexpr int F_TILE_PAD = 16;
constexpr int OT_BYTES = 32 * 128; // one warp's output tile: 32 rows x 16 u64

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

template<int LUT>
NTH_D uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return d;
}
// LUT bytes: a = 0xF0, b = 0xCC, c = 0xAA

constexpr int LUT_SEL_C = (0xF0 & ~0xAA & 0xFF) | (0xCC & 0xAA); // (a & ~c) | (b & c)
constexpr int LUT_OR_AND = 0xF0 | (0xCC & 0xAA);                 // a | (b & c)
constexpr int LUT_XOR_AND = (0xF0 ^ 0xCC) & 0xAA;                // (a ^ b) & c
constexpr int LUT_XOR3 = 0xF0 ^ 0xCC ^ 0xAA;                     // a ^ b ^ c

// F <- srol(F) ^ e.xy ; R <- sror(R ^ e.zw) with one combined table entry (12 ALU-pipe ops + 1 IMAD)
NTH_D void roll_step(State& s, const uint4 e)
{
  {
    const uint32_t lo = s.flo, hi = s.fhi;
    const uint32_t hi1 = __funnelshift_l(lo, hi, 1);              // (hi:lo << 1) high word
    const uint32_t nhi = lop3<LUT_SEL_C>(hi1, hi >> 30, 2u);      // bit 33 <- old bit 63
    const uint32_t nlo = lop3<LUT_OR_AND>(lo + lo, hi, 1u);       // bit 0  <- old bit 32
    s.flo = nlo ^ e.x;
    s.fhi = nhi ^ e.y;
  }
  {
    const uint32_t lo = s.rlo ^ e.z, hi = s.rhi ^ e.w;
    s.rlo = __funnelshift_r(lo, hi, 1);                           // bit 31 <- old bit 32
    const uint32_t y = __funnelshift_r(hi, hi >> 1, 1);           // bit 31 <- old bit 33 (hi bit 1)
    s.rhi = lop3<LUT_SEL_C>(y, lo, 1u);                           // bit 32 <- old bit 0
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

// REDUCE: fused consumer (the loop of the reference's examples/benchmark.cpp:34-39): instead of storing
// the hashes, count the visited windows and accumulate the 64-bit sum and xor of all their hash values.
template<int H, bool REDUCE>
__global__ void __launch_bounds__(KMER_NT, 3)
kmer_fast_kernel(const __grid_constant__ KmerParams P, const __grid_constant__ CUtensorMap omap)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + F_BAR_OFF);
  uint8_t* tile = smem + F_TILE_OFF;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t i0 = (uint64_t)blockIdx.x * KMER_NT;
  const uint64_t i1 = min(i0 + (uint64_t)KMER_NT, P.g.n_items);
  const uint32_t k = P.k, n = P.g.seg; // every item has exactly seg windows

  // uniform geometry: item i = (read i / segs, segment i % segs)
  auto item_byte = [&](uint64_t i) {
    const uint64_t r = P.g.segs > 1 ? i / P.g.segs : i;
    return r * P.g.read_len + (i - r * P.g.segs) * (uint64_t)n;
  };
  const bool active = i0 + tid < i1;
  const uint64_t lo_byte = item_byte(i0), hi_byte = item_byte(i1 - 1) + n + k - 1;
  const uint64_t g0 = (lo_byte ? lo_byte - 1 : 0) & ~15ull, g1 = hi_byte;
  if (g1 - g0 > P.tile_cap) __trap();
  const uint64_t my_byte = active ? item_byte(i0 + tid) : g0 + 1; // idle lanes hash a dummy row that TMA clips
  const uint64_t my_out = (i0 + tid) * (uint64_t)n;

  // ---- stage the CTA's byte range (TMA bulk copy) and build the tables -------------------------
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (tid < F_TILE_PAD) tile[tid] = 'A';
  __syncthreads();
  const uint64_t bulk_end = min((g1 + 15) & ~15ull, P.n_bases & ~15ull);
  const uint32_t bulk_bytes = bulk_end > g0 ? (uint32_t)(bulk_end - g0) : 0u;
  // the 4 KB tetramer table is parked in warp 0's (still unused) output tile for the warm-up phase
  const uint32_t ot_base = (sbase + F_TILE_OFF + F_TILE_PAD + P.tile_cap + 16 + 1023u) & ~1023u;
  if (tid == 0) {
    mbar_expect_tx(bar, bulk_bytes + OT_BYTES);
    if (bulk_bytes) bulk_g2s(tile + F_TILE_PAD, P.bases + g0, bulk_bytes, bar);
    bulk_g2s(smem + (ot_base - sbase), P.t4, OT_BYTES, bar);
    // pull the tile of the CTA that will run on this SM slot a few waves from now into L2
    const uint64_t i_nxt = i0 + (uint64_t)P.prefetch_ctas * KMER_NT;
    if (P.prefetch_ctas && i_nxt < P.g.n_items) {
      const uint64_t nxt = item_byte(i_nxt) & ~15ull, len = (g1 - g0 + 15) & ~15ull;
      if (nxt + len <= (P.n_bases & ~15ull)) bulk_prefetch_l2(P.bases + nxt, (uint32_t)len);
    }
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
    smem[F_LUT_OFF + tid] = is_acgtu(tid) ? 0 : 1; // KMER_NT == 256 threads, one LUT byte each
  }
  for (uint64_t g = max(bulk_end, g0) + tid; g < g1; g += KMER_NT) tile[F_TILE_PAD + (g - g0)] = P.bases[g];
  mbar_wait(bar, 0);
  __syncthreads();

  const uint32_t ps = sbase + F_TILE_OFF + F_TILE_PAD + (uint32_t)(my_byte - g0); // shared address of base 0
  const uint32_t lut = sbase + F_LUT_OFF;

  // ---- warm-up: k in-only steps over bases -1 .. k-2 (base -1 is cancelled by the first roll), ----
  // ---- four bases per step through the tetramer table, then k % 4 single steps               ----
  State s = { 0u, 0u, 0u, 0u };
  uint32_t bad = 0;
  uint32_t run = 0;                   // REDUCE: hashable bases in a row, ending at the newest one
  uint64_t acc_sum = 0, acc_xor = 0;  // REDUCE accumulators
  uint32_t acc_cnt = 0;
  const uint32_t ot = ot_base + warp * OT_BYTES;
  {
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
      const uint32_t v0 = lds_u8(lut + __byte_perm(x, 0u, 0x4440u)), v1 = lds_u8(lut + __byte_perm(x, 0u, 0x4441u));
      const uint32_t v2 = lds_u8(lut + __byte_perm(x, 0u, 0x4442u)), v3 = lds_u8(lut + (x >> 24));
      bad |= v0 | v1 | v2 | v3;
      if (REDUCE) {
        run = v0 ? 0 : run + 1;
        run = v1 ? 0 : run + 1;
        run = v2 ? 0 : run + 1;
        run = v3 ? 0 : run + 1;
      }
      const uint32_t y2 = (x >> 1) & 0x03030303u;          // 2-bit codes, first base in byte 0
      const uint32_t off = ((y2 * 0x40100401u) >> 20) & 0xFF0u; // 16 * (c0<<6 | c1<<4 | c2<<2 | c3)
      roll4_in(s, lds_v4(ot_base + off));
    }
    for (uint32_t j = 4 * nq; j < k; ++j) {
      const uint32_t c = lds_u8(a_w + j);
      const uint32_t v = lds_u8(lut + c);
      bad |= v;
      if (REDUCE) run = v ? 0 : run + 1;
      roll_step(s, lds_v4(sbase + F_IN_OFF + ((c & 6u) << 3)));
    }
  }
  __syncthreads(); // the tetramer table is dead from here on: its bytes become warp 0's output tile

  // ---- main loop: word streams + pair table, 16 u64 per tile row ---------------------------------
  // in-stream starts at base k-1, out-stream at base -1; both are read as aligned words + PRMT realign
  const uint32_t a_in = ps + k - 1, a_out = ps - 1;
  uint32_t wp_in = a_in & ~3u, wp_out = a_out & ~3u;
  const uint32_t sel_in = 0x3210u + 0x1111u * (a_in & 3u), sel_out = 0x3210u + 0x1111u * (a_out & 3u);
  uint32_t w_in = lds_u32(wp_in), w_out = lds_u32(wp_out);

  const uint32_t rbx = (ot + lane * 128) ^ ((lane & 7) << 4); // row base with the 128B-swizzle term folded in
  const int row0 = (int)(i0 + warp * 32);
  const uint32_t pair = sbase + F_PAIR_OFF;

  // up to four windows (cnt = 4, or 2 at the very end of a row): consumes one realigned word of each stream
  auto roll4 = [&](uint64_t (&hv)[4], uint32_t cnt) {
    wp_in += 4;
    wp_out += 4;
    const uint32_t w_in_n = lds_u32(wp_in), w_out_n = lds_u32(wp_out);
    const uint32_t x_in = __byte_perm(w_in, w_in_n, sel_in), x_out = __byte_perm(w_out, w_out_n, sel_out);
    w_in = w_in_n;
    w_out = w_out_n;
    // per byte: code_in at bits 5-6, code_out at bits 3-4  =>  byte = 8 * (4*code_in + code_out) = table offset
    const uint32_t c4 = lop3<LUT_SEL_C>(x_out << 2, x_in << 4, 0x60606060u) & 0x78787878u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i >= 2 && cnt < 4) break; // warp-uniform
      const uint32_t v = lds_u8(lut + __byte_perm(x_in, 0u, 0x4440u | i));
      bad |= v;
      if (REDUCE) run = v ? 0 : run + 1;
      const uint32_t ea = __byte_perm(c4, 0u, 0x4440u | i) + pair;
      const uint2 ef = lds_v2(ea), er = lds_v2(ea + (F_PAIR_R_OFF - F_PAIR_OFF));
      roll_step(s, make_uint4(ef.x, ef.y, er.x, er.y));
      hv[i] = canonical2(s);
      if (REDUCE && run >= k) { // the window is one the reference visits
        ++acc_cnt;
        acc_sum += hv[i];
        acc_xor ^= hv[i];
#pragma unroll
        for (int q = 1; q < H; ++q) {
          const uint64_t e = ext_hash(hv[i], P.mult[q]);
          acc_sum += e;
          acc_xor ^= e;
        }
      }
    }
  };

  if constexpr (REDUCE) {
    for (uint32_t p0 = 0; p0 < n; p0 += 4) {
      uint64_t hv[4];
      roll4(hv, n - p0);
    }
    if (!active) acc_cnt = 0, acc_sum = 0, acc_xor = 0;
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
  }

  constexpr uint32_t STEPS = 16 / H; // windows per tile row
  for (uint32_t p0 = 0; p0 < n; p0 += STEPS) {
#pragma unroll
    for (uint32_t q = 0; q < STEPS / 4; ++q) {
      if (p0 + 4 * q < n) { // n is even on this path: a row ends with a group of 4 or of 2
        uint64_t hv[4];
        roll4(hv, n - (p0 + 4 * q));
        if (q == 0 && p0) { // the previous tile must have left shared memory before it is overwritten;
          if (lane == 0) bulk_wait_read0(); // waiting here (not at the top) hides the TMA read behind 4 rolls
          __syncwarp();
        }
        if (H == 1) {
          st_shared_v2_u64(rbx ^ ((2 * q) << 4), hv[0], hv[1]);
          st_shared_v2_u64(rbx ^ ((2 * q + 1) << 4), hv[2], hv[3]);
        } else if (H == 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i) st_shared_v2_u64(rbx ^ ((4 * q + i) << 4), hv[i], ext_hash(hv[i], P.mult[1]));
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            st_shared_v2_u64(rbx ^ ((2 * i) << 4), hv[i], ext_hash(hv[i], P.mult[1]));
            st_shared_v2_u64(rbx ^ ((2 * i + 1) << 4), ext_hash(hv[i], P.mult[2]), ext_hash(hv[i], P.mult[3]));
          }
        }
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&omap, ot, (int)(p0 * H), row0);
      bulk_commit();
    }
  }

  const bool dirty = active && bad != 0;
  const bool any_dirty = __any_sync(0xffffffffu, dirty);
  if (lane == 0) {
    if (any_dirty) bulk_wait_all0(); // zeros must land after the tile they overwrite
    else bulk_wait_read0();          // shared memory must outlive the TMA read
  }
  __syncwarp();
  if (dirty) { // exact clean-up: windows touching a non-ACGTU byte are not emitted (kmer.cpp:232-235, :255-258)
    uint32_t run = 0;
    for (uint32_t j = 0; j < n + k - 1; ++j) {
      run = lds_u8(lut + lds_u8(ps + j)) ? 0 : run + 1;
      if (j >= k - 1 && run < k) {
        const uint64_t w = my_out + (j - (k - 1));
        for (uint32_t q = 0; q < H; ++q) P.out[w * H + q] = 0;
        if (P.valid_bits) atomicAnd(&P.valid_bits[(P.valid_row0 + w) >> 5], ~(1u << ((P.valid_row0 + w) & 31)));
      }
    }
  }
}

uint32_t fast_smem_bytes(uint32_t tile_cap, bool reduce)
{
  return F_TILE_OFF + F_TILE_PAD + tile_cap + 16 + 1024 + (reduce ? 1 : KMER_NT / 32) * OT_BYTES;
}

// Tensor map of the output seen as [n_items rows] x [seg*h u64], boxes of 32 rows x 16 u64.
cudaError_t make_out_map(const KmerParams& P, CUtensorMap* map)
{
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
    if (e != cudaSuccess) return e;
    if (!fp || qr != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
    encode = (EncodeFn)fp;
  }
  const cuuint64_t dims[2] = { (cuuint64_t)P.g.seg * P.h, P.g.n_items };
  const cuuint64_t strides[1] = { (cuuint64_t)P.g.seg * P.h * 8 };
  const cuuint32_t box[2] = { 16, 32 };
  const cuuint32_t estr[2] = { 1, 1 };
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, P.out, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// 256 x {TF, TR}: combined in-only contribution of four consecutive bases c0..c3 (codes, c0 first):
//   TF = srol^3 S[c0] ^ srol^2 S[c1] ^ srol S[c2] ^ S[c3]
//   TR = Skc[c0] ^ srol Skc[c1] ^ srol^2 Skc[c2] ^ srol^3 Skc[c3],  Skc[c] = srol^k S[complement c]
// One small device buffer per (device, k), created on first use and kept for the life of the process.
cudaError_t get_t4_table(uint32_t k, const uint4** out)
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
  e = cudaMalloc(&d, OT_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaMemcpy(d, tab.data(), OT_BYTES, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(d);
    return e;
  }
  cache[{ dev, k }] = d;
  *out = d;
  return cudaSuccess;
}

template<int H, bool REDUCE>
cudaError_t launch_fast_t(const KmerParams& P, const CUtensorMap& map, cudaStream_t st)
{
  auto fn = kmer_fast_kernel<H, REDUCE>;
  const uint32_t smem_bytes = fast_smem_bytes(P.tile_cap, REDUCE);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (P.g.n_items + KMER_NT - 1) / KMER_NT;
  fn<<<(unsigned)ctas, KMER_NT, smem_bytes, st>>>(P, map);
  return cudaGetLastError();
}

} // namespace

bool kmer_fast_ok(const KmerParams& P)
{
  const KmerGeom& g = P.g;
  return !g.item_byte && !P.out_fwd && (P.h == 1 || P.h == 2 || P.h == 4) && g.seg && ((uint64_t)g.seg * P.h) % 2 == 0 && g.seg % 2 == 0 &&
         g.nk % g.seg == 0 && g.n_items > 0 && g.n_items < 0x7fffffffull && (P.reduce_out || ((uintptr_t)P.out & 15) == 0) &&
         fast_smem_bytes(P.tile_cap, P.reduce_out != nullptr) <= 227u * 1024u;
}

cudaError_t launch_kmer_fast(const KmerParams& Pin, cudaStream_t st)
{
  KmerParams P = Pin;
  CUtensorMap map;
  memset(&map, 0, sizeof map);
  cudaError_t e = P.reduce_out ? cudaSuccess : make_out_map(P, &map);
  if (e != cudaSuccess) return e;
  e = get_t4_table(P.k, &P.t4);
  if (e != cudaSuccess) return e;
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const char* env = getenv("NTHASH_B200_PREFETCH_CTAS");
    P.prefetch_ctas = env ? (uint32_t)atoi(env) : (uint32_t)sms * 3; // one residency wave ahead
  }
  if (P.reduce_out) {
    switch (P.h) {
      case 1: return launch_fast_t<1, true>(P, map, st);
      case 2: return launch_fast_t<2, true>(P, map, st);
      default: return launch_fast_t<4, true>(P, map, st);
    }
  }
  switch (P.h) {
    case 1: return launch_fast_t<1, false>(P, map, st);
    case 2: return launch_fast_t<2, false>(P, map, st);
    default: return launch_fast_t<4, false>(P, map, st);
  }
}

} // namespace nthb
