// seed_jit.cu — SeedNtHash kernel specialised per seed set at run time (NVRTC -> sm_100a cubin).
//
// Same reference lines as seed_kernel.cu (SeedNtHash::init/roll src/seed.cpp:493-544, ntmsm64
// :130-270); this is the fast path for uniform batches.  The generic seed_kernel interprets the seed
// plan (loops over seeds/groups/positions, byte loads, descriptor loads) and spends ~700 instructions
// per window (profiles/r01_ncu_seed_c4_v1.txt).  Here the seed set is baked into the code:
//  * the window's 2-bit base codes live in registers (a shift register of ceil(k/16) words; one funnel
//    shift per word per window), so a group index is two rotates and two masks — no byte loads;
//  * positions are looked up two at a time in 16-entry tables kept as two 128-byte halves (forward /
//    reverse strand): every entry in its own bank pair, so the per-lane LDS.64 never conflict;
//  * ignore-mode seeds start from the full-window hash, rolled with the k-mer kernel's pair table;
//  * hashes leave through TMA tile stores: each warp fills a [32 items] x [TW windows * H] u64 tile.
// Exactness for non-ACGTU bytes is handled as in seed_kernel.cu: windows holding one are recomputed
// byte-exactly, and the read is handed to seed_emit_kernel for the visiting-order replay.
// If NVRTC is unavailable or the seed set does not fit, the caller falls back to seed_kernel.
#include "engine.hpp"
#include "nthash_dev.cuh"
#include "seed_plan.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <dlfcn.h>
#include <mutex>
#include <nvrtc.h>
#include <sstream>

namespace nthb {

namespace {

// Mirrored verbatim in the device source below.
struct JitParams
{
  const uint8_t* bases;
  uint64_t n_bases;
  uint64_t n_items;
  uint32_t read_len, nk, seg, segs;
  uint64_t* out;
  uint32_t* valid_bits;
  uint8_t* read_dirty;
  const uint8_t* tables;
  const uint32_t* care;
  uint32_t tile_cap, care_words;
  const uint64_t* item_byte; // ragged batches (RAGGED variant): item i starts at byte item_byte[i], its windows are
  const uint64_t* item_out;  // dense rows [item_out[i], item_out[i+1]); item_read[i] = its read (NULL: items are reads)
  const uint64_t* item_read;
  uint64_t* out_fwd;         // STRANDS variant: get_forward_hash() / get_reverse_hash() per window and seed, [rows][M]
  uint64_t* out_rev;
  uint32_t str_aligned;      // both are 32-byte aligned (then whole-sector stores apply wherever a row group is)
  uint32_t two;              // the constant 2 from the constant bank (roll_step: keeps one multiply-add on the FMA pipe)
  uint64_t* reduce_out;      // REDUCE variant (fused consumer): {windows visited, sum, xor} of the clean items; nothing is stored
  uint8_t* item_dirty;       // REDUCE variant: items holding a byte that needs the exact path (left to seed_reduce_dirty_kernel)
};

const char* const JIT_PRELUDE = R"JIT(
typedef unsigned char uint8_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
struct alignas(64) TensorMap { uint64_t opaque[16]; };
struct JitParams
{
  const uint8_t* bases;
  uint64_t n_bases;
  uint64_t n_items;
  uint32_t read_len, nk, seg, segs;
  uint64_t* out;
  uint32_t* valid_bits;
  uint8_t* read_dirty;
  const uint8_t* tables;
  const uint32_t* care;
  uint32_t tile_cap, care_words;
  const uint64_t* item_byte; // ragged batches (RAGGED variant): item i starts at byte item_byte[i], its windows are
  const uint64_t* item_out;  // dense rows [item_out[i], item_out[i+1]); item_read[i] = its read (NULL: items are reads)
  const uint64_t* item_read;
  uint64_t* out_fwd;         // STRANDS variant: get_forward_hash() / get_reverse_hash() per window and seed, [rows][M]
  uint64_t* out_rev;
  uint32_t str_aligned;      // both are 32-byte aligned (then whole-sector stores apply wherever a row group is)
  uint32_t two;              // the constant 2 from the constant bank (roll_step: keeps one multiply-add on the FMA pipe)
  uint64_t* reduce_out;      // REDUCE variant (fused consumer): {windows visited, sum, xor} of the clean items; nothing is stored
  uint8_t* item_dirty;       // REDUCE variant: items holding a byte that needs the exact path (left to seed_reduce_dirty_kernel)
};
#define DI __device__ __forceinline__
DI uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DI void mbar_init(uint32_t bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
DI void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DI void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
DI void mbar_wait(uint32_t bar, uint32_t parity)
{
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
DI void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
DI void tma_store_2d(const void* tmap, uint32_t saddr, int c0, int c1)
{
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(saddr) : "memory");
}
DI void tma_store_3d(const void* tmap, uint32_t saddr, int c0, int c1, int c2)
{
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(saddr) : "memory");
}
DI void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DI void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
DI void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
DI void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DI uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
DI uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
DI uint2 lds_v2(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
DI uint4 lds_v4(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
DI uint2 lds_u2x(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
DI void stg_v4(void* p, uint4 v) { asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
#if HAVE_ST256 // one whole 32-byte sector (STG.E.ENL2.256); needs the PTX of CUDA 12.9
DI void stg_v4q(uint64_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) { asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory"); }
#else
DI void stg_v4q(uint64_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) { asm volatile("st.global.v2.u64 [%0], {%1,%2};\n\tst.global.v2.u64 [%0+16], {%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory"); }
#endif
DI void stg_v2q(uint64_t* p, uint64_t a, uint64_t b) { asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory"); }
DI void stg_v2(void* p, uint2 v) { asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory"); }
DI void sts_v2(uint32_t a, uint64_t x, uint64_t y) { asm volatile("st.shared.v2.u64 [%0], {%1,%2};" ::"r"(a), "l"(x), "l"(y) : "memory"); }
DI void sts_u64(uint32_t a, uint64_t x) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(x) : "memory"); }
DI void sts_u8(uint32_t a, uint32_t x) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }
DI uint32_t rotr(uint32_t x, uint32_t r) { return __funnelshift_r(x, x, r); }
DI uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
struct State { uint32_t flo, fhi, rlo, rhi; };
template<int LUT> DI uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT)); return d; }
// F <- srol(F) ^ e.xy ; R <- sror(R ^ e.zw)   (src/internal.hpp:41-47, :83-88); the k-mer kernel's form: 10 ALU-pipe ops,
// the plain shifts as IMAD.HI on the FMA pipe.  ROLL_V2 (NTHASH_B200_ROLL_V2=1, off by default): 8 ALU + 7 FMA-pipe ops; measured
// 1 % slower on C4 (profiles/r02_ab_roll_step.txt) — see kmer_fast_kernel.cu.
DI uint32_t mad_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
DI void roll_step(State& s, const uint4 e, const uint32_t two)
{
#if !ROLL_V2
  {
    const uint32_t lo = s.flo, hi = s.fhi;
    const uint32_t hi1 = __funnelshift_l(lo, hi, 1);
#if EXT_SHIFTS >= 2
    const uint32_t nhi = lop3<0x50 | 0x88>(hi1, hi >> 30, 2u);
#else
    const uint32_t nhi = lop3<0x50 | 0x88>(hi1, __umulhi(hi, 4u), 2u);
#endif
    const uint32_t nlo = lop3<0xF0 | 0x88>(lo + lo, hi, 1u);
    s.flo = nlo ^ e.x;
    s.fhi = nhi ^ e.y;
  }
  {
    const uint32_t lo = s.rlo ^ e.z, hi = s.rhi ^ e.w;
    s.rlo = __funnelshift_r(lo, hi, 1);
#if EXT_SHIFTS >= 2
    const uint32_t y = __funnelshift_r(hi, hi >> 1, 1);
#else
    const uint32_t y = __funnelshift_r(hi, __umulhi(hi, 0x80000000u), 1);
#endif
    s.rhi = lop3<0x50 | 0x88>(y, lo, 1u);
  }
  return;
#endif
  {
    const uint32_t lo = s.flo, hi = s.fhi;
    const uint32_t ch = lop3<0x50 | 0x88>(hi * 2u, __umulhi(hi, 4u), 2u); // (a & ~c) | (b & c): bit 33 <- old bit 63; bit 32 clear
    const uint32_t nhi = mad_hi(lo, two, ch);                            // bit 32 <- old bit 31
    const uint32_t nlo = lop3<0xF0 | 0x88>(lo * 2u, hi, 1u);             // a | (b & c): bit 0 <- old bit 32
    s.flo = nlo ^ e.x;
    s.fhi = nhi ^ e.y;
  }
  {
    const uint32_t lo = s.rlo ^ e.z, hi = s.rhi ^ e.w;
    s.rlo = __funnelshift_r(lo, hi, 1);
    const uint32_t a = __umulhi(hi, 0x80000000u); // hi >> 1
    const uint32_t y = a * 0x80000001u;           // a + (a << 31): bit 63 <- old bit 33
    s.rhi = lop3<0x50 | 0x88>(y, lo, 1u);
  }
}
// Strided care block, stride D: F <- srol^D(F) ^ ef ; R <- sror^D(R ^ er).  srol^D (D <= 31) is a 64-bit rotate
// followed by a swap of the D bits that crossed the 33|31 split (src/internal.hpp:56-66); sror^D is its inverse.
template<int D>
DI void blk_step(State& a, const uint2 ef, const uint2 er)
{
  const uint32_t vlo = __funnelshift_l(a.fhi, a.flo, D), vhi = __funnelshift_l(a.flo, a.fhi, D);
  const uint32_t y = (vlo ^ (vhi >> 1)) & ((1u << D) - 1u);
  a.flo = xor3(vlo, y, ef.x);
  a.fhi = xor3(vhi, y << 1, ef.y);
  const uint32_t xlo = a.rlo ^ er.x, xhi = a.rhi ^ er.y;
  const uint32_t y2 = (xlo ^ (xhi >> 1)) & ((1u << D) - 1u);
  const uint32_t zlo = xlo ^ y2, zhi = xhi ^ (y2 << 1);
  a.rlo = __funnelshift_r(zlo, zhi, D);
  a.rhi = __funnelshift_r(zhi, zlo, D);
}
// extend_hashes (src/internal.hpp:104-118): t = h0 * mult; t ^ (t >> 27).  The shifts run as IMAD / IMAD.HI on the FMA pipe,
// the two xors (one fused with the or) on the ALU pipe.
DI uint64_t ext_hash(uint64_t h0, uint64_t mult)
{
  const uint64_t t = h0 * mult;
  const uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
#if EXT_SHIFTS // plain shifts (ALU pipe) instead of multiplies (FMA pipe): see EXT_SHIFTS in the generator
  const uint32_t nlo = lo ^ __funnelshift_r(lo, hi, 27);
  const uint32_t nhi = hi ^ (hi >> 27);
#else
  const uint32_t nlo = lop3<0x1E>(lo, __umulhi(lo, 32u), hi * 32u); // lo ^ ((lo >> 27) | (hi << 5))
  const uint32_t nhi = hi ^ __umulhi(hi, 32u);
#endif
  return ((uint64_t)nhi << 32) | nlo;
}
// per byte of x: non-zero iff the byte is not one of ACGTUacgtu (the bytes whose 2-bit code (c >> 1) & 3 is their seed's base)
DI uint32_t swar_bad(uint32_t x)
{
  const uint32_t t1 = lop3<(0xF0 & ~0xCC) & 0xFF>(x, x << 1, 0u);
  const uint32_t e1 = lop3<(0xF0 ^ 0xCC) & 0xAA>(t1, x >> 2, 0x04040404u);
  const uint32_t e2 = lop3<(~(0xF0 | 0xCC)) & 0xAA & 0xFF>(x, x >> 4, 0x01010101u);
  const uint32_t e3 = lop3<(0xF0 ^ 0xCC) & 0xAA>(x, 0x40404040u, 0xC8C8C8C8u);
  return lop3<0xF0 | 0xCC | 0xAA>(e1, e2, e3);
}
DI uint64_t srol_n(uint64_t x, unsigned d)
{
  const uint64_t m33 = (1ULL << 33) - 1, m31 = (1ULL << 31) - 1;
  uint64_t lo = x & m33, hi = x >> 33;
  const unsigned a = d % 33, b = d % 31;
  if (a) lo = ((lo << a) | (lo >> (33 - a))) & m33;
  if (b) hi = ((hi << b) | (hi >> (31 - b))) & m31;
  return (hi << 33) | lo;
}
DI uint64_t seed_of_byte(unsigned c)
{
  switch (c) {
    case 'A': case 'a': case 4: case 5: return 0x3c8bfbb395c60474ULL;
    case 'C': case 'c': case 7: return 0x3193c18562a02b4cULL;
    case 'G': case 'g': case 3: return 0x20323ed082572324ULL;
    case 'T': case 't': case 'U': case 'u': case 1: return 0x295549f54be24456ULL;
    default: return 0;
  }
}
)JIT";

const char* const JIT_KERNEL = R"JIT(
extern "C" __global__ void __launch_bounds__(NT, 2)
seed_jit_kernel(const __grid_constant__ JitParams P, const __grid_constant__ TensorMap omap)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  // [tables][validity LUT 256][mbarrier 16][16-byte pad + staged bases][1024-aligned output tiles]
  const uint32_t sbase = smem_u32(smem);
  const uint32_t lut = sbase + TABLE_BYTES, bar = lut + 256, tile = bar + 16;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t i0 = (uint64_t)blockIdx.x * NT;
  const uint64_t i1 = i0 + NT < P.n_items ? i0 + NT : P.n_items;
  bool active = i0 + tid < i1;
  uint64_t item = i0 + tid; // the item this thread hashes (ragged batches re-deal the CTA's items by length below)
#if RAGGED
  // items of any length: geometry from the item arrays, the CTA's byte range through shared memory
  __shared__ uint64_t s_range[2];
  uint64_t my_byte = 0, my_out = 0;
  uint32_t n = 0;
  if (active) {
    my_byte = P.item_byte[i0 + tid];
    my_out = P.item_out[i0 + tid];
    n = (uint32_t)(P.item_out[i0 + tid + 1] - my_out);
    if (tid == 0) s_range[0] = my_byte;
    if (i0 + tid == i1 - 1) s_range[1] = my_byte + (n ? n + K - 1 : 0);
  }
  {
    // A warp runs as long as its longest item: hand the CTA's items out by length class (32 classes, longest first;
    // counting sort in shared memory that the tables and the validity LUT overwrite afterwards).
    uint32_t* cnt = (uint32_t*)smem;        // 32 class counters + the CTA's longest item
    uint8_t* perm = smem + TABLE_BYTES;     // slot -> thread that first held the item (NT <= 256)
    if (tid < 33) cnt[tid] = 0;
    __syncthreads();
    atomicMax(&cnt[32], n);
    __syncthreads();
    const uint32_t cls = 31u - (uint32_t)(((uint64_t)n * 32u) / ((uint64_t)cnt[32] + 1u));
    const uint32_t pos = atomicAdd(&cnt[cls], 1u);
    __syncthreads();
    if (tid < 32) {
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
    item = i0 + perm[tid];
    active = item < i1;
    my_byte = my_out = 0;
    n = 0;
    if (active) {
      my_byte = P.item_byte[item];
      my_out = P.item_out[item];
      n = (uint32_t)(P.item_out[item + 1] - my_out);
    }
    __syncthreads(); // the scratch becomes tables / LUT below
  }
#else
  const uint32_t n = P.seg;
  auto item_byte = [&](uint64_t i) {
    const uint64_t r = P.segs > 1 ? i / P.segs : i;
    return r * P.read_len + (i - r * P.segs) * (uint64_t)n;
  };
#endif
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
#if !STRIPS
  if (tid < 16) sts_u8(tile + tid, 'A');
#endif
  __syncthreads();
#if RAGGED
  const uint64_t lo_byte = s_range[0], g1 = s_range[1] > lo_byte ? s_range[1] : lo_byte;
  const uint64_t g0 = (lo_byte > 16 ? lo_byte - 16 : 0) & ~15ull; // up to 16 bases before the first item are staged too
  if (g1 - g0 > P.tile_cap) __trap();
  if (!n) my_byte = g0 + 16; // nothing to hash: the warm-up below still runs, on bytes that exist
#elif STRIPS
  // Warp-private strips of nibble-packed bases (the staging of kmer_pack_kernel, kmer_fast_kernel.cu): every warp loads the
  // bytes of its own 32 items with coalesced 16-byte loads, keeps one nibble per base (code (c >> 1) & 3) and one flag
  // per 16-byte chunk (does it hold a byte for the exact path?).  Half the shared memory per base, no CTA-wide tile wait.
  if (tid == 0) {
    mbar_expect_tx(bar, TABLE_BYTES);
    bulk_g2s(sbase, P.tables, TABLE_BYTES, bar);
  }
  const uint64_t iw0 = i0 + warp * 32;
  if (iw0 >= i1) return; // a warp without items (warp 0 always has some)
  const uint64_t iw1 = iw0 + 32 < i1 ? iw0 + 32 : i1;
  const uint64_t wlo = item_byte(iw0), wend = item_byte(iw1 - 1) + n + K - 1;
  const uint64_t a0 = (wlo > 16 ? wlo - 16 : 0) & ~15ull; // up to 16 bases before the first item are staged too
  const uint32_t n_chunks = (uint32_t)((wend - a0 + 15) >> 4);
  const uint32_t strip = tile + warp * P.tile_cap, badw = strip + P.tile_cap - 128u; // [nibbles][32 flag words]
  if (n_chunks * 8u + 48u + 128u > P.tile_cap || n_chunks > 1024u) __trap();
  if (lane < 2) asm volatile("st.shared.u32 [%0], %1;" ::"r"(strip + lane * 4), "r"(0u) : "memory"); // 16 nibbles of lead-in: code 0
  {
    auto load_chunk = [&](uint32_t c) {
      uint4 x = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
      const uint64_t gb = a0 + 16ull * c;
      if (c < n_chunks) {
        if (gb + 16 <= P.n_bases) {
          x = __ldg((const uint4*)(P.bases + gb));
        } else {
          uint32_t w[4] = { 0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u };
          for (uint32_t j = 0; j < 16 && gb + j < P.n_bases; ++j) w[j >> 2] = (w[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | ((uint32_t)P.bases[gb + j] << (8 * (j & 3)));
          x = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      return x;
    };
    auto nib = [](uint32_t v) {
      const uint32_t y = (v >> 1) & 0x03030303u;
      return y | (y >> 4);
    };
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += 128) {
      uint4 x[4];
#pragma unroll
      for (uint32_t u = 0; u < 4; ++u) x[u] = load_chunk(c0 + 32 * u + lane);
#pragma unroll
      for (uint32_t u = 0; u < 4; ++u) {
        if (c0 + 32 * u >= n_chunks) break;
        const uint32_t c = c0 + 32 * u + lane;
        const uint32_t anybad = swar_bad(x[u].x) | swar_bad(x[u].y) | swar_bad(x[u].z) | swar_bad(x[u].w);
        const uint32_t m = __ballot_sync(0xffffffffu, anybad != 0);
        if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(badw + ((c0 + 32 * u) >> 5) * 4), "r"(m) : "memory");
        const uint32_t lo = __byte_perm(nib(x[u].x), nib(x[u].y), 0x6420u), hi = __byte_perm(nib(x[u].z), nib(x[u].w), 0x6420u);
        if (c < n_chunks) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(strip + 8u + 8u * c), "r"(lo), "r"(hi) : "memory");
      }
    }
  }
  __syncwarp();
  const uint64_t my_byte = active ? item_byte(i0 + tid) : wlo;
  const uint64_t my_out = (i0 + tid) * (uint64_t)n;
  const uint32_t q0 = 16u + (uint32_t)(my_byte - a0); // nibble index of the item's base 0
#define CODE_AT(q) ((lds_u32(strip + ((((uint32_t)(q)) >> 3) << 2)) >> ((((uint32_t)(q)) & 7u) << 2)) & 3u)
  uint32_t bad = 0;
  for (uint32_t c = (q0 - 16u) >> 4; c <= (q0 - 16u + n + K - 2) >> 4; ++c) bad |= (lds_u32(badw + (c >> 5) * 4) >> (c & 31)) & 1u;
  mbar_wait(bar, 0); // the tables
#else
  const uint64_t lo_byte = item_byte(i0), g1 = item_byte(i1 - 1) + n + K - 1;
  const uint64_t g0 = (lo_byte > 16 ? lo_byte - 16 : 0) & ~15ull; // up to 16 bases before the first item are staged too
  if (g1 - g0 > P.tile_cap) __trap();
  const uint64_t my_byte = active ? item_byte(i0 + tid) : g0 + 16;
  const uint64_t my_out = (i0 + tid) * (uint64_t)n;
#endif
#if !STRIPS
  const uint64_t nb16 = P.n_bases & ~15ull;
  const uint64_t bulk_end = ((g1 + 15) & ~15ull) < nb16 ? ((g1 + 15) & ~15ull) : nb16;
  const uint32_t bulk_bytes = bulk_end > g0 ? (uint32_t)(bulk_end - g0) : 0u;
  if (tid == 0) {
    mbar_expect_tx(bar, bulk_bytes + TABLE_BYTES);
    if (bulk_bytes) bulk_g2s(tile + 16, P.bases + g0, bulk_bytes, bar);
    bulk_g2s(sbase, P.tables, TABLE_BYTES, bar);
  }
  // 1: byte-exact path for the windows holding the byte (no seed, or a raw byte 1/3/4/5/7 whose 2-bit code is not its seed's base)
  for (uint32_t c = tid; c < 256; c += NT) sts_u8(lut + c, (seed_of_byte(c) != 0 && c > 7) ? 0u : 1u);
  for (uint64_t g = (bulk_end > g0 ? bulk_end : g0) + tid; g < g1; g += NT) sts_u8(tile + 16 + (uint32_t)(g - g0), P.bases[g]);
  mbar_wait(bar, 0);
  __syncthreads();

  uint32_t ps = tile + 16 + (uint32_t)(my_byte - g0); // shared address of the item's base 0
#if RAGGED_DIRECT
  // Direct sector stores: DG consecutive windows fill whole 32-byte sectors when the first one's row is a multiple of DG.
  // The lane therefore starts `skip` = row mod DG windows early (on the bytes in front of its item: the previous read's
  // tail, the pad, at worst the barrier word — junk either way), computes those windows like any other and never stores
  // them; from then on position u of the unrolled loop is row u (mod DG) in every lane.
  const uint32_t ps_o = ps, n_o = n;
  const uint64_t my_out_o = my_out;
  const uint32_t skip = n ? (uint32_t)(my_out & (uint64_t)(DG - 1u)) : 0u;
  ps -= skip;
  my_out -= skip;
  n += skip;
#define PS_O ps_o
#define N_O n_o
#define MYOUT_O my_out_o
#else
#define PS_O ps
#define N_O n
#define MYOUT_O my_out
#endif
#endif
#if !RAGGED
#define PS_O ps
#define N_O n
#define MYOUT_O my_out
#endif
  const uint32_t tb = sbase;                                 // group tables: two conflict-free 128-byte halves each
#if RAGGED
  // per warp: [32 descriptors x 16 B][32 private rows x ROW_PITCH]; rows go out through coalesced stores (see MAIN_TILES)
  const uint32_t desc0 = ((tile + 16 + P.tile_cap + 16 + 15u) & ~15u) + warp * (32u * (ROW_PITCH + 16u)), rows0 = desc0 + 32u * 16u;
  const uint32_t hw = lane / LANES_PER_ROW, cpos = (lane % LANES_PER_ROW) * CHUNK;
  // the warp's 32 row pieces -> global memory: a group of LANES_PER_ROW lanes per row, CHUNK bytes per lane; each row's
  // address and byte count come from its 16-byte descriptor; loads are issued in batches of eight before the stores
  auto copy_out = [&]() {
    constexpr uint32_t RPI = 32u / LANES_PER_ROW; // rows per store instruction
#pragma unroll
    for (uint32_t r0 = 0; r0 < 32u; r0 += 8u * RPI) {
      uint4 d[8], v[8];
#pragma unroll
      for (uint32_t j = 0; j < 8u; ++j) {
        const uint32_t row = r0 + j * RPI + hw;
        d[j] = lds_v4(desc0 + row * 16u);
        v[j] = make_uint4(0u, 0u, 0u, 0u);
        if (cpos + CHUNK <= ROW_PITCH) { // lanes past the row's end have nothing to copy (and must not read the next warp's memory)
          if (CHUNK == 16u) {
            v[j] = lds_v4(rows0 + row * ROW_PITCH + cpos);
          } else {
            const uint2 t2 = lds_v2(rows0 + row * ROW_PITCH + cpos);
            v[j] = make_uint4(t2.x, t2.y, 0u, 0u);
          }
        }
      }
#pragma unroll
      for (uint32_t j = 0; j < 8u; ++j) {
        uint8_t* ga = (uint8_t*)(((uint64_t)d[j].y << 32) | d[j].x) + cpos;
        if (cpos + CHUNK <= d[j].z) {
          if (CHUNK == 16u) stg_v4(ga, v[j]);
          else stg_v2(ga, make_uint2(v[j].x, v[j].y));
        }
      }
    }
  };
#elif STRIPS
  const uint32_t ot0 = ((tile + (NT / 32u) * P.tile_cap + 1023u) & ~1023u) + warp * (NBUF * OT_BYTES);
  const int row0 = (int)(i0 + warp * 32);
  const uint32_t lterm = lane * 64 + (((lane >> 1) & 3) << 4); // box3: row inside a block + its 64-byte-swizzle term
#else
  const uint32_t ot0 = ((tile + 16 + P.tile_cap + 16 + 1023u) & ~1023u) + warp * (NBUF * OT_BYTES);
  const int row0 = (int)(i0 + warp * 32);
  const uint32_t lterm = lane * 64 + (((lane >> 1) & 3) << 4); // box3: row inside a block + its 64-byte-swizzle term
#endif

  // warm-up: shift bases -OLDER .. k-2 into the code window (the bases before the item only serve strided blocks,
  // where they cancel); roll the full-window hash over bases -1 .. k-2 (in-only)
  DECL_W
  DECL_BLOCKS
  DECL_STR
#if REDUCE
  uint64_t acc_sum = 0ull, acc_xor = 0ull;
#endif
  State full = { 0u, 0u, 0u, 0u };
#if STRIPS
  for (int j = -(int)OLDER; j < (int)K - 1; ++j) {
    const uint32_t c2 = CODE_AT(q0 + (uint32_t)j);
    SHIFT_IN(c2)
#if ANY_IGNORE
    if (j >= -1) roll_step(full, lds_v4(sbase + INTAB_OFF + (c2 << 4)), P.two);
#endif
  }
#else
  uint32_t bad = 0;
  for (int j = -(int)OLDER; j < (int)K - 1; ++j) {
    const uint32_t c = lds_u8(ps + (uint32_t)j);
    if (j >= 0) bad |= lds_u8(lut + c);
    SHIFT_IN(c >> 1)
#if ANY_IGNORE
    if (j >= -1) roll_step(full, lds_v4(sbase + INTAB_OFF + ((c & 6u) << 3)), P.two);
#endif
  }
#endif
  WARMUP_BLOCKS

#if RAGGED
  for (uint32_t p0 = 0; __any_sync(0xffffffffu, p0 < n); p0 += UNROLL) {
    MAIN_TILES
  }
  const bool dirty = active && n != 0 && bad != 0;
#else
  uint32_t buf = 0;
  for (uint32_t p0 = 0; p0 < n; p0 += UNROLL) {
    MAIN_TILES
  }

#if REDUCE
  {
    // Fused consumer: an item without a byte for the exact path has every window visited by the reference's loop (the
    // windows init / roll skip all hold such a byte: seed.cpp:151, :524-530), so its count / sum / xor are final; an item
    // with one contributes nothing here and is redone, window by window, by seed_reduce_dirty_kernel.
    const bool dirty = active && bad != 0;
    uint64_t cnt = (active && !dirty) ? (uint64_t)n : 0ull;
    if (!active || dirty) acc_sum = acc_xor = 0ull;
    for (int o = 16; o; o >>= 1) {
      cnt += __shfl_down_sync(0xffffffffu, cnt, o);
      acc_sum += __shfl_down_sync(0xffffffffu, acc_sum, o);
      acc_xor ^= __shfl_down_sync(0xffffffffu, acc_xor, o);
    }
    if (lane == 0) {
      atomicAdd((unsigned long long*)P.reduce_out, (unsigned long long)cnt);
      atomicAdd((unsigned long long*)P.reduce_out + 1, (unsigned long long)acc_sum);
      atomicXor((unsigned long long*)P.reduce_out + 2, (unsigned long long)acc_xor);
    }
    if (dirty) {
      P.item_dirty[item] = 1;
      P.read_dirty[P.segs > 1 ? item / P.segs : item] = 1;
    }
    return;
  }
#endif
  STR_TAIL
  const bool dirty = active && bad != 0;
  const bool any_dirty = __any_sync(0xffffffffu, dirty);
  if (lane == 0) {
    if (any_dirty) bulk_wait_all0();
    else bulk_wait_read0();
  }
  __syncwarp();
#endif
  if (dirty) {
    // windows holding a zero-seed byte: byte-exact values (seed.cpp:149-166); then flag the read for the replay
    uint32_t run = 0;
    for (uint32_t j = 0; j < N_O + K - 1; ++j) {
#if STRIPS
      const unsigned cj = P.bases[my_byte + j];
      run = (seed_of_byte(cj) != 0 && cj > 7) ? run + 1 : 0;
#else
      run = lds_u8(lut + lds_u8(PS_O + j)) ? 0 : run + 1;
#endif
      if (j >= K - 1 && run < K) {
        const uint32_t p = j - (K - 1);
        const uint64_t row = MYOUT_O + p;
        for (uint32_t s = 0; s < M; ++s) {
          uint64_t f = 0, r = 0;
          const uint32_t* care = P.care + (size_t)s * P.care_words;
          for (uint32_t q = 0; q < K; ++q) {
            if (care[q >> 5] >> (q & 31) & 1u) {
#if STRIPS
              const unsigned c = P.bases[my_byte + p + q];
#else
              const unsigned c = lds_u8(PS_O + p + q);
#endif
              f ^= srol_n(seed_of_byte(c), K - 1 - q);
              r ^= srol_n(seed_of_byte(c & 7u), q);
            }
          }
          const uint64_t h0 = f + r;
          P.out[row * HT + s * HPS] = h0;
          for (uint32_t q = 1; q < HPS; ++q) P.out[row * HT + s * HPS + q] = ext_hash(h0, MULT[q]);
#if STRANDS
          P.out_fwd[row * M + s] = f;
          P.out_rev[row * M + s] = r;
#endif
        }
      }
    }
    const uint64_t i = item;
#if RAGGED
    P.read_dirty[P.item_read ? P.item_read[i] : i] = 1;
#else
    P.read_dirty[P.segs > 1 ? i / P.segs : i] = 1;
#endif
  }
}
)JIT";

struct Nvrtc
{
  bool ok = false;
  nvrtcResult (*create)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*compile)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*cubin_size)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*cubin)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*log_size)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*log)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*destroy)(nvrtcProgram*) = nullptr;
  int version = 0; // major * 100 + minor
};

const Nvrtc& nvrtc()
{
  static Nvrtc n = [] {
    Nvrtc x;
    void* h = nullptr;
    // the toolkit's own NVRTC first: a process that has imported torch already holds torch's bundled (older) libnvrtc.so.12,
    // whose ptxas rejects the 256-bit st.global.v4.u64 the strand stores use
    for (const char* name : { "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so" }) {
      h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) return x;
    x.create = (decltype(x.create))dlsym(h, "nvrtcCreateProgram");
    x.compile = (decltype(x.compile))dlsym(h, "nvrtcCompileProgram");
    x.cubin_size = (decltype(x.cubin_size))dlsym(h, "nvrtcGetCUBINSize");
    x.cubin = (decltype(x.cubin))dlsym(h, "nvrtcGetCUBIN");
    x.log_size = (decltype(x.log_size))dlsym(h, "nvrtcGetProgramLogSize");
    x.log = (decltype(x.log))dlsym(h, "nvrtcGetProgramLog");
    x.destroy = (decltype(x.destroy))dlsym(h, "nvrtcDestroyProgram");
    x.ok = x.create && x.compile && x.cubin_size && x.cubin && x.log_size && x.log && x.destroy;
    if (auto ver = (nvrtcResult (*)(int*, int*))dlsym(h, "nvrtcVersion")) {
      int a = 0, b = 0;
      if (ver(&a, &b) == NVRTC_SUCCESS) x.version = a * 100 + b;
    }
    return x;
  }();
  return n;
}

std::string hex64(uint64_t v)
{
  char b[32];
  snprintf(b, sizeof b, "0x%016llxULL", (unsigned long long)v);
  return b;
}

} // namespace

constexpr bool SEED_JIT_STRIPS_DEFAULT = false;

struct SeedJit
{
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kernel = nullptr;
  uint8_t* d_tables = nullptr;
  uint32_t table_bytes = 0, tw = 0, row_bytes = 0, ot_bytes = 0, ht = 0, nt = 256, nbuf = 1;
  bool box3 = false;            // output through 3-D tensor stores of 64-byte blocks (rows must be 64-byte multiples)
  bool ragged = false;          // ragged-batch variant (item arrays, per-lane rows, coalesced stores)
  bool ragged_direct = false;   // ... its direct form: whole-sector stores from registers, no rows in shared memory
  uint32_t row_pitch = 0;       // ragged variant: bytes between the lanes' private rows
  const SeedPlanHost* plan = nullptr;
  mutable SeedJit* alt = nullptr;  // the 2-D tile variant for other row lengths, compiled on first use
  mutable SeedJit* alt_ragged = nullptr; // the ragged variant, compiled on first use
  mutable SeedJit* alt_str = nullptr;    // uniform batches with strand outputs (3-D box stores for the hashes), compiled on first use
  mutable SeedJit* alt_str2d = nullptr;  // the same over the 2-D tile variant
  mutable SeedJit* alt_reduce = nullptr; // fused consumer (count / sum / xor, nothing stored), compiled on first use
  bool reduce = false;
  bool strips = false;           // uniform variants: warp-private nibble strips (tile_cap = bytes per warp then)
  bool strands = false;
  mutable std::mutex mu;
  std::string source; // kept for inspection (nthash_seed_plan_jit_source)
};

void seed_jit_destroy(SeedJit* j)
{
  if (!j) return;
  seed_jit_destroy(j->alt);
  seed_jit_destroy(j->alt_ragged);
  seed_jit_destroy(j->alt_str);
  seed_jit_destroy(j->alt_str2d);
  seed_jit_destroy(j->alt_reduce);
  if (j->lib) cudaLibraryUnload(j->lib);
  cudaFree(j->d_tables);
  delete j;
}

const char* seed_jit_source(const SeedJit* j) { return j ? j->source.c_str() : ""; }

// Generates, compiles and loads the kernel for one seed set.  Returns nullptr (with a reason) when the
// specialised path does not apply; the caller then uses the generic kernel.
// mode 0: dense 2-D tiles (TMA 2-D tensor stores), 1: [blocks][rows][8 u64] tiles (TMA 3-D tensor stores; the default
// for uniform batches), 2: ragged batches (per-lane rows, coalesced stores, item arrays), 3: fused consumer for uniform
// batches (count / sum / xor accumulated in registers, no output tiles)
static SeedJit* seed_jit_build_variant(const SeedPlanHost& plan, std::string& why, bool load, int mode, bool strands = false);

SeedJit* seed_jit_build(const SeedPlanHost& plan, std::string& why, bool load)
{
  return seed_jit_build_variant(plan, why, load, getenv("NTHASH_B200_SEED_JIT_NO_BOX") ? 0 : 1);
}

static SeedJit* seed_jit_build_variant(const SeedPlanHost& plan, std::string& why, bool load, int mode, bool strands)
{
  const bool box3 = mode == 1, ragged = mode == 2, reduce = mode == 3;
  // uniform variants: warp-private nibble strips instead of the CTA's TMA-staged ASCII tile (NTHASH_B200_SEED_JIT_STRIPS=0/1)
  const char* strips_env = getenv("NTHASH_B200_SEED_JIT_STRIPS");
  const bool strips = !ragged && (strips_env ? atoi(strips_env) != 0 : SEED_JIT_STRIPS_DEFAULT);
  const uint32_t k = plan.k, m = plan.n_seeds, hps = plan.h, ht = m * hps;
  if (k > 256) { why = "k > 256"; return nullptr; }
  if (strands && (ragged || m > 8)) { why = "strand outputs: uniform batches of at most 8 seeds"; return nullptr; }
  if (ht > 64) { why = "more than 64 hashes per window"; return nullptr; }
  // ---- decomposition of every seed's lookup positions: strided blocks (arithmetic progressions of >= 6
  //      positions, stride 1..4, rolled like the whole window) + pairs for the rest ----
  struct Block { uint32_t seed, q0, d, m, id; };
  std::vector<Block> blocks;
  std::vector<std::vector<uint32_t>> rest(m);
  uint32_t phases = 0, older = 1;
  const bool use_blocks = !getenv("NTHASH_B200_SEED_JIT_NO_BLOCKS");
  for (uint32_t s = 0; s < m; ++s) {
    std::vector<char> in(k, 0);
    for (uint32_t q : plan.lookups[s]) in[q] = 1;
    while (use_blocks) {
      Block best = { s, 0, 0, 0, 0 };
      for (uint32_t d = 1; d <= 4; ++d)
        for (uint32_t q = 0; q < k; ++q) {
          if (!in[q] || (q >= d && in[q - d])) continue; // not the start of a maximal progression
          uint32_t len = 0;
          while (q + len * d < k && in[q + len * d]) ++len;
          if (len > best.m) best = { s, q, d, len, 0 };
        }
      if (best.m < 6 || phases + best.d > 8) break;
      best.id = (uint32_t)blocks.size();
      blocks.push_back(best);
      phases += best.d;
      if (best.q0 < best.d) older = std::max(older, best.d - best.q0);
      for (uint32_t j = 0; j < best.m; ++j) in[best.q0 + j * best.d] = 0;
    }
    for (uint32_t q = 0; q < k; ++q)
      if (in[q]) rest[s].push_back(q);
  }
  // window base j sits at code position off + j of a shift register of kw words; `older` positions below it
  const uint32_t kw = (k + older + 15) / 16, off = 16 * kw - k;
  auto rot_to = [&](int j, uint32_t target_bit) { // expression that brings the code of window offset j to bits target_bit..+1
    const uint32_t pos = (uint32_t)((int)off + j);
    std::ostringstream o;
    o << "rotr(W" << pos / 16 << ", " << ((2 * (pos % 16) + 32 - target_bit) % 32) << "u)";
    return o.str();
  };
  uint32_t lcm_d = 1;
  for (const Block& bk : blocks) {
    uint32_t a = lcm_d, b2 = bk.d;
    while (b2) { const uint32_t t = a % b2; a = b2; b2 = t; }
    lcm_d = lcm_d / a * bk.d;
  }

  // ---- tables: [pair F 128][pair R 128][in-only 64] then per block / per pair group: F half 128 B | R half 128 B
  //      (16 entries x 8 bytes = every entry in its own bank pair: LDS.64 with any mix of indices is conflict-free) ----
  const uint64_t seed_of_code[4] = { SEED_A, SEED_C, SEED_T, SEED_G };
  std::vector<uint8_t> tab(320, 0);
  for (unsigned ci = 0; ci < 4; ++ci) {
    for (unsigned co = 0; co < 4; ++co) {
      const uint64_t f = seed_of_code[ci] ^ srol_n(seed_of_code[co], k);
      const uint64_t r = srol_n(seed_of_code[ci ^ 2], k) ^ seed_of_code[co ^ 2];
      memcpy(&tab[(ci * 4 + co) * 8], &f, 8);
      memcpy(&tab[128 + (ci * 4 + co) * 8], &r, 8);
    }
    const uint64_t fi = seed_of_code[ci], ri = srol_n(seed_of_code[ci ^ 2], k);
    memcpy(&tab[256 + ci * 16], &fi, 8);
    memcpy(&tab[256 + ci * 16 + 8], &ri, 8);
  }
  auto put = [&](uint32_t base, uint32_t e, uint64_t f, uint64_t r) {
    memcpy(&tab[base + e * 8], &f, 8);
    memcpy(&tab[base + 128 + e * 8], &r, 8);
  };
  // block tables: roll table indexed (in << 2 | out), in-only table indexed (in)
  std::vector<uint32_t> blk_roll_off(blocks.size()), blk_in_off(blocks.size());
  std::ostringstream decl_blocks, warm_blocks;
  for (const Block& bk : blocks) {
    const uint32_t q0 = bk.q0, d = bk.d, mm = bk.m;
    blk_roll_off[bk.id] = (uint32_t)tab.size();
    tab.resize(tab.size() + 256, 0);
    blk_in_off[bk.id] = (uint32_t)tab.size();
    tab.resize(tab.size() + 256, 0);
    for (uint32_t e = 0; e < 16; ++e) {
      const uint32_t co = e & 3, ci = e >> 2;
      put(blk_roll_off[bk.id], e, srol_n(seed_of_code[co], d + k - 1 - q0) ^ srol_n(seed_of_code[ci], k - 1 - q0 - (mm - 1) * d),
          srol_n(seed_of_code[co ^ 2], q0) ^ srol_n(seed_of_code[ci ^ 2], q0 + mm * d));
    }
    for (uint32_t ci = 0; ci < 4; ++ci)
      put(blk_in_off[bk.id], ci, srol_n(seed_of_code[ci], k - 1 - q0 - (mm - 1) * d), srol_n(seed_of_code[ci ^ 2], q0 + mm * d));
    for (uint32_t ph = 0; ph < d; ++ph) {
      decl_blocks << " State b" << bk.id << "_" << ph << " = { 0u, 0u, 0u, 0u };";
      // G(ph - d) by m in-only steps over bases ph - d + q0 + t*d (t = 0..m-1)
      warm_blocks << "  for (uint32_t t = 0; t < " << mm << "u; ++t) { const uint32_t ix = "
                  << (strips ? "CODE_AT(q0 + (uint32_t)(" : "(lds_u8(ps + (uint32_t)(") << (int)ph - (int)d + (int)q0 << ") + t * " << d
                  << (strips ? "u) << 3; blk_step<" : "u) & 6u) << 2; blk_step<") << d << ">(b" << bk.id << "_" << ph << ", lds_v2(tb + " << blk_in_off[bk.id]
                  << "u + ix), lds_v2(tb + " << blk_in_off[bk.id] + 128 << "u + ix)); } \\\n";
    }
  }
  // pair groups of the remaining positions
  struct Pair { uint32_t qa, qb, off; bool two; };
  std::vector<std::vector<Pair>> pairs(m);
  for (uint32_t s = 0; s < m; ++s) {
    const std::vector<uint32_t>& lp = rest[s];
    for (size_t a2 = 0; a2 < lp.size(); a2 += 2) {
      Pair pr = { lp[a2], a2 + 1 < lp.size() ? lp[a2 + 1] : 0, (uint32_t)tab.size(), a2 + 1 < lp.size() };
      tab.resize(tab.size() + 256, 0);
      for (uint32_t e = 0; e < (pr.two ? 16u : 4u); ++e) {
        const uint32_t ca = e & 3, cb = e >> 2;
        uint64_t f = srol_n(seed_of_code[ca], k - 1 - pr.qa), r = srol_n(seed_of_code[ca ^ 2], pr.qa);
        if (pr.two) {
          f ^= srol_n(seed_of_code[cb], k - 1 - pr.qb);
          r ^= srol_n(seed_of_code[cb ^ 2], pr.qb);
        }
        put(pr.off, e, f, r);
      }
      pairs[s].push_back(pr);
    }
  }
  const uint32_t table_bytes = (uint32_t)((tab.size() + 15) & ~size_t(15));
  tab.resize(table_bytes, 0);

  // ---- output tile geometry: TW windows per tile row ----
  //  box3: rows of whole 64-byte blocks ([blocks][32 rows][8 u64] under the TMA 64-byte swizzle), 192-256 bytes preferred
  //        (the k-mer kernel's finding: longer pieces cost more occupancy than they gain on the write path);
  //  else: one dense [32 rows][TW*HT u64] tile, an odd number of 16-byte chunks per row preferred (conflict-free STS.128)
  uint32_t tw = reduce ? 4 : 0, best = 0;
  for (uint32_t t = 1; t <= 32 && !reduce; ++t) {
    const uint32_t rb = t * ht * 8;
    if (rb > 512 || t * ht > 256) continue;
    uint32_t score;
    if (box3) {
      if (rb % 64) continue;
      score = 1000 - (rb >= 192 ? rb - 192 : 2 * (192 - rb));
    } else if (ragged) { // a row piece is copied out by 16 (or, for odd hash counts, 32) lanes: at most 256 bytes
      if (rb > 256 || rb % (ht % 2 ? 8 : 16)) continue;
      score = rb;
    } else {
      if (rb % 16) continue;
      score = ((rb / 16) % 2 ? 1000 : 0) + (rb <= 256 ? rb : 512 - rb); // long rows write better (DRAM), up to ~256 B
    }
    if (score > best) { best = score; tw = t; }
  }
  if (const char* e = getenv("NTHASH_B200_SEED_JIT_TW")) { // experiments
    const uint32_t t = (uint32_t)atoi(e);
    if (!ragged && !reduce && t >= 1 && (t * ht * 8) % (box3 ? 64 : 16) == 0 && t * ht <= 256) tw = t;
  }
  if (!tw) { why = "no tile row shape for this number of hashes"; return nullptr; }
  const uint32_t row_bytes = tw * ht * 8, ot_bytes = reduce ? 0u : box3 ? (row_bytes / 64) * 2048u : (32 * row_bytes + 127) & ~127u;
  uint32_t unroll = tw; // windows per generated loop body: a multiple of the tile and of every block stride
  {
    uint32_t a = unroll, b2 = lcm_d;
    while (b2) { const uint32_t t = a % b2; a = b2; b2 = t; }
    unroll = unroll / a * lcm_d;
  }
  if (strands) { // strand hashes leave in groups of four windows (whole 32-byte sectors): the loop body must hold whole groups
    uint32_t a = unroll, b2 = 4;
    while (b2) { const uint32_t t = a % b2; a = b2; b2 = t; }
    unroll = unroll / a * 4;
  }
  // ragged batches, direct form: DG = 4 / gcd(ht, 4) consecutive windows are whole 32-byte sectors; they are kept in registers
  // (DG * ht u64, at most 16) and leave as full-sector stores — no private rows, no descriptors, no copy loop.  Only the two
  // ends of a row share sectors with its neighbours (16-/8-byte stores there).  NTHASH_B200_SEED_JIT_RAGGED_ROWS=1: the older form.
  uint32_t dg = 0;
  if (ragged && !getenv("NTHASH_B200_SEED_JIT_RAGGED_ROWS")) {
    const uint32_t g4 = ht % 4 == 0 ? 4 : ht % 2 == 0 ? 2 : 1;
    if ((4 / g4) * ht <= 16) dg = 4 / g4;
  }
  if (dg) { // loop body: whole groups and whole block periods, four windows or more
    uint32_t a = dg, b2 = lcm_d;
    while (b2) { const uint32_t t = a % b2; a = b2; b2 = t; }
    unroll = dg / a * lcm_d;
    while (unroll < 4) unroll *= 2;
  }
  if (unroll > 24) { why = "strided blocks with incompatible strides"; return nullptr; }
  // CTA size / output buffering: defaults found on C4 (profiles/r01_seed_jit_sweeps.txt), overridable for experiments
  // (box3 on C4: 128 threads with three 6 KB tiles per warp 0.89 of the HBM peak, two 0.85, one 0.76; 2-D tiles 0.63)
  // round 2 (profiles/r02_seed_jit_sweep.txt, after the ALU diet of roll_step / ext_hash): 160 threads with two tiles per
  // warp 10.43 ms, 96 x 3 10.45, 128 x 2 10.59, 128 x 3 (the round-1 default) 11.00: resident warps beat a third tile now
  uint32_t nt = box3 ? 160 : ragged ? 128 : 256, nbuf = box3 ? 2 : 1;
  if (reduce) nt = 192; // no output tiles: shared memory holds the bases only, registers decide the occupancy
  if (const char* e = getenv("NTHASH_B200_SEED_JIT_NT")) nt = (uint32_t)atoi(e);
  if (const char* e = getenv("NTHASH_B200_SEED_JIT_NBUF")) nbuf = (uint32_t)atoi(e);
  if (nt < 32 || nt > 1024 || nt % 32 || nbuf < 1 || nbuf > 4) { why = "bad NT/NBUF override"; return nullptr; }

  // ---- the per-window body, one text per position u in the unrolled loop (block phases are static there) ----
  auto window_body = [&](uint32_t u) {
    std::ostringstream body;
    for (uint32_t s = 0; s < m; ++s) {
      body << "        { /* seed " << s << (plan.ignore_mode[s] ? " (ignore-mode)" : " (care-mode)") << " */ \\\n";
      if (plan.ignore_mode[s]) body << "          uint32_t flo = full.flo, fhi = full.fhi, rlo = full.rlo, rhi = full.rhi; \\\n";
      else body << "          uint32_t flo = 0u, fhi = 0u, rlo = 0u, rhi = 0u; \\\n";
      for (const Block& bk : blocks) {
        if (bk.seed != s) continue;
        const uint32_t ph = u % bk.d;
        body << "          { const uint32_t ix = (" << rot_to((int)bk.q0 - (int)bk.d, 3) << " & 0x18u) | (" << rot_to((int)(bk.q0 + (bk.m - 1) * bk.d), 5)
             << " & 0x60u); blk_step<" << bk.d << ">(b" << bk.id << "_" << ph << ", lds_v2(tb + " << blk_roll_off[bk.id] << "u + ix), lds_v2(tb + "
             << blk_roll_off[bk.id] + 128 << "u + ix)); flo ^= b" << bk.id << "_" << ph << ".flo; fhi ^= b" << bk.id << "_" << ph << ".fhi; rlo ^= b"
             << bk.id << "_" << ph << ".rlo; rhi ^= b" << bk.id << "_" << ph << ".rhi; } \\\n";
      }
      const std::vector<Pair>& pv = pairs[s];
      for (size_t gi = 0; gi < pv.size(); ++gi) {
        body << "          const uint32_t ix" << gi << " = (" << rot_to((int)pv[gi].qa, 3) << " & 0x18u)";
        if (pv[gi].two) body << " | (" << rot_to((int)pv[gi].qb, 5) << " & 0x60u)";
        body << "; const uint2 ef" << gi << " = lds_v2(tb + " << pv[gi].off << "u + ix" << gi << "), er" << gi << " = lds_v2(tb + " << pv[gi].off + 128
             << "u + ix" << gi << "); \\\n";
      }
      size_t gi = 0;
      for (; gi + 1 < pv.size(); gi += 2)
        body << "          flo = xor3(flo, ef" << gi << ".x, ef" << gi + 1 << ".x); fhi = xor3(fhi, ef" << gi << ".y, ef" << gi + 1 << ".y); rlo = xor3(rlo, er"
             << gi << ".x, er" << gi + 1 << ".x); rhi = xor3(rhi, er" << gi << ".y, er" << gi + 1 << ".y); \\\n";
      if (gi < pv.size())
        body << "          flo ^= ef" << gi << ".x; fhi ^= ef" << gi << ".y; rlo ^= er" << gi << ".x; rhi ^= er" << gi << ".y; \\\n";
      if (strands)
        body << "          sf[" << (u % 4) * m + s << "] = ((uint64_t)fhi << 32) | flo; sr[" << (u % 4) * m + s << "] = ((uint64_t)rhi << 32) | rlo; \\\n";
      body << "          const uint64_t h0 = (((uint64_t)fhi << 32) | flo) + (((uint64_t)rhi << 32) | rlo); \\\n";
      body << "          hv[" << s * hps << "] = h0; \\\n";
      for (uint32_t q = 1; q < hps; ++q) body << "          hv[" << s * hps + q << "] = ext_hash(h0, " << hex64(ext_mult(q, k)) << "); \\\n";
      body << "        } \\\n";
    }
    return body.str();
  };

  std::ostringstream src;
  src << "#define HAVE_ST256 " << (nvrtc().version >= 1209 ? 1 : 0) << "\n#define ROLL_V2 " << (getenv("NTHASH_B200_ROLL_V2") ? 1 : 0) << "\n";
  // plain shifts (ALU pipe) instead of shift-by-multiply (FMA pipe) in ext_hash (>= 1) and roll_step (>= 2): level 2 by default.  The
  // multiplies were the better trade while the ALU pipe was the bottleneck; under sustained load the kernel sits on the power cap
  // and the cheaper instruction wins: C4 10.36 / 11.33 ms (best / median of 16 back-to-back launches) -> 10.17 / 11.03, fused
  // consumer 8.28 -> 8.26 (profiles/r02_ab_seed_shifts.txt)
  src << "#define EXT_SHIFTS " << (getenv("NTHASH_B200_SEED_JIT_EXT_SHIFTS") ? atoi(getenv("NTHASH_B200_SEED_JIT_EXT_SHIFTS")) : 2) << "\n";
  src << JIT_PRELUDE;
  src << "#define NT " << nt << "u\n#define NBUF " << nbuf << "u\n#define BULK_WAIT_READ asm volatile(\"cp.async.bulk.wait_group.read "
      << nbuf - 1 << ";\" ::: \"memory\");\n";
  src << "#define K " << k << "u\n#define M " << m << "u\n#define HPS " << hps << "u\n#define HT " << ht << "u\n#define TW " << tw
      << "u\n#define UNROLL " << unroll << "u\n#define OLDER " << std::min<uint32_t>(off, 16) << "u\n#define ROW_BYTES " << row_bytes
      << "u\n#define OT_BYTES " << ot_bytes << "u\n#define TABLE_BYTES " << table_bytes << "u\n#define ANY_IGNORE " << (plan.any_ignore ? 1 : 0)
      << "\n#define PAIRF_OFF 0u\n#define PAIRR_OFF 128u\n#define INTAB_OFF 256u\n";
  {
    // ragged variant: 16-byte chunks, two rows per store instruction when rows are 16-byte aligned (even hash count),
    // else 8-byte chunks, one row per instruction; rows padded to an odd number of 16-byte chunks (conflict-free STS.128)
    const uint32_t chunk = ht % 2 ? 8 : 16, pitch = (row_bytes + 15) / 16 * 16 + ((((row_bytes + 15) / 16) % 2) ? 0 : 16);
    src << "#define RAGGED " << (ragged ? 1 : 0) << "\n#define CHUNK " << chunk << "u\n#define LANES_PER_ROW " << 256 / chunk << "u\n#define ROW_PITCH " << pitch << "u\n";
  }
  src << "#define RAGGED_DIRECT " << (dg ? 1 : 0) << "\n#define DG " << (dg ? dg : 1) << "u\n";
  src << "#define STRANDS " << (strands ? 1 : 0) << "\n#define REDUCE " << (reduce ? 1 : 0) << "\n#define STRIPS " << (strips ? 1 : 0) << "\n";
  if (strands) {
    // STR_FLUSH(w0): the strand hashes of windows w0 .. w0+3 (4*M u64 per array, contiguous in both arrays) as M whole
    // 32-byte sectors when the group starts on one, else value by value; STR_TAIL: the last n % 4 windows of the row
    src << "#define DECL_STR uint64_t sf[" << 4 * m << "], sr[" << 4 * m << "]; const bool str_al = P.str_aligned && (((my_out * M) & 3ull) == 0);\n";
    src << "#define STR_FLUSH(w0) if (active) { uint64_t* pf = P.out_fwd + (my_out + (w0)) * M; uint64_t* pr = P.out_rev + (my_out + (w0)) * M; if (str_al) {";
    for (uint32_t c = 0; c < m; ++c)
      src << " stg_v4q(pf + " << 4 * c << ", sf[" << 4 * c << "], sf[" << 4 * c + 1 << "], sf[" << 4 * c + 2 << "], sf[" << 4 * c + 3 << "]);"
          << " stg_v4q(pr + " << 4 * c << ", sr[" << 4 * c << "], sr[" << 4 * c + 1 << "], sr[" << 4 * c + 2 << "], sr[" << 4 * c + 3 << "]);";
    src << " } else {";
    for (uint32_t j = 0; j < 4 * m; ++j) src << " pf[" << j << "] = sf[" << j << "]; pr[" << j << "] = sr[" << j << "];";
    src << " } }\n";
    src << "#define STR_TAIL { const uint32_t rem = n & 3u; if (active && rem) { uint64_t* pf = P.out_fwd + (my_out + (n - rem)) * M; uint64_t* pr = P.out_rev + (my_out + (n - rem)) * M;";
    for (uint32_t w = 0; w < 3; ++w) {
      src << " if (rem > " << w << "u) {";
      for (uint32_t q = 0; q < m; ++q) src << " pf[" << w * m + q << "] = sf[" << w * m + q << "]; pr[" << w * m + q << "] = sr[" << w * m + q << "];";
      src << " }";
    }
    src << " } }\n";
  } else {
    src << "#define DECL_STR\n#define STR_FLUSH(w0)\n#define STR_TAIL\n";
  }
  src << "__device__ const uint64_t MULT[" << (hps > 1 ? hps : 1) << "] = { 0";
  for (uint32_t q = 1; q < hps; ++q) src << ", " << hex64(ext_mult(q, k));
  src << " };\n#define DECL_W";
  for (uint32_t w = 0; w < kw; ++w) src << " uint32_t W" << w << " = 0u;";
  src << "\n#define DECL_BLOCKS" << decl_blocks.str() << "\n";
  src << "#define WARMUP_BLOCKS \\\n" << warm_blocks.str() << "\n";
  src << "#define SHIFT_IN(t) {";
  for (uint32_t w = 0; w + 1 < kw; ++w) src << " W" << w << " = __funnelshift_r(W" << w << ", W" << w + 1 << ", 2);";
  src << " W" << kw - 1 << " = __funnelshift_r(W" << kw - 1 << ", (t), 2); }\n";
  // full-window roll for ignore-mode seeds: the outgoing base is the previous window's base 0 (position off, before the shift)
  src << "#define FULL_ROLL";
  if (plan.any_ignore)
    src << " { const uint32_t po = " << (strips ? "(c2 << 5)" : "((c << 4) & 0x60u)") << " | (rotr(W" << off / 16 << ", " << ((2 * (off % 16) + 32 - 3) % 32)
        << "u) & 0x18u); const uint2 ef = lds_v2(sbase + PAIRF_OFF + po), er = lds_v2(sbase + PAIRR_OFF + po);"
           " roll_step(full, make_uint4(ef.x, ef.y, er.x, er.y), P.two); }";
  if (strips) src << "\n#define WIN_PRE(P) const uint32_t c2 = CODE_AT(q0 + (K - 1) + (P)); FULL_ROLL SHIFT_IN(c2)\n";
  else src << "\n#define WIN_PRE(P) const uint32_t c = lds_u8(ps + (K - 1) + (P)); bad |= lds_u8(lut + c); FULL_ROLL SHIFT_IN(c >> 1)\n";
  // STORE_WINDOW_i(base): window i of the tile row -> shared memory
  for (uint32_t i = 0; i < tw; ++i) {
    src << "#define STORE_WINDOW_" << i << "(base) {";
    for (uint32_t q = 0; q < ht;) {
      const uint32_t col = i * ht + q;           // u64 column inside the tile row
      const bool two = col % 2 == 0 && q + 1 < ht; // a whole 16-byte chunk
      std::ostringstream addr;
      if (box3) { // block col/8 of the lane's row; 16-byte chunk (col/2)&3 XOR-swizzled with the lane term folded into `base`
        addr << "((base) ^ " << (((col / 2) & 3u) << 4) << "u) + " << (col / 8) * 2048u + (col & 1u) * 8u << "u";
      } else {
        addr << "(base) + " << col * 8 << "u";
      }
      if (two) src << " sts_v2(" << addr.str() << ", hv[" << q << "], hv[" << q + 1 << "]);";
      else src << " sts_u64(" << addr.str() << ", hv[" << q << "]);";
      q += two ? 2 : 1;
    }
    src << " }\n";
  }
  // one tile = TW windows; a full tile (the common case: rows are usually a multiple of TW) runs without the
  // per-window bound checks, a partial one (end of a row) keeps them
  auto tile_text = [&](uint32_t t, bool full) {
    std::ostringstream o;
    if (reduce) { // fused consumer: the window's values go straight into the accumulators
      for (uint32_t i = 0; i < tw; ++i) {
        const uint32_t u = t * tw + i;
        o << "      " << (full ? "{" : "if (p0 + " + std::to_string(u) + "u < n) {") << " \\\n        WIN_PRE(p0 + " << u << "u) \\\n        uint64_t hv[HT]; \\\n"
          << window_body(u) << "       ";
        for (uint32_t q = 0; q < ht; ++q) o << " acc_sum += hv[" << q << "]; acc_xor ^= hv[" << q << "];";
        o << " \\\n      } \\\n";
      }
      return o.str();
    }
    o << "      const uint32_t ot = ot0 + buf * OT_BYTES, rowaddr = ot + " << (box3 ? "lterm" : "lane * ROW_BYTES") << "; \\\n";
    for (uint32_t i = 0; i < tw; ++i) {
      const uint32_t u = t * tw + i;
      o << "      " << (full ? "{" : "if (p0 + " + std::to_string(u) + "u < n) {") << " \\\n        WIN_PRE(p0 + " << u << "u) \\\n        uint64_t hv[HT]; \\\n"
        << window_body(u);
      if (i == 0) // this buffer's previous tile must have left shared memory; waiting only now hides the TMA read behind one window
        o << "        if (p0 + " << t * tw << "u >= NBUF * TW) { if (lane == 0) BULK_WAIT_READ __syncwarp(); } \\\n";
      o << "        STORE_WINDOW_" << i << "(rowaddr) \\\n";
      if (strands && u % 4 == 3) o << "        STR_FLUSH(p0 + " << u - 3 << "u) \\\n";
      o << "      } \\\n";
    }
    o << "      fence_proxy_async_smem(); __syncwarp(); \\\n      if (lane == 0) { ";
    if (box3) o << "tma_store_3d(&omap, ot, 0, row0, (int)((p0 + " << t * tw << "u) * HT / 8u));";
    else o << "tma_store_2d(&omap, ot, (int)((p0 + " << t * tw << "u) * HT), row0);";
    o << " bulk_commit(); } \\\n      buf = buf + 1 == NBUF ? 0 : buf + 1; \\\n";
    return o.str();
  };
  // ragged variant: lanes differ in length; every lane fills its private row, then the warp copies the rows out
  auto ragged_tile_text = [&](uint32_t t) {
    std::ostringstream o;
    o << "      const uint32_t rowaddr = rows0 + lane * ROW_PITCH; \\\n";
    for (uint32_t i = 0; i < tw; ++i) {
      const uint32_t u = t * tw + i;
      o << "      if (p0 + " << u << "u < n) { \\\n        WIN_PRE(p0 + " << u << "u) \\\n        uint64_t hv[HT]; \\\n" << window_body(u)
        << "        STORE_WINDOW_" << i << "(rowaddr) \\\n      } \\\n";
    }
    o << "      { const uint32_t w0 = p0 + " << t * tw << "u, cntw = w0 < n ? (n - w0 < TW ? n - w0 : TW) : 0u; \\\n"
      << "        sts_v2(desc0 + lane * 16u, (uint64_t)(P.out + (my_out + w0) * HT), (uint64_t)(cntw * HT * 8u)); } \\\n"
      << "      __syncwarp(); \\\n      copy_out(); \\\n      __syncwarp(); \\\n";
    return o.str();
  };
  // ragged variant, direct form: DG windows -> registers gv[] -> whole sectors (see RAGGED_DIRECT in the kernel text)
  auto ragged_direct_text = [&](uint32_t gi) {
    std::ostringstream o;
    o << "      uint64_t gv[" << dg * ht << "]; \\\n";
    for (uint32_t i = 0; i < dg; ++i) {
      const uint32_t u = gi * dg + i;
      std::string body = window_body(u); // writes hv[q]: redirect to gv[i * ht + q]
      for (size_t at = 0; (at = body.find("hv[", at)) != std::string::npos;) {
        const size_t e = body.find(']', at);
        const uint32_t q = (uint32_t)std::stoul(body.substr(at + 3, e - at - 3));
        const std::string rep = "gv[" + std::to_string(i * ht + q) + "]";
        body.replace(at, e - at + 1, rep);
        at += rep.size();
      }
      o << "      if (p0 + " << u << "u < n) { \\\n        WIN_PRE(p0 + " << u << "u) \\\n" << body << "      } \\\n";
    }
    o << "      if (active) { const uint32_t w0 = p0 + " << gi * dg << "u; uint64_t* o_ = P.out + (my_out + w0) * HT; \\\n"
      << "        if (w0 >= skip && w0 + " << dg << "u <= n) {";
    for (uint32_t c = 0; c < dg * ht / 4; ++c)
      o << " stg_v4q(o_ + " << 4 * c << ", gv[" << 4 * c << "], gv[" << 4 * c + 1 << "], gv[" << 4 * c + 2 << "], gv[" << 4 * c + 3 << "]);";
    o << " } else { \\\n";
    for (uint32_t i = 0; i < dg; ++i) { // a row's first / last group: the widest aligned pieces of each window that is the lane's
      o << "          if (w0 + " << i << "u >= skip && w0 + " << i << "u < n) {";
      for (uint32_t c = i * ht; c < (i + 1) * ht;) {
        if (c % 4 == 0 && c + 4 <= (i + 1) * ht) {
          o << " stg_v4q(o_ + " << c << ", gv[" << c << "], gv[" << c + 1 << "], gv[" << c + 2 << "], gv[" << c + 3 << "]);";
          c += 4;
        } else if (c % 2 == 0 && c + 2 <= (i + 1) * ht) {
          o << " stg_v2q(o_ + " << c << ", gv[" << c << "], gv[" << c + 1 << "]);";
          c += 2;
        } else {
          o << " o_[" << c << "] = gv[" << c << "];";
          c += 1;
        }
      }
      o << " } \\\n";
    }
    o << "        } \\\n      } \\\n";
    return o.str();
  };
  src << "#define MAIN_TILES \\\n";
  for (uint32_t gi = 0; dg && gi < unroll / dg; ++gi)
    src << "    if (__any_sync(0xffffffffu, p0 + " << gi * dg << "u < n)) { \\\n" << ragged_direct_text(gi) << "    } \\\n";
  for (uint32_t t = 0; ragged && !dg && t < unroll / tw; ++t)
    src << "    if (__any_sync(0xffffffffu, p0 + " << t * tw << "u < n)) { \\\n" << ragged_tile_text(t) << "    } \\\n";
  for (uint32_t t = 0; !ragged && t < unroll / tw; ++t) {
    src << "    if (p0 + " << (t + 1) * tw << "u <= n) { \\\n" << tile_text(t, true) << "    } else if (p0 + " << t * tw << "u < n) { \\\n"
        << tile_text(t, false) << "    } \\\n";
  }
  src << "\n";
  src << JIT_KERNEL;

  const Nvrtc& rt = nvrtc();
  if (!rt.ok) { why = "libnvrtc could not be loaded"; return nullptr; }
  SeedJit* j = new SeedJit();
  j->source = src.str();
  j->table_bytes = table_bytes;
  j->tw = tw;
  j->row_bytes = row_bytes;
  j->ot_bytes = ot_bytes;
  j->ht = ht;
  j->nt = nt;
  j->nbuf = nbuf;
  j->box3 = box3;
  j->ragged = ragged;
  j->ragged_direct = dg != 0;
  j->strands = strands;
  j->reduce = reduce;
  j->strips = strips;
  j->row_pitch = (row_bytes + 15) / 16 * 16 + ((((row_bytes + 15) / 16) % 2) ? 0 : 16);
  j->plan = &plan;
  nvrtcProgram prog = nullptr;
  if (rt.create(&prog, j->source.c_str(), "seed_jit_kernel.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
    why = "nvrtcCreateProgram failed";
    delete j;
    return nullptr;
  }
  const char* opts[] = { "--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo" };
  const nvrtcResult cr = rt.compile(prog, 3, opts);
  if (cr != NVRTC_SUCCESS) {
    size_t ls = 0;
    rt.log_size(prog, &ls);
    std::string log(ls, '\0');
    if (ls) rt.log(prog, &log[0]);
    why = "NVRTC compilation failed: " + log.substr(0, 1500);
    rt.destroy(&prog);
    delete j;
    return nullptr;
  }
  size_t cs = 0;
  rt.cubin_size(prog, &cs);
  std::vector<char> cubin(cs);
  rt.cubin(prog, cubin.data());
  rt.destroy(&prog);
  if (const char* dump = getenv("NTHASH_B200_SEED_JIT_DUMP")) { // inspection: <prefix>.cu and <prefix>.cubin
    const std::string pre = std::string(dump) + "_mode" + std::to_string(mode) + (strands ? "s" : "");
    if (FILE* f = fopen((pre + ".cu").c_str(), "w")) { fputs(j->source.c_str(), f); fclose(f); }
    if (FILE* f = fopen((pre + ".cubin").c_str(), "wb")) { fwrite(cubin.data(), 1, cubin.size(), f); fclose(f); }
  }
  if (!load) return j; // build check on a machine without a GPU
  cudaError_t e = cudaLibraryLoadData(&j->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&j->kernel, j->lib, "seed_jit_kernel");
  if (e == cudaSuccess) e = cudaMalloc(&j->d_tables, table_bytes);
  if (e == cudaSuccess) e = cudaMemcpy(j->d_tables, tab.data(), table_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    why = std::string("loading the specialised kernel failed: ") + cudaGetErrorString(e);
    cudaGetLastError();
    seed_jit_destroy(j);
    return nullptr;
  }
  return j;
}

uint32_t seed_jit_smem_bytes(const SeedJit* j, uint32_t tile_cap)
{
  if (j->strips) { // tile_cap = bytes of one warp's strip (nibbles + flag words)
    const uint32_t out_bytes = j->reduce ? 0u : (j->nt / 32) * j->nbuf * j->ot_bytes;
    return j->table_bytes + 256 + 16 + (j->nt / 32) * tile_cap + 1024 + out_bytes;
  }
  const uint32_t out_bytes = j->reduce ? 0u : j->ragged ? (j->ragged_direct ? 256u : (j->nt / 32) * 32u * (j->row_pitch + 16u) + 256u) : (j->nt / 32) * j->nbuf * j->ot_bytes;
  return j->table_bytes + 256 + 16 + 16 + tile_cap + 16 + 1024 + out_bytes;
}

// Bytes of bases one CTA of the specialised kernel stages (its CTA size may differ from KMER_NT).
static uint32_t seed_jit_tile_cap(const SeedJit* j, const SeedParams& P)
{
  const KmerGeom& g = P.g;
  uint64_t b;
  if (j->strips) { // per warp: 32 consecutive items, up to 16 bases before the first one, 16-byte chunks, 8 bytes of nibbles per chunk
    const uint64_t span = g.segs == 1 ? 31ull * g.read_len + g.seg + P.k - 1 : 32ull * g.seg + (32ull / g.segs + 2) * (P.k - 1);
    const uint64_t chunks = (span + 16 + 15 + 15) / 16 + 1;
    b = chunks * 8 + 48 + 128;
    return chunks > 1024 ? 0xffffffffu : (uint32_t)((b + 15) & ~15ull);
  }
  if (g.item_byte) b = ((uint64_t)P.tile_cap * j->nt + KMER_NT - 1) / KMER_NT + 2 * (uint64_t)P.k + 96; // the caller sized it for KMER_NT items
  else if (g.segs == 1) b = (uint64_t)j->nt * g.read_len + 96;
  else b = (uint64_t)j->nt * g.seg + ((uint64_t)j->nt / g.segs + 2) * (P.k - 1) + 96;
  return b > 0xffffffffull ? 0xffffffffu : (uint32_t)b;
}

// Build check without a GPU: compiles every variant (3-D box, 2-D tiles, ragged) for sm_100a.
bool seed_jit_compile_all(const SeedPlanHost& plan, std::string& why)
{
  for (int mode : { 1, 0, 2, 3, 11, 10 }) { // 1x: with strand outputs
    if (mode == 2 && plan.n_seeds * plan.h > 32) continue;
    if (mode >= 10 && plan.n_seeds > 8) continue;
    std::string w;
    SeedJit* j = seed_jit_build_variant(plan, w, false, mode % 10, mode >= 10);
    if (!j && (mode == 1 || w.rfind("NVRTC compilation failed", 0) == 0)) { // "does not apply" is fine for the side variants
      why = "variant " + std::to_string(mode) + ": " + w;
      return false;
    }
    seed_jit_destroy(j);
  }
  return true;
}

// The variant of the compiled kernel that fits the geometry: ragged batches take the ragged variant; uniform ones the
// 3-D box stores when rows are whole 64-byte blocks, else the 2-D tile variant.  Variants other than the one built with
// the plan are compiled on first use.
static const SeedJit* seed_jit_variant(const SeedJit* j, const KmerGeom& g, bool strands, bool reduce = false)
{
  if (!j) return j;
  if (reduce) { // fused consumer: uniform batches only
    if (g.item_byte) return nullptr;
    std::lock_guard<std::mutex> lock(j->mu);
    if (!j->alt_reduce) {
      std::string why;
      j->alt_reduce = seed_jit_build_variant(*j->plan, why, true, 3);
    }
    return j->alt_reduce;
  }
  if (g.item_byte) {
    if (strands || j->ht > 32 || getenv("NTHASH_B200_SEED_JIT_NO_RAGGED")) return nullptr; // a window must fit a 256-byte row piece
    std::lock_guard<std::mutex> lock(j->mu);
    if (!j->alt_ragged) {
      std::string why;
      j->alt_ragged = seed_jit_build_variant(*j->plan, why, true, 2);
    }
    return j->alt_ragged;
  }
  const bool box = j->box3 && ((uint64_t)g.seg * j->ht) % 8 == 0;
  if (!strands && (box || !j->box3)) return j;
  std::lock_guard<std::mutex> lock(j->mu);
  SeedJit*& slot = strands ? (box ? j->alt_str : j->alt_str2d) : j->alt;
  if (!slot) {
    std::string why;
    slot = seed_jit_build_variant(*j->plan, why, true, box ? 1 : 0, strands);
  }
  return slot;
}

// Uniform batches whose items are all full and whose rows are 16-byte multiples; ragged batches without strand outputs.
bool seed_jit_applies(const SeedJit* j0, const SeedParams& P)
{
  const KmerGeom& g = P.g;
  if (!j0 || !g.n_items || g.n_items >= 0x7fffffffull || ((uintptr_t)P.out & 15)) return false;
  if (P.out_fwd && (g.item_byte || getenv("NTHASH_B200_SEED_JIT_NO_STRANDS"))) return false;
  if (!g.item_byte && (!g.seg || g.nk % g.seg || ((uint64_t)g.seg * j0->ht) % 2)) return false;
  const SeedJit* j = seed_jit_variant(j0, g, P.out_fwd != nullptr);
  return j && seed_jit_smem_bytes(j, seed_jit_tile_cap(j, P)) <= 227u * 1024u;
}

// Fused consumer (P.reduce_out, P.item_dirty set; no outputs): uniform batches whose items are all full.
bool seed_jit_reduce_applies(const SeedJit* j0, const SeedParams& P)
{
  const KmerGeom& g = P.g;
  if (!j0 || !g.n_items || g.n_items >= 0x7fffffffull || g.item_byte || !g.seg || g.nk % g.seg) return false;
  if (getenv("NTHASH_B200_SEED_REDUCE_TWO_PASS")) return false;
  const SeedJit* j = seed_jit_variant(j0, g, false, true);
  return j && seed_jit_smem_bytes(j, seed_jit_tile_cap(j, P)) <= 227u * 1024u;
}

cudaError_t launch_seed_jit(const SeedJit* j0, const SeedParams& P, cudaStream_t st)
{
  const bool reduce = P.reduce_out != nullptr;
  const SeedJit* j = seed_jit_variant(j0, P.g, P.out_fwd != nullptr, reduce);
  if (!j) return cudaErrorNotSupported;
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
  CUtensorMap map;
  memset(&map, 0, sizeof map);
  const uint64_t row_u64 = (uint64_t)P.g.seg * j->ht;
  CUresult cr = CUDA_SUCCESS;
  if (j->ragged || j->reduce) {
    // no tensor map: rows leave through plain coalesced stores / nothing is stored
  } else if (j->box3) { // [row blocks][items][8 u64], boxes of (row_bytes / 64) x 32 x 8 under the 64-byte swizzle
    const cuuint64_t dims[3] = { 8, P.g.n_items, row_u64 / 8 };
    const cuuint64_t strides[2] = { row_u64 * 8, 64 };
    const cuuint32_t box[3] = { 8, 32, j->row_bytes / 64 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, P.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[2] = { row_u64, P.g.n_items };
    const cuuint64_t strides[1] = { row_u64 * 8 };
    const cuuint32_t box[2] = { j->tw * j->ht, 32 };
    const cuuint32_t estr[2] = { 1, 1 };
    cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, P.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;
  JitParams jp;
  jp.bases = P.bases;
  jp.n_bases = P.n_bases;
  jp.n_items = P.g.n_items;
  jp.read_len = P.g.read_len;
  jp.nk = P.g.nk;
  jp.seg = P.g.seg;
  jp.segs = P.g.segs;
  jp.out = P.out;
  jp.valid_bits = P.valid_bits;
  jp.read_dirty = P.read_dirty;
  jp.tables = j->d_tables;
  jp.care = reinterpret_cast<const uint32_t*>(P.plan_blob + P.care_off);
  jp.tile_cap = seed_jit_tile_cap(j, P);
  jp.care_words = P.care_words;
  jp.two = 2;
  jp.reduce_out = P.reduce_out;
  jp.item_dirty = P.item_dirty;
  jp.item_byte = P.g.item_byte;
  jp.item_out = P.g.item_out;
  jp.item_read = P.item_read;
  jp.out_fwd = P.out_fwd;
  jp.out_rev = P.out_rev;
  jp.str_aligned = (((uintptr_t)P.out_fwd | (uintptr_t)P.out_rev) & 31) == 0 ? 1u : 0u;
  const uint32_t smem = seed_jit_smem_bytes(j, jp.tile_cap);
  cudaError_t e = cudaFuncSetAttribute((const void*)j->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = { &jp, &map };
  const uint64_t ctas = (P.g.n_items + j->nt - 1) / j->nt;
  return cudaLaunchKernel((const void*)j->kernel, dim3((unsigned)ctas), dim3(j->nt), args, smem, st);
}

} // namespace nthb
