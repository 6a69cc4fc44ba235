// nthash_dev.cuh — device/host primitives of the B200 ntHash engine.
//
// Arithmetic contract (what the kernels must reproduce bit for bit), with the
// reference locations it replaces (/root/reference = bcgsc/ntHash 2.4.0):
//   srol1 / sror1      src/internal.hpp:41-47, :83-88   split rotate of bits [63:33] | [32:0]
//   srol_n             src/internal.hpp:343-348         srol_table(c, d) == srol_n(seed(c), d)
//   seed constants     src/internal.hpp:124-128
//   ext_hash           src/internal.hpp:104-118         NTM64 multi-hash extension
// Nothing here is shared with oracle/: the oracle is the checker, this is the product.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define NTH_HD __host__ __device__ __forceinline__
#define NTH_D __device__ __forceinline__

namespace nthb {

constexpr uint64_t SEED_A = 0x3c8bfbb395c60474ULL;
constexpr uint64_t SEED_C = 0x3193c18562a02b4cULL;
constexpr uint64_t SEED_G = 0x20323ed082572324ULL;
constexpr uint64_t SEED_T = 0x295549f54be24456ULL;
constexpr uint64_t MULTISEED = 0x90b45d39fb6da1faULL;
constexpr int MULTISHIFT = 27;

// Rotate the 31-bit word [63:33] and the 33-bit word [32:0] left by one.
NTH_HD uint64_t srol1(uint64_t x)
{
  return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | ((x >> 63) << 33) | ((x >> 32) & 1ULL);
}

// Inverse of srol1.
NTH_HD uint64_t sror1(uint64_t x)
{
  return ((x >> 1) & 0x7FFFFFFEFFFFFFFFULL) | (((x >> 33) & 1ULL) << 63) | ((x & 1ULL) << 32);
}

// d-fold srol1 for any d (the two words have periods 33 and 31).
NTH_HD uint64_t srol_n(uint64_t x, unsigned d)
{
  const uint64_t m33 = (1ULL << 33) - 1, m31 = (1ULL << 31) - 1;
  uint64_t lo = x & m33, hi = x >> 33;
  const unsigned a = d % 33, b = d % 31;
  if (a) lo = ((lo << a) | (lo >> (33 - a))) & m33;
  if (b) hi = ((hi << b) | (hi >> (31 - b))) & m31;
  return (hi << 33) | lo;
}

// Seed of a byte as SEED_TAB maps it (ACGTU either case; raw bytes 1,3,4,5,7 are the
// complement slots the reference reaches through `c & 7`); 0 for everything else.
NTH_HD uint64_t seed_of_byte(unsigned c)
{
  switch (c) {
    case 'A': case 'a': case 4: case 5: return SEED_A;
    case 'C': case 'c': case 7: return SEED_C;
    case 'G': case 'g': case 3: return SEED_G;
    case 'T': case 't': case 'U': case 'u': case 1: return SEED_T;
    default: return 0;
  }
}

// NtHash's notion of a hashable base (raw bytes 1,3,4,5,7 are rejected).
NTH_HD bool is_acgtu(unsigned c)
{
  switch (c) {
    case 'A': case 'a': case 'C': case 'c': case 'G': case 'g':
    case 'T': case 't': case 'U': case 'u': return true;
    default: return false;
  }
}

// Extra hash i >= 1 from the canonical hash: multiplier is i ^ (k * MULTISEED).
NTH_HD uint64_t ext_mult(unsigned i, unsigned k) { return (uint64_t)i ^ ((uint64_t)k * MULTISEED); }
NTH_HD uint64_t ext_hash(uint64_t h0, uint64_t mult)
{
  uint64_t t = h0 * mult;
  return t ^ (t >> MULTISHIFT);
}

// ---------------------------------------------------------------- PTX helpers --
#ifdef __CUDACC__
NTH_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

NTH_D void mbar_init(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
NTH_D void fence_mbar_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
NTH_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
NTH_D void mbar_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t done;
  do {
    asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  } while (!done);
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP).  dst, src and bytes
// must be multiples of 16.
NTH_D void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                 smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
NTH_D void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
// 32-byte (one full sector) global store; SASS: STG.E.ENL2.256 (sm_100+).
NTH_D void st_global_v4_u64(uint64_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d)
{
  asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
NTH_D void st_shared_v2_u64(uint32_t saddr, uint64_t a, uint64_t b)
{
  asm volatile("st.shared.v2.u64 [%0], {%1,%2};" ::"r"(saddr), "l"(a), "l"(b) : "memory");
}
// ---- TMA tile store (shared -> global through a tensor map; SASS: UTMASTG) ----
NTH_D void tma_store_2d(const void* tmap, uint32_t saddr, int c0, int c1)
{
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0),
               "r"(c1), "r"(saddr)
               : "memory");
}
NTH_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
NTH_D void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
NTH_D void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
NTH_D void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
NTH_D void st_global_v2_u64(uint64_t* p, uint64_t a, uint64_t b)
{
  asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
#endif

} // namespace nthb
