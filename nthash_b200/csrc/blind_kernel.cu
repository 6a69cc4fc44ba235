// blind_kernel.cu — BlindNtHash::roll(char_in) / peek(char_in) over n independent states.
//
// Reference: BlindNtHash::roll, src/kmer.cpp:355-364 (next_forward_hash :84-94,
// next_reverse_hash :164-174, extend_hashes src/internal.hpp:104-118).  The reference object owns
// a deque of the current k-mer only to know its first base; a traversal engine keeps that base
// (out_base) itself, so a state is (fwd, rev).  Like the reference, no validity check is made on
// char_in: any byte outside SEED_TAB's non-zero entries contributes a zero seed.
// One thread per state; 16-32 bytes in, 16+8h bytes out: a pure streaming kernel.
#include "engine.hpp"
#include "nthash_dev.cuh"

namespace nthb {

namespace {

struct BlindConsts
{
  uint32_t k, h;
  uint64_t s[5], sk[5]; // seed and srol^k(seed) for A, C, G, T, none
};

// SEED_TAB as an index: 0..3 = A,C,G,T (incl. lower case, U, and the raw complement slots), 4 = zero seed
NTH_D int seed_index(unsigned c)
{
  switch (c) {
    case 'A': case 'a': case 4: case 5: return 0;
    case 'C': case 'c': case 7: return 1;
    case 'G': case 'g': case 3: return 2;
    case 'T': case 't': case 'U': case 'u': case 1: return 3;
    default: return 4;
  }
}

template<bool PEEK4>
__global__ void __launch_bounds__(256)
blind_kernel(uint64_t* __restrict__ fwd, uint64_t* __restrict__ rev, const uint8_t* __restrict__ out_base,
             const uint8_t* __restrict__ in_base, uint64_t n, BlindConsts c, uint64_t* __restrict__ out)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t f0 = fwd[i], r0 = rev[i];
  const unsigned co = out_base[i];
  const uint64_t f_base = srol1(f0) ^ c.sk[seed_index(co)];
  const uint64_t r_base = r0 ^ c.s[seed_index(co & 7u)];
  const int n_ext = PEEK4 ? 4 : 1;
#pragma unroll
  for (int e = 0; e < n_ext; ++e) {
    const unsigned ci = PEEK4 ? (unsigned)"ACGT"[e] : (unsigned)in_base[i];
    const uint64_t f = f_base ^ c.s[seed_index(ci)];
    const uint64_t r = sror1(r_base ^ c.sk[seed_index(ci & 7u)]);
    const uint64_t h0 = f + r;
    uint64_t* o = out + (i * n_ext + e) * c.h;
    o[0] = h0;
    for (uint32_t q = 1; q < c.h; ++q) o[q] = ext_hash(h0, ext_mult(q, c.k));
    if (!PEEK4) {
      fwd[i] = f;
      rev[i] = r;
    }
  }
}

// BlindSeedNtHash::roll(char_in) keeps the k-mer itself (seed.cpp:701-718): drop its first base, append char_in.
// One thread per state; the hashing of the new windows is the SeedNtHash batch path on n reads of length k.
__global__ void __launch_bounds__(256)
blind_seed_shift_kernel(uint8_t* __restrict__ kmers, const uint8_t* __restrict__ in_base, uint64_t n, uint32_t k)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t* w = kmers + i * k;
  for (uint32_t j = 0; j + 1 < k; ++j) w[j] = w[j + 1];
  w[k - 1] = in_base[i];
}

} // namespace

cudaError_t launch_blind_seed_shift(uint8_t* kmers, const uint8_t* in_base, uint64_t n, uint32_t k, cudaStream_t st)
{
  if (n == 0) return cudaSuccess;
  const uint64_t blocks = (n + 255) / 256;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  blind_seed_shift_kernel<<<(unsigned)blocks, 256, 0, st>>>(kmers, in_base, n, k);
  return cudaGetLastError();
}

cudaError_t launch_blind(uint64_t* fwd, uint64_t* rev, const uint8_t* out_base, const uint8_t* in_base, uint64_t n,
                         uint32_t k, uint32_t h, uint64_t* out, bool peek4, cudaStream_t st)
{
  if (n == 0) return cudaSuccess;
  const uint64_t blocks = (n + 255) / 256;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  BlindConsts c;
  c.k = k;
  c.h = h;
  const uint64_t base[5] = { SEED_A, SEED_C, SEED_G, SEED_T, 0 };
  for (int x = 0; x < 5; ++x) {
    c.s[x] = base[x];
    c.sk[x] = srol_n(base[x], k);
  }
  if (peek4) blind_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(fwd, rev, out_base, in_base, n, c, out);
  else blind_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(fwd, rev, out_base, in_base, n, c, out);
  return cudaGetLastError();
}

} // namespace nthb
