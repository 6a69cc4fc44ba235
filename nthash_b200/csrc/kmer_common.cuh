// kmer_common.cuh — pieces shared by the general and the fast NtHash kernels.
#pragma once
#include "engine.hpp"
#include "nthash_dev.cuh"

namespace nthb {

// forward / reverse strand hashes as (hi,lo) 32-bit register pairs
struct State
{
  uint32_t flo, fhi, rlo, rhi;
};

// F <- srol(F) ^ a ^ b on a (hi,lo) register pair.
NTH_D void fwd_step(State& s, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi)
{
  const uint32_t lo = s.flo, hi = s.fhi;
  const uint32_t nlo = (lo << 1) | (hi & 1u);
  const uint32_t nhi = (__funnelshift_l(lo, hi, 1) & ~2u) | ((hi >> 30) & 2u);
  s.flo = nlo ^ alo ^ blo;
  s.fhi = nhi ^ ahi ^ bhi;
}

// R <- sror(R ^ a ^ b)
NTH_D void rev_step(State& s, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi)
{
  const uint32_t lo = s.rlo ^ alo ^ blo, hi = s.rhi ^ ahi ^ bhi;
  s.rlo = __funnelshift_r(lo, hi, 1);
  s.rhi = ((hi >> 1) & 0x7FFFFFFEu) | (lo & 1u) | ((hi & 2u) << 30);
}

NTH_D uint64_t canonical(const State& s)
{
  return (((uint64_t)s.fhi << 32) | s.flo) + (((uint64_t)s.rhi << 32) | s.rlo);
}


} // namespace nthb
