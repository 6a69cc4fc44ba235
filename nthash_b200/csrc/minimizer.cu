// minimizer.cu — minimizer selection on top of the NtHash rows: the consumer most k-mer pipelines put right behind the
// hash path (nthash.hpp:14-17, :56-57 name "data structures" and btllib; minimizer sketches are what those tools keep).
//
// Definition used here (restated for the parity tests in tests/oracle_lib.py: minimizer_bits):
//   a window = w consecutive k-mers of one read (dense rows j .. j+w-1); its minimizer is the k-mer with the smallest
//   canonical hash hashes()[0] among those the reference's loop visits (valid_bits), the leftmost one on ties; a window
//   without any visited k-mer has none.  Row j's bit is set iff k-mer j is the minimizer of at least one window.
// Two kernels per chunk of reads that fits L2 (the hash rows of a chunk are written by kmer_fast_kernel and read back
// here while they are still cache-resident): mark the last row of every read, then one thread per window.
#include "engine.hpp"

#include <algorithm>

namespace nthb {

namespace {

constexpr int MZ_T = 256, MZ_PER = 4, MZ_TILE = MZ_T * MZ_PER; // windows per CTA

__global__ void mark_read_ends_uniform(uint32_t* end_bits, uint64_t n_reads, uint32_t nk)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint64_t row = (r + 1) * nk - 1;
  atomicOr(&end_bits[row >> 5], 1u << (row & 31));
}

__global__ void mark_read_ends_ragged(uint32_t* end_bits, const uint64_t* koff, uint64_t n_reads)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint64_t a = koff[r], b = koff[r + 1];
  if (b > a) atomicOr(&end_bits[(b - 1) >> 5], 1u << ((b - 1) & 31));
}

__device__ __forceinline__ uint64_t bits_at(const uint32_t* words, uint64_t bit) // 64 bits starting at `bit`
{
  const uint64_t w = bit >> 5;
  const uint32_t s = (uint32_t)(bit & 31);
  const uint64_t lo = words[w] | ((uint64_t)words[w + 1] << 32);
  return s ? (lo >> s) | ((uint64_t)words[w + 2] << (64 - s)) : lo;
}

// rows / valid / end_bits are chunk-local (row 0 = the chunk's first row; the bitmaps have two readable words of slack);
// min_bits is the caller's global bitmap, row0 the chunk's first row in it.
__global__ void __launch_bounds__(MZ_T)
minimizer_select_kernel(const uint64_t* __restrict__ rows, const uint32_t* __restrict__ valid, const uint32_t* __restrict__ end_bits,
                        uint64_t n_rows, uint32_t w, uint32_t* min_bits, uint64_t row0)
{
  __shared__ uint64_t sh[MZ_TILE + 64];
  __shared__ uint32_t sbits[MZ_TILE / 32 + 4];
  const uint64_t j0 = (uint64_t)blockIdx.x * MZ_TILE;
  const uint32_t span = (uint32_t)min((uint64_t)MZ_TILE + w - 1, n_rows - j0);
  for (uint32_t i = threadIdx.x; i < span; i += MZ_T) sh[i] = rows[j0 + i];
  for (uint32_t i = threadIdx.x; i < MZ_TILE / 32 + 4; i += MZ_T) sbits[i] = 0;
  __syncthreads();
  const uint64_t wmask = w == 64 ? ~0ull : (1ull << w) - 1, emask = wmask >> 1; // rows j..j+w-1; ends among j..j+w-2
  for (uint32_t q = 0; q < MZ_PER; ++q) {
    const uint32_t lj = q * MZ_T + threadIdx.x;
    const uint64_t j = j0 + lj;
    if (j + w > n_rows) continue;
    if (bits_at(end_bits, j) & emask) continue; // the window would cross into the next read
    uint64_t vb = bits_at(valid, j) & wmask;
    if (!vb) continue;
    uint64_t best = 0;
    uint32_t bpos = 0;
    bool found = false;
    while (vb) {
      const uint32_t t = __ffsll((long long)vb) - 1;
      vb &= vb - 1;
      const uint64_t v = sh[lj + t];
      if (!found || v < best) {
        best = v;
        bpos = t;
        found = true;
      }
    }
    atomicOr(&sbits[(lj + bpos) >> 5], 1u << ((lj + bpos) & 31));
  }
  __syncthreads();
  // the tile's words sit at an arbitrary bit offset of the global bitmap: OR them in (neighbouring tiles overlap by w-1 rows)
  const uint64_t gbit0 = row0 + j0;
  const uint32_t sh0 = (uint32_t)(gbit0 & 31);
  for (uint32_t i = threadIdx.x; i < MZ_TILE / 32 + 3; i += MZ_T) {
    const uint32_t cur = sbits[i], prev = i ? sbits[i - 1] : 0u;
    const uint32_t v = sh0 ? (cur << sh0) | (prev >> (32 - sh0)) : cur;
    if (v) atomicOr(&min_bits[(gbit0 >> 5) + i], v);
  }
}

__global__ void popcount_kernel(const uint32_t* bits, uint64_t n_bits, unsigned long long* count)
{
  const uint64_t nw = (n_bits + 31) / 32;
  uint32_t c = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t v = bits[i];
    if (i + 1 == nw && (n_bits & 31)) v &= (1u << (n_bits & 31)) - 1u;
    c += __popc(v);
  }
  for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long)c);
}

} // namespace

cudaError_t launch_mark_read_ends(uint32_t* d_end_bits, uint64_t rows, const uint64_t* d_koff, uint64_t n_reads, uint32_t uniform_nk,
                                  cudaStream_t st)
{
  cudaError_t e = cudaMemsetAsync(d_end_bits, 0, ((rows + 31) / 32 + 3) * 4, st);
  if (e != cudaSuccess || n_reads == 0) return e;
  const unsigned nb = (unsigned)((n_reads + 255) / 256);
  if (d_koff) mark_read_ends_ragged<<<nb, 256, 0, st>>>(d_end_bits, d_koff, n_reads);
  else mark_read_ends_uniform<<<nb, 256, 0, st>>>(d_end_bits, n_reads, uniform_nk);
  return cudaGetLastError();
}

cudaError_t launch_minimizer_select(const uint64_t* d_rows, const uint32_t* d_valid, const uint32_t* d_end_bits, uint64_t n_rows, uint32_t w,
                                    uint32_t* d_min_bits, uint64_t row0, cudaStream_t st)
{
  if (n_rows < w) return cudaSuccess;
  const uint64_t tiles = (n_rows + MZ_TILE - 1) / MZ_TILE;
  minimizer_select_kernel<<<(unsigned)tiles, MZ_T, 0, st>>>(d_rows, d_valid, d_end_bits, n_rows, w, d_min_bits, row0);
  return cudaGetLastError();
}

cudaError_t launch_popcount(const uint32_t* d_bits, uint64_t n_bits, uint64_t* d_count, cudaStream_t st)
{
  cudaError_t e = cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st);
  if (e != cudaSuccess || n_bits == 0) return e;
  popcount_kernel<<<(unsigned)std::min<uint64_t>(((n_bits + 31) / 32 + 255) / 256, 148 * 8), 256, 0, st>>>(d_bits, n_bits, reinterpret_cast<unsigned long long*>(d_count));
  return cudaGetLastError();
}

} // namespace nthb
