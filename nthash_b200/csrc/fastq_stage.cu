// fastq_stage.cu — the step before the hash path: FASTQ text that already sits in device
// memory -> the concatenated `bases` + `read_off` layout every batch entry point takes.  No host-side parsing and no
// library calls: the newline positions come from a three-launch select (per-tile counts, one-block scan of the tile
// sums, ordered write), read_off from the same scan over the sequence-line lengths, and the sequence lines (every 4th
// line, starting at the 2nd) are copied out by one warp per read.  This is staging, not the hot path.  Records are the
// plain four-line form (`@id`, sequence, `+`, qualities), LF or CRLF line ends.
#include "engine.hpp"

namespace nthb {

namespace {

constexpr uint32_t FQ_T = 256;                  // threads per block
constexpr uint32_t FQ_TILE_BYTES = FQ_T * 16;   // select: 16 bytes of text per thread
constexpr uint32_t FQ_VALS = 8;                 // scan: values per thread

// bytes of w equal to '\n', one flag bit per byte (bit 7 of the byte)
__device__ __forceinline__ uint32_t nl_flags(uint32_t w)
{
  const uint32_t x = w ^ 0x0A0A0A0Au;
  return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}

// 16 bytes of text starting at byte i (any alignment, bounds-checked) as four little-endian words; bytes past the end read as 0
__device__ __forceinline__ uint4 text16(const uint8_t* text, uint64_t n_bytes, uint64_t i)
{
  if (i + 16 <= n_bytes && ((uintptr_t)(text + i) & 15) == 0) return *reinterpret_cast<const uint4*>(text + i);
  uint32_t w[4] = { 0, 0, 0, 0 };
  for (uint32_t j = 0; j < 16 && i + j < n_bytes; ++j) w[j >> 2] |= (uint32_t)text[i + j] << (8 * (j & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// exclusive scan of one value per thread across the block; returns this thread's offset, *total = the block's sum
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t* total)
{
  __shared__ uint64_t wsum[FQ_T / 32];
  __shared__ uint64_t btotal;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t x = v;
  for (uint32_t o = 1; o < 32; o <<= 1) {
    const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint64_t s = lane < FQ_T / 32 ? wsum[lane] : 0, t = s;
    for (uint32_t o = 1; o < 32; o <<= 1) {
      const uint64_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    if (lane < FQ_T / 32) wsum[lane] = t - s;
    if (lane == 31) btotal = t;
  }
  __syncthreads();
  const uint64_t r = wsum[warp] + x - v;
  if (total) *total = btotal;
  __syncthreads(); // wsum / btotal may be reused by the caller's next scan
  return r;
}

// ---- select: positions of the newlines, in order ----
__global__ void __launch_bounds__(FQ_T) nl_tile_counts(const uint8_t* text, uint64_t n_bytes, uint64_t* tile_count)
{
  const uint64_t i = ((uint64_t)blockIdx.x * FQ_T + threadIdx.x) * 16;
  uint32_t c = 0;
  if (i < n_bytes) {
    const uint4 v = text16(text, n_bytes, i);
    c = __popc(nl_flags(v.x)) + __popc(nl_flags(v.y)) + __popc(nl_flags(v.z)) + __popc(nl_flags(v.w));
  }
  uint64_t total;
  block_excl_scan(c, &total);
  if (threadIdx.x == 0) tile_count[blockIdx.x] = total;
}

// one block: sums[0..n) -> exclusive offsets in place, the grand total to *total
__global__ void __launch_bounds__(FQ_T) scan_tile_sums(uint64_t* sums, uint64_t n, uint64_t* total)
{
  uint64_t carry = 0;
  for (uint64_t base = 0; base < n; base += (uint64_t)FQ_T * FQ_VALS) {
    uint64_t v[FQ_VALS], mine = 0;
    for (uint32_t j = 0; j < FQ_VALS; ++j) {
      const uint64_t idx = base + (uint64_t)threadIdx.x * FQ_VALS + j;
      v[j] = idx < n ? sums[idx] : 0;
      mine += v[j];
    }
    uint64_t chunk_total;
    uint64_t off = carry + block_excl_scan(mine, &chunk_total);
    for (uint32_t j = 0; j < FQ_VALS; ++j) {
      const uint64_t idx = base + (uint64_t)threadIdx.x * FQ_VALS + j;
      if (idx < n) sums[idx] = off;
      off += v[j];
    }
    carry += chunk_total;
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(FQ_T) nl_write(const uint8_t* text, uint64_t n_bytes, const uint64_t* tile_off, uint64_t* nl)
{
  const uint64_t i = ((uint64_t)blockIdx.x * FQ_T + threadIdx.x) * 16;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (i < n_bytes) v = text16(text, n_bytes, i);
  const uint32_t f[4] = { nl_flags(v.x), nl_flags(v.y), nl_flags(v.z), nl_flags(v.w) };
  const uint32_t c = __popc(f[0]) + __popc(f[1]) + __popc(f[2]) + __popc(f[3]);
  uint64_t o = tile_off[blockIdx.x] + block_excl_scan(c, nullptr);
  for (uint32_t w = 0; w < 4; ++w)
    for (uint32_t m = f[w]; m; m &= m - 1) nl[o++] = i + 4 * w + ((__ffs(m) - 1) >> 3);
}

// ---- exclusive sum of per-read lengths: out[0..n] (out[n] = total) ----
__global__ void __launch_bounds__(FQ_T) len_tile_sums(const uint64_t* len, uint64_t n, uint64_t* tile_sum)
{
  uint64_t mine = 0;
  for (uint32_t j = 0; j < FQ_VALS; ++j) {
    const uint64_t idx = ((uint64_t)blockIdx.x * FQ_T + threadIdx.x) * FQ_VALS + j;
    if (idx < n) mine += len[idx];
  }
  uint64_t total;
  block_excl_scan(mine, &total);
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(FQ_T) len_scan_write(const uint64_t* len, uint64_t n, const uint64_t* tile_off, const uint64_t* total, uint64_t* out)
{
  uint64_t v[FQ_VALS], mine = 0;
  const uint64_t first = ((uint64_t)blockIdx.x * FQ_T + threadIdx.x) * FQ_VALS;
  for (uint32_t j = 0; j < FQ_VALS; ++j) {
    v[j] = first + j < n ? len[first + j] : 0;
    mine += v[j];
  }
  uint64_t off = tile_off[blockIdx.x] + block_excl_scan(mine, nullptr);
  for (uint32_t j = 0; j < FQ_VALS; ++j) {
    if (first + j < n) out[first + j] = off;
    off += v[j];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;
}

// len[r] = length of the sequence line of record r (CR stripped); start[r] = its first byte
__global__ void fastq_seq_lines(const uint8_t* text, uint64_t n_bytes, const uint64_t* nl, uint64_t n_nl, uint64_t n_reads,
                                uint64_t* start, uint64_t* len)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint64_t s = nl[4 * r] + 1;
  uint64_t e = 4 * r + 1 < n_nl ? nl[4 * r + 1] : n_bytes; // a last line without a newline ends at the end of the text
  if (e > s && text[e - 1] == '\r') --e;
  start[r] = s;
  len[r] = e - s;
}

__global__ void fastq_copy_reads(const uint8_t* text, const uint64_t* start, const uint64_t* read_off, uint64_t n_reads, uint8_t* bases)
{
  const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (r >= n_reads) return;
  const uint64_t s = start[r], o = read_off[r], n = read_off[r + 1] - o;
  for (uint64_t j = lane; j < n; j += 32) bases[o + j] = text[s + j];
}

} // namespace

// Returns the number of reads and bases through host pointers (synchronises the stream once, to size the second pass).
cudaError_t fastq_extract(const uint8_t* d_text, uint64_t n_bytes, uint8_t* d_bases, uint64_t bases_capacity, uint64_t* d_read_off,
                          uint64_t reads_capacity, uint64_t* n_reads_out, uint64_t* n_bases_out, cudaStream_t st)
{
  *n_reads_out = 0;
  *n_bases_out = 0;
  if (n_bytes == 0) return cudaMemsetAsync(d_read_off, 0, sizeof(uint64_t), st);
  uint64_t *d_nl = nullptr, *d_cnt = nullptr, *d_start = nullptr, *d_len = nullptr;
  void* d_tmp = nullptr; // the tile sums of whichever scan is in flight
  auto done = [&](cudaError_t e) {
    for (void* p : { (void*)d_nl, (void*)d_cnt, (void*)d_start, (void*)d_len, d_tmp })
      if (p) cudaFreeAsync(p, st);
    return e;
  };
  // 1. newline positions: per-tile counts, scan of the tile sums (one block), ordered write
  const uint64_t n_tiles = (n_bytes + FQ_TILE_BYTES - 1) / FQ_TILE_BYTES;
  if (n_tiles > 0x7fffffffull) return cudaErrorInvalidValue;
  uint64_t* d_tile = nullptr;
  cudaError_t e = cudaMallocAsync(&d_cnt, sizeof(uint64_t), st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_tile, n_tiles * sizeof(uint64_t), st);
  d_tmp = d_tile;
  if (e != cudaSuccess) return done(e);
  nl_tile_counts<<<(unsigned)n_tiles, FQ_T, 0, st>>>(d_text, n_bytes, d_tile);
  scan_tile_sums<<<1, FQ_T, 0, st>>>(d_tile, n_tiles, d_cnt);
  uint64_t n_nl = 0;
  uint8_t last = 0;
  e = cudaMemcpyAsync(&n_nl, d_cnt, sizeof n_nl, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&last, d_text + n_bytes - 1, 1, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_nl, (n_nl + 1) * sizeof(uint64_t), st);
  if (e != cudaSuccess) return done(e);
  nl_write<<<(unsigned)n_tiles, FQ_T, 0, st>>>(d_text, n_bytes, d_tile, d_nl);
  e = cudaGetLastError();
  if (e != cudaSuccess) return done(e);
  const uint64_t n_lines = n_nl + (last != '\n' ? 1 : 0);
  const uint64_t n_reads = n_lines / 4;
  if (n_reads == 0) return done(cudaMemsetAsync(d_read_off, 0, sizeof(uint64_t), st));
  if (n_reads > reads_capacity) return done(cudaErrorInvalidValue);
  // 2. sequence line of every record, 3. read_off = exclusive sum of the lengths (+ total), 4. copy
  e = cudaMallocAsync(&d_start, n_reads * sizeof(uint64_t), st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_len, (n_reads + 1) * sizeof(uint64_t), st);
  if (e != cudaSuccess) return done(e);
  fastq_seq_lines<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(d_text, n_bytes, d_nl, n_nl, n_reads, d_start, d_len);
  // read_off = exclusive sum of the lengths (+ total): the same three launches over the lengths
  cudaFreeAsync(d_tmp, st);
  d_tmp = nullptr;
  const uint64_t n_ltiles = (n_reads + (uint64_t)FQ_T * FQ_VALS - 1) / ((uint64_t)FQ_T * FQ_VALS);
  uint64_t* d_ltile = nullptr;
  e = cudaMallocAsync(&d_ltile, n_ltiles * sizeof(uint64_t), st);
  d_tmp = d_ltile;
  if (e != cudaSuccess) return done(e);
  len_tile_sums<<<(unsigned)n_ltiles, FQ_T, 0, st>>>(d_len, n_reads, d_ltile);
  scan_tile_sums<<<1, FQ_T, 0, st>>>(d_ltile, n_ltiles, d_cnt);
  len_scan_write<<<(unsigned)n_ltiles, FQ_T, 0, st>>>(d_len, n_reads, d_ltile, d_cnt, d_read_off);
  e = cudaGetLastError();
  uint64_t n_bases = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n_bases, d_read_off + n_reads, sizeof n_bases, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return done(e);
  if (n_bases > bases_capacity) return done(cudaErrorInvalidValue);
  const uint64_t threads = n_reads * 32;
  fastq_copy_reads<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_text, d_start, d_read_off, n_reads, d_bases);
  e = cudaGetLastError();
  *n_reads_out = n_reads;
  *n_bases_out = n_bases;
  return done(e);
}

} // namespace nthb
