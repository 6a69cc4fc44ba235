// prep_kernels.cu — ragged-batch bookkeeping on the device: the dense output offsets
// koff[r] = sum_{r'<r} max(0, len_r' - k + 1) (the position arithmetic the reference does
// implicitly with get_pos(), include/nthash/nthash.hpp:170), and the expansion of long reads
// into seg-window items.  Plain three-pass block scan; this is plumbing, not the hot loop.
#include "engine.hpp"

#include <algorithm>

namespace nthb {

namespace {

constexpr int SCAN_T = 256, SCAN_E = 8, SCAN_B = SCAN_T * SCAN_E;

__device__ __forceinline__ uint64_t read_value(const uint64_t* read_off, uint64_t r, uint32_t k, uint32_t seg,
                                               uint64_t& len)
{
  len = read_off[r + 1] - read_off[r];
  const uint64_t nk = len >= k ? len - k + 1 : 0;
  // item mode: a read without windows still gets one (empty) item, so that its bytes are counted when a CTA's staged
  // span is bounded by items x (seg + k - 1) bytes (many reads shorter than k next to a long one used to overflow it)
  return seg ? (nk ? (nk + seg - 1) / seg : 1) : nk;
}

__device__ __forceinline__ uint64_t block_reduce_sum_max(uint64_t v, uint64_t& mx)
{
  __shared__ uint64_t ws[SCAN_T / 32], wm[SCAN_T / 32];
  for (int o = 16; o; o >>= 1) {
    v += __shfl_down_sync(0xffffffffu, v, o);
    mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    ws[threadIdx.x >> 5] = v;
    wm[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < SCAN_T / 32; ++w) {
      v += ws[w];
      mx = max(mx, wm[w]);
    }
  }
  return v;
}

__global__ void __launch_bounds__(SCAN_T)
scan_block_sums(const uint64_t* read_off, uint64_t n, uint32_t k, uint32_t seg, uint64_t* bsum, uint64_t* bmax)
{
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_B + (uint64_t)threadIdx.x * SCAN_E;
  uint64_t s = 0, mx = 0;
  for (int e = 0; e < SCAN_E; ++e) {
    if (base + e < n) {
      uint64_t len;
      s += read_value(read_off, base + e, k, seg, len);
      mx = max(mx, len);
    }
  }
  s = block_reduce_sum_max(s, mx);
  if (threadIdx.x == 0) {
    bsum[blockIdx.x] = s;
    bmax[blockIdx.x] = mx;
  }
}

// one block: bsum -> exclusive offsets in place; stats[0] = total, stats[1] = max len
__global__ void __launch_bounds__(1024) scan_of_sums(uint64_t* bsum, const uint64_t* bmax, uint64_t nb, uint64_t* stats)
{
  __shared__ uint64_t ws[32];
  __shared__ uint64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  uint64_t mx = 0;
  for (uint64_t c0 = 0; c0 < nb; c0 += 1024) {
    const uint64_t i = c0 + threadIdx.x;
    const uint64_t v = i < nb ? bsum[i] : 0;
    if (i < nb) mx = max(mx, bmax[i]);
    uint64_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint64_t t = ws[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        const uint64_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (threadIdx.x >= o) t += y;
      }
      ws[threadIdx.x] = t;
    }
    __syncthreads();
    const uint64_t warp_off = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
    const uint64_t carry = carry_s;
    if (i < nb) bsum[i] = carry + warp_off + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
    __syncthreads();
  }
  for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w) mx = max(mx, ws[w]);
    stats[0] = carry_s;
    stats[1] = mx;
  }
}

__global__ void __launch_bounds__(SCAN_T)
scan_write(const uint64_t* read_off, uint64_t n, uint32_t k, uint32_t seg, const uint64_t* boff, uint64_t* excl)
{
  __shared__ uint64_t ws[SCAN_T / 32];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_B + (uint64_t)threadIdx.x * SCAN_E;
  uint64_t v[SCAN_E], tsum = 0;
  for (int e = 0; e < SCAN_E; ++e) {
    uint64_t len;
    v[e] = base + e < n ? read_value(read_off, base + e, k, seg, len) : 0;
    tsum += v[e];
  }
  uint64_t x = tsum;
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
  __syncthreads();
  uint64_t off = boff[blockIdx.x] + x - tsum;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += ws[w];
  for (int e = 0; e < SCAN_E; ++e) {
    if (base + e < n) excl[base + e] = off;
    off += v[e];
    if (base + e == n - 1) excl[n] = off;
  }
}

// `cap` = entries the table was sized for (a host-side upper bound of the item count, so that nothing has to be read
// back): entries [ioff[n], cap] are padded with empty items at the end of the batch.
__global__ void item_fill(const uint64_t* read_off, const uint64_t* koff, const uint64_t* ioff, uint64_t n,
                          uint32_t k, uint32_t seg, uint64_t* item_byte, uint64_t* item_out, uint64_t* item_read,
                          uint64_t cap)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (uint64_t i = ioff[n] + r; i <= cap; i += (uint64_t)gridDim.x * blockDim.x) {
    item_byte[i] = read_off[n];
    item_out[i] = koff[n];
    item_read[i] = n ? n - 1 : 0;
  }
  if (r >= n) return;
  const uint64_t i0 = ioff[r], i1 = ioff[r + 1];
  for (uint64_t i = i0; i < i1; ++i) {
    item_byte[i] = read_off[r] + (i - i0) * seg;
    item_out[i] = koff[r] + (i - i0) * seg;
    item_read[i] = r;
  }
}

// 2-bit packed bases -> the ASCII bytes the hash kernels read.  Base g of the packed stream is bits 2*(g & 3).. of byte
// g >> 2 (0 = A, 1 = C, 2 = G, 3 = T); bit g of `invalid` (optional) marks a base that is not ACGT and comes out as 'N',
// which is what keeps the reference's byte-level rules (windows with such a base are skipped by NtHash, hashed as a
// zero seed by SeedNtHash).  Thread t writes output bytes [16 t, 16 t + 16) with one 16-byte store.
__global__ void unpack2bit_kernel(const uint8_t* packed, const uint32_t* invalid, uint64_t first_base, uint64_t n_bases, uint8_t* out)
{
  const uint64_t j0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (j0 >= n_bases) return;
  uint32_t w[4] = { 0, 0, 0, 0 };
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint64_t g = first_base + j0 + j;
    uint32_t c = 'A';
    if (j0 + j < n_bases) {
      const uint32_t code = (packed[g >> 2] >> (2 * (g & 3))) & 3u;
      c = (0x54474341u >> (8 * code)) & 0xFFu; // "ACGT"
      if (invalid && (invalid[g >> 5] >> (g & 31) & 1u)) c = 'N';
    }
    w[j >> 2] |= c << (8 * (j & 3));
  }
  *reinterpret_cast<uint4*>(out + j0) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- compaction of the dense rows to the windows the reference's loop visits, in its order --------------------
constexpr int CMP_T = 256;                   // threads per block: 8 warps x 32 bitmap words x 32 rows = 8192 rows
constexpr uint64_t CMP_ROWS = CMP_T * 32ull;

__global__ void compact_count(const uint32_t* valid, uint64_t rows, uint64_t* block_sums)
{
  const uint64_t w = (uint64_t)blockIdx.x * CMP_T + threadIdx.x, nw = (rows + 31) / 32;
  uint32_t v = w < nw ? valid[w] : 0u;
  if (w + 1 == nw && (rows & 31)) v &= (1u << (rows & 31)) - 1u;
  uint32_t c = __popc(v);
  for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  __shared__ uint32_t ws[CMP_T / 32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < CMP_T / 32; ++i) t += ws[i];
    block_sums[blockIdx.x] = t;
  }
}

// one block: exclusive scan of the block sums in place, total to *count
// `append`: *count holds the number of entries already written by earlier chunks: offsets start there and it is advanced
__global__ void compact_scan(uint64_t* block_sums, uint64_t nb, uint64_t* count, bool append)
{
  __shared__ uint64_t carry;
  __shared__ uint64_t ws[32];
  if (threadIdx.x == 0) carry = append ? *count : 0;
  __syncthreads();
  for (uint64_t base = 0; base < nb; base += blockDim.x) {
    const uint64_t i = base + threadIdx.x;
    const uint64_t v = i < nb ? block_sums[i] : 0;
    uint64_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    uint64_t off = carry;
    for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) off += ws[w];
    if (i < nb) block_sums[i] = off + x - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

// every warp walks its 32 bitmap words; for one word the 32 lanes are 32 consecutive rows (coalesced reads), the kept
// ones land at consecutive compact rows
__global__ void compact_write(const uint64_t* out, const uint32_t* valid, uint64_t rows, uint32_t H, const uint64_t* block_off,
                              uint64_t* compact, uint64_t* row_index, uint64_t row_offset, uint64_t capacity)
{
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t w0 = (uint64_t)blockIdx.x * CMP_T + warp * 32, nw = (rows + 31) / 32;
  // this lane's word -> popcount -> exclusive scan inside the warp, then across the block's warps
  uint32_t v = w0 + lane < nw ? valid[w0 + lane] : 0u;
  if (w0 + lane + 1 == nw && (rows & 31)) v &= (1u << (rows & 31)) - 1u;
  const uint32_t c = __popc(v);
  uint32_t x = c;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  __shared__ uint32_t ws[CMP_T / 32];
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  uint64_t base = block_off[blockIdx.x];
  for (uint32_t i = 0; i < warp; ++i) base += ws[i];
  const uint32_t excl = x - c;
  for (uint32_t j = 0; j < 32; ++j) {
    const uint32_t word = __shfl_sync(0xffffffffu, v, j);
    if (!word) continue;
    const uint64_t dst0 = base + __shfl_sync(0xffffffffu, excl, j);
    if (word >> lane & 1u) {
      const uint64_t r = (w0 + j) * 32 + lane, d = dst0 + __popc(word & ((1u << lane) - 1u));
      if (d >= capacity) continue;
      if (compact)
        for (uint32_t q = 0; q < H; ++q) compact[d * H + q] = out[r * H + q];
      if (row_index) row_index[d] = row_offset + r;
    }
  }
}

// count / sum / xor over the rows whose validity bit is set (all H values of a row): the consumer of the reference's
// benchmark loop for outputs that were written to device memory first (SeedNtHash: no fused form yet)
__global__ void __launch_bounds__(256)
reduce_rows_kernel(const uint64_t* out, const uint32_t* valid, uint64_t valid_row0, uint64_t rows, uint32_t H, uint32_t magic,
                   unsigned long long* res)
{
  uint64_t sum = 0, x = 0;
  uint32_t cnt = 0;
  const uint32_t lane = threadIdx.x & 31, span = 32 * H; // a warp walks 32 rows = 32*H contiguous values at a time
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5, w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (uint64_t r0 = w * 32; r0 < rows; r0 += warps * 32) {
    const uint64_t nvals = min((uint64_t)span, (rows - r0) * H);
    for (uint32_t i = lane; i < nvals; i += 32) {
      const uint32_t q = H == 1 ? i : __umulhi(i, magic); // i / H, exact for i < 2^16 (H = 1: the reciprocal 2^32 does not fit 32 bits)
      const uint64_t row = r0 + q, bit = valid_row0 + row;
      if (valid[bit >> 5] >> (bit & 31) & 1u) {
        const uint64_t v = out[r0 * H + i];
        sum += v;
        x ^= v;
        cnt += (i - q * H) == 0;
      }
    }
  }
  for (int o = 16; o; o >>= 1) {
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    sum += __shfl_down_sync(0xffffffffu, sum, o);
    x ^= __shfl_down_sync(0xffffffffu, x, o);
  }
  if (lane == 0 && (cnt | sum | x)) {
    atomicAdd(res, (unsigned long long)cnt);
    atomicAdd(res + 1, (unsigned long long)sum);
    atomicXor(res + 2, (unsigned long long)x);
  }
}

} // namespace

cudaError_t launch_reduce_rows(const uint64_t* d_out, const uint32_t* d_valid, uint64_t valid_row0, uint64_t rows, uint32_t H,
                               uint64_t* d_result, cudaStream_t st)
{
  if (rows == 0) return cudaSuccess;
  if (H == 0 || H > 2040) return cudaErrorInvalidValue; // 32*H must stay below 2^16 for the reciprocal division
  const uint32_t magic = (uint32_t)((0x100000000ull + H - 1) / H);
  const uint64_t blocks = std::min<uint64_t>((rows + 255) / 256, 148ull * 16);
  reduce_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_out, d_valid, valid_row0, rows, H, magic, reinterpret_cast<unsigned long long*>(d_result));
  return cudaGetLastError();
}

// Keeps the rows whose validity bit is set, in order: compact[n][H] (+ their dense row numbers), n to *d_count.
cudaError_t launch_compact_rows(const uint64_t* d_out, const uint32_t* d_valid, uint64_t rows, uint32_t H, uint64_t* d_compact,
                                uint64_t* d_row_index, uint64_t* d_count, cudaStream_t st)
{
  return launch_compact_rows_at(d_out, d_valid, rows, H, d_compact, d_row_index, d_count, false, 0, ~0ull, st);
}

// append = true: entries go behind the *d_count already there (chunked producers on one stream), row numbers are shifted
// by row_offset, and nothing is written past `capacity` entries (the count keeps counting).
cudaError_t launch_compact_rows_at(const uint64_t* d_out, const uint32_t* d_valid, uint64_t rows, uint32_t H, uint64_t* d_compact,
                                   uint64_t* d_row_index, uint64_t* d_count, bool append, uint64_t row_offset, uint64_t capacity,
                                   cudaStream_t st)
{
  if (rows == 0) return append ? cudaSuccess : cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st);
  const uint64_t nb = (rows + CMP_ROWS - 1) / CMP_ROWS;
  uint64_t* sums = nullptr;
  cudaError_t e = cudaMallocAsync(&sums, nb * sizeof(uint64_t), st);
  if (e != cudaSuccess) return e;
  compact_count<<<(unsigned)nb, CMP_T, 0, st>>>(d_valid, rows, sums);
  compact_scan<<<1, 1024, 0, st>>>(sums, nb, d_count, append);
  compact_write<<<(unsigned)nb, CMP_T, 0, st>>>(d_out, d_valid, rows, H, sums, d_compact, d_row_index, row_offset, capacity);
  e = cudaGetLastError();
  cudaFreeAsync(sums, st);
  return e;
}

// d_out must be 16-byte aligned and writable up to the next multiple of 16 bytes past n_bases.
cudaError_t launch_unpack2bit(const uint8_t* d_packed, const uint32_t* d_invalid, uint64_t first_base, uint64_t n_bases, uint8_t* d_out,
                              cudaStream_t st)
{
  if (n_bases == 0) return cudaSuccess;
  const uint64_t threads = (n_bases + 15) / 16;
  unpack2bit_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_packed, d_invalid, first_base, n_bases, d_out);
  return cudaGetLastError();
}

// valid_bits words [0, ceil(*d_rows / 32)) <- all ones, the row count taken from device memory (no read-back)
__global__ void fill_valid_kernel(uint32_t* valid, const uint64_t* d_rows)
{
  const uint64_t words = (*d_rows + 31) / 32;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (uint64_t)gridDim.x * blockDim.x) valid[w] = 0xFFFFFFFFu;
}

cudaError_t launch_fill_valid(uint32_t* d_valid, const uint64_t* d_rows, uint64_t rows_bound, cudaStream_t st)
{
  const uint64_t words = (rows_bound + 31) / 32;
  if (words == 0) return cudaSuccess;
  fill_valid_kernel<<<(unsigned)std::min<uint64_t>((words + 1023) / 1024, 148 * 8), 256, 0, st>>>(d_valid, d_rows);
  return cudaGetLastError();
}

cudaError_t launch_koff_scan(const uint64_t* read_off, uint64_t n_reads, uint32_t k, uint32_t seg, uint64_t* excl,
                             uint64_t* stats, cudaStream_t st)
{
  if (n_reads == 0) {
    cudaError_t e = cudaMemsetAsync(excl, 0, sizeof(uint64_t), st);
    if (e != cudaSuccess) return e;
    return cudaMemsetAsync(stats, 0, 2 * sizeof(uint64_t), st);
  }
  const uint64_t nb = (n_reads + SCAN_B - 1) / SCAN_B;
  uint64_t* tmp = nullptr;
  cudaError_t e = cudaMallocAsync(&tmp, 2 * nb * sizeof(uint64_t), st);
  if (e != cudaSuccess) return e;
  scan_block_sums<<<(unsigned)nb, SCAN_T, 0, st>>>(read_off, n_reads, k, seg, tmp, tmp + nb);
  scan_of_sums<<<1, 1024, 0, st>>>(tmp, tmp + nb, nb, stats);
  scan_write<<<(unsigned)nb, SCAN_T, 0, st>>>(read_off, n_reads, k, seg, tmp, excl);
  e = cudaGetLastError();
  cudaFreeAsync(tmp, st);
  return e;
}

// Length-class order of every block of 256 consecutive items (the CTAs of kmer_fast_kernel on ragged batches): 32 classes
// relative to the block's longest item, longest first, counting sort.  perm[slot] = the block-local index of the item that
// thread `slot` takes.  Items past n_items (the padding of the last block) count as empty.
__global__ void __launch_bounds__(256) item_perm_kernel(const uint64_t* __restrict__ item_out, uint64_t n_items, uint8_t* __restrict__ perm)
{
  __shared__ uint32_t cnt[33];
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  const uint64_t i = (uint64_t)blockIdx.x * 256 + tid;
  const uint32_t n = i < n_items ? (uint32_t)(item_out[i + 1] - item_out[i]) : 0u;
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
  perm[(uint64_t)blockIdx.x * 256 + cnt[cls] + pos] = (uint8_t)tid;
}

__global__ void block_span_max_kernel(const uint64_t* __restrict__ read_off, uint64_t n_reads, unsigned long long* d_max)
{
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, r0 = b * 256;
  if (r0 >= n_reads) return;
  const uint64_t r1 = r0 + 256 < n_reads ? r0 + 256 : n_reads;
  atomicMax(d_max, (unsigned long long)(read_off[r1] - read_off[r0]));
}

cudaError_t launch_block_span_max(const uint64_t* read_off, uint64_t n_reads, uint64_t* d_max, cudaStream_t st)
{
  const uint64_t blocks = (n_reads + 255) / 256;
  if (blocks == 0) return cudaSuccess;
  block_span_max_kernel<<<(unsigned)((blocks + 255) / 256), 256, 0, st>>>(read_off, n_reads, reinterpret_cast<unsigned long long*>(d_max));
  return cudaGetLastError();
}

cudaError_t launch_item_perm(const uint64_t* item_out, uint64_t n_items, uint8_t* perm, cudaStream_t st)
{
  if (n_items == 0) return cudaSuccess;
  const uint64_t blocks = (n_items + 255) / 256;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  item_perm_kernel<<<(unsigned)blocks, 256, 0, st>>>(item_out, n_items, perm);
  return cudaGetLastError();
}

cudaError_t launch_item_fill(const uint64_t* read_off, const uint64_t* koff, uint64_t n_reads, uint32_t k,
                             uint32_t seg, uint64_t* item_byte, uint64_t* item_out, uint64_t* item_read,
                             uint64_t cap, cudaStream_t st)
{
  // item offsets per read = exclusive scan of ceil(nk / seg)
  uint64_t* ioff = nullptr;
  cudaError_t e = cudaMallocAsync(&ioff, (n_reads + 3) * sizeof(uint64_t), st);
  if (e != cudaSuccess) return e;
  e = launch_koff_scan(read_off, n_reads, k, seg, ioff, ioff + n_reads + 1, st);
  if (e == cudaSuccess) {
    const unsigned bs = 256;
    item_fill<<<(unsigned)((n_reads + bs - 1) / bs), bs, 0, st>>>(read_off, koff, ioff, n_reads, k, seg,
                                                                 item_byte, item_out, item_read, cap);
    e = cudaGetLastError();
  }
  cudaFreeAsync(ioff, st);
  return e;
}

} // namespace nthb
