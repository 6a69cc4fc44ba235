// fastq_stage.cu — the step before the hash path: FASTQ text that already sits in device
// memory -> the concatenated `bases` + `read_off` layout every batch entry point takes.  No host-side parsing: the
// newline positions come from a device-wide select, the sequence lines (every 4th line, starting at the 2nd) are
// copied out by one warp per read.  Library scans (CUB DeviceSelect / DeviceScan) do the prefix work: this is staging,
// not the hot path.  Records are the plain four-line form (`@id`, sequence, `+`, qualities), LF or CRLF line ends.
#include "engine.hpp"

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

namespace nthb {

namespace {

struct IsNewline
{
  const uint8_t* text;
  __device__ bool operator()(uint64_t i) const { return text[i] == '\n'; }
};

__global__ void count_newlines(const uint8_t* text, uint64_t n_bytes, unsigned long long* count)
{
  uint32_t c = 0;
  const uint64_t head = min(n_bytes, (uint64_t)((16 - ((uintptr_t)text & 15)) & 15)), n16 = (n_bytes - head) / 16;
  const uint4* t16 = reinterpret_cast<const uint4*>(text + head);
  auto nls = [](uint32_t w) { // bytes of w equal to '\n'
    const uint32_t x = w ^ 0x0A0A0A0Au;
    return __popc(~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u);
  };
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = t16[i];
    c += nls(v.x) + nls(v.y) + nls(v.z) + nls(v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (uint64_t i = 0; i < head; ++i) c += text[i] == '\n';
    for (uint64_t i = head + n16 * 16; i < n_bytes; ++i) c += text[i] == '\n';
  }
  for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long)c);
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
  void* d_tmp = nullptr;
  auto done = [&](cudaError_t e) {
    for (void* p : { (void*)d_nl, (void*)d_cnt, (void*)d_start, (void*)d_len, d_tmp })
      if (p) cudaFreeAsync(p, st);
    return e;
  };
  // 1. count the newlines, then list their positions
  cudaError_t e = cudaMallocAsync(&d_cnt, sizeof(uint64_t), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, sizeof(uint64_t), st);
  if (e != cudaSuccess) return done(e);
  count_newlines<<<1184, 256, 0, st>>>(d_text, n_bytes, reinterpret_cast<unsigned long long*>(d_cnt));
  uint64_t max_nl = 0;
  e = cudaMemcpyAsync(&max_nl, d_cnt, sizeof max_nl, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_nl, (max_nl + 1) * sizeof(uint64_t), st);
  if (e != cudaSuccess) return done(e);
  thrust::counting_iterator<uint64_t> idx(0);
  size_t tmp_bytes = 0;
  e = cub::DeviceSelect::If(nullptr, tmp_bytes, idx, d_nl, d_cnt, (int64_t)n_bytes, IsNewline{ d_text }, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_tmp, tmp_bytes, st);
  if (e == cudaSuccess) e = cub::DeviceSelect::If(d_tmp, tmp_bytes, idx, d_nl, d_cnt, (int64_t)n_bytes, IsNewline{ d_text }, st);
  uint64_t n_nl = 0;
  uint8_t last = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n_nl, d_cnt, sizeof n_nl, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&last, d_text + n_bytes - 1, 1, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
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
  e = cudaMemsetAsync(d_len + n_reads, 0, sizeof(uint64_t), st);
  cudaFreeAsync(d_tmp, st);
  d_tmp = nullptr;
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len, d_read_off, (int64_t)(n_reads + 1), st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_tmp, tmp_bytes, st);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_len, d_read_off, (int64_t)(n_reads + 1), st);
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
