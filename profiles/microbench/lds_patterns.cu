// Microbenchmark (round 1): L1TEX data-pipe wavefronts per shared-memory load instruction for the
// access shapes of the k-mer kernel: a 16-byte table lookup where the warp touches only 4 distinct
// entries, 8-byte and 4-byte variants, and byte loads at a 150-byte lane stride.
// Run under: ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,gpu__time_duration.sum
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template<int MODE>
__global__ void __launch_bounds__(256) k_lds(uint32_t* sink, int iters)
{
  __shared__ __align__(16) uint8_t sm[48 * 1024];
  for (int i = threadIdx.x; i < 12 * 1024; i += 256) ((uint32_t*)sm)[i] = i * 2654435761u;
  __syncthreads();
  uint32_t acc = 0, idx = (threadIdx.x * 7 + blockIdx.x) & 3;
  const uint32_t lane_base = (threadIdx.x * 150) % (40 * 1024);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { uint4 v = *(const uint4*)(sm + (0x41 + idx * 2) * 16); acc ^= v.x ^ v.y ^ v.z ^ v.w; idx = (idx + (v.x & 1) + 1) & 3; }
    if (MODE == 1) { uint2 v = *(const uint2*)(sm + (0x41 + idx * 2) * 16); acc ^= v.x ^ v.y; idx = (idx + (v.x & 1) + 1) & 3; }
    if (MODE == 2) { uint32_t v = *(const uint32_t*)(sm + (0x41 + idx * 2) * 16); acc ^= v; idx = (idx + (v & 1) + 1) & 3; }
    if (MODE == 3) { uint32_t v = sm[lane_base + it % 150]; acc += v; }
    if (MODE == 4) { uint32_t v = *(const uint32_t*)(sm + ((lane_base + (it % 37) * 4) & ~3u)); acc += v; }
    if (MODE == 5) { uint4 v = *(const uint4*)(sm + idx * 16); acc ^= v.x ^ v.y ^ v.z ^ v.w; idx = (idx + (v.x & 1) + 1) & 3; } // 4 adjacent entries (64 B)
  }
  sink[blockIdx.x * 256 + threadIdx.x] = acc;
}

int main()
{
  uint32_t* sink;
  cudaMalloc(&sink, 148 * 8 * 256 * 4);
  const char* names[] = { "LDS.128 4 distinct entries", "LDS.64 4 distinct entries", "LDS.32 4 distinct entries", "LDS.U8 lane stride 150", "LDS.32 lane stride 150", "LDS.128 4 adjacent entries" };
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int m = 0; m < 6; ++m) {
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      switch (m) {
        case 0: k_lds<0><<<148 * 8, 256>>>(sink, 4096); break;
        case 1: k_lds<1><<<148 * 8, 256>>>(sink, 4096); break;
        case 2: k_lds<2><<<148 * 8, 256>>>(sink, 4096); break;
        case 3: k_lds<3><<<148 * 8, 256>>>(sink, 4096); break;
        case 4: k_lds<4><<<148 * 8, 256>>>(sink, 4096); break;
        case 5: k_lds<5><<<148 * 8, 256>>>(sink, 4096); break;
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("%-30s %.3f ms (%d warp-loads per SM)\n", names[m], ms, 8 * 8 * 4096);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
