// Microbenchmark (round 1): how many L1TEX data-pipe wavefronts does a lane-strided store cost?
// Each lane owns a 960-byte row (120 u64, the C2 shape) and writes it front to back, as the k-mer
// kernel's lanes do.  Variants: 32-byte (STG.256), 16-byte (STG.128), 8-byte (STG.64) stores, with
// and without cache hints.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template<int MODE>
__global__ void __launch_bounds__(256) k_store(uint64_t* out, uint64_t n_rows, uint64_t salt)
{
  const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= n_rows) return;
  uint64_t* p = out + r * 120;
  uint64_t v = r * 0x9e3779b97f4a7c15ull + salt;
#pragma unroll 2
  for (int i = 0; i < 120; i += 4) {
    uint64_t a = v, b = v + 1, c = v + 2, d = v + 3;
    v = v * 6364136223846793005ull + 1442695040888963407ull;
    if (MODE == 0) asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
    if (MODE == 1) {
      asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(p + i), "l"(a), "l"(b) : "memory");
      asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(p + i + 2), "l"(c), "l"(d) : "memory");
    }
    if (MODE == 2) { p[i] = a; p[i + 1] = b; p[i + 2] = c; p[i + 3] = d; }
    if (MODE == 3) asm volatile("st.global.cs.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
    if (MODE == 4) asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
  }
}


// Round 2: is the lane-strided STG.256 cap a per-request cost?  G lanes (2 or 4) share G rows: every instruction writes
// G * 32 contiguous bytes of ONE row per lane group (what a shuffle transpose between the lanes of a group would give the
// DIRECT output path), G instructions cover the group's G rows.  Same bytes, same order per row, 1/G as many requests per line.
template<int G>
__global__ void __launch_bounds__(256) k_store_group(uint64_t* out, uint64_t n_rows, uint64_t salt)
{
  const uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= n_rows) return;
  const uint32_t sub = (uint32_t)(t % G);
  const uint64_t r0 = t - sub; // the group's first row
  uint64_t v = t * 0x9e3779b97f4a7c15ull + salt;
#pragma unroll 1
  for (int i = 0; i < 120; i += 4 * G) {   // G sectors of every row per round
#pragma unroll
    for (int g = 0; g < G; ++g) {
      uint64_t* p = out + (r0 + g) * 120 + i + 4 * sub;
      if (i + 4 * (int)sub < 120)
        asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v), "l"(v + 1), "l"(v + 2), "l"(v + 3) : "memory");
      v = v * 6364136223846793005ull + 1442695040888963407ull;
    }
  }
}

// coalesced reference: the warp writes 1024 contiguous bytes per instruction
__global__ void __launch_bounds__(256) k_store_coalesced(uint64_t* out, uint64_t n_rows, uint64_t salt)
{
  const uint64_t w = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5; // warp id
  const uint32_t lane = threadIdx.x & 31;
  if (w * 32 >= n_rows) return;
  uint64_t* p = out + w * 32 * 120 + lane * 4;
  uint64_t v = w * 0x9e3779b97f4a7c15ull + salt + lane;
#pragma unroll 2
  for (int i = 0; i < 30; ++i) {
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p + i * 128), "l"(v), "l"(v + 1), "l"(v + 2), "l"(v + 3) : "memory");
    v = v * 6364136223846793005ull + 1442695040888963407ull;
  }
}

int main()
{
  const uint64_t n_rows = 10000000;
  uint64_t* out;
  cudaMalloc(&out, n_rows * 120 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const unsigned blocks = (unsigned)((n_rows + 255) / 256);
  const char* names[] = { "STG.256 lane-strided", "2x STG.128 lane-strided", "4x STG.64 lane-strided", "STG.256 .cs", "STG.256 L1::no_allocate", "STG.256 coalesced", "STG.256, lane pairs: 64 B per row", "STG.256, lane quads: 128 B per row", "STG.256, 8 lanes: 256 B per row" };
  for (int m = 0; m < 9; ++m) {
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      switch (m) {
        case 0: k_store<0><<<blocks, 256>>>(out, n_rows, rep); break;
        case 1: k_store<1><<<blocks, 256>>>(out, n_rows, rep); break;
        case 2: k_store<2><<<blocks, 256>>>(out, n_rows, rep); break;
        case 3: k_store<3><<<blocks, 256>>>(out, n_rows, rep); break;
        case 4: k_store<4><<<blocks, 256>>>(out, n_rows, rep); break;
        case 5: k_store_coalesced<<<blocks, 256>>>(out, n_rows, rep); break;
        case 6: k_store_group<2><<<blocks, 256>>>(out, n_rows, rep); break;
        case 7: k_store_group<4><<<blocks, 256>>>(out, n_rows, rep); break;
        case 8: k_store_group<8><<<blocks, 256>>>(out, n_rows, rep); break;
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    printf("%-40s %.3f ms  %.1f GB/s\n", names[m], best, n_rows * 960.0 / best / 1e6);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
