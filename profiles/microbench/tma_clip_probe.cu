// Probe (round 1): which clipped TMA tensor stores does sm_100a accept?  box larger than a tensor dimension,
// negative start coordinates, rank 4.  Prints the CUDA error of each variant and the rows that were written.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_clip_probe tma_clip_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_store(const __grid_constant__ CUtensorMap map, int rank, int c1, int c2, int c3)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* t = (uint64_t*)smem;
  for (int i = threadIdx.x; i < 3 * 32 * 8; i += 32) t[i] = 1000 + i; // [3 blocks][32 rows][8]
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (threadIdx.x == 0) {
    if (rank == 3)
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map), "r"(0), "r"(c1),
                   "r"(c2), "r"(smem_u32(smem))
                   : "memory");
    else
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(&map), "r"(0),
                   "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(smem))
                   : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
  EncodeFn encode = (EncodeFn)fp;
  const int N = 1 << 16;
  uint64_t* out;
  cudaMalloc(&out, N * 8);
  std::vector<uint64_t> h(N);
  struct V { const char* name; int rank; cuuint64_t dims[4]; cuuint64_t strides[3]; cuuint32_t box[4]; int c1, c2, c3; };
  const V vs[] = {
    { "3D rows 64 >= box 32            ", 3, { 8, 64, 6, 1 }, { 384, 64, 0 }, { 8, 32, 3, 1 }, 0, 0, 0 },
    { "3D rows 4 < box 32              ", 3, { 8, 4, 6, 1 }, { 384, 64, 0 }, { 8, 32, 3, 1 }, 0, 0, 0 },
    { "3D rows 64, start row -5        ", 3, { 8, 64, 6, 1 }, { 384, 64, 0 }, { 8, 32, 3, 1 }, -5, 0, 0 },
    { "4D rows 64 >= box 32            ", 4, { 8, 64, 6, 4 }, { 384, 64, 32768 }, { 8, 32, 3, 1 }, 0, 0, 1 },
    { "4D rows 4 < box 32, start -2    ", 4, { 8, 4, 6, 4 }, { 384, 64, 32768 }, { 8, 32, 3, 1 }, -2, 3, 1 },
    { "4D rows 1, start -7             ", 4, { 8, 1, 6, 4 }, { 384, 64, 32768 }, { 8, 32, 3, 1 }, -7, 0, 2 },
    { "4D rows 4, strides not monotone ", 4, { 8, 4, 6, 4 }, { 1600, 64, 7760 }, { 8, 32, 3, 1 }, 0, 0, 0 },
  };
  for (const V& v : vs) {
    cudaMemset(out, 0, N * 8);
    CUtensorMap map;
    const cuuint32_t estr[4] = { 1, 1, 1, 1 };
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, v.rank, out, v.dims, v.strides, v.box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s encode failed %d\n", v.name, (int)r); continue; }
    cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    k_store<<<1, 32, 16384>>>(map, v.rank, v.c1, v.c2, v.c3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s -> %s", v.name, cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    cudaMemcpy(h.data(), out, N * 8, cudaMemcpyDeviceToHost);
    size_t nz = 0, first = 0, last = 0;
    for (size_t i = 0; i < (size_t)N; ++i)
      if (h[i]) { if (!nz) first = i; last = i; ++nz; }
    printf("  wrote %zu u64, first at %zu (val %llu), last at %zu\n", nz, first, (unsigned long long)h[first], last);
  }
  return 0;
}
