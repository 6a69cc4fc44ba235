// Microbenchmark (round 1): HBM write bandwidth of TMA 2-D tile stores as a function of the tile
// geometry, to find which output pattern the k-mer kernel should produce.  Every warp owns 32 rows of
// an [n_rows][cols] u64 matrix and stores it tile by tile from shared memory (contents irrelevant).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_store_patterns tma_store_patterns.cu
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// rows_per_box x cols_per_box tile per store; a warp covers `rows_per_warp` rows and all `cols`
template<int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_tma(const __grid_constant__ CUtensorMap map, int cols, int box_cols,
                                                   int box_rows, int rows_per_warp, int n_rows, int spin)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_bytes = box_cols * box_rows * 8;
  uint8_t* tile = smem + warp * ((tile_bytes + 1023) & ~1023);
  for (int i = lane; i < tile_bytes / 8; i += 32) ((uint64_t*)tile)[i] = i + blockIdx.x;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int row0 = (blockIdx.x * WARPS + warp) * rows_per_warp;
  if (row0 >= n_rows) return;
  uint32_t x = lane;
  for (int r = 0; r < rows_per_warp; r += box_rows) {
    for (int c = 0; c < cols; c += box_cols) {
      for (int s = 0; s < spin; ++s) x = x * 1664525u + 1013904223u; // stand-in for the hashing work
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map), "r"(c),
                     "r"(row0 + r), "r"(smem_u32(tile))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      __syncwarp();
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  if (x == 12345u) tile[0] = 1;
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
  const size_t total_u64 = 1200000000ull; // 9.6 GB
  uint64_t* out;
  cudaMalloc(&out, total_u64 * 8 + (1 << 20));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  struct Cfg { const char* name; int cols, box_cols, box_rows, rows_per_warp, spin; CUtensorMapSwizzle sw; CUtensorMapL2promotion l2; };
  const Cfg cfgs[] = {
    { "pitch 960B  box 16x32 (kernel now)  ", 120, 16, 32, 32, 0, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE },
    { "pitch 960B  box 16x32 L2 promo 256B ", 120, 16, 32, 32, 0, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B },
    { "pitch 1024B box 16x32 (aligned rows)", 128, 16, 32, 32, 0, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE },
    { "pitch 960B  box  8x32 (64B rows)    ", 120, 8, 32, 32, 0, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE },
    { "pitch 960B  box 24x32 (192B rows)   ", 120, 24, 32, 32, 0, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE },
    { "pitch 960B  box 40x32 (320B rows)   ", 120, 40, 32, 32, 0, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE },
    { "pitch 960B  box 120x4 (whole rows)  ", 120, 120, 4, 32, 0, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE },
    { "pitch 960B  box 120x8 (whole rows)  ", 120, 120, 8, 32, 0, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE },
    { "pitch 128B  box 16x32 (contiguous)  ", 16, 16, 32, 32 * 8, 0, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE },
    { "pitch 960B  box 16x32 + spin 400    ", 120, 16, 32, 32, 400, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE },
    { "pitch 960B  box 16x32 + spin 1600   ", 120, 16, 32, 32, 1600, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE },
  };
  for (const Cfg& c : cfgs) {
    const uint64_t n_rows = total_u64 / c.cols;
    CUtensorMap map;
    const cuuint64_t dims[2] = { (cuuint64_t)c.cols, n_rows };
    const cuuint64_t strides[1] = { (cuuint64_t)c.cols * 8 };
    const cuuint32_t box[2] = { (cuuint32_t)c.box_cols, (cuuint32_t)c.box_rows };
    const cuuint32_t estr[2] = { 1, 1 };
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        c.sw, c.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s encode failed %d\n", c.name, (int)r); continue; }
    constexpr int WARPS = 8;
    const int tile_bytes = ((c.box_cols * c.box_rows * 8 + 1023) & ~1023);
    const int smem = WARPS * tile_bytes + 1024;
    cudaFuncSetAttribute(k_tma<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const unsigned blocks = (unsigned)((n_rows + WARPS * c.rows_per_warp - 1) / (WARPS * c.rows_per_warp));
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      k_tma<WARPS><<<blocks, WARPS * 32, smem>>>(map, c.cols, c.box_cols, c.box_rows, c.rows_per_warp, (int)n_rows, c.spin);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    printf("%s smem/CTA %6d  %.3f ms  %.1f GB/s  (%s)\n", c.name, smem, best, total_u64 * 8.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
