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

// 3-D variant: the output row is seen as [cols/inner][inner]; one store moves `nblk` adjacent inner-blocks of
// 32 rows (box = inner x 32 x nblk), i.e. nblk*inner*8 contiguous bytes per row, laid out block-major in
// shared memory so that a lane writing its row's block j stays conflict-free under the 64B/128B swizzle.
template<int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_tma3(const __grid_constant__ CUtensorMap map, int cols, int inner, int nblk,
                                                    int n_rows, int spin, int nbuf)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_bytes = inner * 32 * nblk * 8;
  const int tile_pitch = (tile_bytes + 1023) & ~1023;
  uint8_t* tile = smem + warp * tile_pitch * nbuf;
  for (int i = lane; i < tile_pitch * nbuf / 8; i += 32) ((uint64_t*)tile)[i] = i + blockIdx.x;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int row0 = (blockIdx.x * WARPS + warp) * 32;
  if (row0 >= n_rows) return;
  uint32_t x = lane;
  int b = 0;
  for (int c = 0; c < cols / inner; c += nblk) {
    for (int s = 0; s < spin; ++s) x = x * 1664525u + 1013904223u;
    if (lane == 0) {
      if (nbuf == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map), "r"(0),
                   "r"(row0), "r"(c), "r"(smem_u32(tile + b * tile_pitch))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    b = (b + 1) % nbuf;
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  if (x == 12345u) tile[0] = 1;
}

// per-lane 1-D bulk stores: every lane pushes `seg` contiguous bytes of its own row per step
template<int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_bulk1d(uint64_t* out, int cols, int seg_cols, int n_rows, int spin)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_bytes = seg_cols * 8 + 16; // padded row in shared memory
  uint8_t* tile = smem + warp * ((row_bytes * 32 + 127) & ~127);
  for (int i = lane; i < row_bytes * 32 / 8; i += 32) ((uint64_t*)tile)[i] = i + blockIdx.x;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const long long row = (long long)(blockIdx.x * WARPS + warp) * 32 + lane;
  if (row >= n_rows) return;
  uint32_t x = lane;
  for (int c = 0; c < cols; c += seg_cols) {
    for (int s = 0; s < spin; ++s) x = x * 1664525u + 1013904223u;
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + row * cols + c),
                 "r"(smem_u32(tile + lane * row_bytes)), "r"(seg_cols * 8)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
  // ---- 3-D boxes: inner x 32 rows x nblk column blocks ----
  struct Cfg3 { const char* name; int cols, inner, nblk, spin, nbuf, warps; CUtensorMapSwizzle sw; };
  const Cfg3 cfg3[] = {
    { "3D pitch 960B inner 8 (64B) x2 = 128B/row     ", 120, 8, 1, 0, 1, 8, CU_TENSOR_MAP_SWIZZLE_64B },
    { "3D pitch 960B inner 8 x3 = 192B/row           ", 120, 8, 3, 0, 1, 8, CU_TENSOR_MAP_SWIZZLE_64B },
    { "3D pitch 960B inner 8 x5 = 320B/row           ", 120, 8, 5, 0, 1, 8, CU_TENSOR_MAP_SWIZZLE_64B },
    { "3D pitch 960B inner 8 x5 = 320B/row, 5 warps  ", 120, 8, 5, 0, 1, 5, CU_TENSOR_MAP_SWIZZLE_64B },
    { "3D pitch 960B inner 8 x5 = 320B/row, 2 bufs   ", 120, 8, 5, 0, 2, 4, CU_TENSOR_MAP_SWIZZLE_64B },
    { "3D pitch 960B inner 8 x15 = whole row         ", 120, 8, 15, 0, 1, 4, CU_TENSOR_MAP_SWIZZLE_64B },
    { "3D pitch 960B inner 4 (32B) x10 = 320B/row    ", 120, 4, 10, 0, 1, 8, CU_TENSOR_MAP_SWIZZLE_32B },
    { "3D pitch 960B inner 2 (16B) x20 = 320B/row    ", 120, 2, 20, 0, 1, 8, CU_TENSOR_MAP_SWIZZLE_NONE },
    { "3D pitch 1920B inner 16 (128B) x2 = 256B/row  ", 240, 16, 2, 0, 1, 8, CU_TENSOR_MAP_SWIZZLE_128B },
    { "3D pitch 1920B inner 16 (128B) x3 = 384B/row  ", 240, 16, 3, 0, 1, 6, CU_TENSOR_MAP_SWIZZLE_128B },
    { "3D pitch 1920B inner 16 (128B) x5 = 640B/row  ", 240, 16, 5, 0, 1, 4, CU_TENSOR_MAP_SWIZZLE_128B },
    { "3D pitch 3840B inner 16 (128B) x1 = 128B/row  ", 480, 16, 1, 0, 1, 8, CU_TENSOR_MAP_SWIZZLE_128B },
    { "3D pitch 3840B inner 16 (128B) x2 = 256B/row  ", 480, 16, 2, 0, 1, 8, CU_TENSOR_MAP_SWIZZLE_128B },
    { "3D pitch 3840B inner 16 (128B) x3 = 384B/row  ", 480, 16, 3, 0, 1, 6, CU_TENSOR_MAP_SWIZZLE_128B },
    { "3D pitch 960B inner 8 x5 + spin 1000          ", 120, 8, 5, 1000, 1, 8, CU_TENSOR_MAP_SWIZZLE_64B },
    { "3D pitch 960B inner 8 x5 + spin 1000, 2 bufs  ", 120, 8, 5, 1000, 2, 4, CU_TENSOR_MAP_SWIZZLE_64B },
  };
  for (const Cfg3& c : cfg3) {
    const uint64_t n_rows = total_u64 / c.cols;
    CUtensorMap map;
    const cuuint64_t dims[3] = { (cuuint64_t)c.inner, n_rows, (cuuint64_t)(c.cols / c.inner) };
    const cuuint64_t strides[2] = { (cuuint64_t)c.cols * 8, (cuuint64_t)c.inner * 8 };
    const cuuint32_t box[3] = { (cuuint32_t)c.inner, 32, (cuuint32_t)c.nblk };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        c.sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s encode failed %d\n", c.name, (int)r); continue; }
    const int tile_bytes = ((c.inner * 32 * c.nblk * 8 + 1023) & ~1023) * c.nbuf;
    const int smem = c.warps * tile_bytes + 1024;
    const unsigned blocks = (unsigned)((n_rows + c.warps * 32 - 1) / (c.warps * 32));
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      switch (c.warps) {
        case 8: cudaFuncSetAttribute(k_tma3<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                k_tma3<8><<<blocks, 256, smem>>>(map, c.cols, c.inner, c.nblk, (int)n_rows, c.spin, c.nbuf); break;
        case 6: cudaFuncSetAttribute(k_tma3<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                k_tma3<6><<<blocks, 192, smem>>>(map, c.cols, c.inner, c.nblk, (int)n_rows, c.spin, c.nbuf); break;
        case 5: cudaFuncSetAttribute(k_tma3<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                k_tma3<5><<<blocks, 160, smem>>>(map, c.cols, c.inner, c.nblk, (int)n_rows, c.spin, c.nbuf); break;
        default: cudaFuncSetAttribute(k_tma3<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                k_tma3<4><<<blocks, 128, smem>>>(map, c.cols, c.inner, c.nblk, (int)n_rows, c.spin, c.nbuf); break;
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    printf("%s smem/CTA %6d  %.3f ms  %.1f GB/s  (%s)\n", c.name, smem, best, total_u64 * 8.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  // ---- per-lane 1-D bulk stores ----
  const int segs1d[] = { 16, 30, 40, 60, 120 };
  for (int sc : segs1d) {
    const int cols = 120;
    const uint64_t n_rows = total_u64 / cols;
    constexpr int WARPS = 4;
    const int smem = WARPS * (((sc * 8 + 16) * 32 + 127) & ~127) + 1024;
    cudaFuncSetAttribute(k_bulk1d<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const unsigned blocks = (unsigned)((n_rows + WARPS * 32 - 1) / (WARPS * 32));
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      k_bulk1d<WARPS><<<blocks, WARPS * 32, smem>>>(out, cols, sc, (int)n_rows, 0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    printf("1D bulk per lane, pitch 960B, %4d B per store      smem/CTA %6d  %.3f ms  %.1f GB/s  (%s)\n", sc * 8, smem, best,
           total_u64 * 8.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
