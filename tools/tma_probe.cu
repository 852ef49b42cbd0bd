// Probe (run under gpurun): 2-D tensor-map TMA load of an int32 tile, the way tps_pass_tile_kernel stages its labels.
// Finding on B200 (round 2): boxes may start at negative or beyond-the-image coordinates (zero fill), but the
// innermost start coordinate must keep the global address 16-byte aligned -- x = -2 or 2 (int32) raises
// cudaErrorIllegalInstruction at the UTMALDG (the last case of the second pass below shows it).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/tma_probe.cu && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int COLS = 68, ROWS = 17;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int x0, int y0, int* out) {
  __shared__ __align__(128) int tile[ROWS][COLS];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(ROWS * COLS * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(&tile[0][0])), "l"(&map), "r"(x0), "r"(y0), "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < ROWS * COLS; i += blockDim.x) out[i] = tile[i / COLS][i % COLS];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int W = 320, H = 240;
  std::vector<int> img(W * H);
  for (int i = 0; i < W * H; i++) img[i] = 1000 + i;
  int *d_img, *d_out;
  cudaMalloc(&d_img, W * H * 4); cudaMalloc(&d_out, ROWS * COLS * 4);
  cudaMemcpy(d_img, img.data(), W * H * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
  for (int dtype_variant = 0; dtype_variant < 2; dtype_variant++) {
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    const cuuint32_t box[2] = {COLS, ROWS};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = ((EncodeFn)fn)(&map, dtype_variant ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d_img, dims,
                                 strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode(dtype %d) rc=%d\n", dtype_variant, (int)rc);
    // the innermost coordinate must be a multiple of 4 elements (16 bytes): (-2, -1) and (2, 0) fault, see below
    const int tests[6][2] = {{64, 16}, {0, 0}, {-4, -1}, {300, 230}, {320, 240}, {dtype_variant ? 2 : 0, 0}};
    for (auto& t : tests) {
      cudaMemset(d_out, 0xff, ROWS * COLS * 4);
      probe<<<1, 128>>>(map, t[0], t[1], d_out);
      cudaError_t s = cudaDeviceSynchronize();
      std::vector<int> out(ROWS * COLS);
      cudaMemcpy(out.data(), d_out, ROWS * COLS * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < ROWS; r++)
        for (int c = 0; c < COLS; c++) {
          const int x = t[0] + c, y = t[1] + r;
          const int want = (x >= 0 && x < W && y >= 0 && y < H) ? img[y * W + x] : 0;
          bad += out[r * COLS + c] != want;
        }
      printf("  load at (%d,%d): %s, mismatches %d\n", t[0], t[1], cudaGetErrorString(s), bad);
      if (s != cudaSuccess) { cudaGetLastError(); return 1; }
    }
  }
  return 0;
}
