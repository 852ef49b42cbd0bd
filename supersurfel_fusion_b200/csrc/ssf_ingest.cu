// Ingest in front of the segmentation (sm_100a): the depth bilateral filter, the 16-bit
// depth decode and the grey image.
//
// Replaces, in SupersurfelFusion::processFrame (reference: core/src/supersurfel_fusion.cu:171-181),
//   cv::cuda::bilateralFilter(depth, depth, -1, 0.03, 4.5)   (:180, OpenCV cudaimgproc)
//   cv::cuda::cvtColor(rgb, gray, CV_RGB2GRAY)               (:175, OpenCV cudev)
// and, in the dataset node, depth.convertTo(CV_32FC1, depth_scale)
//   (node/supersurfel_fusion_rgbd_benchmark_node.cpp:609-610),
// so that no third-party GPU library sits in the frame loop and the whole frame stays one
// CUDA graph.  The arithmetic is the published algorithm of those OpenCV functions, restated
// in oracle/oracle_ingest.cpp (see there for what pins it).  Unlike the reference call the
// filter is out of place: the reference filters src == dst, which races between thread
// blocks.
#include "ssf_engine.h"
#include "ssf_math.cuh"

namespace ssf {

constexpr int BIL_TX = 32, BIL_TY = 16;
constexpr int BIL_MAX_RADIUS = 12;

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

// One thread per pixel; the CTA stages its 32x16 tile plus a `radius` apron (reflect-101 at
// the image border) in shared memory, so every depth value is read from L2 once per tile
// and the (2 radius + 1)^2 taps run out of shared memory.  Same tap order and the same
// separately rounded fp32 operations as the oracle.
__global__ void __launch_bounds__(BIL_TX * BIL_TY) bilateral_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                    int W, int H, int radius, float s2, float c2) {
  pdl_sync();
  extern __shared__ float tile[];
  const int tw = BIL_TX + 2 * radius, th = BIL_TY + 2 * radius;
  const int x0 = blockIdx.x * BIL_TX - radius, y0 = blockIdx.y * BIL_TY - radius;
  for (int i = threadIdx.y * BIL_TX + threadIdx.x; i < tw * th; i += BIL_TX * BIL_TY) {
    const int tx = i % tw, ty = i / tw;
    tile[i] = src[(size_t)reflect101(y0 + ty, H) * W + reflect101(x0 + tx, W)];
  }
  __syncthreads();
  const int x = blockIdx.x * BIL_TX + threadIdx.x, y = blockIdx.y * BIL_TY + threadIdx.y;
  if (x >= W || y >= H) return;
  const int lx = threadIdx.x + radius, ly = threadIdx.y + radius;
  const float center = tile[ly * tw + lx];
  const float r2 = (float)(radius * radius);
  float sum1 = 0.f, sum2 = 0.f;
  for (int dy = -radius; dy <= radius; dy++) {
    const float* row = tile + (ly + dy) * tw + lx;
    for (int dx = -radius; dx <= radius; dx++) {
      const float space2 = (float)(dx * dx + dy * dy);
      if (space2 > r2) continue;
      const float value = row[dx];
      const float dv = fabsf(value - center);
      const float weight = expf(space2 * s2 + (dv * dv) * c2);
      sum1 = sum1 + weight * value;
      sum2 = sum2 + weight;
    }
  }
  dst[(size_t)y * W + x] = sum1 / sum2;
}

__global__ void gray_kernel(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ gray, size_t n) {
  pdl_sync();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
  gray[i] = (uint8_t)((r * 4899 + g * 9617 + b * 1868 + (1 << 13)) >> 14);
}

__global__ void depth16_kernel(const uint16_t* __restrict__ d16, float* __restrict__ out, size_t n, float scale) {
  pdl_sync();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)d16[i] * scale;
}

int bilateral_radius(int kernel_size, float sigma_spatial) {
  if (!(sigma_spatial > 0.f)) sigma_spatial = 1.f;
  int radius = kernel_size <= 0 ? (int)lrint((double)sigma_spatial * 1.5) : kernel_size / 2;
  return radius < 1 ? 1 : radius;
}

// cv::cuda::bilateralFilter(src, dst, kernel_size, sigma_color, sigma_spatial), out of place
int launch_bilateral(Engine* e, const float* src_dev, float* dst_dev, int kernel_size, float sigma_color,
                     float sigma_spatial) {
  if (!(sigma_color > 0.f)) sigma_color = 1.f;
  if (!(sigma_spatial > 0.f)) sigma_spatial = 1.f;
  const int radius = bilateral_radius(kernel_size, sigma_spatial);
  if (radius > BIL_MAX_RADIUS) return SSF_ERR_INVALID_ARG;
  const float s2 = -0.5f / (sigma_spatial * sigma_spatial);
  const float c2 = -0.5f / (sigma_color * sigma_color);
  const size_t smem = (size_t)(BIL_TX + 2 * radius) * (BIL_TY + 2 * radius) * sizeof(float);
  launch_pdl(e, bilateral_kernel, dim3((e->W + BIL_TX - 1) / BIL_TX, (e->H + BIL_TY - 1) / BIL_TY), dim3(BIL_TX, BIL_TY), smem,
             src_dev, dst_dev, e->W, e->H, radius, s2, c2);
  e->launches++;
  return SSF_OK;
}

void launch_gray(Engine* e, const uint8_t* rgb_dev, uint8_t* gray_dev) {
  launch_pdl(e, gray_kernel, dim3((unsigned)((e->npix + 255) / 256)), dim3(256), 0, rgb_dev, gray_dev, e->npix);
  e->launches++;
}

void launch_depth16(Engine* e, const uint16_t* d16_dev, float* out_dev, float scale) {
  launch_pdl(e, depth16_kernel, dim3((unsigned)((e->npix + 255) / 256)), dim3(256), 0, d16_dev, out_dev, e->npix, scale);
  e->launches++;
}

}  // namespace ssf
