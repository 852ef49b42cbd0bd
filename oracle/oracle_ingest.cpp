// TEST INFRASTRUCTURE ONLY -- CPU oracle of the ingest step in front of the hot path.
//
// Restates what SupersurfelFusion::processFrame does to its inputs before segmentation
// (core/src/supersurfel_fusion.cu:171-181):
//   cv::cuda::cvtColor(rgb, gray, CV_RGB2GRAY)                       (:175)
//   cv::cuda::bilateralFilter(depth, depth, -1, 0.03, 4.5)           (:180)
// and the depth decode of the dataset node in front of it
//   depth.convertTo(depth, CV_32FC1, depth_scale)     (node/supersurfel_fusion_rgbd_benchmark_node.cpp:609-610)
//
// PARITY UNPINNED for the bilateral filter: it lives in OpenCV 3.4's cudaimgproc module
// (README.md:31-32), an un-vendored dependency whose source is not under /root/reference and
// which no reference test pins.  This file restates the published algorithm of that module
// (modules/cudaimgproc/src/bilateral_filter.cpp + src/cuda/bilateral_filter.cu):
//   radius = cvRound(1.5 sigma_spatial) when kernel_size <= 0, window (2 radius + 1)^2, taps
//   outside the disc of that radius skipped, weight = exp(-d2 / (2 sigma_s^2) - dv^2 / (2 sigma_c^2))
//   INCLUDING the centre tap, BORDER_REFLECT_101, out = sum(w v) / sum(w), fp32 throughout.
// tests/golden/make_bilateral_golden.py pins it against the CPU twin of the same function
// (cv2.bilateralFilter of OpenCV 4.13, same radius / disc / border rules, colour weights through
// an interpolated table) to 1e-4 m.  The reference calls the filter IN PLACE (src == dst), which
// races between thread blocks; like the CUDA path this restatement is out of place
// (SURVEY.md appendix B14).
#include <math.h>
#include <stdint.h>

#include "oracle.h"

namespace {
inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}
}  // namespace

extern "C" {

void orc_bilateral_filter(const float* depth, int width, int height, int kernel_size, float sigma_color,
                          float sigma_spatial, float* out) {
  if (!(sigma_color > 0.f)) sigma_color = 1.f;
  if (!(sigma_spatial > 0.f)) sigma_spatial = 1.f;
  int radius = kernel_size <= 0 ? (int)lrint((double)sigma_spatial * 1.5) : kernel_size / 2;   // cvRound
  if (radius < 1) radius = 1;
  const float s2 = -0.5f / (sigma_spatial * sigma_spatial);
  const float c2 = -0.5f / (sigma_color * sigma_color);
  const float r2 = (float)(radius * radius);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < height; y++) {
    for (int x = 0; x < width; x++) {
      const float center = depth[(size_t)y * width + x];
      float sum1 = 0.f, sum2 = 0.f;
      for (int cy = y - radius; cy <= y + radius; cy++) {
        const int yy = reflect101(cy, height);
        for (int cx = x - radius; cx <= x + radius; cx++) {
          const float space2 = (float)((x - cx) * (x - cx) + (y - cy) * (y - cy));
          if (space2 > r2) continue;
          const float value = depth[(size_t)yy * width + reflect101(cx, width)];
          const float dv = fabsf(value - center);
          const float weight = expf(space2 * s2 + (dv * dv) * c2);
          sum1 = sum1 + weight * value;
          sum2 = sum2 + weight;
        }
      }
      out[(size_t)y * width + x] = sum1 / sum2;
    }
  }
}

// cvtColor RGB2GRAY for 8-bit images: fixed point, 14 fractional bits, rounded
// (OpenCV color conversion: R2Y = 4899, G2Y = 9617, B2Y = 1868, CV_DESCALE(x, 14)).
void orc_rgb_to_gray(const uint8_t* rgb, int n_pixels, uint8_t* gray) {
  for (int i = 0; i < n_pixels; i++) {
    const int r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
    gray[i] = (uint8_t)((r * 4899 + g * 9617 + b * 1868 + (1 << 13)) >> 14);
  }
}

// Mat::convertTo(CV_32FC1, scale) of a 16-bit depth image: fp32 product per pixel.
void orc_depth16_to_metres(const uint16_t* depth16, int n_pixels, float scale, float* out) {
  for (int i = 0; i < n_pixels; i++) out[i] = (float)depth16[i] * scale;
}

}  // extern "C"
