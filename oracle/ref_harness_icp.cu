// TEST INFRASTRUCTURE ONLY -- reference-kernel harness, stand-alone launches of the
// reference's computeSymmetricICPSystem<128> kernel template
// (core/include/supersurfel_fusion/dense_registration_kernels.cuh:175-291).  Separate
// translation unit: that header overloads atomicAdd inside namespace supersurfel_fusion,
// which would hide ::atomicAdd(float*, float) from supersurfel_fusion_kernels.cu.
#include "ref_harness.h"

#include <supersurfel_fusion/dense_registration_kernels.cuh>

#include <cstring>

extern "C" {

// one launch of computeSymmetricICPSystem<128> (dense_registration.cu:301-324)
void ref_icp_system(RefEngine* e, const float* R9, const float* t3, int n, float* out29) {
  sf::MotionTrackingData* mtd;
  cudaMallocManaged(&mtd, sizeof(sf::MotionTrackingData));
  cudaMemset(mtd, 0, sizeof(sf::MotionTrackingData));
  sf::computeSymmetricICPSystem<128><<<(n + 127) / 128, 128>>>(
      mtd, RAW(e->model.positions), RAW(e->model.colors), RAW(e->model.orientations), RAW(e->frame.colors),
      RAW(e->frame.orientations), RAW(e->frame.confidences), mat_from(R9), make_float3(t3[0], t3[1], t3[2]),
      e->cam.fx, e->cam.fy, e->cam.cx, e->cam.cy, e->tps->getTexIndex()->getTextureObject(),
      e->texDepth->getTextureObject(), e->cam.width, e->cam.height, n);
  cudaDeviceSynchronize();
  memcpy(out29, mtd, 29 * sizeof(float));
  cudaFree(mtd);
}

// time `launches` back-to-back system builds (memset + kernel, as the reference issues
// them) with CUDA events; returns milliseconds per launch
float ref_icp_system_time(RefEngine* e, const float* R9, const float* t3, int n, int launches) {
  sf::MotionTrackingData* mtd;
  cudaMalloc(&mtd, sizeof(sf::MotionTrackingData));
  const Mat33 R = mat_from(R9);
  const float3 t = make_float3(t3[0], t3[1], t3[2]);
  cudaEventRecord(e->ev[0]);
  for (int i = 0; i < launches; i++) {
    cudaMemsetAsync(mtd, 0, sizeof(sf::MotionTrackingData));
    sf::computeSymmetricICPSystem<128><<<(n + 127) / 128, 128>>>(
        mtd, RAW(e->model.positions), RAW(e->model.colors), RAW(e->model.orientations), RAW(e->frame.colors),
        RAW(e->frame.orientations), RAW(e->frame.confidences), R, t, e->cam.fx, e->cam.fy, e->cam.cx, e->cam.cy,
        e->tps->getTexIndex()->getTextureObject(), e->texDepth->getTextureObject(), e->cam.width, e->cam.height, n);
  }
  cudaEventRecord(e->ev[1]);
  cudaEventSynchronize(e->ev[1]);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
  cudaFree(mtd);
  return ms / launches;
}

}  // extern "C"
