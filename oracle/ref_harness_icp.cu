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

// DenseRegistration::align (dense_registration.cu:52-243) exactly as closeGlobalLoop calls it
// (supersurfel_fusion.cu:776-792): a keyframe's supersurfels, uploaded from the host, against
// the harness' current frame.
int ref_align(RefEngine* e, const RefSurfelsHost* src, int n, const float* Rinit9, const float* tinit3, float* R9,
              float* t3) {
  thrust::host_vector<float3> hp(n), hc(n);
  thrust::host_vector<Mat33> ho(n);
  thrust::host_vector<float> hf(n);
  for (int i = 0; i < n; i++) {
    hp[i] = make_float3(src->positions[3 * i], src->positions[3 * i + 1], src->positions[3 * i + 2]);
    hc[i] = make_float3(src->colors[3 * i], src->colors[3 * i + 1], src->colors[3 * i + 2]);
    ho[i] = mat_from(src->orientations + 9 * i);
    hf[i] = src->confidences[i];
  }
  thrust::device_vector<float3> dp = hp, dc = hc;
  thrust::device_vector<Mat33> dorient = ho;
  thrust::device_vector<float> dconf = hf;
  Mat33 R;
  float3 t;
  const bool ok = e->icp->align(dp, dc, dorient, dconf, e->frame.positions, e->frame.colors, e->frame.orientations,
                                e->frame.confidences, n, e->texDepth, e->tps->getTexIndex(), mat_from(Rinit9),
                                make_float3(tinit3[0], tinit3[1], tinit3[2]), e->cam, R, t);
  for (int r = 0; r < 3; r++) { R9[3 * r] = R.rows[r].x; R9[3 * r + 1] = R.rows[r].y; R9[3 * r + 2] = R.rows[r].z; }
  t3[0] = t.x; t3[1] = t.y; t3[2] = t.z;
  return ok ? 1 : 0;
}

}  // extern "C"
