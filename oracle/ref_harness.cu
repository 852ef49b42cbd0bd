// TEST INFRASTRUCTURE ONLY -- reference-kernel harness (oracle/_ref/libssf_ref.so).
//
// Drives the reference's OWN, UNMODIFIED hot-path code on a GPU so that (a) the CPU
// oracle and the CUDA product can be pinned against what the reference actually
// computes, and (b) "the reference's own build on one GPU of the same box" can be timed
// next to ours.  What is compiled, from where it lies under /root/reference (see
// oracle/Makefile target `ref`; nothing is copied into this repository):
//   core/src/TPS_RGBD.cu, core/src/TPS_RGBD_kernels.cu            -> class TPS_RGBD
//   core/src/dense_registration.cu, ..._kernels.cu + vendored Eigen -> class DenseRegistration
//   core/src/cached_allocator.cpp                                   -> thrust scratch allocator
//   core/src/supersurfel_fusion_kernels.cu (textually included below: its own header
//     drags in the out-of-scope VO / ferns / MOD classes, so its include guard is
//     pre-defined and the four headers it really needs are included instead)
// OpenCV is replaced by the type shim in oracle/ref_shim (GpuMat, Ptr, Rect, Size).
// The only code written here is the driver that replays the launch sequence of
// SupersurfelFusion::processFrame / generateSupersurfels
// (core/src/supersurfel_fusion.cu:166-530, 551-593) around those classes, with the
// out-of-scope neighbours (sparse VO, MOD, ferns, loop closure) left out exactly as the
// CPU oracle leaves them out.  The full reference library cannot be built here (ROS,
// OpenCV-CUDA, g2o, SuiteSparse, darknet are absent).
#include <cuda.h>
#include <cuda_runtime.h>

#include "ref_harness.h"

#include <thrust/copy.h>
#include <thrust/execution_policy.h>
#include <thrust/sort.h>

// --- the reference's surfel kernels, included as they are ---------------------------
#define SUPERSURFEL_FUSION_KERNELS_CUH
#include <supersurfel_fusion/cuda_utils_dev.cuh>
#include <supersurfel_fusion/matrix_math.cuh>
#include <supersurfel_fusion/reduce_dev.cuh>
#include <core/src/supersurfel_fusion_kernels.cu>

#include <chrono>
#include <cstring>
#include <vector>

namespace {

__global__ void shim_rgb_to_rgba(const unsigned char* src, size_t sstep, unsigned char* dst, size_t dstep, int w, int h) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const unsigned char* s = src + y * sstep + 3 * x;
  unsigned char* d = dst + y * dstep + 4 * x;
  d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = 255;
}

}  // namespace

namespace cv { namespace cuda {
void cvtColor(const GpuMat& src, GpuMat& dst, int, int) {
  dst.create(src.rows, src.cols, CV_8UC4);
  dim3 b(32, 8), g((src.cols + 31) / 32, (src.rows + 7) / 8);
  shim_rgb_to_rgba<<<g, b>>>(src.data, src.step, dst.data, dst.step, src.cols, src.rows);
}
}}  // namespace cv::cuda

static void resize_set(sf::Supersurfels& s, size_t n) {
  s.positions.resize(n); s.colors.resize(n); s.stamps.resize(n); s.orientations.resize(n);
  s.shapes.resize(n); s.dims.resize(n); s.confidences.resize(n);
}

static void upload_set(sf::Supersurfels& s, const RefSurfelsHost* h, size_t n) {
  cudaMemcpy(thrust::raw_pointer_cast(s.positions.data()), h->positions, n * 12, cudaMemcpyHostToDevice);
  cudaMemcpy(thrust::raw_pointer_cast(s.colors.data()), h->colors, n * 12, cudaMemcpyHostToDevice);
  cudaMemcpy(thrust::raw_pointer_cast(s.stamps.data()), h->stamps, n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(thrust::raw_pointer_cast(s.orientations.data()), h->orientations, n * 36, cudaMemcpyHostToDevice);
  cudaMemcpy(thrust::raw_pointer_cast(s.shapes.data()), h->shapes, n * 24, cudaMemcpyHostToDevice);
  cudaMemcpy(thrust::raw_pointer_cast(s.dims.data()), h->dims, n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(thrust::raw_pointer_cast(s.confidences.data()), h->confidences, n * 4, cudaMemcpyHostToDevice);
}

static void download_set(const sf::Supersurfels& s, const RefSurfelsHost* h, size_t n) {
  cudaMemcpy(h->positions, thrust::raw_pointer_cast(s.positions.data()), n * 12, cudaMemcpyDeviceToHost);
  cudaMemcpy(h->colors, thrust::raw_pointer_cast(s.colors.data()), n * 12, cudaMemcpyDeviceToHost);
  cudaMemcpy(h->stamps, thrust::raw_pointer_cast(s.stamps.data()), n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(h->orientations, thrust::raw_pointer_cast(s.orientations.data()), n * 36, cudaMemcpyDeviceToHost);
  cudaMemcpy(h->shapes, thrust::raw_pointer_cast(s.shapes.data()), n * 24, cudaMemcpyDeviceToHost);
  cudaMemcpy(h->dims, thrust::raw_pointer_cast(s.dims.data()), n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(h->confidences, thrust::raw_pointer_cast(s.confidences.data()), n * 4, cudaMemcpyDeviceToHost);
}

// replay of generateSupersurfels (supersurfel_fusion.cu:551-593)
static void ref_generate_impl(RefEngine* e) {
  const size_t S = e->nbSuperpixels;
  resize_set(e->frame, S);
  e->frame.memset(S);
  sf::computeSupersurfelCoeffs<<<e->grdIm, e->blkIm>>>(
      RAW(e->frame.positions), RAW(e->frame.shapes), RAW(e->frame.colors), RAW(e->frame.confidences),
      e->tps->getTexRGBA()->getTextureObject(), e->texDepth->getTextureObject(),
      e->tps->getTexIndex()->getTextureObject(), e->tps->getTexInliers()->getTextureObject(),
      e->tps->getTexBound()->getTextureObject(), e->cam.width, e->cam.height, e->cam.fx, e->cam.fy, e->cam.cx,
      e->cam.cy);
  cudaDeviceSynchronize();
  CudaCheckError();
  sf::computeSupersurfels<<<e->grdList, e->blkList>>>(
      RAW(e->frame.positions), RAW(e->frame.colors), RAW(e->frame.stamps), RAW(e->frame.orientations),
      RAW(e->frame.shapes), RAW(e->frame.dims), RAW(e->frame.confidences), e->p.range_min, e->p.range_max, e->stamp,
      (int)S);
  cudaDeviceSynchronize();
  CudaCheckError();
}

// replay of the model-update block (supersurfel_fusion.cu:351-483)
static void ref_fuse_impl(RefEngine* e) {
  const int S = e->nbSuperpixels;
  if (e->nbSupersurfels > 0) {
    cudaMemset(e->nbRemovedDev, 0, sizeof(int));
    thrust::device_vector<bool> matched(S, false);
    if (e->nbVisible > 0) {
      thrust::device_vector<float2> idx_scores(S, make_float2(-1.0f, 0.05f));
      sf::findBestMatches<<<(e->nbVisible + 127) / 128, 128>>>(
          RAW(e->frame.positions), RAW(e->frame.colors), RAW(e->frame.orientations), RAW(e->frame.confidences),
          RAW(e->model.positions), RAW(e->model.colors), RAW(e->model.orientations), RAW(e->model.confidences),
          RAW(matched), RAW(idx_scores), e->tps->getTexIndex()->getTextureObject(), e->pose.R, e->pose.t, e->cam.fx,
          e->cam.fy, e->cam.cx, e->cam.cy, e->p.range_min, e->p.range_max, e->cam.width, e->cam.height, e->nbVisible);
      cudaDeviceSynchronize();
      CudaCheckError();
      sf::updateSupersurfels<<<e->grdList, e->blkList>>>(
          RAW(e->frame.positions), RAW(e->frame.colors), RAW(e->frame.shapes), RAW(e->frame.confidences),
          RAW(e->model.positions), RAW(e->model.colors), RAW(e->model.stamps), RAW(e->model.orientations),
          RAW(e->model.shapes), RAW(e->model.dims), RAW(e->model.confidences), RAW(matched), RAW(idx_scores),
          e->pose.R, e->pose.t, e->stamp, S);
      cudaDeviceSynchronize();
      CudaCheckError();
    }
    sf::insertSupersurfels<<<e->grdList, e->blkList>>>(
        RAW(e->frame.positions), RAW(e->frame.colors), RAW(e->frame.orientations), RAW(e->frame.shapes),
        RAW(e->frame.dims), RAW(e->frame.confidences), RAW(e->model.positions), RAW(e->model.colors),
        RAW(e->model.stamps), RAW(e->model.orientations), RAW(e->model.shapes), RAW(e->model.dims),
        RAW(e->model.confidences), e->pose.R, e->pose.t, e->stamp, RAW(matched), e->nbSupersurfelsDev, S,
        e->p.nb_supersurfels_max);
    cudaDeviceSynchronize();
    CudaCheckError();
    cudaMemcpy(&e->nbSupersurfels, e->nbSupersurfelsDev, sizeof(int), cudaMemcpyDeviceToHost);

    thrust::device_vector<int> states(e->nbSupersurfels, 0);
    e->nbVisible = 0;
    int* nb_visible_d;
    cudaMalloc((void**)&nb_visible_d, sizeof(int));
    cudaMemcpy(nb_visible_d, &e->nbVisible, sizeof(int), cudaMemcpyHostToDevice);
    Mat33 R_view = transpose(e->pose.R);
    float3 t_view = -R_view * e->pose.t;
    sf::filterModel<<<(e->nbSupersurfels + 127) / 128, 128>>>(
        RAW(e->model.positions), RAW(e->model.stamps), RAW(e->model.confidences), RAW(states), e->stamp, e->p.delta_t,
        e->p.conf_thresh, e->texDepth->getTextureObject(), e->nbRemovedDev, nb_visible_d, R_view, t_view, e->cam.fx,
        e->cam.fy, e->cam.cx, e->cam.cy, e->p.range_min, e->p.range_max, e->cam.width, e->cam.height,
        e->nbSupersurfels);
    cudaDeviceSynchronize();
    CudaCheckError();
    cudaMemcpy(&e->nbRemoved, e->nbRemovedDev, sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(&e->nbVisible, nb_visible_d, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(nb_visible_d);
    thrust::sort_by_key(thrust::cuda::par(e->allocator), states.begin(), states.end(), e->model.begin());
    e->nbSupersurfels -= e->nbRemoved;
    cudaMemcpy(e->nbSupersurfelsDev, &e->nbSupersurfels, sizeof(int), cudaMemcpyHostToDevice);
  } else {
    thrust::copy(e->frame.begin(), e->frame.end(), e->model.begin());
    e->nbSupersurfels = S;
    e->nbVisible = e->nbSupersurfels;
    e->nbRemoved = 0;
    cudaMemcpy(e->nbSupersurfelsDev, &e->nbSupersurfels, sizeof(int), cudaMemcpyHostToDevice);
  }
}

static bool ref_icp_impl(RefEngine* e, const Mat33& R_view, const float3& t_view, Mat33& R_rel, float3& t_rel) {
  thrust::host_vector<float3> none_a, none_b;   // the sparse-feature lists are always empty (supersurfel_fusion.cu:244-295)
  return e->icp->featureConstrainedSymmetricICP(e->model.positions, e->model.colors, e->model.orientations,
                                                e->frame.colors, e->frame.orientations, e->frame.confidences, none_a,
                                                none_b, e->nbVisible, e->texDepth, e->tps->getTexIndex(), R_view,
                                                t_view, e->cam, R_rel, t_rel);
}

static void mat_to(const Mat33& m, float* r) {
  r[0] = m.rows[0].x; r[1] = m.rows[0].y; r[2] = m.rows[0].z; r[3] = m.rows[1].x; r[4] = m.rows[1].y;
  r[5] = m.rows[1].z; r[6] = m.rows[2].x; r[7] = m.rows[2].y; r[8] = m.rows[2].z;
}

// pose <- pose o rel with the quaternion round trip of supersurfel_fusion.cu:313-328,
// done with the reference's own rotMatToQuat-free path: Eigen is only linked inside
// dense_registration.cu, so the float quaternion renormalisation is restated here.
static void compose_pose(Transform3& pose, const Mat33& R_rel, const float3& t_rel) {
  pose.t = pose.R * t_rel + pose.t;
  pose.R = pose.R * R_rel;
  float m[3][3] = {{pose.R.rows[0].x, pose.R.rows[0].y, pose.R.rows[0].z},
                   {pose.R.rows[1].x, pose.R.rows[1].y, pose.R.rows[1].z},
                   {pose.R.rows[2].x, pose.R.rows[2].y, pose.R.rows[2].z}};
  float q[4];
  float t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.f) {
    t = sqrtf(t + 1.f); q[3] = 0.5f * t; t = 0.5f / t;
    q[0] = (m[2][1] - m[1][2]) * t; q[1] = (m[0][2] - m[2][0]) * t; q[2] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrtf(m[i][i] - m[j][j] - m[k][k] + 1.f); q[i] = 0.5f * t; t = 0.5f / t;
    q[3] = (m[k][j] - m[j][k]) * t; q[j] = (m[j][i] + m[i][j]) * t; q[k] = (m[k][i] + m[i][k]) * t;
  }
  const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int a = 0; a < 4; a++) q[a] /= n;
  const float tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  pose.R = make_mat33(1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy,
                      tyz + twx, 1 - (txx + tyy));
}

extern "C" {

RefEngine* ref_create(const RefParams* p) {
  RefEngine* e = new RefEngine;
  e->p = *p;
  e->cam.fx = p->fx; e->cam.fy = p->fy; e->cam.cx = p->cx; e->cam.cy = p->cy;
  e->cam.height = p->height; e->cam.width = p->width;
  // SupersurfelFusion::initialize (supersurfel_fusion.cu:85-136)
  e->blkIm = dim3(32, 32);
  e->grdIm = dim3((p->width + 31) / 32, (p->height + 31) / 32);
  const int gx = (p->width + p->cell_size - 1) / p->cell_size, gy = (p->height + p->cell_size - 1) / p->cell_size;
  e->nbSuperpixels = gx * gy;
  e->blkList = dim3(128);
  e->grdList = dim3((e->nbSuperpixels + 127) / 128);
  e->filteredDepth.create(p->height, p->width, CV_32FC1);
  e->texDepth = new sf::Texture<float>(e->filteredDepth);
  e->tps = new sf::TPS_RGBD(p->cell_size, p->lambda_pos, p->lambda_bound, p->lambda_size, p->lambda_disp,
                            p->thresh_disp, p->seg_iter, p->seg_use_ransac != 0, p->nb_samples, p->filter_iter,
                            p->filter_alpha, p->filter_beta, p->filter_threshold);
  e->nbVisible = e->nbSupersurfels = e->nbRemoved = 0;
  resize_set(e->model, p->nb_supersurfels_max);
  e->model.memset(p->nb_supersurfels_max);
  resize_set(e->frame, e->nbSuperpixels);
  cudaMalloc(&e->nbSupersurfelsDev, sizeof(int));
  cudaMalloc(&e->nbRemovedDev, sizeof(int));
  cudaMemset(e->nbSupersurfelsDev, 0, sizeof(int));
  cudaMemset(e->nbRemovedDev, 0, sizeof(int));
  e->pose.R = make_mat33(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
  e->pose.t = make_float3(0.f, 0.f, 0.f);
  e->icp = new sf::DenseRegistration(p->icp_iter, p->icp_cov_thresh);
  e->stamp = 0;
  e->rgb.create(p->height, p->width, CV_8UC3);
  e->depth.create(p->height, p->width, CV_32FC1);
  for (int i = 0; i < 5; i++) cudaEventCreate(&e->ev[i]);
  return e;
}

void ref_destroy(RefEngine* e) {
  if (!e) return;
  cudaDeviceSynchronize();
  delete e->tps;
  delete e->icp;
  cudaFree(e->nbSupersurfelsDev);
  cudaFree(e->nbRemovedDev);
  for (int i = 0; i < 5; i++) cudaEventDestroy(e->ev[i]);
  delete e;
}

int ref_nb_superpixels(RefEngine* e) { return e->nbSuperpixels; }

// tps->compute / filter / computeDepthImage (supersurfel_fusion.cu:189-191)
void ref_tps(RefEngine* e, const unsigned char* rgb, const float* depth) {
  e->rgb.upload_dense(rgb);
  e->depth.upload_dense(depth);
  e->tps->compute(e->rgb, e->depth);
  e->tps->filter();
  e->tps->computeDepthImage(e->filteredDepth);
  cudaDeviceSynchronize();
}

void ref_get_segmentation(RefEngine* e, int* labels, int* bound, unsigned char* inliers, float* superpixels,
                          float* slanted) {
  cudaDeviceSynchronize();
  if (labels) e->tps->getIndexImage().download_dense(labels);
  if (bound) e->tps->getBoundaryImage().download_dense(bound);
  if (inliers) e->tps->getInliersImage().download_dense(inliers);
  if (slanted) e->filteredDepth.download_dense(slanted);
  if (superpixels)
    cudaMemcpy(superpixels, thrust::raw_pointer_cast(e->tps->getSuperpixels().data()),
               (size_t)e->nbSuperpixels * sizeof(sf::SuperpixelRGBD), cudaMemcpyDeviceToHost);
}

// overwrite the segmentation images in place (device buffers keep their textures)
void ref_set_segmentation(RefEngine* e, const int* labels, const int* bound, const unsigned char* inliers,
                          const float* slanted, const unsigned char* rgba) {
  if (labels) const_cast<cv::cuda::GpuMat&>(e->tps->getIndexImage()).upload_dense(labels);
  if (bound) const_cast<cv::cuda::GpuMat&>(e->tps->getBoundaryImage()).upload_dense(bound);
  if (inliers) const_cast<cv::cuda::GpuMat&>(e->tps->getInliersImage()).upload_dense(inliers);
  if (rgba) const_cast<cv::cuda::GpuMat&>(e->tps->getRGBAImage()).upload_dense(rgba);
  if (slanted) e->filteredDepth.upload_dense(slanted);
  cudaDeviceSynchronize();
}

void ref_generate(RefEngine* e, int stamp) {
  e->stamp = stamp;
  ref_generate_impl(e);
}

void ref_get_frame(RefEngine* e, const RefSurfelsHost* out) { download_set(e->frame, out, e->nbSuperpixels); }
void ref_set_frame(RefEngine* e, const RefSurfelsHost* in) { upload_set(e->frame, in, e->nbSuperpixels); }
void ref_get_model(RefEngine* e, const RefSurfelsHost* out, int n) { download_set(e->model, out, n); }
void ref_set_model(RefEngine* e, const RefSurfelsHost* in, int n, int n_visible) {
  upload_set(e->model, in, n);
  e->nbSupersurfels = n;
  e->nbVisible = n_visible;
  cudaMemcpy(e->nbSupersurfelsDev, &n, sizeof(int), cudaMemcpyHostToDevice);
}
void ref_set_pose(RefEngine* e, const float* R9, const float* t3) {
  e->pose.R = mat_from(R9);
  e->pose.t = make_float3(t3[0], t3[1], t3[2]);
}
void ref_get_pose(RefEngine* e, float* R9, float* t3) {
  mat_to(e->pose.R, R9);
  t3[0] = e->pose.t.x; t3[1] = e->pose.t.y; t3[2] = e->pose.t.z;
}
void ref_get_counts(RefEngine* e, int* c4) {
  c4[0] = e->nbSupersurfels; c4[1] = e->nbVisible; c4[2] = e->nbRemoved; c4[3] = e->stamp;
}

// DenseRegistration::featureConstrainedSymmetricICP, the reference's own host loop
int ref_icp(RefEngine* e, const float* Rview9, const float* tview3, float* Rrel9, float* trel3) {
  Mat33 R_rel;
  float3 t_rel;
  const bool ok = ref_icp_impl(e, mat_from(Rview9), make_float3(tview3[0], tview3[1], tview3[2]), R_rel, t_rel);
  mat_to(R_rel, Rrel9);
  trel3[0] = t_rel.x; trel3[1] = t_rel.y; trel3[2] = t_rel.z;
  return ok ? 1 : 0;
}

void ref_fuse(RefEngine* e, int stamp) {
  e->stamp = stamp;
  ref_fuse_impl(e);
}

// processFrame without its out-of-scope neighbours (supersurfel_fusion.cu:166-530)
void ref_process_frame(RefEngine* e, const unsigned char* rgb, const float* depth, const float* prior12,
                       RefStats* st) {
  auto w0 = std::chrono::steady_clock::now();
  cudaEventRecord(e->ev[0]);
  e->rgb.upload_dense(rgb);       // rgb.upload / depth.upload (:173-174)
  e->depth.upload_dense(depth);
  e->tps->compute(e->rgb, e->depth);
  e->tps->filter();
  e->tps->computeDepthImage(e->filteredDepth);
  cudaEventRecord(e->ev[1]);
  ref_generate_impl(e);
  cudaEventRecord(e->ev[2]);
  if (prior12) ref_set_pose(e, prior12, prior12 + 9);
  int icp_ran = 0, icp_ok = 0;
  if (e->nbVisible > 0) {
    icp_ran = 1;
    Mat33 R_view = transpose(e->pose.R);
    float3 t_view = -R_view * e->pose.t;
    Mat33 R_rel = make_mat33(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
    float3 t_rel = make_float3(0.f, 0.f, 0.f);
    if (ref_icp_impl(e, R_view, t_view, R_rel, t_rel)) {
      icp_ok = 1;
      compose_pose(e->pose, R_rel, t_rel);
    }
  }
  cudaEventRecord(e->ev[3]);
  ref_fuse_impl(e);
  cudaEventRecord(e->ev[4]);
  cudaEventSynchronize(e->ev[4]);
  auto w1 = std::chrono::steady_clock::now();
  if (st) {
    st->stamp = e->stamp; st->nb_supersurfels = e->nbSupersurfels; st->nb_visible = e->nbVisible;
    st->nb_removed = e->nbRemoved; st->icp_ran = icp_ran; st->icp_valid = icp_ok;
    cudaEventElapsedTime(&st->ms_tps, e->ev[0], e->ev[1]);
    cudaEventElapsedTime(&st->ms_generate, e->ev[1], e->ev[2]);
    cudaEventElapsedTime(&st->ms_icp, e->ev[2], e->ev[3]);
    cudaEventElapsedTime(&st->ms_fuse, e->ev[3], e->ev[4]);
    cudaEventElapsedTime(&st->ms_total, e->ev[0], e->ev[4]);
    st->wall_ms = std::chrono::duration<float, std::milli>(w1 - w0).count();
  }
  e->stamp++;
}

}  // extern "C"
