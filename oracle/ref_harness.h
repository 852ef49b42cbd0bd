// TEST INFRASTRUCTURE ONLY -- shared declarations of the reference-kernel harness
// (see ref_harness.cu for what is compiled and why).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <opencv2/core.hpp>
#include <supersurfel_fusion/TPS_RGBD.hpp>
#include <supersurfel_fusion/cached_allocator.hpp>
#include <supersurfel_fusion/cam_param.hpp>
#include <supersurfel_fusion/cuda_error_check.h>
#include <supersurfel_fusion/dense_registration.hpp>
#include <supersurfel_fusion/matrix_math.cuh>
#include <supersurfel_fusion/supersurfels.hpp>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

namespace sf = supersurfel_fusion;

struct RefParams {
  float fx, fy, cx, cy;
  int height, width;
  int cell_size;
  float lambda_pos, lambda_bound, lambda_size, lambda_disp, thresh_disp;
  int seg_iter, seg_use_ransac, nb_samples, filter_iter;
  float filter_alpha, filter_beta, filter_threshold;
  float range_min, range_max;
  int delta_t;
  float conf_thresh;
  int nb_supersurfels_max, icp_iter;
  double icp_cov_thresh;
};

struct RefSurfelsHost {
  float* positions; float* colors; int* stamps; float* orientations; float* shapes; float* dims; float* confidences;
};

struct RefStats {
  int stamp, nb_supersurfels, nb_visible, nb_removed, icp_ran, icp_valid;
  float ms_tps, ms_generate, ms_icp, ms_fuse, ms_total, wall_ms;
};

struct RefEngine {
  RefParams p;
  sf::CamParam cam;
  sf::TPS_RGBD* tps;
  sf::DenseRegistration* icp;
  sf::Supersurfels model, frame;
  sf::CachedAllocator allocator;
  cv::cuda::GpuMat rgb, depth, filteredDepth;
  cv::Ptr<sf::Texture<float>> texDepth;
  int nbSuperpixels, nbSupersurfels, nbVisible, nbRemoved, stamp;
  int *nbSupersurfelsDev, *nbRemovedDev;
  Transform3 pose;
  dim3 blkIm, grdIm, blkList, grdList;
  cudaEvent_t ev[5];
};

#define RAW(v) thrust::raw_pointer_cast(&(v)[0])

static inline Mat33 mat_from(const float* r) { return make_mat33(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8]); }
