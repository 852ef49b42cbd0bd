/* libssf -- C-ABI of the B200-native supersurfel tracking-and-fusion hot path.
 *
 * Drop-in boundary for the reference's C++ class supersurfel_fusion::SupersurfelFusion
 * (reference: core/include/supersurfel_fusion/supersurfel_fusion.hpp:40-143).  Every
 * entry point cites the reference interface it replaces.  Plain C types only: no
 * torch, thrust, OpenCV or Eigen types cross this boundary.
 *
 * Conventions
 *  - every function returns an int status: SSF_OK (0) or a negative SSF_ERR_* code;
 *    the reference instead prints and exit(-1)s (cuda_error_check.h:30-66).
 *    ssf_last_error() returns a human-readable message for the last failure.
 *  - a handle is bound to one CUDA device and one stream and is NOT thread-safe
 *    (the reference is driven from a single ROS callback thread,
 *    node/supersurfel_fusion_node.cpp:74-85).  N handles on N devices from N host
 *    threads / processes is the supported multi-GPU mode.
 *  - "host-or-device pointer": the argument may point to pageable/pinned host
 *    memory or to device memory of the handle's device (copies use
 *    cudaMemcpyDefault).
 *  - supersurfel arrays use the member layout of the reference's Supersurfels
 *    container (core/include/supersurfel_fusion/supersurfels.hpp:32-41):
 *      positions float[N][3], colors float[N][3] (RGB 0..255), stamps int[N][2],
 *      orientations float[N][9] (rows e1,e2,normal), shapes float[N][6]
 *      (xx,xy,xz,yy,yz,zz), dims float[N][2], confidences float[N].
 *  - poses are (R row-major 3x3, t) mapping camera to world, as Transform3
 *    (core/include/supersurfel_fusion/matrix_types.h:38-42).
 */
#ifndef SSF_H
#define SSF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSF_OK 0
#define SSF_ERR_INVALID_ARG (-1)
#define SSF_ERR_CUDA (-2)
#define SSF_ERR_NO_DEVICE (-3)
#define SSF_ERR_IO (-4)
#define SSF_ERR_STATE (-5)

typedef struct SsfEngine* SsfHandle;

/* core/include/supersurfel_fusion/cam_param.hpp:27-31 */
typedef struct SsfCamParam {
  float fx, fy, cx, cy;
  int height, width;
} SsfCamParam;

/* The hot-path arguments of SupersurfelFusion::initialize
 * (supersurfel_fusion.hpp:46-74), same names, same defaults.  The sparse-VO
 * arguments (nb_features ... untracked_threshold) belong to an out-of-scope
 * neighbour and are not part of this struct.  enable_loop_closure / enable_mod keep the
 * reference's defaults (true, supersurfel_fusion.hpp:72-73) and are RECORDED, not acted on:
 * the ferns loop detector + deformation-graph optimisation and the MOD / YOLO detector are
 * out-of-scope neighbours of this path.  A caller that runs them feeds their results in
 * through the hooks below (ssf_invalidate_frame_supersurfels, ssf_align, ssf_apply_deformation,
 * ssf_set_pose, ssf_transform_model); with no caller driving the hooks the library behaves as
 * the reference does with both switches off. */
typedef struct SsfConfig {
  SsfCamParam cam;
  int cell_size;          /* 16 */
  float lambda_pos;       /* 50 */
  float lambda_bound;     /* 1000 */
  float lambda_size;      /* 10000 */
  float lambda_disp;      /* 1e6 */
  float thresh_disp;      /* 1e-4 */
  int seg_iter;           /* 10 */
  int seg_use_ransac;     /* 1 */
  int nb_samples;         /* 16 */
  int filter_iter;        /* 4 */
  float filter_alpha;     /* 0.1 */
  float filter_beta;      /* 1.0 */
  float filter_threshold; /* 0.05 */
  float range_min;        /* 0.2 */
  float range_max;        /* 5.0 */
  int delta_t;            /* 20 */
  float conf_thresh;      /* 2500 */
  int nb_supersurfels_max;/* 50000 */
  int icp_iter;           /* 10 */
  double icp_cov_thresh;  /* 0.04 */
  int enable_loop_closure;/* 1 (recorded; see above) */
  int enable_mod;         /* 1 (recorded; see above) */
} SsfConfig;

/* Host-or-device view of supersurfel arrays (any member may be NULL = skip). */
typedef struct SsfSurfels {
  float* positions;
  float* colors;
  int32_t* stamps;
  float* orientations;
  float* shapes;
  float* dims;
  float* confidences;
} SsfSurfels;

/* What the reference prints per frame (supersurfel_fusion.cu:516-528,
 * dense_registration.cu:336-341) returned as data. */
typedef struct SsfFrameStats {
  int32_t stamp;            /* stamp of the processed frame */
  int32_t nb_supersurfels;  /* model size after the update */
  int32_t nb_visible;       /* active (in-view) prefix length */
  int32_t nb_removed;
  int32_t nb_matched;       /* frame superpixels fused into the model */
  int32_t nb_inserted;
  int32_t icp_ran;          /* 0 on the bootstrap frame */
  int32_t icp_valid;
  int32_t icp_iters;
  float icp_inliers;
  double icp_error;         /* sqrt(r / inliers) of the last built system */
  float gpu_ms;             /* device time of the frame (CUDA events) */
  /* Per-stage device times, the breakdown the reference prints per frame
   * (supersurfel_fusion.cu:516-528); filled by the synchronous entry points when
   * SSF_FLAG_STAGE_TIMING is set (event nodes at the stage boundaries of the frame graph), else 0.
   * Not available in pipelined mode, where the stages of consecutive frames overlap. */
  float ms_ingest;          /* [bilateral filter,] RGBA / disparity conversion, grid seeding */
  float ms_segmentation;    /* tps->compute + filter + computeDepthImage */
  float ms_extraction;      /* generateSupersurfels */
  float ms_registration;    /* featureConstrainedSymmetricICP + pose composition */
  float ms_fusion;          /* association, update, insert, cull + partition */
} SsfFrameStats;

/* ---- lifecycle ----------------------------------------------------------- */
/* Defaults of initialize() (supersurfel_fusion.hpp:46-74); camera = TUM fr1. */
int ssf_config_default(SsfConfig* cfg);
/* SupersurfelFusion() + initialize() (supersurfel_fusion.cu:49-164). */
int ssf_create(const SsfConfig* cfg, int device, SsfHandle* out);
/* ~SupersurfelFusion() (supersurfel_fusion.cu:40-47). */
int ssf_destroy(SsfHandle h);
/* Run all work of this handle on an existing CUDA stream (cudaStream_t / CUstream
 * passed as void*); NULL restores the handle's own stream. */
int ssf_set_stream(SsfHandle h, void* cuda_stream);
const char* ssf_last_error(SsfHandle h);
/* isInitialized() (supersurfel_fusion.hpp:85) */
int ssf_is_initialized(SsfHandle h);

/* ---- the per-frame entry point ------------------------------------------- */
/* processFrame(rgb_h 8UC3 RGB, depth_h 32FC1 metres) (supersurfel_fusion.cu:166-530).
 * rgb/depth: host-or-device pointers, strides in BYTES (cv::Mat::step).
 * pose_prior_Rt12: optional camera pose prior (R row-major 9 floats, then t), the
 * stand-in for the sparse-VO pose (supersurfel_fusion.cu:225-228); NULL keeps the
 * previous fused pose.  Depth is expected already bilateral-filtered unless
 * SSF_FLAG_BILATERAL is set.  Synchronous, like the reference. */
#define SSF_FLAG_BILATERAL 1u
#define SSF_FLAG_STAGE_TIMING 2u /* fill SsfFrameStats.ms_* (synchronous entry points only) */
int ssf_process_frame(SsfHandle h, const uint8_t* rgb, size_t rgb_stride, const float* depth,
                      size_t depth_stride, const float* pose_prior_Rt12, uint32_t flags);
/* Same with the 16-bit depth image of the TUM / live drivers: decodes
 * depth.convertTo(CV_32FC1, depth_scale) on the device
 * (node/supersurfel_fusion_rgbd_benchmark_node.cpp:609-610; node/supersurfel_fusion_node.cpp
 * does the same with its depth_scale parameter).  depth_stride in BYTES. */
int ssf_process_frame_depth16(SsfHandle h, const uint8_t* rgb, size_t rgb_stride, const uint16_t* depth16,
                              size_t depth_stride, float depth_scale, const float* pose_prior_Rt12,
                              uint32_t flags);
/* Same, inputs already resident on the handle's device (dense, stride = width). */
int ssf_process_frame_device(SsfHandle h, const uint8_t* rgb_dev, const float* depth_dev,
                             const float* pose_prior_Rt12, uint32_t flags);
int ssf_get_frame_stats(SsfHandle h, SsfFrameStats* out);
/* Pipelined form of processFrame for streams of frames: ssf_submit_frame() enqueues a frame and
 * returns at once, ssf_wait_frame() blocks until the OLDEST submitted frame is done and returns
 * its stats and pose.  The frame's kernel sequence (ingest -> segmentation iterations -> extraction
 * -> registration + fusion; only the last part touches the model) is cut into
 * ssf_get_pipeline_depth() stages of about equal cost (default 6, SSF_PIPELINE_STAGES=1..6), each
 * on its own stream, and that many frames may be in flight, one per stage.  Same kernels in the
 * same order per frame: results are identical to ssf_process_frame; latency per frame is
 * unchanged, the frame rate is set by the longest stage.  The input buffers must stay valid until
 * the frame has been waited for (pinned host memory or device memory for a truly asynchronous
 * copy).  The synchronous entry points and the getters require that no frame is in flight; a
 * caller that must edit the frame between stages (ssf_invalidate_frame_supersurfels, the MOD
 * hook) drives the stage entry points below instead.
 * Pose priors: a prior given to ssf_submit_frame is fixed at submit time, i.e. BEFORE the up to
 * depth-1 older frames in flight have been fused -- it cannot be derived from the previous
 * frame's fused pose (an external odometry source that runs ahead of the fusion is fine; NULL,
 * "keep the previous fused pose", is always exact because the last stage reads the pose on the
 * device).  A visual-odometry front end that needs frame k's fused pose to predict frame k+1
 * must use the synchronous ssf_process_frame.
 * Error handling: if a submit fails after part of the frame was enqueued the handle is retired
 * (every later call returns SSF_ERR_STATE; ssf_last_error tells why); destroy and re-create it.
 * ssf_prepare() builds and uploads every CUDA graph the given flags need (synchronous frame
 * graph and all slot x stage graphs) so that the first frames do not pay for stream capture and
 * graph instantiation; without it the graphs are built lazily by the first calls. */
int ssf_prepare(SsfHandle h, uint32_t flags);
int ssf_submit_frame(SsfHandle h, const uint8_t* rgb, size_t rgb_stride, const float* depth,
                     size_t depth_stride, const float* pose_prior_Rt12, uint32_t flags);
int ssf_wait_frame(SsfHandle h, SsfFrameStats* out, float R[9], float t[3]);
int ssf_get_pipeline_depth(SsfHandle h, int* stages);
/* How a configuration's frame is cut into pipeline stages (no device needed): the frame is the
 * step sequence 0 = ingest, 1 .. T = segmentation steps (T = seg_iter + 2: colour iterations,
 * disparity-plane initialisation, colour + disparity iterations, smoothing + render), T + 1 =
 * extraction, T + 2 = registration + fusion; stage p runs steps [first[p], first[p + 1]).
 * first must hold stages + 1 ints.  Returns the number of stages used (>= 1), *nb_steps = T + 3. */
int ssf_plan_pipeline(const SsfConfig* cfg, int stages, int persistent_segmentation, int* first, int* nb_steps);
/* The per-step cost estimates (microseconds on a B200 at 640x480; only their ratios matter) that
 * ssf_plan_pipeline balances.  Writes one int per frame step, returns the number of steps. */
int ssf_plan_weights(const SsfConfig* cfg, int persistent_segmentation, int* weights, int capacity);

/* ---- ingest (supersurfel_fusion.cu:171-181) -------------------------------- */
/* cv::cuda::bilateralFilter(depth, depth, kernel_size, sigma_color, sigma_spatial)
 * (supersurfel_fusion.cu:180 calls it with -1, 0.03, 4.5), out of place; host-or-device
 * pointers, H x W floats.  SSF_FLAG_BILATERAL runs it inside ssf_process_frame*. */
int ssf_bilateral_filter(SsfHandle h, const float* depth, size_t depth_stride, int kernel_size,
                         float sigma_color, float sigma_spatial, float* out);
/* The filtered depth of the last frame processed with SSF_FLAG_BILATERAL (what the
 * reference downloads for its VO, supersurfel_fusion.cu:181). */
int ssf_get_filtered_depth(SsfHandle h, float* depth);
/* cv::cuda::cvtColor(rgb, gray, CV_RGB2GRAY) of the last frame's colour image
 * (supersurfel_fusion.cu:175-177; consumed by the out-of-scope VO / MOD). */
int ssf_get_gray(SsfHandle h, uint8_t* gray);

/* ---- getters (supersurfel_fusion.hpp:85-91) ------------------------------ */
int ssf_get_pose(SsfHandle h, float R[9], float t[3]);          /* getPose() */
int ssf_set_pose(SsfHandle h, const float R[9], const float t[3]); /* loop-closure / relocalisation hook */
int ssf_get_stamp(SsfHandle h, int* stamp);                      /* getStamp() */
int ssf_set_stamp(SsfHandle h, int stamp);
int ssf_get_counts(SsfHandle h, int* nb_supersurfels, int* nb_visible, int* nb_removed); /* getnbSupersurfels() */
int ssf_get_nb_superpixels(SsfHandle h, int* nb_superpixels);
/* getModel() / getFrame(): copy the first n rows into caller buffers
 * (host-or-device); the node does the same through thrust::host_vector
 * (node/supersurfel_fusion_node.cpp:306-310). */
int ssf_copy_model(SsfHandle h, const SsfSurfels* dst, int n);
int ssf_copy_frame(SsfHandle h, const SsfSurfels* dst);
/* getModel() / getFrame() without a copy, for consumers that live on the same GPU: the engine's
 * own planar storage.  Plane p of element i is base[p * stride + i]; planes: 0-2 position,
 * 3-5 colour (RGB 0..255), 6-7 stamps (int32 bit patterns), 8-16 orientation (rows e1, e2,
 * normal), 17-22 shape (xx xy xz yy yz zz), 23-24 dims, 25 confidence, 26-28 CIELab of the colour.
 * count = nb_supersurfels (model) or the number of superpixels (frame).  Read-only; valid until
 * the next frame is processed. */
typedef struct SsfPlanarView {
  const float* base;   /* device pointer */
  int stride;          /* elements per plane */
  int count;           /* valid elements */
  int planes;          /* 29 */
} SsfPlanarView;
int ssf_get_model_view(SsfHandle h, SsfPlanarView* out);
int ssf_get_frame_view(SsfHandle h, SsfPlanarView* out);
/* tps->getIndexImage() / getBoundaryImage() / getInliersImage() / getDispImage()
 * (TPS_RGBD.hpp:77-81), the slanted-plane depth (computeDepthImage, TPS_RGBD.cu:507-525)
 * and tps->getSuperpixels() as S x 12 floats (TPS_RGBD.hpp:33-38).  NULL = skip. */
int ssf_get_segmentation(SsfHandle h, int32_t* labels, int32_t* bound, uint8_t* inliers, float* disp,
                         float* slanted_depth, float* superpixels, uint8_t* rgba);
/* computeSuperpixelSegIm (supersurfel_fusion.cu:635-640): H x W x 3 preview, boundaries white. */
int ssf_render_preview(SsfHandle h, uint8_t* bgr);
/* computeSlantedPlaneIm (supersurfel_fusion.cu:642-650): slanted depth, float H x W. */
int ssf_get_slanted_depth(SsfHandle h, float* depth);
/* exportModel(filename) (supersurfel_fusion.cu:595-633), same text format. */
int ssf_export_model(SsfHandle h, const char* path);
/* extractLocalPointCloud (supersurfel_fusion.cu:884-927): stable supersurfels within
 * `radius` of the camera, in camera coordinates.  capacity rows available in the
 * caller buffers (host-or-device); *count receives the number written. */
int ssf_extract_local_point_cloud(SsfHandle h, float radius, float* positions, float* normals,
                                  int capacity, int* count);
/* MOD hook: mask[S] != 0 marks a frame supersurfel dynamic => confidence = -1
 * (what motion_detection.cu:573 does to frame.confidences). */
int ssf_invalidate_frame_supersurfels(SsfHandle h, const uint8_t* mask);
/* Loop-closure hook: rigidly move the whole model (applyTransformSuperSurfel,
 * supersurfel_fusion_kernels.cu:469-488). */
int ssf_transform_model(SsfHandle h, const float R[9], const float t[3]);

/* Loop-closure hook: DeformationGraph::applyGraphToModel -> applyDeformation
 * (deformation_graph.cu:840-861, deformation_graph_kernels.cu:27-73): warp the first
 * model_size model supersurfels by an embedded deformation graph computed by the caller.
 * nodes_*: nb_nodes rows (positions [3], rotations [9] row-major, translations [3]);
 * neighbours_weights [model_size][4], neighbours_idx [model_size][4] (the four nearest
 * nodes of every supersurfel, VertexMap of deformation_graph_types.hpp).  Host-or-device. */
int ssf_apply_deformation(SsfHandle h, const float* nodes_positions, const float* nodes_rotations,
                          const float* nodes_translations, int nb_nodes, const float* neighbours_weights,
                          const int32_t* neighbours_idx, int model_size);
/* ---- consumer formats (the wire / disk formats after the path) ----------------------- */
/* publishModelMarker / publishFrameMarker geometry (node/supersurfel_fusion_node.cpp:303-413,
 * 415-520): per supersurfel 6 points [3] (two triangles of the +-3 sqrt(dims) quad along e1/e2)
 * and 6 RGBA colours [4] (colour / 255, alpha 1); zeros / black below conf_thresh.
 * which: 0 = model (nb_supersurfels rows), 1 = frame.  *count = rows available. */
int ssf_get_markers(SsfHandle h, int which, float conf_thresh, float* points, float* colors, int capacity,
                    int* count);
/* One line of the TUM trajectory file the benchmark node writes
 * (node/supersurfel_fusion_rgbd_benchmark_node.cpp:727-729):
 * "timestamp tx ty tz qx qy qz qw\n" of the current pose. */
int ssf_format_tum_pose(SsfHandle h, const char* timestamp, char* line, size_t line_size);

/* ---- stage entry points (for parity tests and callers that drive stages) --- */
/* State injection: host-or-device arrays in the layouts above. */
int ssf_set_model(SsfHandle h, const SsfSurfels* src, int nb_supersurfels, int nb_visible);
int ssf_set_frame(SsfHandle h, const SsfSurfels* src);
int ssf_set_segmentation(SsfHandle h, const int32_t* labels, const int32_t* bound, const uint8_t* inliers,
                         const float* slanted_depth, const uint8_t* rgba);
/* tps->compute + filter + computeDepthImage (supersurfel_fusion.cu:189-191). */
int ssf_tps_segment(SsfHandle h, const uint8_t* rgb, size_t rgb_stride, const float* depth,
                    size_t depth_stride);
/* RANSAC candidate planes of the last ssf_tps_segment: S x nb_samples x 4 floats. */
int ssf_get_ransac_samples(SsfHandle h, float* samples);
/* generateSupersurfels() (supersurfel_fusion.cu:551-593). */
int ssf_generate_supersurfels(SsfHandle h);
/* One launch of the symmetric ICP system build at the view transform (R,t)
 * (computeSymmetricICPSystem<128>, dense_registration_kernels.cuh:175-291):
 * out29 = JtJ[21] (upper triangle, row-major), Jtr[6], r, inliers
 * (MotionTrackingData, dense_registration_types.hpp:55-60).  n_src <= 0 uses the
 * model's visible prefix. */
int ssf_icp_system(SsfHandle h, const float R[9], const float t[3], int n_src, float out29[29]);
/* Enqueue-only variant for timing: `launches` back-to-back builds on the handle's
 * stream, no host sync, result left on the device. */
int ssf_icp_system_enqueue(SsfHandle h, const float R[9], const float t[3], int n_src, int launches);
/* featureConstrainedSymmetricICP (dense_registration.cu:245-424), whole
 * Gauss-Newton loop on the device.  R_init/t_init = view transform (inverse of the
 * prior pose); NULL = derive from the handle's current pose.  Does not modify the
 * pose.  Returns R_rel/t_rel (identity/zero when invalid). */
int ssf_icp(SsfHandle h, const float* R_init, const float* t_init, float out29[29], float R_rel[9],
            float t_rel[3], int* iters, int* valid);
/* The same loop driven step by step, for tile-parallel registration of one large frame across
 * GPUs: every rank holds the whole frame state and a copy of the model,
 * builds the system over ITS slice of the visible prefix, the 29 floats are summed across
 * ranks in rank order (see supersurfel_fusion_b200/multi.py), and every rank applies the
 * identical Gauss-Newton step.  src_begin must be a multiple of 4.
 *   ssf_icp_begin  : dense_registration.cu:262-299 (loop set-up)
 *   ssf_icp_build  : one computeSymmetricICPSystem launch over [src_begin, src_begin+src_count)
 *   ssf_icp_solve  : dense_registration.cu:326-391 with the given (reduced) system
 *   ssf_icp_finish : dense_registration.cu:394-421 (+ supersurfel_fusion.cu:313-328 when apply_to_pose) */
int ssf_icp_begin(SsfHandle h, const float* R_init, const float* t_init);
int ssf_icp_build(SsfHandle h, int src_begin, int src_count, float out29[29]);
int ssf_icp_solve(SsfHandle h, const float sys29[29], int* done);
int ssf_icp_finish(SsfHandle h, int apply_to_pose, float R_rel[9], float t_rel[3], int* iters, int* valid);
/* Fused build + exchange + solve over NVLink peer memory (no NCCL, no host round trip per
 * iteration): every rank's system kernel stores its 29 partial sums straight into every
 * peer's exchange buffer (P2P stores over NVLink), waits for the peers' flags, sums the
 * slices in rank order and runs the Gauss-Newton step -- so all ranks iterate in lock step
 * inside their own kernels.  Set-up: exchange ssf_peer_handle() blobs between the ranks
 * (any transport; bench/tests use torch.distributed.all_gather) and call ssf_connect_peers()
 * once.  ssf_icp_tiled() is collective: every rank must call it with the same R_init/t_init. */
#define SSF_PEER_HANDLE_BYTES 64
#define SSF_MAX_PEERS 8
int ssf_peer_handle(SsfHandle h, void* handle64);
int ssf_connect_peers(SsfHandle h, int rank, int world, const void* handles);
int ssf_icp_tiled(SsfHandle h, const float* R_init, const float* t_init, int src_begin, int src_count,
                  float out29[29], float R_rel[9], float t_rel[3], int* iters, int* valid);
/* DenseRegistration::align (dense_registration.cu:52-243), the loop-closure registration:
 * `source` = a keyframe's supersurfels in ITS camera frame (positions, colors, orientations,
 * confidences are read; host-or-device), target = the handle's current frame (supersurfels,
 * label map, slanted depth).  R_init/t_init map keyframe-camera to current-camera coordinates.
 * Whole loop in one kernel launch.  R/t receive the reference's return value (identity/zero
 * when invalid); pairs = matched pairs of the last iteration; out29 = its (centred, scaled)
 * system.  Uses the handle's icp_iter / icp_cov_thresh like the reference's shared
 * DenseRegistration object (supersurfel_fusion.cu:137,776). */
int ssf_align(SsfHandle h, const SsfSurfels* source, int source_size, const float R_init[9],
              const float t_init[3], float R[9], float t[3], int* valid, int* iters, int* pairs,
              float out29[29]);
/* Model update block of processFrame (supersurfel_fusion.cu:351-483) at the
 * handle's current pose and stamp. */
int ssf_fuse(SsfHandle h);

/* ---- timing helpers (CUDA events on the handle's stream) ------------------ */
int ssf_timer_start(SsfHandle h);
int ssf_timer_stop(SsfHandle h, float* ms); /* records, synchronises, returns elapsed */
int ssf_synchronize(SsfHandle h);
/* Number of kernels this library launched on the handle since creation. */
int ssf_get_launch_count(SsfHandle h, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* SSF_H */
