/* TEST INFRASTRUCTURE ONLY -- C API of the CPU oracle (liboracle.so).
 *
 * A CPU restatement of the reference's per-frame hot path, used as the parity
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product (supersurfel_fusion_b200/, libssf.so)
 * never links, loads or calls anything declared here.
 *
 * Array layouts are the reference's Supersurfels members
 * (core/include/supersurfel_fusion/supersurfels.hpp:32-41) copied to the host:
 *   positions    float[N][3]       colors  float[N][3] (RGB 0..255)
 *   stamps       int  [N][2]       orientations float[N][9] (rows e1,e2,normal)
 *   shapes       float[N][6] (xx,xy,xz,yy,yz,zz)   dims float[N][2]
 *   confidences  float[N]
 * Images are dense row-major (pitch == width).
 */
#ifndef SSF_ORACLE_H
#define SSF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcCam { float fx, fy, cx, cy; int height, width; } OrcCam;

typedef struct OrcSurfels {
  float* positions; float* colors; int* stamps; float* orientations;
  float* shapes; float* dims; float* confidences;
} OrcSurfels;

/* initialize() arguments that the hot path uses
 * (core/include/supersurfel_fusion/supersurfel_fusion.hpp:46-74). */
typedef struct OrcConfig {
  OrcCam cam;
  int cell_size;
  float lambda_pos, lambda_bound, lambda_size, lambda_disp, thresh_disp;
  int seg_iter, seg_use_ransac, nb_samples;
  int filter_iter;
  float filter_alpha, filter_beta, filter_threshold;
  float range_min, range_max;
  int delta_t;
  float conf_thresh;
  int nb_supersurfels_max;
  int icp_iter;
  double icp_cov_thresh;
} OrcConfig;

void orc_config_default(OrcConfig* cfg);

/* ---- ICP (dense_registration_kernels.cuh:175-291, dense_registration.cu:245-424) */
/* One system build at transform (R,t): out29 = JtJ[21] Jtr[6] r inliers. */
void orc_icp_system(int n_src, const float* src_pos, const float* src_col, const float* src_orient,
                    const float* tgt_col, const float* tgt_orient, const float* tgt_conf,
                    const float* R9, const float* t3, const OrcCam* cam,
                    const int32_t* labels, const float* depth, float* out29);

typedef struct OrcIcpStats {
  int valid, iters;
  float inliers;        /* of the last built system */
  double error;         /* sqrt(r/inliers) of the last built system */
  float last_system[29];
} OrcIcpStats;

/* Full Gauss-Newton loop.  R_init/t_init is the view transform (inverse prior pose).
 * On success writes R_rel/t_rel (else identity/zero) and returns 1. */
int orc_icp(int n_src, const float* src_pos, const float* src_col, const float* src_orient,
            const float* tgt_col, const float* tgt_orient, const float* tgt_conf,
            const float* R_init9, const float* t_init3, const OrcCam* cam,
            const int32_t* labels, const float* depth, int nb_iter, double cov_thresh,
            float* R_rel9, float* t_rel3, OrcIcpStats* stats);

/* DenseRegistration::align (dense_registration.cu:52-243): keyframe supersurfels (source, with
 * confidences) against the current frame; returns 1 and (R, t) when valid. */
int orc_align(int n_src, const float* src_pos, const float* src_col, const float* src_orient,
              const float* src_conf, const float* tgt_col, const float* tgt_orient, const float* tgt_conf,
              const float* R_init9, const float* t_init3, const OrcCam* cam, const int32_t* labels,
              const float* depth, int nb_iter, double cov_thresh, float* R9, float* t3, OrcIcpStats* stats);

/* pose <- pose o (R_rel,t_rel) with quaternion renormalisation (supersurfel_fusion.cu:313-328) */
void orc_compose_pose(float* R9, float* t3, const float* R_rel9, const float* t_rel3);

/* ---- supersurfel extraction (supersurfel_fusion_kernels.cu:113-224) */
void orc_generate_supersurfels(const OrcCam* cam, int n_superpixels, const uint8_t* rgba,
                               const float* slanted_depth, const int32_t* labels,
                               const uint8_t* inliers, const int32_t* bound,
                               float z_min, float z_max, int stamp, OrcSurfels* frame);

/* ---- fusion (supersurfel_fusion.cu:351-483, supersurfel_fusion_kernels.cu:348-467,522-682) */
typedef struct OrcFuseCounts {
  int nb_supersurfels, nb_visible, nb_removed, nb_matched, nb_inserted;
  /* why filterModel removed them (supersurfel_fusion_kernels.cu:429-449), a split of nb_removed for the tests:
   * stale = (age > delta_t && conf < conf_thresh && stamp > delta_t), invalid = conf <= 0 only,
   * occluded = in view and in front of the slanted depth (p.z < 0.8 z) */
  int nb_removed_stale, nb_removed_invalid, nb_removed_occluded;
} OrcFuseCounts;
/* model arrays have capacity nb_max; counts in/out. */
void orc_fuse(const OrcCam* cam, int n_superpixels, const OrcSurfels* frame, OrcSurfels* model,
              int nb_max, const float* R9, const float* t3, const int32_t* labels,
              const float* slanted_depth, float z_min, float z_max, int stamp, int delta_t,
              float conf_thresh, OrcFuseCounts* counts /* in: nb_supersurfels, nb_visible */);

/* ---- TPS segmentation (TPS_RGBD.cu:101-525 and kernels) */
typedef struct OrcTps OrcTps; /* persistent state: RNG streams survive across frames */
OrcTps* orc_tps_create(const OrcConfig* cfg);
void orc_tps_destroy(OrcTps*);
/* rgb: H*W*3 (R,G,B); depth: H*W metres (0 = missing). */
void orc_tps_compute(OrcTps*, const uint8_t* rgb, const float* depth);
/* outputs (any pointer may be NULL) */
void orc_tps_get(const OrcTps*, int32_t* labels, int32_t* bound, uint8_t* inliers, float* disp,
                 float* superpixels /* S x 12 floats: xy_rg, theta_b, size */,
                 float* slanted_depth, uint8_t* rgba);
int orc_tps_nb_superpixels(const OrcTps*);
/* stage hooks for the parity tests */
void orc_tps_get_samples(const OrcTps*, float* samples /* S*nb_samples*4 */);

/* ---- whole engine (supersurfel_fusion.cu:166-530 minus VO/MOD/ferns) */
typedef struct OrcEngine OrcEngine;
typedef struct OrcFrameStats {
  int stamp, nb_supersurfels, nb_visible, nb_removed;
  int icp_ran, icp_valid, icp_iters;
  float icp_inliers;
  double icp_error;
  int nb_matched, nb_inserted, nb_removed_stale, nb_removed_invalid, nb_removed_occluded;
} OrcFrameStats;
OrcEngine* orc_engine_create(const OrcConfig* cfg);
void orc_engine_destroy(OrcEngine*);
/* prior_Rt12: optional pose prior (R row-major 9 + t 3), NULL = previous pose. */
void orc_engine_process_frame(OrcEngine*, const uint8_t* rgb, const float* depth, const float* prior_Rt12,
                              OrcFrameStats* stats);
/* Same with the MOD hook: mask[S] != 0 marks a frame supersurfel dynamic -> confidence = -1 right after
 * generateSupersurfels and before the registration, where the reference's detectMotion writes
 * frame.confidences (supersurfel_fusion.cu:194-213, motion_detection.cu:573).  mask may be NULL. */
void orc_engine_process_frame_masked(OrcEngine*, const uint8_t* rgb, const float* depth, const float* prior_Rt12,
                                     const uint8_t* mask, OrcFrameStats* stats);
void orc_engine_get_pose(const OrcEngine*, float* R9, float* t3);
void orc_engine_set_pose(OrcEngine*, const float* R9, const float* t3);
/* applyTransformSuperSurfel over the whole model (loop-closure hook) */
void orc_engine_transform_model(OrcEngine*, const float* R9, const float* t3);
/* extractLocalPointCloud with the engine's pose, conf_thresh and the given radius; returns the count */
int orc_engine_local_cloud(const OrcEngine*, float radius, float* out_pos, float* out_nrm);
void orc_engine_get_model(const OrcEngine*, OrcSurfels* out /* caller buffers, nb_supersurfels rows */);
void orc_engine_get_frame(const OrcEngine*, OrcSurfels* out /* caller buffers, S rows */);
OrcTps* orc_engine_tps(OrcEngine*);
void orc_set_num_threads(int n);

/* ---- consumers of the model (oracle_consumers.cpp) */
void orc_apply_deformation(float* positions, float* orientations, float* shapes, const float* node_pos,
                           const float* node_rot, const float* node_trans, const float* weights,
                           const int32_t* nn, int model_size);
void orc_markers(const float* positions, const float* colors, const float* orientations, const float* dims,
                 const float* confidences, int n, float conf_thresh, float* points, float* out_colors);
int orc_format_tum_pose(const float* R9, const float* t3, const char* timestamp, char* line, int line_size);
/* extractLocalPointCloudKernel (supersurfel_fusion_kernels.cu:490-520) + the view transform of its caller
 * (supersurfel_fusion.cu:896-897): stable supersurfels (conf >= conf_thresh) within `radius` of the camera,
 * in camera coordinates, ascending model order (the reference's order is by atomic ticket).  Returns the count. */
int orc_extract_local_point_cloud(int n, const float* positions, const float* orientations, const float* confidences,
                                  float conf_thresh, const float* R_pose9, const float* t_pose3, float radius,
                                  float* out_pos, float* out_nrm);
/* applyTransformSuperSurfel (supersurfel_fusion_kernels.cu:467-488): rigid motion of every supersurfel with
 * confidence > 0 (position, orientation rows, shape). */
void orc_transform_model(int n, float* positions, float* orientations, float* shapes, const float* confidences,
                         const float* R9, const float* t3);

/* ---- ingest in front of the path (supersurfel_fusion.cu:171-181; oracle_ingest.cpp) */
void orc_bilateral_filter(const float* depth, int width, int height, int kernel_size, float sigma_color,
                          float sigma_spatial, float* out);
void orc_rgb_to_gray(const uint8_t* rgb, int n_pixels, uint8_t* gray);
void orc_depth16_to_metres(const uint16_t* depth16, int n_pixels, float scale, float* out);

#ifdef __cplusplus
}
#endif
#endif
