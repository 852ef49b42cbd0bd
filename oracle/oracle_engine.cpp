// TEST INFRASTRUCTURE ONLY -- CPU oracle, per-frame sequencing.
// Restates SupersurfelFusion::initialize / processFrame
// (core/src/supersurfel_fusion.cu:49-164, 166-530) without its out-of-scope
// neighbours: the sparse-VO pose prior is an input (default: previous pose, which is
// what the reference falls back to when VO fails, sparse_vo.cpp:131-176), MOD, ferns
// and loop closure are absent, and the depth image is taken as already
// bilateral-filtered (cv::cuda::bilateralFilter is un-vendored
// third-party code that runs before the path).
#include "oracle.h"
#include "oracle_math.h"
#include <vector>

using namespace orc;

struct OrcEngine {
  OrcConfig cfg;
  OrcTps* tps;
  int S;
  // frame / model storage
  std::vector<float> f_pos, f_col, f_ori, f_shp, f_dim, f_cnf;
  std::vector<int> f_stp;
  std::vector<float> m_pos, m_col, m_ori, m_shp, m_dim, m_cnf;
  std::vector<int> m_stp;
  std::vector<int32_t> labels, bound;
  std::vector<uint8_t> inliers, rgba;
  std::vector<float> slanted;
  int nbSupersurfels, nbVisible, nbRemoved, stamp;
  float R[9], t[3];

  OrcSurfels frame() { return OrcSurfels{f_pos.data(), f_col.data(), f_stp.data(), f_ori.data(), f_shp.data(), f_dim.data(), f_cnf.data()}; }
  OrcSurfels model() { return OrcSurfels{m_pos.data(), m_col.data(), m_stp.data(), m_ori.data(), m_shp.data(), m_dim.data(), m_cnf.data()}; }
};

extern "C" void orc_config_default(OrcConfig* c) {
  // supersurfel_fusion.hpp:46-74 defaults; camera = rgbd_benchmark/fr1_cam.yaml
  c->cam = OrcCam{525.0f, 525.0f, 319.5f, 239.5f, 480, 640};
  c->cell_size = 16;
  c->lambda_pos = 50.0f; c->lambda_bound = 1000.0f; c->lambda_size = 10000.0f; c->lambda_disp = 1000000.0f;
  c->thresh_disp = 0.0001f;
  c->seg_iter = 10; c->seg_use_ransac = 1; c->nb_samples = 16;
  c->filter_iter = 4; c->filter_alpha = 0.1f; c->filter_beta = 1.0f; c->filter_threshold = 0.05f;
  c->range_min = 0.2f; c->range_max = 5.0f;
  c->delta_t = 20; c->conf_thresh = 2500.0f; c->nb_supersurfels_max = 50000;
  c->icp_iter = 10; c->icp_cov_thresh = 0.04;
}

extern "C" OrcEngine* orc_engine_create(const OrcConfig* cfg) {
  OrcEngine* e = new OrcEngine;
  e->cfg = *cfg;
  e->tps = orc_tps_create(cfg);
  e->S = orc_tps_nb_superpixels(e->tps);
  const size_t S = e->S, M = cfg->nb_supersurfels_max, N = (size_t)cfg->cam.width * cfg->cam.height;
  e->f_pos.assign(3 * S, 0); e->f_col.assign(3 * S, 0); e->f_stp.assign(2 * S, 0); e->f_ori.assign(9 * S, 0);
  e->f_shp.assign(6 * S, 0); e->f_dim.assign(2 * S, 0); e->f_cnf.assign(S, 0);
  // model.memset (supersurfel_fusion.cu:119-126)
  e->m_pos.assign(3 * M, 0); e->m_col.assign(3 * M, 0); e->m_stp.assign(2 * M, 0); e->m_ori.assign(9 * M, 0);
  e->m_shp.assign(6 * M, 0); e->m_dim.assign(2 * M, 0); e->m_cnf.assign(M, 0);
  e->labels.assign(N, 0); e->bound.assign(N, 0); e->inliers.assign(N, 0); e->rgba.assign(4 * N, 0);
  e->slanted.assign(N, 0);
  e->nbSupersurfels = e->nbVisible = e->nbRemoved = 0;
  e->stamp = 0;
  const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; i++) e->R[i] = I[i];
  e->t[0] = e->t[1] = e->t[2] = 0.f;
  return e;
}
extern "C" void orc_engine_destroy(OrcEngine* e) {
  if (!e) return;
  orc_tps_destroy(e->tps);
  delete e;
}

extern "C" void orc_engine_process_frame(OrcEngine* e, const uint8_t* rgb, const float* depth,
                                         const float* prior, OrcFrameStats* stats) {
  orc_engine_process_frame_masked(e, rgb, depth, prior, nullptr, stats);
}

extern "C" void orc_engine_process_frame_masked(OrcEngine* e, const uint8_t* rgb, const float* depth,
                                                const float* prior, const uint8_t* mask, OrcFrameStats* stats) {
  const OrcConfig& c = e->cfg;
  // tps->compute / filter / computeDepthImage (supersurfel_fusion.cu:189-191)
  orc_tps_compute(e->tps, rgb, depth);
  orc_tps_get(e->tps, e->labels.data(), e->bound.data(), e->inliers.data(), nullptr, nullptr,
              e->slanted.data(), e->rgba.data());
  // generateSupersurfels (:194, :551-593)
  OrcSurfels frame = e->frame();
  orc_generate_supersurfels(&c.cam, e->S, e->rgba.data(), e->slanted.data(), e->labels.data(),
                            e->inliers.data(), e->bound.data(), c.range_min, c.range_max, e->stamp, &frame);
  // MOD hook: detectMotion marks dynamic frame supersurfels invalid (:198-213, motion_detection.cu:573)
  if (mask)
    for (int f = 0; f < e->S; f++)
      if (mask[f]) e->f_cnf[f] = -1.0f;
  // pose prior (:225-228)
  if (prior) {
    for (int i = 0; i < 9; i++) e->R[i] = prior[i];
    for (int i = 0; i < 3; i++) e->t[i] = prior[9 + i];
  }
  OrcIcpStats is{};
  int icp_ran = 0;
  // frame-to-model registration (:232-328)
  if (e->nbVisible > 0) {
    icp_ran = 1;
    Mat33 R = mkmat(mk3(e->R[0], e->R[1], e->R[2]), mk3(e->R[3], e->R[4], e->R[5]), mk3(e->R[6], e->R[7], e->R[8]));
    Mat33 Rv = transpose(R);
    f3 tv = -(Rv * mk3(e->t[0], e->t[1], e->t[2]));
    float Rv9[9] = {Rv.rows[0].x, Rv.rows[0].y, Rv.rows[0].z, Rv.rows[1].x, Rv.rows[1].y, Rv.rows[1].z,
                    Rv.rows[2].x, Rv.rows[2].y, Rv.rows[2].z};
    float tv3[3] = {tv.x, tv.y, tv.z};
    float Rrel[9], trel[3];
    int ok = orc_icp(e->nbVisible, e->m_pos.data(), e->m_col.data(), e->m_ori.data(), e->f_col.data(),
                     e->f_ori.data(), e->f_cnf.data(), Rv9, tv3, &c.cam, e->labels.data(), e->slanted.data(),
                     c.icp_iter, c.icp_cov_thresh, Rrel, trel, &is);
    if (ok) orc_compose_pose(e->R, e->t, Rrel, trel);
  }
  // model update (:351-483)
  OrcSurfels model = e->model();
  OrcFuseCounts fc{};
  fc.nb_supersurfels = e->nbSupersurfels;
  fc.nb_visible = e->nbVisible;
  orc_fuse(&c.cam, e->S, &frame, &model, c.nb_supersurfels_max, e->R, e->t, e->labels.data(),
           e->slanted.data(), c.range_min, c.range_max, e->stamp, c.delta_t, c.conf_thresh, &fc);
  e->nbSupersurfels = fc.nb_supersurfels;
  e->nbVisible = fc.nb_visible;
  e->nbRemoved = fc.nb_removed;
  if (stats) {
    stats->stamp = e->stamp;
    stats->nb_supersurfels = e->nbSupersurfels;
    stats->nb_visible = e->nbVisible;
    stats->nb_removed = e->nbRemoved;
    stats->icp_ran = icp_ran;
    stats->icp_valid = is.valid;
    stats->icp_iters = is.iters;
    stats->icp_inliers = is.inliers;
    stats->icp_error = is.error;
    stats->nb_matched = fc.nb_matched;
    stats->nb_inserted = fc.nb_inserted;
    stats->nb_removed_stale = fc.nb_removed_stale;
    stats->nb_removed_invalid = fc.nb_removed_invalid;
    stats->nb_removed_occluded = fc.nb_removed_occluded;
  }
  e->stamp++;  // :521
}

extern "C" void orc_engine_get_pose(const OrcEngine* e, float* R9, float* t3) {
  for (int i = 0; i < 9; i++) R9[i] = e->R[i];
  for (int i = 0; i < 3; i++) t3[i] = e->t[i];
}
extern "C" void orc_engine_set_pose(OrcEngine* e, const float* R9, const float* t3) {
  for (int i = 0; i < 9; i++) e->R[i] = R9[i];
  for (int i = 0; i < 3; i++) e->t[i] = t3[i];
}
extern "C" void orc_engine_transform_model(OrcEngine* e, const float* R9, const float* t3) {
  orc_transform_model(e->nbSupersurfels, e->m_pos.data(), e->m_ori.data(), e->m_shp.data(), e->m_cnf.data(), R9, t3);
}
extern "C" int orc_engine_local_cloud(const OrcEngine* e, float radius, float* out_pos, float* out_nrm) {
  return orc_extract_local_point_cloud(e->nbSupersurfels, e->m_pos.data(), e->m_ori.data(), e->m_cnf.data(),
                                       e->cfg.conf_thresh, e->R, e->t, radius, out_pos, out_nrm);
}
static void copy_out(const OrcSurfels& s, OrcSurfels* o, size_t n) {
  if (o->positions) std::memcpy(o->positions, s.positions, n * 12);
  if (o->colors) std::memcpy(o->colors, s.colors, n * 12);
  if (o->stamps) std::memcpy(o->stamps, s.stamps, n * 8);
  if (o->orientations) std::memcpy(o->orientations, s.orientations, n * 36);
  if (o->shapes) std::memcpy(o->shapes, s.shapes, n * 24);
  if (o->dims) std::memcpy(o->dims, s.dims, n * 8);
  if (o->confidences) std::memcpy(o->confidences, s.confidences, n * 4);
}
extern "C" void orc_engine_get_model(const OrcEngine* e, OrcSurfels* out) {
  copy_out(const_cast<OrcEngine*>(e)->model(), out, e->nbSupersurfels);
}
extern "C" void orc_engine_get_frame(const OrcEngine* e, OrcSurfels* out) {
  copy_out(const_cast<OrcEngine*>(e)->frame(), out, e->S);
}
extern "C" OrcTps* orc_engine_tps(OrcEngine* e) { return e->tps; }
