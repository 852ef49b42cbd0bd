// Header-only C++ shim: the reference's class name and method names over the libssf C-ABI.
//
// A maintainer of the reference replaces
//   #include <supersurfel_fusion/supersurfel_fusion.hpp>   (reference: core/include/...:40-143)
// by this header and links libssf.so instead of libsfusion; node code such as
// node/supersurfel_fusion_node.cpp:74-196 (initialize / processFrame / getPose / getModel /
// getnbSupersurfels / getStamp / exportModel / computeSuperpixelSegIm) keeps compiling with
// cv::Mat arguments replaced by {data, step} pairs (ImageView below is layout-compatible
// with the fields of cv::Mat that the reference reads).  No OpenCV / thrust / Eigen types.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "ssf.h"

namespace supersurfel_fusion {

struct CamParam {  // cam_param.hpp:27-31
  float fx, fy, cx, cy;
  int height, width;
};

struct Mat33 { float rows[3][3]; };                 // matrix_types.h:33-36
struct Transform3 { Mat33 R; float t[3]; };         // matrix_types.h:38-42

struct ImageView {       // what processFrame reads of a cv::Mat: data pointer and row stride
  const void* data;
  size_t step;           // bytes per row
};

// Host copy of a supersurfel set, member layout of supersurfels.hpp:32-41
struct SupersurfelsHost {
  std::vector<float> positions, colors, orientations, shapes, dims, confidences;
  std::vector<int32_t> stamps;
  void resize(size_t n) {
    positions.resize(3 * n); colors.resize(3 * n); stamps.resize(2 * n); orientations.resize(9 * n);
    shapes.resize(6 * n); dims.resize(2 * n); confidences.resize(n);
  }
  SsfSurfels view() {
    return SsfSurfels{positions.data(), colors.data(), stamps.data(), orientations.data(), shapes.data(),
                      dims.data(), confidences.data()};
  }
  size_t size() const { return confidences.size(); }
};

class SupersurfelFusion {
 public:
  explicit SupersurfelFusion(int device = 0) : h_(nullptr), device_(device) {}
  ~SupersurfelFusion() { if (h_) ssf_destroy(h_); }
  SupersurfelFusion(const SupersurfelFusion&) = delete;
  SupersurfelFusion& operator=(const SupersurfelFusion&) = delete;

  // supersurfel_fusion.hpp:46-74 -- same order, names and defaults.  enable_loop_closure / enable_mod
  // are recorded only: the loop detector and MOD are the caller's, see the note in ssf.h (SsfConfig).
  void initialize(const CamParam& cam_param, int cell_size = 16, float lambda_pos = 50.0f,
                  float lambda_bound = 1000.0f, float lambda_size = 10000.0f, float lambda_disp = 1000000.0f,
                  float thresh_disp = 0.0001f, int seg_iter = 10, bool seg_use_ransac = true, int nb_samples = 16,
                  int filter_iter = 4, float filter_alpha = 0.1f, float filter_beta = 1.0f,
                  float filter_threshold = 0.05f, float range_min = 0.2f, float range_max = 5.0f, int delta_t = 20,
                  float conf_thresh = 2500.0f, int nb_supersurfels_max = 50000, int icp_iter = 10,
                  double icp_cov_thresh = 0.04, int /*nb_features*/ = 2000, float /*features_scale_factor*/ = 1.2f,
                  int /*features_nb_levels*/ = 8, int /*ini_th_fast*/ = 20, int /*min_th_fast*/ = 7,
                  int /*untracked_threshold*/ = 10, bool enable_loop_closure = true, bool enable_mod = true) {
    SsfConfig c;
    ssf_config_default(&c);
    c.cam = SsfCamParam{cam_param.fx, cam_param.fy, cam_param.cx, cam_param.cy, cam_param.height, cam_param.width};
    c.cell_size = cell_size; c.lambda_pos = lambda_pos; c.lambda_bound = lambda_bound; c.lambda_size = lambda_size;
    c.lambda_disp = lambda_disp; c.thresh_disp = thresh_disp; c.seg_iter = seg_iter;
    c.seg_use_ransac = seg_use_ransac ? 1 : 0; c.nb_samples = nb_samples; c.filter_iter = filter_iter;
    c.filter_alpha = filter_alpha; c.filter_beta = filter_beta; c.filter_threshold = filter_threshold;
    c.range_min = range_min; c.range_max = range_max; c.delta_t = delta_t; c.conf_thresh = conf_thresh;
    c.nb_supersurfels_max = nb_supersurfels_max; c.icp_iter = icp_iter; c.icp_cov_thresh = icp_cov_thresh;
    c.enable_loop_closure = enable_loop_closure ? 1 : 0; c.enable_mod = enable_mod ? 1 : 0;
    if (h_) { ssf_destroy(h_); h_ = nullptr; }
    check(ssf_create(&c, device_, &h_), "ssf_create");
    cfg_ = c;
  }

  // processFrame(const cv::Mat& rgb_h /*8UC3 RGB*/, const cv::Mat& depth_h /*32FC1 m*/)
  // filter_depth = true runs the reference's cv::cuda::bilateralFilter(depth, depth, -1, 0.03, 4.5)
  // (supersurfel_fusion.cu:180) inside the library; false expects an already filtered image.
  void processFrame(const ImageView& rgb_h, const ImageView& depth_h, const Transform3* pose_prior = nullptr,
                    bool filter_depth = true) {
    float prior[12];
    pack(pose_prior, prior);
    check(ssf_process_frame(h_, static_cast<const uint8_t*>(rgb_h.data), rgb_h.step,
                            static_cast<const float*>(depth_h.data), depth_h.step, pose_prior ? prior : nullptr,
                            filter_depth ? SSF_FLAG_BILATERAL : 0u),
          "ssf_process_frame");
  }
  // the 16-bit depth image the nodes receive, decoded on the device
  // (depth.convertTo(CV_32FC1, depth_scale), node/supersurfel_fusion_rgbd_benchmark_node.cpp:609-610)
  void processFrame(const ImageView& rgb_h, const ImageView& depth16_h, float depth_scale,
                    const Transform3* pose_prior = nullptr, bool filter_depth = true) {
    float prior[12];
    pack(pose_prior, prior);
    check(ssf_process_frame_depth16(h_, static_cast<const uint8_t*>(rgb_h.data), rgb_h.step,
                                    static_cast<const uint16_t*>(depth16_h.data), depth16_h.step, depth_scale,
                                    pose_prior ? prior : nullptr, filter_depth ? SSF_FLAG_BILATERAL : 0u),
          "ssf_process_frame_depth16");
  }
  // Pipelined processFrame for streams: submit frame k+1 before waiting for frame k; the
  // segmentation of the newer frame overlaps the registration + fusion of the older one.
  void submitFrame(const ImageView& rgb_h, const ImageView& depth_h, const Transform3* pose_prior = nullptr,
                   bool filter_depth = true) {
    float prior[12];
    pack(pose_prior, prior);
    check(ssf_submit_frame(h_, static_cast<const uint8_t*>(rgb_h.data), rgb_h.step,
                           static_cast<const float*>(depth_h.data), depth_h.step, pose_prior ? prior : nullptr,
                           filter_depth ? SSF_FLAG_BILATERAL : 0u),
          "ssf_submit_frame");
  }
  Transform3 waitFrame(SsfFrameStats* stats = nullptr) {
    Transform3 tf;
    check(ssf_wait_frame(h_, stats, &tf.R.rows[0][0], tf.t), "ssf_wait_frame");
    return tf;
  }
  void generateSupersurfels() { check(ssf_generate_supersurfels(h_), "ssf_generate_supersurfels"); }
  void exportModel(const std::string& filename) { check(ssf_export_model(h_, filename.c_str()), "ssf_export_model"); }
  void computeSuperpixelSegIm(std::vector<uint8_t>& seg_im_bgr) {
    seg_im_bgr.resize((size_t)cfg_.cam.width * cfg_.cam.height * 3);
    check(ssf_render_preview(h_, seg_im_bgr.data()), "ssf_render_preview");
  }
  void computeSlantedPlaneIm(std::vector<float>& slanted_plane_im) {
    slanted_plane_im.resize((size_t)cfg_.cam.width * cfg_.cam.height);
    check(ssf_get_slanted_depth(h_, slanted_plane_im.data()), "ssf_get_slanted_depth");
  }
  bool isInitialized() { return h_ && ssf_is_initialized(h_); }
  void getFrame(SupersurfelsHost& out) {
    int s = 0;
    check(ssf_get_nb_superpixels(h_, &s), "ssf_get_nb_superpixels");
    out.resize(s);
    SsfSurfels v = out.view();
    check(ssf_copy_frame(h_, &v), "ssf_copy_frame");
  }
  void getModel(SupersurfelsHost& out) {
    const int n = getnbSupersurfels();
    out.resize(n);
    SsfSurfels v = out.view();
    check(ssf_copy_model(h_, &v, n), "ssf_copy_model");
  }
  int getnbSupersurfels() {
    int n = 0;
    check(ssf_get_counts(h_, &n, nullptr, nullptr), "ssf_get_counts");
    return n;
  }
  int getStamp() {
    int s = 0;
    check(ssf_get_stamp(h_, &s), "ssf_get_stamp");
    return s;
  }
  Transform3 getPose() {
    Transform3 tf;
    check(ssf_get_pose(h_, &tf.R.rows[0][0], tf.t), "ssf_get_pose");
    return tf;
  }
  void extractLocalPointCloud(std::vector<float>& positions, std::vector<float>& normals) {
    const int cap = getnbSupersurfels() > 0 ? getnbSupersurfels() : 1;
    positions.resize(3 * (size_t)cap);
    normals.resize(3 * (size_t)cap);
    int n = 0;
    check(ssf_extract_local_point_cloud(h_, cfg_.range_max, positions.data(), normals.data(), cap, &n),
          "ssf_extract_local_point_cloud");
    positions.resize(3 * (size_t)n);
    normals.resize(3 * (size_t)n);
  }
  // DenseRegistration::align (dense_registration.hpp:44-59) as closeGlobalLoop uses it
  // (supersurfel_fusion.cu:776-792): keyframe supersurfels -> current frame
  bool align(SupersurfelsHost& keyframe, const Transform3& init, Transform3& result) {
    SsfSurfels v = keyframe.view();
    int valid = 0;
    check(ssf_align(h_, &v, (int)keyframe.size(), &init.R.rows[0][0], init.t, &result.R.rows[0][0], result.t, &valid,
                    nullptr, nullptr, nullptr),
          "ssf_align");
    return valid != 0;
  }
  // DeformationGraph::applyGraphToModel (deformation_graph.cu:840-861)
  void applyDeformation(const std::vector<float>& nodes_positions, const std::vector<float>& nodes_rotations,
                        const std::vector<float>& nodes_translations, const std::vector<float>& neighbours_weights,
                        const std::vector<int32_t>& neighbours_idx) {
    check(ssf_apply_deformation(h_, nodes_positions.data(), nodes_rotations.data(), nodes_translations.data(),
                                (int)(nodes_positions.size() / 3), neighbours_weights.data(), neighbours_idx.data(),
                                (int)(neighbours_weights.size() / 4)),
          "ssf_apply_deformation");
  }
  // what publishModelMarker / publishFrameMarker build on the host (supersurfel_fusion_node.cpp:303-520)
  void getMarkers(bool frame, std::vector<float>& points, std::vector<float>& colors) {
    int n = 0;
    if (frame) check(ssf_get_nb_superpixels(h_, &n), "ssf_get_nb_superpixels");
    else n = getnbSupersurfels();
    points.resize(18 * (size_t)n);
    colors.resize(24 * (size_t)n);
    if (n == 0) return;
    check(ssf_get_markers(h_, frame ? 1 : 0, frame ? 0.0f : cfg_.conf_thresh, points.data(), colors.data(), n, nullptr),
          "ssf_get_markers");
  }
  // one line of estimated.txt (supersurfel_fusion_rgbd_benchmark_node.cpp:727-729)
  std::string tumPoseLine(const std::string& timestamp) {
    char line[256];
    check(ssf_format_tum_pose(h_, timestamp.c_str(), line, sizeof(line)), "ssf_format_tum_pose");
    return std::string(line);
  }
  SsfFrameStats getFrameStats() {
    SsfFrameStats st;
    check(ssf_get_frame_stats(h_, &st), "ssf_get_frame_stats");
    return st;
  }
  SsfHandle handle() { return h_; }

 private:
  static void pack(const Transform3* tf, float* prior) {
    if (!tf) return;
    for (int i = 0; i < 9; i++) prior[i] = tf->R.rows[i / 3][i % 3];
    for (int i = 0; i < 3; i++) prior[9 + i] = tf->t[i];
  }
  void check(int rc, const char* what) {
    if (rc != SSF_OK) throw std::runtime_error(std::string(what) + " failed: " + (h_ ? ssf_last_error(h_) : "no handle"));
  }
  SsfHandle h_;
  int device_;
  SsfConfig cfg_;
};

}  // namespace supersurfel_fusion
