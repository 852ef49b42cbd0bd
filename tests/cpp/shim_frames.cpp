// Test program for include/supersurfel_fusion.hpp: drives the reference's class surface
// (initialize / processFrame / getPose / getnbSupersurfels / getStamp / getModel / getFrame /
// extractLocalPointCloud / exportModel, core/include/supersurfel_fusion/supersurfel_fusion.hpp:46-101)
// the way node/supersurfel_fusion_rgbd_benchmark_node.cpp:573-744 does, on raw frames written by
// tests/test_gpu_cpp_shim.py, and prints what it got as "key value..." lines for the test to compare
// with the ctypes path.
//   shim_frames <width> <height> <fx> <fy> <cx> <cy> <n_frames> <prefix> <export_path>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "supersurfel_fusion.hpp"

using supersurfel_fusion::CamParam;
using supersurfel_fusion::ImageView;
using supersurfel_fusion::SupersurfelFusion;
using supersurfel_fusion::SupersurfelsHost;
using supersurfel_fusion::Transform3;

static bool read_file(const std::string& path, void* dst, size_t bytes) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  const size_t got = fread(dst, 1, bytes, f);
  fclose(f);
  return got == bytes;
}

int main(int argc, char** argv) {
  if (argc < 10) return 2;
  const int W = atoi(argv[1]), H = atoi(argv[2]);
  const CamParam cam{(float)atof(argv[3]), (float)atof(argv[4]), (float)atof(argv[5]), (float)atof(argv[6]), H, W};
  const int n_frames = atoi(argv[7]);
  const std::string prefix = argv[8], export_path = argv[9];
  try {
    SupersurfelFusion fusion;
    if (fusion.isInitialized()) return 3;
    // launch/supersurfel_fusion_rgbd_benchmark.launch values, positional like the node's call (conf_thresh lowered
    // so that three small frames already hold stable supersurfels for exportModel / extractLocalPointCloud)
    fusion.initialize(cam, 16, 10.0f, 1000.0f, 1000.0f, 1e8f, 1e-4f, 10, true, 16, 3, 0.1f, 1.0f, 0.05f, 0.2f, 5.0f, 20,
                      400.0f, 20000, 10, 0.05);
    if (!fusion.isInitialized()) return 4;
    std::vector<uint8_t> rgb((size_t)W * H * 3);
    std::vector<float> depth((size_t)W * H);
    for (int k = 0; k < n_frames; k++) {
      if (!read_file(prefix + "_rgb_" + std::to_string(k) + ".bin", rgb.data(), rgb.size())) return 5;
      if (!read_file(prefix + "_depth_" + std::to_string(k) + ".bin", depth.data(), depth.size() * 4)) return 5;
      fusion.processFrame(ImageView{rgb.data(), (size_t)W * 3}, ImageView{depth.data(), (size_t)W * 4}, nullptr,
                          /*filter_depth=*/false);
      const Transform3 tf = fusion.getPose();
      printf("pose %d", k);
      for (int i = 0; i < 9; i++) printf(" %.9g", tf.R.rows[i / 3][i % 3]);
      for (int i = 0; i < 3; i++) printf(" %.9g", tf.t[i]);
      printf("\n");
      const SsfFrameStats st = fusion.getFrameStats();
      printf("stats %d %d %d %d %d %d %d\n", k, fusion.getStamp(), fusion.getnbSupersurfels(), st.nb_visible,
             st.nb_removed, st.icp_valid, st.icp_iters);
    }
    SupersurfelsHost model, frame;
    fusion.getModel(model);
    fusion.getFrame(frame);
    double sum_pos = 0.0, sum_conf = 0.0;
    for (float v : model.positions) sum_pos += v;
    for (float v : model.confidences) sum_conf += v;
    printf("model %zu %.9g %.9g\n", model.size(), sum_pos, sum_conf);
    printf("frame %zu\n", frame.size());
    std::vector<float> cloud_p, cloud_n;
    fusion.extractLocalPointCloud(cloud_p, cloud_n);
    printf("cloud %zu\n", cloud_p.size() / 3);
    std::vector<float> mp, mc;
    fusion.getMarkers(false, mp, mc);
    printf("markers %zu\n", mp.size() / 18);
    printf("tum %s", fusion.tumPoseLine("1305031102.175304").c_str());
    fusion.exportModel(export_path);
  } catch (const std::exception& ex) {
    fprintf(stderr, "shim_frames: %s\n", ex.what());
    return 1;
  }
  return 0;
}
