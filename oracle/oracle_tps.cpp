// TEST INFRASTRUCTURE ONLY -- CPU oracle, TPS RGB-D superpixel segmentation.
// Restates the launch sequence of core/src/TPS_RGBD.cu:101-525 and the kernels of
// core/src/TPS_RGBD_kernels.cu:27-614 and
// core/include/supersurfel_fusion/TPS_RGBD_kernels.cuh:178-651.
//
// The reference's label-update passes are racy (SURVEY.md section 7 "hard parts",
// appendix B1); the oracle fixes ONE legal serialisation, which the CUDA path
// reproduces bit for bit:
//  * within a pass every decision reads the pass-start label map, boundary map and
//    superpixel means (the per-block shared-memory snapshot of
//    TPS_RGBD_kernels.cuh:272-292 made global);
//  * all neighbour +-1 boundary updates of the pass are applied, then every
//    relabelled pixel's own "= b" overrides (the order the __syncthreads at
//    :422/:631 intends);
//  * the running sums are order-free: x, y, r, g, b, n and the integer disparity
//    moments are exact integers; the three moments that involve the disparity
//    (sum x*d, sum y*d, sum d) quantise d to 2^-30 fixed point (the reference adds
//    fp32 products with atomicAdd in scheduling order, :445-466);
//  * the plane-smoothing filter is a true Jacobi iteration (the reference updates
//    in place while neighbours are being read, TPS_RGBD_kernels.cu:585,612); its
//    out-of-bounds read for the bottom-right node (appendix B4) is "no neighbour".
// Decision arithmetic is fp32, IEEE, evaluated left to right exactly as written in
// the reference, with no fused multiply-add (-ffp-contract=off).
#include "oracle.h"
#include "oracle_math.h"
#include <algorithm>
#include <vector>

using namespace orc;

struct OrcRng;
extern "C" OrcRng* orc_rng_create(int n, unsigned long long seed);
extern "C" void orc_rng_destroy(OrcRng*);
extern "C" unsigned int orc_rng_u32(OrcRng*, int id);
extern "C" float orc_rng_uniform(OrcRng*, int id);

namespace {

const double kDispFix = 1073741824.0;        // 2^30
const double kDispClamp = 137438953472.0;    // 2^37 (disparity >= 128 1/m saturates)

inline int64_t quant_disp(float d) {
  double s = (double)d * kDispFix;
  if (s != s) s = 0.0;
  s = fmin(fmax(s, -kDispClamp), kDispClamp);
  return (int64_t)llrint(s);
}

// TPS_RGBD.hpp:33-38
struct Superpixel { f4 xy_rg, theta_b, size; };

// TPS_RGBD.hpp:40-44 with integer accumulators
struct Sums {
  int64_t x, y, r, g, b, n;
  int64_t dx, dy, dxx, dyy, dxy, dn;  // exact integers
  int64_t dxd, dyd, dd;               // 2^-30 fixed point
};

// TPS_RGBD_kernels.cu:27-59
inline bool solvePlaneEquations(f4& theta, const float x1, const float y1, const float z1, const float d1,
                                const float x2, const float y2, const float z2, const float d2,
                                const float x3, const float y3, const float z3, const float d3) {
  const float epsilonValue = 1e-20;
  float denominatorA = (x1 * z2 - x2 * z1) * (y2 * z3 - y3 * z2) - (x2 * z3 - x3 * z2) * (y1 * z2 - y2 * z1);
  if (!std::isfinite(denominatorA) && denominatorA < epsilonValue) return false;
  theta.x = ((z2 * d1 - z1 * d2) * (y2 * z3 - y3 * z2) - (z3 * d2 - z2 * d3) * (y1 * z2 - y2 * z1)) / denominatorA;
  float denominatorB = y1 * z2 - y2 * z1;
  if (denominatorB > epsilonValue) {
    theta.y = (z2 * d1 - z1 * d2 - theta.x * (x1 * z2 - x2 * z1)) / denominatorB;
  } else {
    denominatorB = y2 * z3 - y3 * z2;
    theta.y = (z3 * d2 - z2 * d3 - theta.x * (x2 * z3 - x3 * z2)) / denominatorB;
  }
  if (z1 > epsilonValue) theta.z = (d1 - theta.x * x1 - theta.y * y1) / z1;
  else if (z2 > epsilonValue) theta.z = (d2 - theta.x * x2 - theta.y * y2) / z2;
  else theta.z = (d3 - theta.x * x3 - theta.y * y3) / z3;
  return true;
}

inline float int_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

}  // namespace

struct OrcTps {
  OrcConfig cfg;
  int W, H, cell, gx, gy, S, nbSamples;
  std::vector<int32_t> labels, bound;
  std::vector<uint8_t> inliers, rgba;
  std::vector<float> disp, slanted;
  std::vector<Superpixel> sp;
  std::vector<Sums> sums;
  std::vector<f4> samples;
  OrcRng* rng;

  int label_at(int x, int y) const { return (x >= 0 && x < W && y >= 0 && y < H) ? labels[y * W + x] : -1; }
  int tex_label(float x, float y) const { return labels[tex_coord(y, H) * W + tex_coord(x, W)]; }
  float tex_disp(float x, float y) const { return disp[tex_coord(y, H) * W + tex_coord(x, W)]; }

  // TPS_RGBD_kernels.cu:224-242
  void mergeRGB() {
    for (int k = 0; k < S; k++) {
      const Sums& c = sums[k];
      Superpixel& s = sp[k];
      float n = (float)c.n;
      s.xy_rg.x = (float)c.x / n;
      s.xy_rg.y = (float)c.y / n;
      s.xy_rg.z = (float)c.r / n;
      s.xy_rg.w = (float)c.g / n;
      s.theta_b.w = (float)c.b / n;
      s.size.x = n;
    }
  }
  // TPS_RGBD_kernels.cu:244-276
  void mergeRGBD() {
    mergeRGB();
    for (int k = 0; k < S; k++) {
      const Sums& c = sums[k];
      const float dx = (float)c.dx, dy = (float)c.dy, dxx = (float)c.dxx, dyy = (float)c.dyy,
                  dxy = (float)c.dxy, dn = (float)c.dn;
      const float dxd = (float)((double)c.dxd * (1.0 / kDispFix));
      const float dyd = (float)((double)c.dyd * (1.0 / kDispFix));
      const float dd = (float)((double)c.dd * (1.0 / kDispFix));
      f4 theta = {0.f, 0.f, 0.f, 0.f};
      if (!solvePlaneEquations(theta, dxx, dxy, dx, dxd, dxy, dyy, dy, dyd, dx, dy, dn, dd)) {
        theta.x = 0.f; theta.y = 0.f; theta.z = int_as_float(0xFFE00000u);
      }
      sp[k].theta_b.x = theta.x; sp[k].theta_b.y = theta.y; sp[k].theta_b.z = theta.z;
    }
  }

  // TPS_RGBD_kernels.cuh:178-233
  bool isUnchangeable(int x, int y) const {
    const int index = labels[y * W + x];
    int jump = 0;
    bool prev = (label_at(x - 1, y - 1) == index);
    const int ox[7] = {0, 1, 1, 1, 0, -1, -1};
    const int oy[7] = {-1, -1, 0, 1, 1, 1, 0};
    for (int k = 0; k < 7; k++) {
      bool cur = (label_at(x + ox[k], y + oy[k]) == index);
      if (prev != cur) { jump++; prev = cur; }
    }
    return jump > 2;
  }

  struct Change { int x, y, old_index, new_index, b; };

  // One label-update pass: updateTPSRGB_kernel (TPS_RGBD_kernels.cuh:476-651) when
  // !use_disp, updateTPSRGBD_kernel (:235-474) when use_disp.
  void pass(int OX, int OY, bool use_disp) {
    const float lambda_pos = cfg.lambda_pos, lambda_bound = cfg.lambda_bound, lambda_size = cfg.lambda_size,
                lambda_disp = cfg.lambda_disp, thresh_disp = cfg.thresh_disp;
    const int min_size = (int)((float)(cell * cell) / 4.f);  // TPS_RGBD.cu:198 (float -> int parameter)
    std::vector<Change> changes;
    struct InlierEdit { int p; uint8_t v; };
    std::vector<InlierEdit> inlier_edits;
    const int nx[4] = {0, -1, 1, 0};
    const int ny[4] = {-1, 0, 0, 1};

    // launch geometry of TPS_RGBD.cu:185-186: 16x16 blocks over (W/2, H/2) threads, a
    // block returns early when its first pixel is outside the image (:254-255)
    const int raw_w = 16 * ((W / 2 + 15) / 16), raw_h = 16 * ((H / 2 + 15) / 16);
    for (int raw_y = 0; raw_y < raw_h; raw_y++) {
      const int y = 2 * raw_y + OY;
      if (32 * (raw_y / 16) + OY >= H || y >= H) continue;
      for (int raw_x = 0; raw_x < raw_w; raw_x++) {
        const int x = 2 * raw_x + (raw_x + OX) % 2;
        if (32 * (raw_x / 16) >= W || x >= W) continue;
        const int p = y * W + x;
        const int bounds = bound[p];
        const int index = labels[p];
        int new_index = index;
        const Superpixel prev_sp = sp[index];

        float disp_v = 0.f;
        uint8_t prev_inlier = 0, inlier = 0xff;
        float disp_energy = 0.f;
        if (use_disp) {
          disp_v = disp[p];
          prev_inlier = inliers[p];
          float dp = prev_sp.theta_b.x * (float)x + prev_sp.theta_b.y * (float)y + prev_sp.theta_b.z;
          disp_energy = (dp - disp_v) * (dp - disp_v);
          if (!std::isfinite(disp_energy) || disp_energy > thresh_disp || dp < 0.f) {
            disp_energy = thresh_disp;
            inlier = 0;
          }
        }

        if (bounds && !isUnchangeable(x, y)) {
          const float cr = (float)rgba[4 * p], cg = (float)rgba[4 * p + 1], cb = (float)rgba[4 * p + 2];
          const float px = (float)x, py = (float)y;
          const float size = prev_sp.size.x;
          const float s = size / (size - 1.f);
          const float dpx = s * (px - prev_sp.xy_rg.x), dpy = s * (py - prev_sp.xy_rg.y);
          const float dcx = s * (cr - prev_sp.xy_rg.z), dcy = s * (cg - prev_sp.xy_rg.w),
                      dcz = s * (cb - prev_sp.theta_b.w);
          const float dsize = size - (float)min_size;
          float best_energy = (dcx * dcx + dcy * dcy + dcz * dcz) + lambda_pos * (dpx * dpx + dpy * dpy);
          if (use_disp) best_energy = best_energy + lambda_disp * disp_energy;
          best_energy = best_energy - lambda_size * fminf(dsize, 0.f);
          best_energy = best_energy + lambda_bound * (float)bounds;

          int nl[4];
          for (int k = 0; k < 4; k++) nl[k] = label_at(x + nx[k], y + ny[k]);
          for (int k = 0; k < 4; k++) {
            const int i_n = nl[k];
            if (i_n == -1 || i_n == index) continue;
            const Superpixel n_sp = sp[i_n];
            const float ex = px - n_sp.xy_rg.x, ey = py - n_sp.xy_rg.y;
            const float fx = cr - n_sp.xy_rg.z, fy = cg - n_sp.xy_rg.w, fz = cb - n_sp.theta_b.w;
            const float nsize = n_sp.size.x + 1.f - (float)min_size;
            float n_disp_energy = 0.f;
            uint8_t n_inlier = 0xff;
            if (use_disp) {
              float dp = n_sp.theta_b.x * (float)x + n_sp.theta_b.y * (float)y + n_sp.theta_b.z;
              n_disp_energy = (dp - disp_v) * (dp - disp_v);
              if (!std::isfinite(n_disp_energy) || n_disp_energy > thresh_disp || dp < 0.f) {
                n_disp_energy = thresh_disp;
                n_inlier = 0;
              }
            }
            int b = 0;
            for (int q = 0; q < 4; q++)
              if (nl[q] != i_n) b++;
            float energy = (fx * fx + fy * fy + fz * fz) + lambda_pos * (ex * ex + ey * ey);
            if (use_disp) energy = energy + lambda_disp * n_disp_energy;
            energy = energy - lambda_size * fminf(nsize, 0.f);
            energy = energy + lambda_bound * (float)b;
            if (energy < best_energy) {
              best_energy = energy;
              new_index = i_n;
              if (use_disp) inlier = n_inlier;
            }
          }
          if (new_index != index) {
            int b = 0;
            for (int k = 0; k < 4; k++)
              if (nl[k] != new_index) b++;
            changes.push_back(Change{x, y, index, new_index, b});
          }
        }

        if (use_disp) {
          // TPS_RGBD_kernels.cuh:443-472; sums are not read during a pass, so they
          // can be edited in place.
          const bool moved = (index != new_index);
          if (inlier && (!prev_inlier || moved)) {
            Sums& c = sums[new_index];
            c.dx += x; c.dy += y; c.dxx += (int64_t)x * x; c.dyy += (int64_t)y * y; c.dxy += (int64_t)x * y;
            const int64_t q = quant_disp(disp_v);
            c.dxd += (int64_t)x * q; c.dyd += (int64_t)y * q; c.dd += q; c.dn += 1;
          }
          if (prev_inlier && (!inlier || moved)) {
            Sums& c = sums[index];
            c.dx -= x; c.dy -= y; c.dxx -= (int64_t)x * x; c.dyy -= (int64_t)y * y; c.dxy -= (int64_t)x * y;
            const int64_t q = quant_disp(disp_v);
            c.dxd -= (int64_t)x * q; c.dyd -= (int64_t)y * q; c.dd -= q; c.dn -= 1;
          }
          if (inlier != prev_inlier) inlier_edits.push_back(InlierEdit{p, inlier});
        }
      }
    }

    // apply: neighbour boundary deltas (from the snapshot labels) ...
    for (const Change& c : changes) {
      for (int k = 0; k < 4; k++) {
        const int qx = c.x + nx[k], qy = c.y + ny[k];
        const int i_n = label_at(qx, qy);  // labels not yet modified: snapshot
        if (i_n == c.new_index) bound[qy * W + qx]--;
        else if (i_n == c.old_index) bound[qy * W + qx]++;
      }
    }
    // ... then own overrides, labels and colour/position sums
    for (const Change& c : changes) {
      const int p = c.y * W + c.x;
      bound[p] = c.b;
      Sums& o = sums[c.old_index];
      Sums& n = sums[c.new_index];
      const int r = rgba[4 * p], g = rgba[4 * p + 1], b = rgba[4 * p + 2];
      o.x -= c.x; o.y -= c.y; o.r -= r; o.g -= g; o.b -= b; o.n -= 1;
      n.x += c.x; n.y += c.y; n.r += r; n.g += g; n.b += b; n.n += 1;
    }
    for (const Change& c : changes) labels[c.y * W + c.x] = c.new_index;
    for (const InlierEdit& e : inlier_edits) inliers[e.p] = e.v;
  }

  // TPS_RGBD_kernels.cu:324-401
  void initSamples() {
    const int nbWalks = 10;
    const float radius = (float)cell / 2.f;
    const float dxs[4] = {-1.f, 0.f, 1.f, 0.f};
    const float dys[4] = {0.f, -1.f, 0.f, 1.f};
    for (int index = 0; index < S; index++)
      for (int s = 0; s < nbSamples; s++) {
        const int idx = index * nbSamples + s;
        const float cx = sp[index].xy_rg.x, cy = sp[index].xy_rg.y;
        float x = cx, y = cy;
        int i = tex_label(x, y);
        int k = 0;
        while (i != index && k++ < 10) {
          x = (float)((double)cx + ((double)radius * 2.) * (double)(orc_rng_uniform(rng, idx) - 1.f));
          y = (float)((double)cy + ((double)radius * 2.) * (double)(orc_rng_uniform(rng, idx) - 1.f));
          i = tex_label(x, y);
        }
        f3 xyd[3];
        float d = tex_disp(x, y);
        xyd[0] = xyd[1] = xyd[2] = mk3(x, y, d);
        for (int j = 0; j < 3; j++)
          for (int w = 0; w < nbWalks; w++) {
            const int dir = (int)(orc_rng_u32(rng, idx) & 3u);
            const float next_x = x + dxs[dir], next_y = y + dys[dir];
            i = tex_label(x, y);
            if (i == index && next_x >= 0 && next_x < (float)W && next_y >= 0 && next_y < (float)H) {
              x = next_x; y = next_y;
              const float dd = tex_disp(x, y);
              if (std::isfinite(dd)) xyd[j] = mk3(x, y, dd);
            }
          }
        f4 sample = {0.f, 0.f, 0.f, 0.f};
        if (!solvePlaneEquations(sample, xyd[0].x, xyd[0].y, 1.f, xyd[0].z, xyd[1].x, xyd[1].y, 1.f, xyd[1].z,
                                 xyd[2].x, xyd[2].y, 1.f, xyd[2].z)) {
          sample.x = 0.f; sample.y = 0.f; sample.z = xyd[2].z;
        }
        sample.w = 0.f;
        samples[idx] = sample;
      }
  }
  // TPS_RGBD_kernels.cu:403-433
  void evalSamples() {
    const float sigma2 = cfg.thresh_disp;
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        const int index = labels[y * W + x];
        const float d = disp[y * W + x];
        for (int k = 0; k < nbSamples; k++) {
          f4& theta = samples[index * nbSamples + k];
          if (std::isfinite(theta.z)) {
            const float dp = theta.x * (float)x + theta.y * (float)y + theta.z;
            const float dd = (d - dp) * (d - dp);
            if (dd < sigma2) theta.w += 1.f;
          }
        }
      }
  }
  // TPS_RGBD_kernels.cu:435-467
  void selectSamples() {
    for (int idx = 0; idx < S; idx++) {
      f4 best = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < nbSamples; k++) {
        const f4 theta = samples[idx * nbSamples + k];
        if (theta.w > best.w) best = theta;
      }
      sp[idx].theta_b.x = best.x; sp[idx].theta_b.y = best.y; sp[idx].theta_b.z = best.z;
      Sums& c = sums[idx];
      c.dx = c.dy = c.dxx = c.dyy = c.dxy = c.dxd = c.dyd = c.dd = c.dn = 0;
    }
  }
  // TPS_RGBD_kernels.cu:112-155 (ransac) / :157-190 (plain)
  void initDispCoeffs(bool ransac) {
    const float threshold = cfg.thresh_disp;
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        const int p = y * W + x;
        const int index = labels[p];
        const float d = disp[p];
        uint8_t inlier = 0;
        if (std::isfinite(d)) {
          bool ok = true;
          if (ransac) {
            const f4 theta = sp[index].theta_b;
            const float dp = theta.x * (float)x + theta.y * (float)y + theta.z;
            const float dd = (dp - d) * (dp - d);
            ok = std::isfinite(dd) && dd < threshold && dp > 0.f;
          }
          if (ok) {
            inlier = 0xff;
            Sums& c = sums[index];
            c.dx += x; c.dy += y; c.dxx += (int64_t)x * x; c.dyy += (int64_t)y * y; c.dxy += (int64_t)x * y;
            const int64_t q = quant_disp(d);
            c.dxd += (int64_t)x * q; c.dyd += (int64_t)y * q; c.dd += q; c.dn += 1;
          }
        }
        inliers[p] = inlier;
      }
  }

  // TPS_RGBD.cu:101-478
  void compute(const uint8_t* rgb, const float* depth) {
    const int N = W * H;
    // cudaMemset of superpixels and coeffs (TPS_RGBD.cu:126-128)
    std::fill(sp.begin(), sp.end(), Superpixel{{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}});
    std::fill(sums.begin(), sums.end(), Sums{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0});
    // cvtColor(BGR2BGRA) is a 3->4 channel copy, alpha 255 (TPS_RGBD.cu:136)
    for (int p = 0; p < N; p++) {
      rgba[4 * p] = rgb[3 * p]; rgba[4 * p + 1] = rgb[3 * p + 1]; rgba[4 * p + 2] = rgb[3 * p + 2]; rgba[4 * p + 3] = 255;
    }
    // depth2disp32F_kernel (TPS_RGBD_kernels.cu:278-296)
    for (int p = 0; p < N; p++) disp[p] = 1.f / depth[p];
    // initSuperpixelsRGBD_kernel (TPS_RGBD_kernels.cu:61-110)
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        const int p = y * W + x;
        const int index = gx * (y / cell) + (x / cell);
        labels[p] = index;
        int b = 0;
        if ((x + 1) % cell <= 1) b++;
        if ((y + 1) % cell <= 1) b++;
        bound[p] = b;
        Sums& c = sums[index];
        c.x += x; c.y += y; c.r += rgba[4 * p]; c.g += rgba[4 * p + 1]; c.b += rgba[4 * p + 2]; c.n += 1;
      }
    // inliersMat is not cleared per frame in the reference; it is fully rewritten by
    // initDispCoeffs* before its first read.
    mergeRGB();
    const int nbIters = cfg.seg_iter;
    for (int k = 0; k < nbIters / 2; k++) {
      pass(0, 0, false); mergeRGB();
      pass(1, 1, false); mergeRGB();
      pass(0, 1, false); mergeRGB();
      pass(1, 0, false); mergeRGB();
    }
    if (cfg.seg_use_ransac) {
      initSamples();
      evalSamples();
      selectSamples();
      initDispCoeffs(true);
    } else {
      initDispCoeffs(false);
    }
    mergeRGBD();
    for (int k = nbIters / 2; k < nbIters; k++) {
      pass(0, 0, true); mergeRGBD();
      pass(1, 1, true); mergeRGBD();
      pass(0, 1, true); mergeRGBD();
      pass(1, 0, true); mergeRGBD();
    }
  }

  // TPS_RGBD.cu:480-505, TPS_RGBD_kernels.cu:510-614
  void filter() {
    struct Node { f3 X, Z; float px, py; };
    std::vector<Node> data(S), next(S);
    for (int i = 0; i < S; i++) {
      const Superpixel& s = sp[i];
      f3 X = mk3(s.xy_rg.x * s.theta_b.x + s.xy_rg.y * s.theta_b.y + s.theta_b.z, s.theta_b.x, s.theta_b.y);
      data[i] = Node{X, X, s.xy_rg.x, s.xy_rg.y};
    }
    const float alpha = cfg.filter_alpha, beta = cfg.filter_beta, threshold = cfg.filter_threshold;
    const int v[4] = {-1, 0, 0, 1};
    const int u[4] = {0, -1, 1, 0};
    for (int it = 0; it < cfg.filter_iter; it++) {
      next = data;
      for (int y = 0; y < gy; y++)
        for (int x = 0; x < gx; x++) {
          const int idx = y * gx + x;
          Cov3 A = mkcov(alpha, 0.f, 0.f, alpha, 0.f, alpha);
          const Node node_i = data[idx];
          f3 R = alpha * node_i.Z;
          for (int j = 0; j < 4; j++) {
            const int yy = y + v[j], xx = x + u[j];
            // appendix B4: the reference tests x (not xx) against gridSizeX
            if (yy >= 0 && yy < gy && xx >= 0 && x < gx) {
              const int nidx = yy * gx + xx;
              if (nidx >= S) continue;  // the reference reads one element out of bounds here
              const Node node_j = data[nidx];
              const f3 Xj = node_j.X;
              const float dx = node_i.px - node_j.px;
              const float dy = node_i.py - node_j.py;
              const float dz = node_i.X.x - Xj.x;
              if (std::isfinite(dz) && dz * dz < threshold * threshold) {
                A.xx += beta * 2.f;
                A.xy += -beta * dx;
                A.xz += -beta * dy;
                A.yy += beta * (2.f + dx * dx);
                A.yz += beta * (dx * dy);
                A.zz += beta * (2.f + dy * dy);
                R.x += beta * (2.f * Xj.x + dx * Xj.y + dy * Xj.z);
                R.y += beta * (-dx * Xj.x + 2.f * Xj.y);
                R.z += beta * (-dy * Xj.x + 2.f * Xj.z);
              }
            }
          }
          Cov3 A_1;
          if (inverse(A, A_1)) next[idx].X = A_1 * R;
        }
      data.swap(next);
    }
    for (int i = 0; i < S; i++) {
      const f3 X = data[i].X;
      Superpixel& s = sp[i];
      s.theta_b.x = X.y;
      s.theta_b.y = X.z;
      s.theta_b.z = X.x - s.xy_rg.x * X.y - s.xy_rg.y * X.z;
    }
  }

  // TPS_RGBD_kernels.cu:469-508 (full-image ROI, scale 1)
  void renderDepth() {
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        const float xx = (float)x, yy = (float)y;
        const f4 theta = sp[labels[y * W + x]].theta_b;
        const float d = xx * theta.x + yy * theta.y + theta.z;
        slanted[y * W + x] = 1.f / d;
      }
  }
};

extern "C" OrcTps* orc_tps_create(const OrcConfig* cfg) {
  OrcTps* t = new OrcTps;
  t->cfg = *cfg;
  t->W = cfg->cam.width; t->H = cfg->cam.height; t->cell = cfg->cell_size;
  t->gx = (t->W + t->cell - 1) / t->cell;
  t->gy = (t->H + t->cell - 1) / t->cell;
  t->S = t->gx * t->gy;
  t->nbSamples = cfg->nb_samples;
  const size_t N = (size_t)t->W * t->H;
  t->labels.assign(N, 0); t->bound.assign(N, 0); t->inliers.assign(N, 0); t->rgba.assign(4 * N, 0);
  t->disp.assign(N, 0.f); t->slanted.assign(N, 0.f);
  t->sp.resize(t->S); t->sums.resize(t->S);
  t->samples.assign((size_t)t->S * t->nbSamples, f4{0, 0, 0, 0});
  // initRandStates_kernel (TPS_RGBD_kernels.cu:318-322), once per image size
  t->rng = orc_rng_create(t->S * t->nbSamples, 1234ULL);
  return t;
}
extern "C" void orc_tps_destroy(OrcTps* t) {
  if (!t) return;
  orc_rng_destroy(t->rng);
  delete t;
}
extern "C" void orc_tps_compute(OrcTps* t, const uint8_t* rgb, const float* depth) {
  t->compute(rgb, depth);   // tps->compute   (supersurfel_fusion.cu:189)
  t->filter();              // tps->filter    (:190)
  t->renderDepth();         // tps->computeDepthImage (:191)
}
extern "C" void orc_tps_get(const OrcTps* t, int32_t* labels, int32_t* bound, uint8_t* inliers, float* disp,
                            float* superpixels, float* slanted_depth, uint8_t* rgba) {
  const size_t N = (size_t)t->W * t->H;
  if (labels) std::memcpy(labels, t->labels.data(), N * 4);
  if (bound) std::memcpy(bound, t->bound.data(), N * 4);
  if (inliers) std::memcpy(inliers, t->inliers.data(), N);
  if (disp) std::memcpy(disp, t->disp.data(), N * 4);
  if (superpixels) std::memcpy(superpixels, t->sp.data(), (size_t)t->S * sizeof(Superpixel));
  if (slanted_depth) std::memcpy(slanted_depth, t->slanted.data(), N * 4);
  if (rgba) std::memcpy(rgba, t->rgba.data(), 4 * N);
}
extern "C" int orc_tps_nb_superpixels(const OrcTps* t) { return t->S; }
extern "C" void orc_tps_get_samples(const OrcTps* t, float* samples) {
  std::memcpy(samples, t->samples.data(), t->samples.size() * sizeof(f4));
}
