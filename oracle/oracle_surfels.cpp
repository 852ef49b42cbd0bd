// TEST INFRASTRUCTURE ONLY -- CPU oracle, supersurfel extraction and model fusion.
// Restates core/src/supersurfel_fusion_kernels.cu:113-224 (computeSupersurfelCoeffs,
// computeSupersurfels), :348-467 (insertSupersurfels, filterModel), :522-682
// (findBestMatches, updateSupersurfels) and the host sequencing of
// core/src/supersurfel_fusion.cu:351-483, 551-593.
//
// Deterministic serialisation of the reference's races:
//  * per-superpixel moment sums are order-free: each fp32 term is quantised to
//    2^-32 fixed point and summed in int64 (the reference uses fp32 atomicAdd in
//    scheduling order, supersurfel_fusion_kernels.cu:148-165);
//  * association keeps the true arg-min, ties to the lowest model id (B11);
//  * insertion appends in ascending frame-superpixel order, and the post-cull
//    reorder is a STABLE partition active | inactive | removed (B13).
#include "oracle.h"
#include "oracle_math.h"
#include <algorithm>
#include <vector>

using namespace orc;

namespace {

const double kFix = 4294967296.0;  // 2^32
const double kFixClamp = 1152921504606846976.0;  // 2^60

inline int64_t quant(float v) {
  double s = (double)v * kFix;
  if (s != s) s = 0.0;  // NaN contributes nothing
  s = fmin(fmax(s, -kFixClamp), kFixClamp);
  return (int64_t)llrint(s);
}
inline float dequant(int64_t v) { return (float)((double)v * (1.0 / kFix)); }

inline f3 ld3(const float* p, int i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
inline void st3(float* p, int i, f3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }
inline Mat33 ldm(const float* p, int i = 0) {
  p += 9 * i;
  return mkmat(mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]), mk3(p[6], p[7], p[8]));
}
inline void stm(float* p, int i, const Mat33& m) {
  p += 9 * i;
  for (int r = 0; r < 3; r++) { p[3 * r] = m.rows[r].x; p[3 * r + 1] = m.rows[r].y; p[3 * r + 2] = m.rows[r].z; }
}
inline Cov3 ldc(const float* p, int i) { p += 6 * i; return mkcov(p[0], p[1], p[2], p[3], p[4], p[5]); }
inline void stc(float* p, int i, const Cov3& c) {
  p += 6 * i; p[0] = c.xx; p[1] = c.xy; p[2] = c.xz; p[3] = c.yy; p[4] = c.yz; p[5] = c.zz;
}

void copy_row(const OrcSurfels& s, int i, OrcSurfels& d, int j) {
  for (int k = 0; k < 3; k++) d.positions[3 * j + k] = s.positions[3 * i + k];
  for (int k = 0; k < 3; k++) d.colors[3 * j + k] = s.colors[3 * i + k];
  for (int k = 0; k < 2; k++) d.stamps[2 * j + k] = s.stamps[2 * i + k];
  for (int k = 0; k < 9; k++) d.orientations[9 * j + k] = s.orientations[9 * i + k];
  for (int k = 0; k < 6; k++) d.shapes[6 * j + k] = s.shapes[6 * i + k];
  for (int k = 0; k < 2; k++) d.dims[2 * j + k] = s.dims[2 * i + k];
  d.confidences[j] = s.confidences[i];
}

}  // namespace

extern "C" void orc_generate_supersurfels(const OrcCam* cam, int S, const uint8_t* rgba,
                                          const float* slanted_depth, const int32_t* labels,
                                          const uint8_t* inliers, const int32_t* bound, float z_min,
                                          float z_max, int stamp, OrcSurfels* frame) {
  const int W = cam->width, H = cam->height;
  // frame.memset (supersurfel_fusion.cu:560)
  std::vector<int64_t> acc((size_t)S * 15, 0);
  std::vector<int64_t> cnt(S, 0);
  // computeSupersurfelCoeffs (supersurfel_fusion_kernels.cu:113-167)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const int p = y * W + x;
      if (!inliers[p]) continue;
      int index = labels[p];
      int b = bound[p];
      float depth = slanted_depth[p];
      if (std::isfinite(depth) && depth > 0.0f && b == 0) {
        f3 pos = mk3(((float)x - cam->cx) * depth / cam->fx, ((float)y - cam->cy) * depth / cam->fy, depth);
        f3 lab = rgbToLab(mk3((float)rgba[4 * p], (float)rgba[4 * p + 1], (float)rgba[4 * p + 2]));
        Cov3 cov = outer_product(pos);
        int64_t* a = &acc[(size_t)index * 15];
        a[0] += quant(pos.x); a[1] += quant(pos.y); a[2] += quant(pos.z);
        a[3] += quant(lab.x); a[4] += quant(lab.y); a[5] += quant(lab.z);
        a[6] += quant(cov.xx); a[7] += quant(cov.xy); a[8] += quant(cov.xz);
        a[9] += quant(cov.yy); a[10] += quant(cov.yz); a[11] += quant(cov.zz);
        cnt[index] += 1;
      }
    }
  // computeSupersurfels (supersurfel_fusion_kernels.cu:169-224)
  for (int k = 0; k < S; k++) {
    const int64_t* a = &acc[(size_t)k * 15];
    f3 position = mk3(dequant(a[0]), dequant(a[1]), dequant(a[2]));
    f3 color = mk3(dequant(a[3]), dequant(a[4]), dequant(a[5]));
    Cov3 shape = mkcov(dequant(a[6]), dequant(a[7]), dequant(a[8]), dequant(a[9]), dequant(a[10]), dequant(a[11]));
    float conf = (float)cnt[k];
    Mat33 orient = mkmat(mk3(0, 0, 0), mk3(0, 0, 0), mk3(0, 0, 0));
    f2 dims = {0.f, 0.f};
    i2 st = {0, 0};

    float z = position.z / conf;
    if (std::isfinite(z) && conf > 100.0f && z > z_min && z < z_max) {
      position.x /= conf; position.y /= conf; position.z = z;
      color.x /= conf; color.y /= conf; color.z /= conf;
      color = labToRgb(color);
      shape = shape / conf - outer_product(position);
      f3 vals;
      eigenDecomposition(shape, orient, vals, 10);
      dims.x = vals.x; dims.y = vals.y;
      st.x = stamp; st.y = stamp;
      if (vals.x / vals.y > 50.0f) conf = -1.0f;
    } else {
      conf = -1.0f;  // raw sums stay in place, as in the reference
    }
    st3(frame->positions, k, position);
    st3(frame->colors, k, color);
    frame->stamps[2 * k] = st.x; frame->stamps[2 * k + 1] = st.y;
    stm(frame->orientations, k, orient);
    stc(frame->shapes, k, shape);
    frame->dims[2 * k] = dims.x; frame->dims[2 * k + 1] = dims.y;
    frame->confidences[k] = conf;
  }
}

extern "C" void orc_fuse(const OrcCam* cam, int S, const OrcSurfels* frame_in, OrcSurfels* model, int nb_max,
                         const float* R9, const float* t3, const int32_t* labels, const float* slanted_depth,
                         float z_min, float z_max, int stamp, int delta_t, float conf_thresh,
                         OrcFuseCounts* counts) {
  const OrcSurfels& frame = *frame_in;
  const int W = cam->width, H = cam->height;
  const Mat33 R = ldm(R9);
  const f3 t = mk3(t3[0], t3[1], t3[2]);
  int nbSupersurfels = counts->nb_supersurfels;
  int nbVisible = counts->nb_visible;
  int nbRemoved = 0, nbMatched = 0, nbInserted = 0;
  counts->nb_removed_stale = counts->nb_removed_invalid = counts->nb_removed_occluded = 0;

  if (nbSupersurfels > 0) {
    std::vector<unsigned char> matched(S, 0);
    if (nbVisible > 0) {
      std::vector<float> score_id(S, -1.0f), score_d(S, 0.05f);
      // findBestMatches (supersurfel_fusion_kernels.cu:522-599)
      const Mat33 Rview = transpose(R);
      const f3 tview = -(Rview * t);
      const Mat33 Rt = transpose(R);
      for (int m = 0; m < nbVisible; m++) {
        if (!(model->confidences[m] > 0.0f)) continue;
        f3 mp = ld3(model->positions, m);
        f3 pv = Rview * mp + tview;
        int px = project_round(pv.x * cam->fx / pv.z + cam->cx);
        int py = project_round(pv.y * cam->fy / pv.z + cam->cy);
        if (!(pv.z > z_min && pv.z < z_max && px >= 0 && px < W && py >= 0 && py < H)) continue;
        int f = labels[py * W + px];
        matched[f] = 1;
        if (!(frame.confidences[f] > 0.0f)) continue;
        f3 fp = R * ld3(frame.positions, f) + t;
        Mat33 frot = ldm(frame.orientations, f) * Rt;
        f3 fn = normalize(frot.rows[2]);
        f3 mn = normalize(ldm(model->orientations, m).rows[2]);
        f3 flab = rgbToLab(ld3(frame.colors, f));
        f3 mlab = rgbToLab(ld3(model->colors, m));
        float dist = length(mp - fp);
        float lab_dist = length(mlab - flab);
        float delta_norm = fabsf(dot(mn, fn));
        if (lab_dist < 15.0f && delta_norm > 0.8f && dist < 0.05f) {
          if (dist < score_d[f]) { score_d[f] = dist; score_id[f] = (float)m; }
        }
      }
      // updateSupersurfels (supersurfel_fusion_kernels.cu:601-682)
      for (int f = 0; f < S; f++) {
        int m = (int)score_id[f];
        if (!(matched[f] && m >= 0)) continue;
        nbMatched++;
        f3 mp = ld3(model->positions, m);
        f3 fp = R * ld3(frame.positions, f) + t;
        Cov3 fshape = mult_ABAt(R, ldc(frame.shapes, f));
        Cov3 mshape = ldc(model->shapes, m);
        f3 flab = rgbToLab(ld3(frame.colors, f));
        f3 mlab = rgbToLab(ld3(model->colors, m));
        float m_conf = model->confidences[m];
        float f_conf = frame.confidences[f];
        float ratio = 1.0f / (m_conf + f_conf);
        model->stamps[2 * m + 1] = stamp;
        f3 fused_color = labToRgb(ratio * (f_conf * flab + m_conf * mlab));
        Cov3 f1, m1, fused_shape, fused1;
        f3 fused_pos;
        float w = ratio * f_conf;
        bool info = false;
        if (inverse(fshape, f1) && inverse(mshape, m1)) {
          fused1 = w * f1 + (1.0f - w) * m1;
          if (inverse(fused1, fused_shape)) {
            fused_pos = fused_shape * ((w * f1) * fp + ((1.0f - w) * m1) * mp);
            info = true;
          }
        }
        if (!info) {
          fused_shape = ratio * (f_conf * fshape + m_conf * mshape);
          fused_pos = ratio * (f_conf * fp + m_conf * mp);
        }
        st3(model->positions, m, fused_pos);
        model->confidences[m] = m_conf + f_conf;
        stc(model->shapes, m, fused_shape);
        Mat33 vecs; f3 vals;
        eigenDecomposition(fused_shape, vecs, vals, 10);
        stm(model->orientations, m, vecs);
        st3(model->colors, m, fused_color);
        model->dims[2 * m] = vals.x; model->dims[2 * m + 1] = vals.y;
      }
    }
    // insertSupersurfels (supersurfel_fusion_kernels.cu:348-395), ascending frame id
    {
      const Mat33 Rt = transpose(R);
      for (int f = 0; f < S; f++) {
        if (!(frame.confidences[f] > 0.0f && !matched[f])) continue;
        int k = nbSupersurfels;
        if (k < nb_max) {
          st3(model->positions, k, R * ld3(frame.positions, f) + t);
          stc(model->shapes, k, mult_ABAt(R, ldc(frame.shapes, f)));
          stm(model->orientations, k, ldm(frame.orientations, f) * Rt);
          model->confidences[k] = frame.confidences[f];
          st3(model->colors, k, ld3(frame.colors, f));
          model->stamps[2 * k] = stamp; model->stamps[2 * k + 1] = stamp;
          model->dims[2 * k] = frame.dims[2 * f]; model->dims[2 * k + 1] = frame.dims[2 * f + 1];
          nbSupersurfels++;
          nbInserted++;
        }
      }
    }
    // filterModel (supersurfel_fusion_kernels.cu:397-467)
    std::vector<int> states(nbSupersurfels, 0);
    nbVisible = 0;
    {
      const Mat33 Rv = transpose(R);
      const f3 tv = -(Rv * t);
      for (int i = 0; i < nbSupersurfels; i++) {
        int state = 0;
        int time_diff = stamp - model->stamps[2 * i + 1];
        float conf = model->confidences[i];
        if ((time_diff > delta_t && conf < conf_thresh && stamp > delta_t) || conf <= 0.0f) {
          if (time_diff > delta_t && conf < conf_thresh && stamp > delta_t) counts->nb_removed_stale++;
          else counts->nb_removed_invalid++;
          model->confidences[i] = -1.0f;
          state = 2;
        } else {
          f3 p = Rv * ld3(model->positions, i) + tv;
          if (p.z > z_min && p.z < z_max) {
            float u = cam->fx * p.x / p.z + cam->cx;
            float v = cam->fy * p.y / p.z + cam->cy;
            if (u >= 0.0f && u < (float)W && v >= 0.0f && v < (float)H) {
              float z = slanted_depth[tex_coord(v, H) * W + tex_coord(u, W)];
              if (p.z < 0.8f * z) { model->confidences[i] = -1.0f; state = 2; counts->nb_removed_occluded++; }
            } else state = 1;
          } else state = 1;
        }
        if (state == 0) nbVisible++;
        if (state == 2) nbRemoved++;
        states[i] = state;
      }
    }
    // sort_by_key on states (supersurfel_fusion.cu:469) restated as a stable partition
    {
      std::vector<int> order;
      order.reserve(nbSupersurfels);
      for (int s = 0; s < 3; s++)
        for (int i = 0; i < nbSupersurfels; i++)
          if (states[i] == s) order.push_back(i);
      std::vector<float> pos(3 * (size_t)nbSupersurfels), col(3 * (size_t)nbSupersurfels),
          ori(9 * (size_t)nbSupersurfels), shp(6 * (size_t)nbSupersurfels), dms(2 * (size_t)nbSupersurfels),
          cnf(nbSupersurfels);
      std::vector<int> stp(2 * (size_t)nbSupersurfels);
      OrcSurfels tmp{pos.data(), col.data(), stp.data(), ori.data(), shp.data(), dms.data(), cnf.data()};
      for (int j = 0; j < nbSupersurfels; j++) copy_row(*model, order[j], tmp, j);
      for (int j = 0; j < nbSupersurfels; j++) copy_row(tmp, j, *model, j);
    }
    nbSupersurfels -= nbRemoved;
  } else {
    // first frame: model <- frame, all S entries (supersurfel_fusion.cu:477-483)
    for (int f = 0; f < S; f++) copy_row(frame, f, *model, f);
    nbSupersurfels = S;
    nbVisible = S;
  }
  counts->nb_supersurfels = nbSupersurfels;
  counts->nb_visible = nbVisible;
  counts->nb_removed = nbRemoved;
  counts->nb_matched = nbMatched;
  counts->nb_inserted = nbInserted;
}
