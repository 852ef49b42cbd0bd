// TEST INFRASTRUCTURE ONLY -- CPU oracle of the small consumers on either side of the model
// (SURVEY.md section 8f ranks 3-4):
//   applyDeformation           core/src/deformation_graph_kernels.cu:27-73
//   quatToRotMat/rotMatToQuat  core/include/supersurfel_fusion/matrix_math.cuh:512-585
//   marker geometry            node/supersurfel_fusion_node.cpp:303-413
//   TUM trajectory line        node/supersurfel_fusion_rgbd_benchmark_node.cpp:616-620,727-729
//   extractLocalPointCloud     core/src/supersurfel_fusion_kernels.cu:490-520, supersurfel_fusion.cu:884-920
//   applyTransformSuperSurfel  core/src/supersurfel_fusion_kernels.cu:467-488
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <cmath>

#include "oracle.h"
#include "oracle_math.h"

using namespace orc;

namespace {

struct Q4 { float x, y, z, w; };

// matrix_math.cuh:529-585, quirks included: index 2 wins when m22 exceeds EITHER other diagonal
Q4 rotMatToQuat(const Mat33& m) {
  Q4 q;
  float s;
  const float trace = m.rows[0].x + m.rows[1].y + m.rows[2].z;
  if (trace > 0) {
    s = sqrtf(trace + 1);
    q.w = 0.5f * s;
    s = 0.5f / s;
    q.x = (m.rows[2].y - m.rows[1].z) * s;
    q.y = (m.rows[0].z - m.rows[2].x) * s;
    q.z = (m.rows[1].x - m.rows[0].y) * s;
  } else {
    int i = 0;
    if (m.rows[1].y > m.rows[0].x) i = 1;
    if (m.rows[2].z > m.rows[0].x || m.rows[2].z > m.rows[1].y) i = 2;
    switch (i) {
      case 0:
        s = sqrtf(1.0f + m.rows[0].x - m.rows[1].y - m.rows[2].z);
        q.x = 0.5f * s; s = 0.5f / s;
        q.w = (m.rows[2].y - m.rows[1].z) * s; q.y = (m.rows[0].y + m.rows[1].x) * s; q.z = (m.rows[0].z + m.rows[2].x) * s;
        break;
      case 1:
        s = sqrtf(1.0f + m.rows[1].y - m.rows[0].x - m.rows[2].z);
        q.y = 0.5f * s; s = 0.5f / s;
        q.w = (m.rows[0].z - m.rows[2].x) * s; q.x = (m.rows[0].y + m.rows[1].x) * s; q.z = (m.rows[1].z + m.rows[2].y) * s;
        break;
      default:
        s = sqrtf(1.0f + m.rows[2].z - m.rows[0].x - m.rows[1].y);
        q.z = 0.5f * s; s = 0.5f / s;
        q.w = (m.rows[1].x - m.rows[0].y) * s; q.x = (m.rows[0].z + m.rows[2].x) * s; q.y = (m.rows[1].z + m.rows[2].y) * s;
        break;
    }
  }
  return q;
}

// matrix_math.cuh:512-527; `wy` really is w*z there (:521)
Mat33 quatToRotMat(const Q4& q) {
  const float x2 = q.x * q.x, y2 = q.y * q.y, z2 = q.z * q.z;
  const float xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z;
  const float wx = q.w * q.x, wy = q.w * q.z, wz = q.w * q.z;
  return mkmat(mk3(1.0f - 2.0f * (y2 + z2), 2.0f * (xy - wz), 2.0f * (xz + wy)),
               mk3(2.0f * (xy + wz), 1.0f - 2.0f * (x2 + z2), 2.0f * (yz - wx)),
               mk3(2.0f * (xz - wy), 2.0f * (yz + wx), 1.0f - 2.0f * (x2 + y2)));
}

}  // namespace

extern "C" {

void orc_apply_deformation(float* positions, float* orientations, float* shapes, const float* node_pos,
                           const float* node_rot, const float* node_trans, const float* weights,
                           const int32_t* nn, int model_size) {
  for (int i = 0; i < model_size; i++) {
    const f3 pi = mk3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
    f3 po = mk3(0.f, 0.f, 0.f);
    Q4 bq = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 4; k++) {
      const int node = nn[4 * i + k];
      const float wk = weights[4 * i + k];
      const f3 gk = mk3(node_pos[3 * node], node_pos[3 * node + 1], node_pos[3 * node + 2]);
      const f3 tk = mk3(node_trans[3 * node], node_trans[3 * node + 1], node_trans[3 * node + 2]);
      const float* r = node_rot + 9 * node;
      const Mat33 Rk = mkmat(mk3(r[0], r[1], r[2]), mk3(r[3], r[4], r[5]), mk3(r[6], r[7], r[8]));
      const Q4 qk = rotMatToQuat(Rk);
      po = po + wk * (Rk * (pi - gk) + gk + tk);
      bq.x += wk * qk.x; bq.y += wk * qk.y; bq.z += wk * qk.z; bq.w += wk * qk.w;
    }
    const float len = sqrtf(bq.x * bq.x + bq.y * bq.y + bq.z * bq.z + bq.w * bq.w);
    bq.x /= len; bq.y /= len; bq.z /= len; bq.w /= len;
    const Mat33 av = quatToRotMat(bq);
    float* o = orientations + 9 * i;
    const Mat33 O = mkmat(mk3(o[0], o[1], o[2]), mk3(o[3], o[4], o[5]), mk3(o[6], o[7], o[8]));
    const Mat33 On = O * transpose(av);
    for (int r2 = 0; r2 < 3; r2++) { o[3 * r2] = On.rows[r2].x; o[3 * r2 + 1] = On.rows[r2].y; o[3 * r2 + 2] = On.rows[r2].z; }
    float* sh = shapes + 6 * i;
    const Cov3 S = mult_ABAt(av, mkcov(sh[0], sh[1], sh[2], sh[3], sh[4], sh[5]));
    sh[0] = S.xx; sh[1] = S.xy; sh[2] = S.xz; sh[3] = S.yy; sh[4] = S.yz; sh[5] = S.zz;
    positions[3 * i] = po.x; positions[3 * i + 1] = po.y; positions[3 * i + 2] = po.z;
  }
}

// extractLocalPointCloudKernel (supersurfel_fusion_kernels.cu:490-520) with the view transform its caller
// builds from the pose (supersurfel_fusion.cu:896-897: R_view = R^T, t_view = -R_view t).  The reference
// appends by atomic ticket; here ascending model order (compare order-insensitively).
int orc_extract_local_point_cloud(int n, const float* positions, const float* orientations, const float* confidences,
                                  float conf_thresh, const float* R_pose9, const float* t_pose3, float radius,
                                  float* out_pos, float* out_nrm) {
  const Mat33 R = mkmat(mk3(R_pose9[0], R_pose9[1], R_pose9[2]), mk3(R_pose9[3], R_pose9[4], R_pose9[5]),
                        mk3(R_pose9[6], R_pose9[7], R_pose9[8]));
  const Mat33 Rv = transpose(R);
  const f3 tv = -(Rv * mk3(t_pose3[0], t_pose3[1], t_pose3[2]));
  int count = 0;
  for (int k = 0; k < n; k++) {
    if (!(confidences[k] >= conf_thresh)) continue;
    const f3 p = Rv * mk3(positions[3 * k], positions[3 * k + 1], positions[3 * k + 2]) + tv;
    if (!(length(p) < radius)) continue;
    const float* o = orientations + 9 * k;
    const f3 nrm = normalize(Rv * mk3(o[6], o[7], o[8]));     // rows[2] = normal
    out_pos[3 * count] = p.x; out_pos[3 * count + 1] = p.y; out_pos[3 * count + 2] = p.z;
    out_nrm[3 * count] = nrm.x; out_nrm[3 * count + 1] = nrm.y; out_nrm[3 * count + 2] = nrm.z;
    count++;
  }
  return count;
}

// applyTransformSuperSurfel (supersurfel_fusion_kernels.cu:467-488)
void orc_transform_model(int n, float* positions, float* orientations, float* shapes, const float* confidences,
                         const float* R9, const float* t3) {
  const Mat33 R = mkmat(mk3(R9[0], R9[1], R9[2]), mk3(R9[3], R9[4], R9[5]), mk3(R9[6], R9[7], R9[8]));
  const f3 t = mk3(t3[0], t3[1], t3[2]);
  const Mat33 Rt = transpose(R);
  for (int k = 0; k < n; k++) {
    if (confidences[k] <= 0.0f) continue;
    const f3 p = R * mk3(positions[3 * k], positions[3 * k + 1], positions[3 * k + 2]) + t;
    positions[3 * k] = p.x; positions[3 * k + 1] = p.y; positions[3 * k + 2] = p.z;
    float* o = orientations + 9 * k;
    const Mat33 On = mkmat(mk3(o[0], o[1], o[2]), mk3(o[3], o[4], o[5]), mk3(o[6], o[7], o[8])) * Rt;
    for (int r = 0; r < 3; r++) { o[3 * r] = On.rows[r].x; o[3 * r + 1] = On.rows[r].y; o[3 * r + 2] = On.rows[r].z; }
    float* sh = shapes + 6 * k;
    const Cov3 S = mult_ABAt(R, mkcov(sh[0], sh[1], sh[2], sh[3], sh[4], sh[5]));
    sh[0] = S.xx; sh[1] = S.xy; sh[2] = S.xz; sh[3] = S.yy; sh[4] = S.yz; sh[5] = S.zz;
  }
}

void orc_markers(const float* positions, const float* colors, const float* orientations, const float* dims,
                 const float* confidences, int n, float conf_thresh, float* points, float* out_colors) {
  for (int i = 0; i < n; i++) {
    float* P = points + (size_t)i * 18;
    float* C = out_colors + (size_t)i * 24;
    if (confidences[i] > conf_thresh) {
      float v0 = 3.0f * sqrtf(dims[2 * i]);
      float v1 = 3.0f * sqrtf(dims[2 * i + 1]);
      const f3 e0 = mk3(orientations[9 * i], orientations[9 * i + 1], orientations[9 * i + 2]);
      const f3 e1 = mk3(orientations[9 * i + 3], orientations[9 * i + 4], orientations[9 * i + 5]);
      f3 pos = mk3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
      if (!std::isfinite(v0)) v0 = 0;
      if (!std::isfinite(v1)) v1 = 0;
      if (!std::isfinite(pos.x) || !std::isfinite(pos.y) || !std::isfinite(pos.z)) pos = mk3(0.f, 0.f, 0.f);
      const f3 a = v0 * e0, b = v1 * e1;
      const f3 p0 = mk3(pos.x + a.x + b.x, pos.y + a.y + b.y, pos.z + a.z + b.z);
      const f3 p1 = mk3(pos.x + a.x - b.x, pos.y + a.y - b.y, pos.z + a.z - b.z);
      const f3 p2 = mk3(pos.x - a.x - b.x, pos.y - a.y - b.y, pos.z - a.z - b.z);
      const f3 p3 = mk3(pos.x - a.x + b.x, pos.y - a.y + b.y, pos.z - a.z + b.z);
      const f3 tri[6] = {p0, p1, p2, p0, p2, p3};
      for (int k = 0; k < 6; k++) {
        P[3 * k] = tri[k].x; P[3 * k + 1] = tri[k].y; P[3 * k + 2] = tri[k].z;
        C[4 * k] = colors[3 * i] / 255; C[4 * k + 1] = colors[3 * i + 1] / 255; C[4 * k + 2] = colors[3 * i + 2] / 255;
        C[4 * k + 3] = 1.f;
      }
    } else {
      for (int k = 0; k < 18; k++) P[k] = 0.f;
      for (int k = 0; k < 6; k++) { C[4 * k] = 0.f; C[4 * k + 1] = 0.f; C[4 * k + 2] = 0.f; C[4 * k + 3] = 1.f; }
    }
  }
}

// "timestamp tx ty tz qx qy qz qw\n": tf::Transform from the float pose, tf::Matrix3x3::getRotation
// in double, operator<< of doubles (6 significant digits)
int orc_format_tum_pose(const float* R9, const float* t3, const char* timestamp, char* line, int line_size) {
  const double m[3][3] = {{R9[0], R9[1], R9[2]}, {R9[3], R9[4], R9[5]}, {R9[6], R9[7], R9[8]}};
  double q[4];
  const double trace = m[0][0] + m[1][1] + m[2][2];
  if (trace > 0.0) {
    double s = sqrt(trace + 1.0);
    q[3] = s * 0.5;
    s = 0.5 / s;
    q[0] = (m[2][1] - m[1][2]) * s; q[1] = (m[0][2] - m[2][0]) * s; q[2] = (m[1][0] - m[0][1]) * s;
  } else {
    const int i = m[0][0] < m[1][1] ? (m[1][1] < m[2][2] ? 2 : 1) : (m[0][0] < m[2][2] ? 2 : 0);
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[i] = s * 0.5;
    s = 0.5 / s;
    q[3] = (m[k][j] - m[j][k]) * s; q[j] = (m[j][i] + m[i][j]) * s; q[k] = (m[k][i] + m[i][k]) * s;
  }
  return snprintf(line, (size_t)line_size, "%s %g %g %g %g %g %g %g\n", timestamp, (double)t3[0], (double)t3[1],
                  (double)t3[2], q[0], q[1], q[2], q[3]);
}

}  // extern "C"
