// TEST INFRASTRUCTURE ONLY -- CPU oracle, frame-to-model symmetric ICP.
// Restates core/include/supersurfel_fusion/dense_registration_kernels.cuh:175-291
// (computeSymmetricICPSystem) and core/src/dense_registration.cu:245-424
// (featureConstrainedSymmetricICP, with the always-empty sparse-feature lists).
// The 6x6 LDLT / LU / quaternion steps restate the published algorithms of the
// reference's vendored Eigen 3.3.7 (third_party/eigen3/Eigen/src/Cholesky/LDLT.h:290-400,
// 556-590; Geometry/Quaternion.h:747-785; Geometry/AngleAxis.h:218-243).
#include "oracle.h"
#include "oracle_math.h"
#include <cfloat>
#include <vector>
#include <omp.h>

using namespace orc;

namespace {

inline f3 ld3(const float* p, int i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
inline Mat33 ldm(const float* p) {
  return mkmat(mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]), mk3(p[6], p[7], p[8]));
}

// Sum of three products as the fused chain fma(a.z, b.z, fma(a.y, b.y, a.x * b.x)): the
// contraction nvcc applies to the reference's dot products (its SASS for
// computeSymmetricICPSystem / makeCorrespondences reads FMUL, FFMA, FFMA, FADD for
// R * p + t), fixed here so that the CUDA path (csrc/ssf_icp.cu, packed FFMA2 chains) and
// this restatement take every gate decision on bit-identical numbers.
inline float dotf(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline f3 mulf(const Mat33& m, f3 v) { return mk3(dotf(m.rows[0], v), dotf(m.rows[1], v), dotf(m.rows[2], v)); }

// Per-source term of dense_registration_kernels.cuh:207-281.  Returns false when the
// element contributes nothing; otherwise fills term[29].
inline bool icp_term(int i, const float* src_pos, const float* src_col, const float* src_orient,
                     const float* tgt_col, const float* tgt_orient, const float* tgt_conf,
                     const Mat33& R, f3 t, const OrcCam& cam, const int32_t* labels,
                     const float* depth, float* term) {
  f3 ps = ld3(src_pos, i);
  ps = mulf(R, ps) + t;
  int u = project_round(ps.x * cam.fx / ps.z + cam.cx);
  int v = project_round(ps.y * cam.fy / ps.z + cam.cy);
  if (!(u >= 0 && u < cam.width && v >= 0 && v < cam.height)) return false;
  int target_id = labels[v * cam.width + u];
  float zt = depth[v * cam.width + u];
  if (!(tgt_conf[target_id] > 0.0f && zt >= 0.2f && zt <= 5.0f)) return false;
  f3 dl = rgbToLab(ld3(src_col, i)) - rgbToLab(ld3(tgt_col, target_id));
  float dist_color = sqrtf(dotf(dl, dl));
  f3 pt = mk3(zt * ((float)u - cam.cx) / cam.fx, zt * ((float)v - cam.cy) / cam.fy, zt);
  f3 nt = mk3(tgt_orient[9 * target_id + 6], tgt_orient[9 * target_id + 7], tgt_orient[9 * target_id + 8]);
  // normalize (vector_math.cuh:247-252): v * rsqrtf(v.v); the host has no rsqrtf, the
  // correctly rounded 1/sqrt stands in (the device intrinsic is within 2 ulp of it)
  f3 m = mulf(R, mk3(src_orient[9 * i + 6], src_orient[9 * i + 7], src_orient[9 * i + 8]));
  f3 ns = m * (1.0f / sqrtf(dotf(m, m)));
  f3 dd = ps - pt;
  if (!(dist_color < 20.0f && sqrtf(dotf(dd, dd)) < 0.1f && fabsf(dotf(nt, ns)) > 0.8f)) return false;
  const float w = 1.0f;
  f3 d = pt - ps;
  f3 c1 = mk3(fmaf(pt.y, ns.z, -(pt.z * ns.y)), fmaf(pt.z, ns.x, -(pt.x * ns.z)), fmaf(pt.x, ns.y, -(pt.y * ns.x)));
  f3 c2 = mk3(fmaf(ps.y, nt.z, -(ps.z * nt.y)), fmaf(ps.z, nt.x, -(ps.x * nt.z)), fmaf(ps.x, nt.y, -(ps.y * nt.x)));
  float dn1 = dotf(d, ns);
  float dn2 = dotf(d, nt);
  float x1[6] = {c1.x, c1.y, c1.z, ns.x, ns.y, ns.z};
  float x2[6] = {c2.x, c2.y, c2.z, nt.x, nt.y, nt.z};
  int k = 0;
  for (int a = 0; a < 6; a++)
    for (int b = a; b < 6; b++) term[k++] = w * (x1[a] * x1[b] + x2[a] * x2[b]);
  for (int a = 0; a < 6; a++) term[21 + a] = w * (dn1 * x1[a] + dn2 * x2[a]);
  term[27] = w * dn2 * dn2;
  term[28] = 1.0f;
  return true;
}

void icp_system(int n_src, const float* src_pos, const float* src_col, const float* src_orient,
                const float* tgt_col, const float* tgt_orient, const float* tgt_conf, const Mat33& R,
                f3 t, const OrcCam& cam, const int32_t* labels, const float* depth, float* out29) {
  // The reference sums fp32 terms by a shared-memory tree and float atomics in
  // scheduling order (reduce_dev.cuh:24-77, dense_registration_kernels.cuh:71-85);
  // the oracle fixes the order-free limit of that: exact fp32 terms, accumulated in
  // double, rounded once to fp32.
  double acc[29];
  for (int k = 0; k < 29; k++) acc[k] = 0.0;
#pragma omp parallel
  {
    double loc[29];
    for (int k = 0; k < 29; k++) loc[k] = 0.0;
#pragma omp for schedule(static)
    for (int i = 0; i < n_src; i++) {
      float term[29];
      if (icp_term(i, src_pos, src_col, src_orient, tgt_col, tgt_orient, tgt_conf, R, t, cam, labels, depth, term))
        for (int k = 0; k < 29; k++) loc[k] += (double)term[k];
    }
#pragma omp critical
    for (int k = 0; k < 29; k++) acc[k] += loc[k];
  }
  for (int k = 0; k < 29; k++) out29[k] = (float)acc[k];
}

// ---- 6x6 dense helpers in double ---------------------------------------------
// LDLT with diagonal pivoting, lower storage (Eigen LDLT.h:290-400), then the
// solve of LDLT.h:556-590 (pseudo-inverse of D with tolerance DBL_MIN).
void ldlt_solve6(const double Ain[6][6], const double b[6], double x[6]) {
  const int n = 6;
  double m[6][6];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) m[i][j] = Ain[i][j];
  int perm[6];
  double temp[6];
  for (int k = 0; k < n; k++) {
    int big = k;
    double best = fabs(m[k][k]);
    for (int i = k + 1; i < n; i++)
      if (fabs(m[i][i]) > best) { best = fabs(m[i][i]); big = i; }
    perm[k] = big;
    if (k != big) {
      int s = n - big - 1;
      for (int j = 0; j < k; j++) std::swap(m[k][j], m[big][j]);
      for (int i = 0; i < s; i++) std::swap(m[big + 1 + i][k], m[big + 1 + i][big]);
      std::swap(m[k][k], m[big][big]);
      for (int i = k + 1; i < big; i++) std::swap(m[i][k], m[big][i]);
    }
    int rs = n - k - 1;
    if (k > 0) {
      for (int j = 0; j < k; j++) temp[j] = m[j][j] * m[k][j];
      double s = 0.0;
      for (int j = 0; j < k; j++) s += m[k][j] * temp[j];
      m[k][k] -= s;
      for (int i = 0; i < rs; i++) {
        double a = 0.0;
        for (int j = 0; j < k; j++) a += m[k + 1 + i][j] * temp[j];
        m[k + 1 + i][k] -= a;
      }
    }
    double akk = m[k][k];
    if (rs > 0 && fabs(akk) > 0.0)
      for (int i = 0; i < rs; i++) m[k + 1 + i][k] /= akk;
  }
  double y[6];
  for (int i = 0; i < n; i++) y[i] = b[i];
  for (int k = 0; k < n; k++) std::swap(y[k], y[perm[k]]);           // P b
  for (int i = 0; i < n; i++)                                          // L^-1
    for (int j = 0; j < i; j++) y[i] -= m[i][j] * y[j];
  for (int i = 0; i < n; i++) {                                        // D^+
    if (fabs(m[i][i]) > DBL_MIN) y[i] /= m[i][i]; else y[i] = 0.0;
  }
  for (int i = n - 1; i >= 0; i--)                                     // L^-T
    for (int j = i + 1; j < n; j++) y[i] -= m[j][i] * y[j];
  for (int k = n - 1; k >= 0; k--) std::swap(y[k], y[perm[k]]);       // P^-1
  for (int i = 0; i < n; i++) x[i] = y[i];
}

// diag(A^-1) by partial-pivot Gauss-Jordan (the reference uses
// JtJ.lu().inverse(), dense_registration.cu:394).
void inverse_diag6(const double Ain[6][6], double diag[6]) {
  const int n = 6;
  double a[6][12];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) { a[i][j] = Ain[i][j]; a[i][n + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < n; c++) {
    int p = c;
    for (int r = c + 1; r < n; r++)
      if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
    if (p != c) for (int j = 0; j < 2 * n; j++) std::swap(a[c][j], a[p][j]);
    double piv = a[c][c];
    for (int j = 0; j < 2 * n; j++) a[c][j] /= piv;
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      double f = a[r][c];
      if (f != 0.0) for (int j = 0; j < 2 * n; j++) a[r][j] -= f * a[c][j];
    }
  }
  for (int i = 0; i < n; i++) diag[i] = a[i][n + i];
}

// Quaternion(Matrix3).normalized().toRotationMatrix() in scalar type T
// (Eigen Quaternion.h:747-785, 560-590).
template <typename T>
void quat_renormalise(T m[3][3]) {
  T q[4];  // x y z w
  T t = m[0][0] + m[1][1] + m[2][2];
  if (t > T(0)) {
    t = std::sqrt(t + T(1));
    q[3] = T(0.5) * t;
    t = T(0.5) / t;
    q[0] = (m[2][1] - m[1][2]) * t;
    q[1] = (m[0][2] - m[2][0]) * t;
    q[2] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + T(1));
    q[i] = T(0.5) * t;
    t = T(0.5) / t;
    q[3] = (m[k][j] - m[j][k]) * t;
    q[j] = (m[j][i] + m[i][j]) * t;
    q[k] = (m[k][i] + m[i][k]) * t;
  }
  T nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int a = 0; a < 4; a++) q[a] /= nrm;
  const T tx = T(2) * q[0], ty = T(2) * q[1], tz = T(2) * q[2];
  const T twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const T txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const T tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  m[0][0] = T(1) - (tyy + tzz); m[0][1] = txy - twz;          m[0][2] = txz + twy;
  m[1][0] = txy + twz;          m[1][1] = T(1) - (txx + tzz); m[1][2] = tyz - twx;
  m[2][0] = txz - twy;          m[2][1] = tyz + twx;          m[2][2] = T(1) - (txx + tyy);
}

}  // namespace

extern "C" void orc_icp_system(int n_src, const float* src_pos, const float* src_col, const float* src_orient,
                               const float* tgt_col, const float* tgt_orient, const float* tgt_conf,
                               const float* R9, const float* t3, const OrcCam* cam, const int32_t* labels,
                               const float* depth, float* out29) {
  icp_system(n_src, src_pos, src_col, src_orient, tgt_col, tgt_orient, tgt_conf, ldm(R9),
             mk3(t3[0], t3[1], t3[2]), *cam, labels, depth, out29);
}

extern "C" int orc_icp(int n_src, const float* src_pos, const float* src_col, const float* src_orient,
                       const float* tgt_col, const float* tgt_orient, const float* tgt_conf,
                       const float* R_init9, const float* t_init3, const OrcCam* cam, const int32_t* labels,
                       const float* depth, int nb_iter, double cov_thresh, float* R_rel9, float* t_rel3,
                       OrcIcpStats* stats) {
  // dense_registration.cu:262-424
  bool valid = true;
  int iter = 0;
  double tf_inc[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  const Mat33 R_init = ldm(R_init9);
  const f3 t_init = mk3(t_init3[0], t_init3[1], t_init3[2]);
  Mat33 R_inc = identity33();
  f3 t_inc = mk3(0, 0, 0);
  double JtJ[6][6] = {{0}};
  double Jtr[6] = {0};
  double prev_error = DBL_MAX;
  double error = 0.0;
  float inliers = 0.0f;
  float sys[29] = {0};
  int iters_done = 0;

  while (iter++ < nb_iter) {
    iters_done++;
    R_inc = mkmat(mk3((float)tf_inc[0][0], (float)tf_inc[0][1], (float)tf_inc[0][2]),
                  mk3((float)tf_inc[1][0], (float)tf_inc[1][1], (float)tf_inc[1][2]),
                  mk3((float)tf_inc[2][0], (float)tf_inc[2][1], (float)tf_inc[2][2]));
    t_inc = mk3((float)tf_inc[0][3], (float)tf_inc[1][3], (float)tf_inc[2][3]);
    Mat33 R_corres = R_inc * R_init;
    f3 t_corres = R_inc * t_init + t_inc;

    icp_system(n_src, src_pos, src_col, src_orient, tgt_col, tgt_orient, tgt_conf, R_corres, t_corres,
               *cam, labels, depth, sys);

    // upper-triangular packing -> full symmetric (dense_registration.cu:326-331)
    int k = 0;
    for (int a = 0; a < 6; a++)
      for (int b = a; b < 6; b++) { JtJ[a][b] = (double)sys[k]; JtJ[b][a] = (double)sys[k]; k++; }
    for (int a = 0; a < 6; a++) Jtr[a] = (double)sys[21 + a];
    error = std::sqrt((double)(sys[27] / sys[28]));
    inliers = sys[28];

    if (inliers < 100.0f) { valid = false; break; }

    double Xp[6];
    ldlt_solve6(JtJ, Jtr, Xp);

    double tran[3] = {Xp[3], Xp[4], Xp[5]};
    double axis[3] = {Xp[0], Xp[1], Xp[2]};
    double axis_norm = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
    double angle = 0.5 * std::atan(axis_norm);
    // DIVERGENCE: the reference divides by a zero norm and
    // returns a NaN pose for exactly zero motion; the oracle (and the CUDA path)
    // treat a zero axis as the identity rotation.
    if (axis_norm > 0.0) { axis[0] /= axis_norm; axis[1] /= axis_norm; axis[2] /= axis_norm; }
    else { axis[0] = 1.0; axis[1] = 0.0; axis[2] = 0.0; angle = 0.0; }
    double c = std::cos(angle);
    for (int a = 0; a < 3; a++) tran[a] *= c;

    // AngleAxisd(angle, axis).toRotationMatrix()
    double Rr[3][3];
    {
      double s = std::sin(angle);
      double sa[3] = {s * axis[0], s * axis[1], s * axis[2]};
      double ca[3] = {(1.0 - c) * axis[0], (1.0 - c) * axis[1], (1.0 - c) * axis[2]};
      double tmp;
      tmp = ca[0] * axis[1]; Rr[0][1] = tmp - sa[2]; Rr[1][0] = tmp + sa[2];
      tmp = ca[0] * axis[2]; Rr[0][2] = tmp + sa[1]; Rr[2][0] = tmp - sa[1];
      tmp = ca[1] * axis[2]; Rr[1][2] = tmp - sa[0]; Rr[2][1] = tmp + sa[0];
      for (int a = 0; a < 3; a++) Rr[a][a] = ca[a] * axis[a] + c;
    }
    // iso_iter = Rot * Trans(tran) * Rot  =>  [Rr*Rr | Rr*tran]
    double tf_iter[4][4] = {{0}};
    double Rit[3][3];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) Rit[a][b] = Rr[a][0] * Rr[0][b] + Rr[a][1] * Rr[1][b] + Rr[a][2] * Rr[2][b];
    quat_renormalise<double>(Rit);
    for (int a = 0; a < 3; a++) {
      for (int b = 0; b < 3; b++) tf_iter[a][b] = Rit[a][b];
      tf_iter[a][3] = Rr[a][0] * tran[0] + Rr[a][1] * tran[1] + Rr[a][2] * tran[2];
    }
    tf_iter[3][3] = 1.0;

    double nt[4][4];
    for (int a = 0; a < 4; a++)
      for (int b = 0; b < 4; b++) {
        double s = 0.0;
        for (int q = 0; q < 4; q++) s += tf_iter[a][q] * tf_inc[q][b];
        nt[a][b] = s;
      }
    std::memcpy(tf_inc, nt, sizeof(nt));

    if (error / prev_error > 0.9995) break;
    prev_error = error;
  }

  // covariance gate on the LAST BUILT JtJ (appendix B10)
  double diag[6];
  inverse_diag6(JtJ, diag);
  for (int a = 0; a < 6; a++)
    if (diag[a] > cov_thresh) { valid = false; break; }

  Mat33 R = identity33();
  f3 t = mk3(0, 0, 0);
  if (valid) {
    // NB: t_inc here is still the value from the top of the last iteration
    // (dense_registration.cu:407 tests it before refreshing it at :411-417).
    if (length(t_inc) > 0.2f) valid = false;
    else {
      R_inc = mkmat(mk3((float)tf_inc[0][0], (float)tf_inc[0][1], (float)tf_inc[0][2]),
                    mk3((float)tf_inc[1][0], (float)tf_inc[1][1], (float)tf_inc[1][2]),
                    mk3((float)tf_inc[2][0], (float)tf_inc[2][1], (float)tf_inc[2][2]));
      t_inc = mk3((float)tf_inc[0][3], (float)tf_inc[1][3], (float)tf_inc[2][3]);
      R = transpose(R_inc);
      t = -(R * t_inc);
    }
  }
  for (int a = 0; a < 3; a++) {
    R_rel9[3 * a] = R.rows[a].x; R_rel9[3 * a + 1] = R.rows[a].y; R_rel9[3 * a + 2] = R.rows[a].z;
  }
  t_rel3[0] = t.x; t_rel3[1] = t.y; t_rel3[2] = t.z;
  if (stats) {
    stats->valid = valid ? 1 : 0;
    stats->iters = iters_done;
    stats->inliers = inliers;
    stats->error = error;
    std::memcpy(stats->last_system, sys, sizeof(sys));
  }
  return valid ? 1 : 0;
}


// ---- loop-closure registration: DenseRegistration::align ----------------------------
// (core/src/dense_registration.cu:52-243; kernels makeCorrespondences,
//  core/src/dense_registration_kernels.cu:27-100, and buildSymmetricPoint2PlaneSystem<128>,
//  core/include/supersurfel_fusion/dense_registration_kernels.cuh:87-173).
// Source = a keyframe's supersurfels, target = the current frame.  Differences from the
// frame-to-model loop that are restated as they are: the source confidence IS tested, the
// depth gate is isfinite() only, normals are re-normalised, the matched pairs are centred and
// scaled before the system is built, every one of nb_iter iterations runs (no convergence
// test), the translation gate is 0.3 m, and the returned transform is built from the
// R_inc / t_inc of the TOP of the last iteration (:90-96 vs :229-238: they are not refreshed
// after the last update, unlike :411-417 of the frame-to-model loop).
extern "C" int orc_align(int n_src, const float* src_pos, const float* src_col, const float* src_orient,
                         const float* src_conf, const float* tgt_col, const float* tgt_orient,
                         const float* tgt_conf, const float* R_init9, const float* t_init3, const OrcCam* cam,
                         const int32_t* labels, const float* depth, int nb_iter, double cov_thresh,
                         float* R9, float* t3, OrcIcpStats* stats) {
  bool valid = true;
  int iter = 0, iters_done = 0;
  double tf_inc[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  const Mat33 R_init = ldm(R_init9);
  const f3 t_init = mk3(t_init3[0], t_init3[1], t_init3[2]);
  Mat33 R_inc = identity33();
  f3 t_inc = mk3(0, 0, 0);
  double JtJ[6][6] = {{0}};
  double Jtr[6] = {0};
  float sys[29] = {0};
  int nb_pairs = 0;
  std::vector<f3> ms_p(n_src), ms_n(n_src), mt_p(n_src), mt_n(n_src);
  // rgbToLab of the keyframe colours is iteration-invariant
  std::vector<f3> src_lab(n_src);
  for (int i = 0; i < n_src; i++) src_lab[i] = rgbToLab(ld3(src_col, i));

  while (iter++ < nb_iter) {
    iters_done++;
    R_inc = mkmat(mk3((float)tf_inc[0][0], (float)tf_inc[0][1], (float)tf_inc[0][2]),
                  mk3((float)tf_inc[1][0], (float)tf_inc[1][1], (float)tf_inc[1][2]),
                  mk3((float)tf_inc[2][0], (float)tf_inc[2][1], (float)tf_inc[2][2]));
    t_inc = mk3((float)tf_inc[0][3], (float)tf_inc[1][3], (float)tf_inc[2][3]);
    const Mat33 R = R_inc * R_init;
    const f3 t = R_inc * t_init + t_inc;

    // makeCorrespondences + remove_if compaction (kept in source order)
    nb_pairs = 0;
    for (int i = 0; i < n_src; i++) {
      if (!(src_conf[i] > 0.0f)) continue;
      const f3 p_view = mulf(R, ld3(src_pos, i)) + t;
      const int u = project_round(p_view.x * cam->fx / p_view.z + cam->cx);
      const int v = project_round(p_view.y * cam->fy / p_view.z + cam->cy);
      if (u < 0 || u >= cam->width || v < 0 || v >= cam->height) continue;
      const int target_id = labels[v * cam->width + u];
      if (!(tgt_conf[target_id] > 0.0f)) continue;
      const f3 dl = src_lab[i] - rgbToLab(ld3(tgt_col, target_id));
      const float dist_color = sqrtf(dotf(dl, dl));
      const float t_depth = depth[v * cam->width + u];
      if (!std::isfinite(t_depth)) continue;
      f3 s_normal = mk3(src_orient[9 * i + 6], src_orient[9 * i + 7], src_orient[9 * i + 8]);
      s_normal = s_normal * (1.0f / sqrtf(dotf(s_normal, s_normal)));
      s_normal = mulf(R, s_normal);
      s_normal = s_normal * (1.0f / sqrtf(dotf(s_normal, s_normal)));
      f3 t_normal = mk3(tgt_orient[9 * target_id + 6], tgt_orient[9 * target_id + 7], tgt_orient[9 * target_id + 8]);
      t_normal = t_normal * (1.0f / sqrtf(dotf(t_normal, t_normal)));
      const f3 t_position = mk3(t_depth * ((float)u - cam->cx) / cam->fx, t_depth * ((float)v - cam->cy) / cam->fy, t_depth);
      const f3 dd = p_view - t_position;
      if (dist_color < 20.0f && sqrtf(dotf(dd, dd)) < 0.1f && fabsf(dotf(s_normal, t_normal)) > 0.8f) {
        ms_p[nb_pairs] = p_view; ms_n[nb_pairs] = s_normal;
        mt_p[nb_pairs] = t_position; mt_n[nb_pairs] = t_normal;
        nb_pairs++;
      }
    }
    if (nb_pairs < 100) { valid = false; break; }

    // centroids and isotropic scale of the matched sets (thrust::reduce over floats in an
    // unspecified tree order in the reference; here the order-free limit: double sums)
    double cs[3] = {0, 0, 0}, ct[3] = {0, 0, 0};
    for (int k = 0; k < nb_pairs; k++) {
      cs[0] += ms_p[k].x; cs[1] += ms_p[k].y; cs[2] += ms_p[k].z;
      ct[0] += mt_p[k].x; ct[1] += mt_p[k].y; ct[2] += mt_p[k].z;
    }
    const float fn = (float)nb_pairs;
    const f3 source_centroid = mk3((float)cs[0] / fn, (float)cs[1] / fn, (float)cs[2] / fn);
    const f3 target_centroid = mk3((float)ct[0] / fn, (float)ct[1] / fn, (float)ct[2] / fn);
    double sc_sum_t = 0.0, sc_sum_s = 0.0;
    for (int k = 0; k < nb_pairs; k++) {
      const f3 a = mt_p[k] - target_centroid, b = ms_p[k] - source_centroid;
      sc_sum_t += (double)dotf(a, a);
      sc_sum_s += (double)dotf(b, b);
    }
    float scale = (float)sc_sum_t;
    scale += (float)sc_sum_s;
    scale = sqrtf(scale / (2.0f * fn));
    scale = 1.0f / scale;

    // buildSymmetricPoint2PlaneSystem<128>
    double acc[29];
    for (int k = 0; k < 29; k++) acc[k] = 0.0;
    for (int k = 0; k < nb_pairs; k++) {
      const f3 ps = scale * (ms_p[k] - source_centroid);
      const f3 pt = scale * (mt_p[k] - target_centroid);
      const f3 ns = ms_n[k] * (1.0f / sqrtf(dotf(ms_n[k], ms_n[k])));
      const f3 nt = mt_n[k] * (1.0f / sqrtf(dotf(mt_n[k], mt_n[k])));
      const f3 d = pt - ps;
      const f3 c1 = mk3(fmaf(pt.y, ns.z, -(pt.z * ns.y)), fmaf(pt.z, ns.x, -(pt.x * ns.z)), fmaf(pt.x, ns.y, -(pt.y * ns.x)));
      const f3 c2 = mk3(fmaf(ps.y, nt.z, -(ps.z * nt.y)), fmaf(ps.z, nt.x, -(ps.x * nt.z)), fmaf(ps.x, nt.y, -(ps.y * nt.x)));
      const float dn1 = dotf(d, ns), dn2 = dotf(d, nt);
      const float x1[6] = {c1.x, c1.y, c1.z, ns.x, ns.y, ns.z};
      const float x2[6] = {c2.x, c2.y, c2.z, nt.x, nt.y, nt.z};
      int q = 0;
      for (int a = 0; a < 6; a++)
        for (int b = a; b < 6; b++) acc[q++] += (double)(x1[a] * x1[b] + x2[a] * x2[b]);
      for (int a = 0; a < 6; a++) acc[21 + a] += (double)(dn1 * x1[a] + dn2 * x2[a]);
      acc[27] += (double)(dn2 * dn2);
      acc[28] += 1.0;
    }
    for (int k = 0; k < 29; k++) sys[k] = (float)acc[k];
    int q = 0;
    for (int a = 0; a < 6; a++)
      for (int b = a; b < 6; b++) { JtJ[a][b] = (double)sys[q]; JtJ[b][a] = (double)sys[q]; q++; }
    for (int a = 0; a < 6; a++) Jtr[a] = (double)sys[21 + a];

    double Xp[6];
    ldlt_solve6(JtJ, Jtr, Xp);
    double tran[3] = {Xp[3], Xp[4], Xp[5]};
    double axis[3] = {Xp[0], Xp[1], Xp[2]};
    const double axis_norm = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
    double angle = (double)(0.5f) * std::atan(axis_norm);
    if (axis_norm > 0.0) { axis[0] /= axis_norm; axis[1] /= axis_norm; axis[2] /= axis_norm; }   // appendix B9 guard
    else { axis[0] = 1.0; axis[1] = 0.0; axis[2] = 0.0; angle = 0.0; }
    const double c = std::cos(angle), sn = std::sin(angle);
    for (int a = 0; a < 3; a++) { tran[a] /= (double)scale; tran[a] *= c; }
    double Rr[3][3];
    {
      const double sa[3] = {sn * axis[0], sn * axis[1], sn * axis[2]};
      const double ca[3] = {(1.0 - c) * axis[0], (1.0 - c) * axis[1], (1.0 - c) * axis[2]};
      double tmp;
      tmp = ca[0] * axis[1]; Rr[0][1] = tmp - sa[2]; Rr[1][0] = tmp + sa[2];
      tmp = ca[0] * axis[2]; Rr[0][2] = tmp + sa[1]; Rr[2][0] = tmp - sa[1];
      tmp = ca[1] * axis[2]; Rr[1][2] = tmp - sa[0]; Rr[2][1] = tmp + sa[0];
      for (int a = 0; a < 3; a++) Rr[a][a] = ca[a] * axis[a] + c;
    }
    // iso_iter = Trans(ct) * Rot * Trans(tran) * Rot * Trans(-cs): linear part Rr*Rr,
    // translation ct + Rr*tran - (Rr*Rr)*cs; the rotation block alone is then re-normalised
    double R2[3][3], tf_iter[4][4] = {{0}};
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) R2[a][b] = Rr[a][0] * Rr[0][b] + Rr[a][1] * Rr[1][b] + Rr[a][2] * Rr[2][b];
    const double csd[3] = {(double)source_centroid.x, (double)source_centroid.y, (double)source_centroid.z};
    const double ctd[3] = {(double)target_centroid.x, (double)target_centroid.y, (double)target_centroid.z};
    for (int a = 0; a < 3; a++)
      tf_iter[a][3] = ctd[a] + (Rr[a][0] * tran[0] + Rr[a][1] * tran[1] + Rr[a][2] * tran[2]) -
                      (R2[a][0] * csd[0] + R2[a][1] * csd[1] + R2[a][2] * csd[2]);
    quat_renormalise<double>(R2);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) tf_iter[a][b] = R2[a][b];
    tf_iter[3][3] = 1.0;
    double nt4[4][4];
    for (int a = 0; a < 4; a++)
      for (int b = 0; b < 4; b++) {
        double sum = 0.0;
        for (int k = 0; k < 4; k++) sum += tf_iter[a][k] * tf_inc[k][b];
        nt4[a][b] = sum;
      }
    std::memcpy(tf_inc, nt4, sizeof(nt4));
  }

  double diag[6];
  inverse_diag6(JtJ, diag);
  for (int a = 0; a < 6; a++)
    if (diag[a] > cov_thresh) { valid = false; break; }
  Mat33 R = identity33();
  f3 t = mk3(0, 0, 0);
  if (valid) {
    if (length(t_inc) > 0.3f) valid = false;
    else {
      R = transpose(R_inc);          // R_inc / t_inc of the top of the last iteration (see header)
      t = -(R * t_inc);
    }
  }
  for (int a = 0; a < 3; a++) {
    R9[3 * a] = R.rows[a].x; R9[3 * a + 1] = R.rows[a].y; R9[3 * a + 2] = R.rows[a].z;
  }
  t3[0] = t.x; t3[1] = t.y; t3[2] = t.z;
  if (stats) {
    stats->valid = valid ? 1 : 0;
    stats->iters = iters_done;
    stats->inliers = (float)nb_pairs;
    stats->error = sys[28] > 0 ? std::sqrt((double)(sys[27] / sys[28])) : 0.0;
    std::memcpy(stats->last_system, sys, sizeof(sys));
  }
  return valid ? 1 : 0;
}

extern "C" void orc_compose_pose(float* R9, float* t3, const float* R_rel9, const float* t_rel3) {
  // supersurfel_fusion.cu:313-328 (Eigen::Quaternionf round trip in float)
  Mat33 R = ldm(R9), Rr = ldm(R_rel9);
  f3 t = mk3(t3[0], t3[1], t3[2]), tr = mk3(t_rel3[0], t_rel3[1], t_rel3[2]);
  t = R * tr + t;
  R = R * Rr;
  float m[3][3] = {{R.rows[0].x, R.rows[0].y, R.rows[0].z},
                   {R.rows[1].x, R.rows[1].y, R.rows[1].z},
                   {R.rows[2].x, R.rows[2].y, R.rows[2].z}};
  quat_renormalise<float>(m);
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) R9[3 * a + b] = m[a][b];
  t3[0] = t.x; t3[1] = t.y; t3[2] = t.z;
}

extern "C" void orc_set_num_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
