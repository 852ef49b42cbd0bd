// TEST INFRASTRUCTURE ONLY -- CPU oracle for the supersurfel hot path.
//
// This header restates, in plain C++ for the host, the fp32 arithmetic the
// reference's device helpers define.  It is a checker: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build or call it.  Nothing under supersurfel_fusion_b200/ includes it.
//
// Parity status: the reference ships no tests, golden vectors or CPU path for
// this code, so the restatement is pinned against the
// reference's own kernels compiled into oracle/_ref (see oracle/Makefile,
// tests/test_ref_harness.py) and is otherwise "parity unpinned".
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).  Build with -ffp-contract=off so that a*b+c is two
// roundings, which is the arithmetic the CUDA path is compiled to as well.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };
struct i2 { int x, y; };

// core/include/supersurfel_fusion/matrix_types.h:26-42
struct Cov3 { float xx, xy, xz, yy, yz, zz; };
struct Mat33 { f3 rows[3]; };

// core/include/supersurfel_fusion/cam_param.hpp:27-31
struct Cam { float fx, fy, cx, cy; int height, width; };

inline f3 mk3(float x, float y, float z) { return f3{x, y, z}; }

// core/include/supersurfel_fusion/vector_math.cuh:164-252 (float3 operators)
inline f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
inline f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
inline f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
inline f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(f3 v) { return sqrtf(dot(v, v)); }
// vector_math.cuh:117-120
inline f3 cross(f3 a, f3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// vector_math.cuh:247-252: v * rsqrtf(dot(v,v)).  The host has no rsqrtf; use the
// correctly rounded 1/sqrt (the device intrinsic is within 2 ulp of it).
inline f3 normalize(f3 v) {
  float inv = 1.0f / sqrtf(dot(v, v));
  return v * inv;
}

// ---- Cov3 (matrix_math.cuh:29-222) ----------------------------------------
inline Cov3 mkcov(float xx, float xy, float xz, float yy, float yz, float zz) {
  return Cov3{xx, xy, xz, yy, yz, zz};
}
// matrix_math.cuh:41-63
inline bool inverse(const Cov3& in, Cov3& out) {
  out.xx = in.zz * in.yy - in.yz * in.yz;
  out.xy = in.xz * in.yz - in.zz * in.xy;
  out.xz = in.xy * in.yz - in.xz * in.yy;
  out.yy = in.zz * in.xx - in.xz * in.xz;
  out.yz = in.xy * in.xz - in.xx * in.yz;
  out.zz = in.xx * in.yy - in.xy * in.xy;
  float det = in.xx * out.xx + in.xy * out.xy + in.xz * out.xz;
  if (fabs((double)det) > 1e-9) {  // fabs(float) vs double literal: compared in double
    out.xx /= det; out.xy /= det; out.xz /= det;
    out.yy /= det; out.yz /= det; out.zz /= det;
    return true;
  }
  return false;
}
inline Cov3 operator+(const Cov3& a, const Cov3& b) {
  return mkcov(a.xx + b.xx, a.xy + b.xy, a.xz + b.xz, a.yy + b.yy, a.yz + b.yz, a.zz + b.zz);
}
inline Cov3 operator-(const Cov3& a, const Cov3& b) {
  return mkcov(a.xx - b.xx, a.xy - b.xy, a.xz - b.xz, a.yy - b.yy, a.yz - b.yz, a.zz - b.zz);
}
inline Cov3 operator*(const Cov3& a, float b) {
  return mkcov(a.xx * b, a.xy * b, a.xz * b, a.yy * b, a.yz * b, a.zz * b);
}
inline Cov3 operator*(float b, const Cov3& a) {
  return mkcov(b * a.xx, b * a.xy, b * a.xz, b * a.yy, b * a.yz, b * a.zz);
}
inline Cov3 operator/(const Cov3& a, float b) {
  return mkcov(a.xx / b, a.xy / b, a.xz / b, a.yy / b, a.yz / b, a.zz / b);
}
// matrix_math.cuh:165-170
inline f3 operator*(const Cov3& m, f3 b) {
  return mk3(m.xx * b.x + m.xy * b.y + m.xz * b.z,
             m.xy * b.x + m.yy * b.y + m.yz * b.z,
             m.xz * b.x + m.yz * b.y + m.zz * b.z);
}
// matrix_math.cuh:184-194
inline Cov3 square(const Cov3& a) {
  Cov3 r;
  r.xx = a.xx * a.xx + a.xy * a.xy + a.xz * a.xz;
  r.xy = a.xx * a.xy + a.xy * a.yy + a.xz * a.yz;
  r.xz = a.xx * a.xz + a.xy * a.yz + a.xz * a.zz;
  r.yy = a.xy * a.xy + a.yy * a.yy + a.yz * a.yz;
  r.yz = a.xy * a.xz + a.yy * a.yz + a.yz * a.zz;
  r.zz = a.xz * a.xz + a.yz * a.yz + a.zz * a.zz;
  return r;
}
// matrix_math.cuh:212-222
inline Cov3 outer_product(f3 v) {
  return mkcov(v.x * v.x, v.x * v.y, v.x * v.z, v.y * v.y, v.y * v.z, v.z * v.z);
}
inline float trace(const Cov3& a) { return a.xx + a.yy + a.zz; }

// ---- Mat33 (matrix_math.cuh:241-478) ---------------------------------------
inline Mat33 mkmat(f3 a, f3 b, f3 c) { Mat33 m; m.rows[0] = a; m.rows[1] = b; m.rows[2] = c; return m; }
inline Mat33 identity33() { return mkmat(mk3(1, 0, 0), mk3(0, 1, 0), mk3(0, 0, 1)); }
// matrix_math.cuh:461-466
inline f3 operator*(const Mat33& a, f3 b) { return mk3(dot(a.rows[0], b), dot(a.rows[1], b), dot(a.rows[2], b)); }
// matrix_math.cuh:364-387
inline Mat33 operator*(const Mat33& a, const Mat33& b) {
  Mat33 r;
  for (int i = 0; i < 3; i++) {
    const f3& ai = a.rows[i];
    r.rows[i] = mk3(ai.x * b.rows[0].x + ai.y * b.rows[1].x + ai.z * b.rows[2].x,
                    ai.x * b.rows[0].y + ai.y * b.rows[1].y + ai.z * b.rows[2].y,
                    ai.x * b.rows[0].z + ai.y * b.rows[1].z + ai.z * b.rows[2].z);
  }
  return r;
}
// matrix_math.cuh:476-483
inline Mat33 transpose(const Mat33& a) {
  return mkmat(mk3(a.rows[0].x, a.rows[1].x, a.rows[2].x),
               mk3(a.rows[0].y, a.rows[1].y, a.rows[2].y),
               mk3(a.rows[0].z, a.rows[1].z, a.rows[2].z));
}
// matrix_math.cuh:442-459
inline Cov3 mult_ABAt(const Mat33& A, const Cov3& B) {
  f3 r1 = mk3(B.xx, B.xy, B.xz), r2 = mk3(B.xy, B.yy, B.yz), r3 = mk3(B.xz, B.yz, B.zz);
  Mat33 BAtt = mkmat(mk3(dot(r1, A.rows[0]), dot(r2, A.rows[0]), dot(r3, A.rows[0])),
                     mk3(dot(r1, A.rows[1]), dot(r2, A.rows[1]), dot(r3, A.rows[1])),
                     mk3(dot(r1, A.rows[2]), dot(r2, A.rows[2]), dot(r3, A.rows[2])));
  return mkcov(dot(A.rows[0], BAtt.rows[0]), dot(A.rows[0], BAtt.rows[1]), dot(A.rows[0], BAtt.rows[2]),
               dot(A.rows[1], BAtt.rows[1]), dot(A.rows[1], BAtt.rows[2]), dot(A.rows[2], BAtt.rows[2]));
}

// ---- colour (vector_math.cuh:543-585) ---------------------------------------
// vector_math.cuh:566-585
inline f3 rgbToLab(f3 c) {
  float r = c.x / 255.0f, g = c.y / 255.0f, b = c.z / 255.0f;
  r = (r > 0.04045f) ? powf((r + 0.055f) / 1.055f, 2.4f) : r / 12.92f;
  g = (g > 0.04045f) ? powf((g + 0.055f) / 1.055f, 2.4f) : g / 12.92f;
  b = (b > 0.04045f) ? powf((b + 0.055f) / 1.055f, 2.4f) : b / 12.92f;
  float x = (r * 0.4124f + g * 0.3575f + b * 0.1805f) / 0.95047f;
  float y = (r * 0.2126f + g * 0.7152f + b * 0.0722f);
  float z = (r * 0.0193f + g * 0.1192f + b * 0.9505f) / 1.08883f;
  x = (x > 0.008856f) ? cbrtf(x) : 7.787f * x + 16.0f / 116.0f;
  y = (y > 0.008856f) ? cbrtf(y) : 7.787f * y + 16.0f / 116.0f;
  z = (z > 0.008856f) ? cbrtf(z) : 7.787f * z + 16.0f / 116.0f;
  return mk3(116.0f * y - 16.0f, 500.0f * (x - y), 200.0f * (y - z));
}
// vector_math.cuh:543-564.  Note the two double literals (1.8758, 1.0570): the g and
// b rows are evaluated in double and rounded once, as written in the reference.
inline f3 labToRgb(f3 c) {
  float y = (c.x + 16.0f) / 116.0f;
  float x = c.y / 500.0f + y;
  float z = y - c.z / 200.0f;
  x = 0.95047f * ((powf(x, 3.0f) > 0.008856f) ? powf(x, 3.0f) : (x - 16.0f / 116.0f) / 7.787f);
  y = 1.0f * ((powf(y, 3.0f) > 0.008856f) ? powf(y, 3.0f) : (y - 16.0f / 116.0f) / 7.787f);
  z = 1.08883f * ((powf(z, 3.0f) > 0.008856f) ? powf(z, 3.0f) : (z - 16.0f / 116.0f) / 7.787f);
  float r = x * 3.2406f - y * 1.5372f - z * 0.4986f;
  float g = (float)((double)(-x * 0.9689f) + (double)y * 1.8758 + (double)(z * 0.0415f));
  float b = (float)((double)(x * 0.0557f - y * 0.2040f) + (double)z * 1.0570);
  r = (r > 0.0031308f) ? (1.055f * powf(r, 1.0f / 2.4f) - 0.055f) : 12.92f * r;
  g = (g > 0.0031308f) ? (1.055f * powf(g, 1.0f / 2.4f) - 0.055f) : 12.92f * g;
  b = (b > 0.0031308f) ? (1.055f * powf(b, 1.0f / 2.4f) - 0.055f) : 12.92f * b;
  return mk3(fmaxf(0.0f, fminf(1.0f, r)) * 255.0f,
             fmaxf(0.0f, fminf(1.0f, g)) * 255.0f,
             fmaxf(0.0f, fminf(1.0f, b)) * 255.0f);
}

// core/src/supersurfel_fusion_kernels.cu:48-111 (eigenDecomposition): n squarings of
// the trace-normalised matrix and of its complement, "row holding the max entry"
// selection, eigenvalue read off at the largest SIGNED eigenvector component.
inline f3 max_row(const Cov3& A) {
  float vmax = fmaxf(fmaxf(fmaxf(fmaxf(fmaxf(A.xx, A.xy), A.xz), A.yy), A.yz), A.zz);
  if (A.xx == vmax || A.xy == vmax || A.xz == vmax) return normalize(mk3(A.xx, A.xy, A.xz));
  if (A.yy == vmax || A.yz == vmax) return normalize(mk3(A.xy, A.yy, A.yz));
  return normalize(mk3(A.xz, A.yz, A.zz));
}
inline float rayleigh_at_max(const Cov3& A, f3 e) {
  float emax = fmaxf(fmaxf(e.x, e.y), e.z);
  if (e.x == emax) return (A.xx * e.x + A.xy * e.y + A.xz * e.z) / e.x;
  if (e.y == emax) return (A.xy * e.x + A.yy * e.y + A.yz * e.z) / e.y;
  return (A.xz * e.x + A.yz * e.y + A.zz * e.z) / e.z;
}
inline void eigenDecomposition(const Cov3& A, Mat33& vecs, f3& vals, int n) {
  Cov3 Ai = A / trace(A);
  Cov3 Bi = mkcov(1.f - Ai.xx, -Ai.xy, -Ai.xz, 1.f - Ai.yy, -Ai.yz, 1.f - Ai.zz);
  for (int i = 0; i < n; ++i) {
    Ai = square(Ai); Ai = Ai / trace(Ai);
    Bi = square(Bi); Bi = Bi / trace(Bi);
  }
  vecs.rows[0] = max_row(Ai);
  vecs.rows[2] = max_row(Bi);
  vecs.rows[1] = cross(vecs.rows[2], vecs.rows[0]);
  vals.x = rayleigh_at_max(A, vecs.rows[0]);
  vals.y = rayleigh_at_max(A, vecs.rows[1]);
  vals.z = rayleigh_at_max(A, vecs.rows[2]);
}

// core/include/supersurfel_fusion/texture_impl.hpp:30-49: pitch-2D texture, point
// filter, clamp addressing, unnormalised coordinates => texel (floor(x), floor(y))
// clamped to the image.
inline int tex_coord(float c, int n) {
  if (!(c >= 0.0f)) return 0;  // negatives and NaN clamp to 0
  if (c >= (float)n) return n - 1;
  return (int)floorf(c);
}

// lroundf as the projection uses it (dense_registration_kernels.cuh:214-215,
// supersurfel_fusion_kernels.cu:564): round half away from zero.  A non-finite or
// huge coordinate can never pass the later distance gate, so it is mapped to an
// out-of-image sentinel instead of relying on float->long conversion overflow.
inline int project_round(float v) {
  if (!std::isfinite(v) || fabsf(v) >= 1.0e9f) return -1000000000;
  return (int)lroundf(v);
}

}  // namespace orc
