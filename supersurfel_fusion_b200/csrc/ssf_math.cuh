// Device arithmetic of the supersurfel hot path (sm_100a).
//
// Semantics follow the reference's device helpers (matrix_math.cuh, vector_math.cuh,
// supersurfel_fusion_kernels.cu:48-111; cited per function) but the code is organised
// around registers-only small-vector types.  Everything here is compiled with
// -fmad=false: decision arithmetic must round exactly like the specification
// (mul and add are separate IEEE operations), and the kernels that use it are
// bound by HBM/L2 traffic or launch latency, not by FP32 issue.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace ssf {

// Programmatic dependent launch (PDL).  A kernel launched with the programmatic-serialisation attribute may
// start once every CTA of its predecessor has executed griddepcontrol.launch_dependents (or exited), and must
// execute griddepcontrol.wait before it touches anything the predecessor wrote (the wait returns when the
// predecessor has completed and its writes are visible).  Both instructions are no-ops for a kernel launched
// without the attribute (launch_kernel, ssf_engine.h).  Measured on the synchronous VGA frame (SSF_PDL,
// tools/pdl_ab.sh): a trigger at the TOP of every kernel loses (0.470 vs 0.465 ms: the successor's CTAs take SM
// slots while the predecessor still needs them), but the 40 fused segmentation passes of a frame triggering AFTER their
// decisions and waiting AFTER their argument arithmetic gain 7 % (0.432 ms): that is the default (mode 3).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_trigger();
  pdl_wait();
}

// ---- TMA bulk copies + mbarrier (the async-proxy path global -> shared memory) ---------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
struct V3 { float x, y, z; };
struct Sym3 { float xx, xy, xz, yy, yz, zz; };   // symmetric 3x3 (reference: Cov3, matrix_types.h:26-31)
struct M3 { V3 r0, r1, r2; };                     // rows (reference: Mat33, matrix_types.h:33-36)

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
// vector_math.cuh:235-238
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// vector_math.cuh:117-120
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// vector_math.cuh:241-244
__device__ __forceinline__ float length(V3 v) { return sqrtf(dot(v, v)); }
// vector_math.cuh:247-252 (v * rsqrtf(v.v))
__device__ __forceinline__ V3 normalize(V3 v) { return v * rsqrtf(dot(v, v)); }

__device__ __forceinline__ M3 m3(V3 a, V3 b, V3 c) { M3 m; m.r0 = a; m.r1 = b; m.r2 = c; return m; }
// matrix_math.cuh:461-466
__device__ __forceinline__ V3 operator*(const M3& m, V3 b) { return v3(dot(m.r0, b), dot(m.r1, b), dot(m.r2, b)); }
// matrix_math.cuh:476-483
__device__ __forceinline__ M3 transpose(const M3& a) {
  return m3(v3(a.r0.x, a.r1.x, a.r2.x), v3(a.r0.y, a.r1.y, a.r2.y), v3(a.r0.z, a.r1.z, a.r2.z));
}
__device__ __forceinline__ V3 row_times(V3 a, const M3& b) {  // one row of matrix_math.cuh:364-387
  return v3(a.x * b.r0.x + a.y * b.r1.x + a.z * b.r2.x,
            a.x * b.r0.y + a.y * b.r1.y + a.z * b.r2.y,
            a.x * b.r0.z + a.y * b.r1.z + a.z * b.r2.z);
}
__device__ __forceinline__ M3 operator*(const M3& a, const M3& b) {
  return m3(row_times(a.r0, b), row_times(a.r1, b), row_times(a.r2, b));
}

__device__ __forceinline__ Sym3 sym3(float xx, float xy, float xz, float yy, float yz, float zz) {
  Sym3 s; s.xx = xx; s.xy = xy; s.xz = xz; s.yy = yy; s.yz = yz; s.zz = zz; return s;
}
__device__ __forceinline__ Sym3 operator+(const Sym3& a, const Sym3& b) {
  return sym3(a.xx + b.xx, a.xy + b.xy, a.xz + b.xz, a.yy + b.yy, a.yz + b.yz, a.zz + b.zz);
}
__device__ __forceinline__ Sym3 operator-(const Sym3& a, const Sym3& b) {
  return sym3(a.xx - b.xx, a.xy - b.xy, a.xz - b.xz, a.yy - b.yy, a.yz - b.yz, a.zz - b.zz);
}
__device__ __forceinline__ Sym3 operator*(float s, const Sym3& a) {
  return sym3(s * a.xx, s * a.xy, s * a.xz, s * a.yy, s * a.yz, s * a.zz);
}
__device__ __forceinline__ Sym3 operator/(const Sym3& a, float s) {
  return sym3(a.xx / s, a.xy / s, a.xz / s, a.yy / s, a.yz / s, a.zz / s);
}
__device__ __forceinline__ float trace(const Sym3& a) { return a.xx + a.yy + a.zz; }
// matrix_math.cuh:165-170
__device__ __forceinline__ V3 operator*(const Sym3& m, V3 b) {
  return v3(m.xx * b.x + m.xy * b.y + m.xz * b.z, m.xy * b.x + m.yy * b.y + m.yz * b.z,
            m.xz * b.x + m.yz * b.y + m.zz * b.z);
}
// matrix_math.cuh:184-194
__device__ __forceinline__ Sym3 square(const Sym3& a) {
  return sym3(a.xx * a.xx + a.xy * a.xy + a.xz * a.xz, a.xx * a.xy + a.xy * a.yy + a.xz * a.yz,
              a.xx * a.xz + a.xy * a.yz + a.xz * a.zz, a.xy * a.xy + a.yy * a.yy + a.yz * a.yz,
              a.xy * a.xz + a.yy * a.yz + a.yz * a.zz, a.xz * a.xz + a.yz * a.yz + a.zz * a.zz);
}
// matrix_math.cuh:212-222
__device__ __forceinline__ Sym3 outer(V3 v) {
  return sym3(v.x * v.x, v.x * v.y, v.x * v.z, v.y * v.y, v.y * v.z, v.z * v.z);
}
// matrix_math.cuh:41-63: adjugate / determinant, rejected when |det| <= 1e-9
__device__ __forceinline__ bool invert(const Sym3& in, Sym3& out) {
  out.xx = in.zz * in.yy - in.yz * in.yz;
  out.xy = in.xz * in.yz - in.zz * in.xy;
  out.xz = in.xy * in.yz - in.xz * in.yy;
  out.yy = in.zz * in.xx - in.xz * in.xz;
  out.yz = in.xy * in.xz - in.xx * in.yz;
  out.zz = in.xx * in.yy - in.xy * in.xy;
  const float det = in.xx * out.xx + in.xy * out.xy + in.xz * out.xz;
  if (fabs((double)det) > 1e-9) {
    out.xx /= det; out.xy /= det; out.xz /= det; out.yy /= det; out.yz /= det; out.zz /= det;
    return true;
  }
  return false;
}
// matrix_math.cuh:442-459: A B A^T for symmetric B
__device__ __forceinline__ Sym3 rotate_sym(const M3& A, const Sym3& B) {
  const V3 b0 = v3(B.xx, B.xy, B.xz), b1 = v3(B.xy, B.yy, B.yz), b2 = v3(B.xz, B.yz, B.zz);
  const V3 t0 = v3(dot(b0, A.r0), dot(b1, A.r0), dot(b2, A.r0));
  const V3 t1 = v3(dot(b0, A.r1), dot(b1, A.r1), dot(b2, A.r1));
  const V3 t2 = v3(dot(b0, A.r2), dot(b1, A.r2), dot(b2, A.r2));
  return sym3(dot(A.r0, t0), dot(A.r0, t1), dot(A.r0, t2), dot(A.r1, t1), dot(A.r1, t2), dot(A.r2, t2));
}

// sRGB (0..255) -> CIELab, vector_math.cuh:566-585
__device__ __forceinline__ float srgb_to_linear(float c) {
  return (c > 0.04045f) ? powf((c + 0.055f) / 1.055f, 2.4f) : c / 12.92f;
}
__device__ __forceinline__ float lab_f(float t) { return (t > 0.008856f) ? cbrtf(t) : 7.787f * t + 16.0f / 116.0f; }
__device__ __forceinline__ V3 rgb_to_lab(V3 c) {
  const float r = srgb_to_linear(c.x / 255.0f), g = srgb_to_linear(c.y / 255.0f), b = srgb_to_linear(c.z / 255.0f);
  const float x = lab_f((r * 0.4124f + g * 0.3575f + b * 0.1805f) / 0.95047f);
  const float y = lab_f(r * 0.2126f + g * 0.7152f + b * 0.0722f);
  const float z = lab_f((r * 0.0193f + g * 0.1192f + b * 0.9505f) / 1.08883f);
  return v3(116.0f * y - 16.0f, 500.0f * (x - y), 200.0f * (y - z));
}
// CIELab -> sRGB (0..255), vector_math.cuh:543-564 (two rows carry a double literal
// in the reference and are therefore evaluated in double)
__device__ __forceinline__ float lab_finv(float t) {
  const float t3 = powf(t, 3.0f);
  return (t3 > 0.008856f) ? t3 : (t - 16.0f / 116.0f) / 7.787f;
}
__device__ __forceinline__ float linear_to_srgb(float c) {
  return (c > 0.0031308f) ? (1.055f * powf(c, 1.0f / 2.4f) - 0.055f) : 12.92f * c;
}
__device__ __forceinline__ V3 lab_to_rgb(V3 c) {
  float y = (c.x + 16.0f) / 116.0f;
  float x = c.y / 500.0f + y;
  float z = y - c.z / 200.0f;
  x = 0.95047f * lab_finv(x);
  y = 1.0f * lab_finv(y);
  z = 1.08883f * lab_finv(z);
  float r = x * 3.2406f - y * 1.5372f - z * 0.4986f;
  float g = (float)((double)(-x * 0.9689f) + (double)y * 1.8758 + (double)(z * 0.0415f));
  float b = (float)((double)(x * 0.0557f - y * 0.2040f) + (double)z * 1.0570);
  r = linear_to_srgb(r); g = linear_to_srgb(g); b = linear_to_srgb(b);
  return v3(fmaxf(0.0f, fminf(1.0f, r)) * 255.0f, fmaxf(0.0f, fminf(1.0f, g)) * 255.0f,
            fmaxf(0.0f, fminf(1.0f, b)) * 255.0f);
}

// Dominant eigenvector by repeated squaring of the trace-normalised matrix
// (supersurfel_fusion_kernels.cu:48-111).  dominant_row() is the "row that holds the
// largest entry" rule with the reference's tie order.
__device__ __forceinline__ V3 dominant_row(const Sym3& A) {
  const float vmax = fmaxf(fmaxf(fmaxf(fmaxf(fmaxf(A.xx, A.xy), A.xz), A.yy), A.yz), A.zz);
  if (A.xx == vmax || A.xy == vmax || A.xz == vmax) return normalize(v3(A.xx, A.xy, A.xz));
  if (A.yy == vmax || A.yz == vmax) return normalize(v3(A.xy, A.yy, A.yz));
  return normalize(v3(A.xz, A.yz, A.zz));
}
__device__ __forceinline__ float eigenvalue_along(const Sym3& A, V3 e) {
  const float emax = fmaxf(fmaxf(e.x, e.y), e.z);   // largest SIGNED component (appendix B6)
  if (e.x == emax) return (A.xx * e.x + A.xy * e.y + A.xz * e.z) / e.x;
  if (e.y == emax) return (A.xy * e.x + A.yy * e.y + A.yz * e.z) / e.y;
  return (A.xz * e.x + A.yz * e.y + A.zz * e.z) / e.z;
}
__device__ __forceinline__ void eigenframe(const Sym3& A, M3& vecs, V3& vals) {
  Sym3 P = A / trace(A);
  Sym3 Q = sym3(1.f - P.xx, -P.xy, -P.xz, 1.f - P.yy, -P.yz, 1.f - P.zz);
#pragma unroll 1
  for (int i = 0; i < 10; ++i) {
    P = square(P); P = P / trace(P);
    Q = square(Q); Q = Q / trace(Q);
  }
  vecs.r0 = dominant_row(P);
  vecs.r2 = dominant_row(Q);
  vecs.r1 = cross(vecs.r2, vecs.r0);
  vals = v3(eigenvalue_along(A, vecs.r0), eigenvalue_along(A, vecs.r1), eigenvalue_along(A, vecs.r2));
}

// Projection rounding: lroundf semantics (half away from zero) with non-finite /
// huge values mapped out of the image (they can never pass the distance gates).
__device__ __forceinline__ int round_px(float v) {
  if (!(fabsf(v) < 1.0e9f)) return -1000000000;
  return (int)lroundf(v);
}
// Point-sampled clamp addressing of the reference's textures (texture_impl.hpp:30-49)
__device__ __forceinline__ int tex_coord(float c, int n) {
  if (!(c >= 0.0f)) return 0;
  if (c >= (float)n) return n - 1;
  return (int)floorf(c);
}

// Order-free fixed-point accumulation (2^-32 for metric moments, 2^-30 for disparity)
__device__ __forceinline__ long long quantize(float v, double scale, double clampv) {
  double s = (double)v * scale;
  if (s != s) s = 0.0;
  s = fmin(fmax(s, -clampv), clampv);
  return __double2ll_rn(s);
}

// Warp-level pre-reduction of order-free integer sums keyed by a label: the lanes of a warp
// (32 consecutive pixels of a row) are cut into runs of equal labels, every run is summed
// with shuffles, and only the first lane of a run issues the atomics.  All 32 lanes must
// call these; lanes with nothing to add pass zeros.
__device__ __forceinline__ int run_head_lane(int label) {
  const int lane = threadIdx.x & 31;
  const int prev = __shfl_up_sync(0xffffffffu, label, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != label);
  return 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
}
template <int N>
__device__ __forceinline__ void run_reduce(long long (&v)[N], int head) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int oh = __shfl_down_sync(0xffffffffu, head, o);
    const bool take = (lane + o < 32) && (oh == head);
#pragma unroll
    for (int k = 0; k < N; k++) {
      const long long vv = __shfl_down_sync(0xffffffffu, v[k], o);
      if (take) v[k] += vv;
    }
  }
}

}  // namespace ssf
