// TPS RGB-D superpixel segmentation for sm_100a: grid seeding, topology-preserving
// boundary relabelling, RANSAC slanted-plane initialisation, plane smoothing and
// slanted-depth rendering.
//
// Replaces TPS_RGBD::compute / filter / computeDepthImage
// (reference: core/src/TPS_RGBD.cu:101-525) and its kernels
// (core/src/TPS_RGBD_kernels.cu:27-614,
//  core/include/supersurfel_fusion/TPS_RGBD_kernels.cuh:178-651).
//
// Mechanism:
//  * the reference relabels with one thread per active pixel on a per-CTA
//    shared-memory snapshot, which races across CTA seams and between the two
//    horizontally adjacent active pixels of a pass.  Here ONE thread owns each
//    adjacent pair: every label, boundary count and inlier flag it reads is either
//    inactive in this pass or its own, so a pass is exactly "decide on the
//    pass-start state, then apply" with no barrier at all;
//  * all running sums are 64-bit integers (exact pixel coordinates / colours, and
//    2^-30 fixed-point disparity), so RED.ADD.64 accumulation is order-free and the
//    label map is reproducible bit for bit (oracle/oracle_tps.cpp states the same
//    serialisation);
//  * seeding is one CTA per grid cell with an in-CTA reduction (no atomics), the
//    plane-smoothing filter is a single double-buffered (Jacobi) CTA, and the depth
//    render writes the interleaved (label, depth) map the later stages gather from.
// Compiled with -fmad=false: energies are compared for strict inequality, so every
// product and sum must round as written.
#include "ssf_engine.h"
#include "ssf_math.cuh"

#include <cooperative_groups.h>
#include <cuda.h>            // CUtensorMap (types only; the encoder is resolved at run time)
#include <curand_kernel.h>

namespace cg = cooperative_groups;

namespace ssf {

constexpr double kDispFix = 1073741824.0;      // 2^30
constexpr double kDispClamp = 137438953472.0;  // 2^37

struct TpsArgs {
  int W, H, cell, gx, gy, S;
  int raw_w, raw_h;        // thread extents of the reference launch (TPS_RGBD.cu:185-186)
  int min_size;
  int debug;               // profiling knob, 0 in production
  int pdl_late;            // SSF_PDL 2 / 3: the fused pass lets its successor launch only once its decisions are made
  int pdl_trig;            // where: 0 after the decisions (default), 1 at the top, 2 once the pass's loads are in flight, 3 after the apply phase (A/B)
  float lambda_pos, lambda_bound, lambda_size, lambda_disp, thresh_disp;
  uchar4* rgba;
  float* disp;
  int* labels;
  int* bound;
  unsigned char* inliers;
  Superpixel* sp;
  SpSums* sums;            // the sums every kernel but the fused pass works on (= sums_cur of the next pass)
  // fused pass (tps_pass_tile_kernel): three rotating sum buffers, see there
  SpSums* sums_nxt;
  SpSums* sums_zero;
  unsigned long long gx_magic;   // ceil(2^32 / gx): k / gx == (k * gx_magic) >> 32 for every superpixel id k
  unsigned long long cell_magic; // ceil(2^32 / cell), the same for pixel coordinates
  long long* trace;        // profiling aid (SSF_TPS_TRACE): per-CTA clock64 stamps of the fused pass, else NULL
};

static TpsArgs tps_args(const Engine* e) {
  TpsArgs a;
  a.W = e->W; a.H = e->H; a.cell = e->cfg.cell_size; a.gx = e->gx; a.gy = e->gy; a.S = e->S;
  a.raw_w = 16 * ((e->W / 2 + 15) / 16);
  a.raw_h = 16 * ((e->H / 2 + 15) / 16);
  a.min_size = (int)((float)(e->cfg.cell_size * e->cfg.cell_size) / 4.f);
  a.debug = 0;
  if (const char* v = getenv("SSF_TPS_DEBUG")) a.debug = atoi(v);
  a.lambda_pos = e->cfg.lambda_pos; a.lambda_bound = e->cfg.lambda_bound; a.lambda_size = e->cfg.lambda_size;
  a.lambda_disp = e->cfg.lambda_disp; a.thresh_disp = e->cfg.thresh_disp;
  a.rgba = e->rgba; a.disp = e->disp; a.labels = e->labels; a.bound = e->bound; a.inliers = e->inliers;
  a.sp = e->sp; a.sums = e->sums;
  a.sums_nxt = a.sums_zero = nullptr;
  a.gx_magic = (0x100000000ull + (unsigned)e->gx - 1) / (unsigned)e->gx;
  a.cell_magic = (0x100000000ull + (unsigned)e->cfg.cell_size - 1) / (unsigned)e->cfg.cell_size;
  a.trace = reinterpret_cast<long long*>(e->tps_trace);
  a.pdl_late = e->pdl_now >= 2;
  a.pdl_trig = e->pdl_now == 4 ? 1 : e->pdl_now == 5 ? 2 : e->pdl_now == 6 ? 3 : 0;
  return a;
}

// The three rotating per-superpixel sum buffers of a frame slot: before pass p (p = passes already
// run on this frame) the true sums are in buffer p % 3, buffer (p + 1) % 3 is all zero.
static SpSums* sums_buffer(const Engine* e, int p) { return e->sums + (size_t)(p % 3) * e->S; }

__device__ __forceinline__ void add64(long long* p, long long v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}

// ---- seeding: cvtColor + depth2disp + initSuperpixelsRGBD + first merge ------------
// (TPS_RGBD.cu:126-180; TPS_RGBD_kernels.cu:61-110, 224-242, 278-296).  One CTA per
// grid cell; the cell's sums come from an in-CTA reduction.
__global__ void __launch_bounds__(256) tps_seed_kernel(TpsArgs a, const uint8_t* __restrict__ rgb, size_t rgb_stride,
                                                       const float* __restrict__ depth, size_t depth_stride) {
  pdl_sync();
  const int cellx = blockIdx.x % a.gx, celly = blockIdx.x / a.gx;
  const int index = blockIdx.x;
  const int x0 = cellx * a.cell, y0 = celly * a.cell;
  int sx = 0, sy = 0, sr = 0, sg = 0, sb = 0, sn = 0;
  for (int i = threadIdx.x; i < a.cell * a.cell; i += blockDim.x) {
    const int x = x0 + i % a.cell, y = y0 + i / a.cell;
    if (x >= a.W || y >= a.H) continue;
    const size_t p = (size_t)y * a.W + x;
    const uint8_t* src = rgb + (size_t)y * rgb_stride + 3 * (size_t)x;
    const uchar4 c = make_uchar4(src[0], src[1], src[2], 255);
    a.rgba[p] = c;
    const float d = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(depth) + (size_t)y * depth_stride +
                                                    4 * (size_t)x);
    a.disp[p] = 1.f / d;
    a.labels[p] = index;
    a.bound[p] = (((x + 1) % a.cell <= 1) ? 1 : 0) + (((y + 1) % a.cell <= 1) ? 1 : 0);
    sx += x; sy += y; sr += c.x; sg += c.y; sb += c.z; sn += 1;
  }
  __shared__ int red[8][6];
  int v[6] = {sx, sy, sr, sg, sb, sn};
#pragma unroll
  for (int k = 0; k < 6; k++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 6; k++) red[threadIdx.x >> 5][k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t[6] = {0, 0, 0, 0, 0, 0};
    for (int w = 0; w < (int)(blockDim.x >> 5); w++)
      for (int k = 0; k < 6; k++) t[k] += red[w][k];
    SpSums s;
    s.x = t[0]; s.y = t[1]; s.r = t[2]; s.g = t[3]; s.b = t[4]; s.n = t[5];
    s.dx = s.dy = s.dxx = s.dyy = s.dxy = s.dn = s.dxd = s.dyd = s.dd = 0; s.pad = 0;
    a.sums[index] = s;
    if (a.sums_nxt) {        // fused passes: the buffer the first pass accumulates into starts at zero
      SpSums z;
      z.x = z.y = z.r = z.g = z.b = z.n = z.dx = z.dy = z.dxx = z.dyy = z.dxy = z.dn = z.dxd = z.dyd = z.dd = z.pad = 0;
      a.sums_nxt[index] = z;
    }
    const float n = (float)s.n;
    Superpixel q;
    q.xy_rg = make_float4((float)s.x / n, (float)s.y / n, (float)s.r / n, (float)s.g / n);
    q.theta_b = make_float4(0.f, 0.f, 0.f, (float)s.b / n);
    q.size = make_float4(n, 0.f, 0.f, 0.f);
    a.sp[index] = q;
  }
}

// ---- sums -> means (+ plane)  (TPS_RGBD_kernels.cu:224-276, 27-59) -------------------
__device__ __forceinline__ bool solve_plane(float& tx, float& ty, float& tz, const float x1, const float y1,
                                            const float z1, const float d1, const float x2, const float y2,
                                            const float z2, const float d2, const float x3, const float y3,
                                            const float z3, const float d3) {
  const float eps = 1e-20;
  const float denA = (x1 * z2 - x2 * z1) * (y2 * z3 - y3 * z2) - (x2 * z3 - x3 * z2) * (y1 * z2 - y2 * z1);
  if (!isfinite(denA) && denA < eps) return false;
  tx = ((z2 * d1 - z1 * d2) * (y2 * z3 - y3 * z2) - (z3 * d2 - z2 * d3) * (y1 * z2 - y2 * z1)) / denA;
  float denB = y1 * z2 - y2 * z1;
  if (denB > eps) {
    ty = (z2 * d1 - z1 * d2 - tx * (x1 * z2 - x2 * z1)) / denB;
  } else {
    denB = y2 * z3 - y3 * z2;
    ty = (z3 * d2 - z2 * d3 - tx * (x2 * z3 - x3 * z2)) / denB;
  }
  if (z1 > eps) tz = (d1 - tx * x1 - ty * y1) / z1;
  else if (z2 > eps) tz = (d2 - tx * x2 - ty * y2) / z2;
  else tz = (d3 - tx * x3 - ty * y3) / z3;
  return true;
}

// means (and, with DISP, the least-squares disparity plane) of one superpixel from its sums
template <bool DISP>
__device__ __forceinline__ Superpixel tps_superpixel_from_sums(const SpSums& c) {
  const float n = (float)c.n;
  Superpixel s;
  s.xy_rg = make_float4((float)c.x / n, (float)c.y / n, (float)c.r / n, (float)c.g / n);
  s.size = make_float4(n, 0.f, 0.f, 0.f);
  const float mb = (float)c.b / n;
  if (DISP) {
    const float dx = (float)c.dx, dy = (float)c.dy, dxx = (float)c.dxx, dyy = (float)c.dyy, dxy = (float)c.dxy,
                dn = (float)c.dn;
    const float dxd = (float)((double)c.dxd * (1.0 / kDispFix));
    const float dyd = (float)((double)c.dyd * (1.0 / kDispFix));
    const float dd = (float)((double)c.dd * (1.0 / kDispFix));
    float tx = 0.f, ty = 0.f, tz = 0.f;
    if (!solve_plane(tx, ty, tz, dxx, dxy, dx, dxd, dxy, dyy, dy, dyd, dx, dy, dn, dd)) {
      tx = 0.f; ty = 0.f; tz = __int_as_float(0xFFE00000);
    }
    s.theta_b = make_float4(tx, ty, tz, mb);
  } else {
    s.theta_b = make_float4(0.f, 0.f, 0.f, mb);   // the plane is unused (and zero) in the colour-only phase
  }
  return s;
}

template <bool DISP>
__device__ __forceinline__ void tps_merge_item(const TpsArgs& a, int k) {
  const Superpixel n = tps_superpixel_from_sums<DISP>(a.sums[k]);
  Superpixel& s = a.sp[k];
  s.xy_rg = n.xy_rg;
  s.size.x = n.size.x;
  if (DISP) s.theta_b = n.theta_b;
  else s.theta_b.w = n.theta_b.w;
}

// Where a pass reads superpixel means from.  Global: the array the merge phase wrote.
struct SpGlobal {
  const Superpixel* sp;
  __device__ __forceinline__ Superpixel get(int k) const { return sp[k]; }
};
// Band cache: the CTA recomputed, from the sums, the superpixels seeded in the grid-cell
// rows around its band of image rows into shared memory (anything
// outside the window is recomputed on the fly; the window is sized so that this cannot
// happen, see the margin in tps_persistent_kernel).
template <bool DISP>
struct SpCached {
  const Superpixel* cache;
  const SpSums* sums;
  int first, count;   // cached superpixel ids [first, first + count)
  __device__ __forceinline__ Superpixel get(int k) const {
    const int slot = k - first;
    if (slot >= 0 && slot < count) return cache[slot];
    return tps_superpixel_from_sums<DISP>(sums[k]);
  }
};

template <bool DISP>
__global__ void tps_merge_kernel(TpsArgs a) {
  pdl_sync();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < a.S) tps_merge_item<DISP>(a, k);
}

// ---- one relabelling pass (TPS_RGBD_kernels.cuh:235-651) ---------------------------
struct Decision {
  int index, new_index, b;
  unsigned char inlier, prev_inlier;
};

// Per-pixel inputs of a decision, fetched together with the label window so that all the
// global loads of a pass are in flight at once (one dependent L2 round trip, not three).
struct PixelIn {
  int bounds;
  uchar4 col;
  float disp;
  unsigned char inlier;
};
template <bool DISP>
__device__ __forceinline__ PixelIn tps_fetch_pixel(const TpsArgs& a, size_t p) {
  PixelIn in;
  in.bounds = a.bound[p];
  in.col = a.rgba[p];
  in.disp = DISP ? a.disp[p] : 0.f;
  in.inlier = DISP ? a.inliers[p] : (unsigned char)0;
  return in;
}

template <bool DISP, typename SP>
__device__ __forceinline__ void tps_decide(const TpsArgs& a, const SP& spsrc, const PixelIn& in, int x, int y,
                                           const int (&L)[3][4], int c, Decision& d, float& disp_v) {
  const int bounds = in.bounds;
  const int index = L[1][c];
  int new_index = index;
  const Superpixel prev = spsrc.get(index);
  unsigned char inlier = 0xff, prev_inlier = 0;
  float disp_energy = 0.f;
  disp_v = 0.f;
  if (DISP) {
    disp_v = in.disp;
    prev_inlier = in.inlier;
    const float dp = prev.theta_b.x * (float)x + prev.theta_b.y * (float)y + prev.theta_b.z;
    disp_energy = (dp - disp_v) * (dp - disp_v);
    if (!isfinite(disp_energy) || disp_energy > a.thresh_disp || dp < 0.f) {
      disp_energy = a.thresh_disp;
      inlier = 0;
    }
  }
  int newb = 0;
  bool movable = false;
  if (bounds) {
    // isUnchangeable: transitions along the open 8-ring TL,T,TR,R,BR,B,BL,L
    const int ring[8] = {L[0][c - 1], L[0][c], L[0][c + 1], L[1][c + 1], L[2][c + 1], L[2][c], L[2][c - 1], L[1][c - 1]};
    int jump = 0;
    bool prevb = (ring[0] == index);
#pragma unroll
    for (int k = 1; k < 8; k++) {
      const bool cur = (ring[k] == index);
      if (prevb != cur) { jump++; prevb = cur; }
    }
    movable = !(jump > 2);
  }
  if (movable) {
    const uchar4 col = in.col;
    const float cr = (float)col.x, cg = (float)col.y, cb = (float)col.z;
    const float px = (float)x, py = (float)y;
    const float size = prev.size.x;
    const float s = size / (size - 1.f);
    const float dpx = s * (px - prev.xy_rg.x), dpy = s * (py - prev.xy_rg.y);
    const float dcx = s * (cr - prev.xy_rg.z), dcy = s * (cg - prev.xy_rg.w), dcz = s * (cb - prev.theta_b.w);
    const float dsize = size - (float)a.min_size;
    float best = (dcx * dcx + dcy * dcy + dcz * dcz) + a.lambda_pos * (dpx * dpx + dpy * dpy);
    if (DISP) best = best + a.lambda_disp * disp_energy;
    best = best - a.lambda_size * fminf(dsize, 0.f);
    best = best + a.lambda_bound * (float)bounds;
    const int nl[4] = {L[0][c], L[1][c - 1], L[1][c + 1], L[2][c]};  // up, left, right, down
    // the four candidates are evaluated as independent straight-line chains (so that they
    // overlap in the pipeline) and then compared in the reference's order: up, left, right, down
    float cand_e[4];
    unsigned char cand_in[4];
    bool cand_ok[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int i_n = nl[k];
      cand_ok[k] = !(i_n == -1 || i_n == index);
      const Superpixel ns = spsrc.get(cand_ok[k] ? i_n : index);
      const float ex = px - ns.xy_rg.x, ey = py - ns.xy_rg.y;
      const float fx = cr - ns.xy_rg.z, fy = cg - ns.xy_rg.w, fz = cb - ns.theta_b.w;
      const float nsize = ns.size.x + 1.f - (float)a.min_size;
      float n_energy = 0.f;
      unsigned char n_inlier = 0xff;
      if (DISP) {
        const float dp = ns.theta_b.x * (float)x + ns.theta_b.y * (float)y + ns.theta_b.z;
        n_energy = (dp - disp_v) * (dp - disp_v);
        const bool out = !isfinite(n_energy) || n_energy > a.thresh_disp || dp < 0.f;
        n_energy = out ? a.thresh_disp : n_energy;
        n_inlier = out ? 0 : 0xff;
      }
      int b = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) b += (nl[q] != i_n);
      float energy = (fx * fx + fy * fy + fz * fz) + a.lambda_pos * (ex * ex + ey * ey);
      if (DISP) energy = energy + a.lambda_disp * n_energy;
      energy = energy - a.lambda_size * fminf(nsize, 0.f);
      energy = energy + a.lambda_bound * (float)b;
      cand_e[k] = energy;
      cand_in[k] = n_inlier;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (cand_ok[k] && cand_e[k] < best) {
        best = cand_e[k];
        new_index = nl[k];
        if (DISP) inlier = cand_in[k];
      }
    }
    if (new_index != index) {
#pragma unroll
      for (int q = 0; q < 4; q++) newb += (nl[q] != new_index);
    }
  }
  d.index = index; d.new_index = new_index; d.b = newb; d.inlier = inlier; d.prev_inlier = prev_inlier;
}

template <bool DISP, typename SP>
__device__ __forceinline__ void tps_pass_item(const TpsArgs& a, const SP& spsrc, int q, int ry, int OX, int OY) {
  const int y = 2 * ry + OY;
  if (ry >= a.raw_h || y >= a.H || 32 * (ry / 16) + OY >= a.H) return;
  const int rx0 = OX ? 2 * q : 2 * q - 1;
  int xs[2];
  bool ok[2];
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int rx = rx0 + j;
    xs[j] = 2 * rx + ((rx + OX) & 1);
    ok[j] = rx >= 0 && rx < a.raw_w && 32 * (rx / 16) < a.W && xs[j] < a.W;
  }
  if (!ok[0] && !ok[1]) return;
  // 3 x 4 label window around the pair (column 1 = first pixel, column 2 = second)
  const int cb = 4 * q + (OX ? 0 : -2);
  int L[3][4];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const int yy = y - 1 + r;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int xx = cb + c;
      L[r][c] = (xx >= 0 && xx < a.W && yy >= 0 && yy < a.H) ? a.labels[(size_t)yy * a.W + xx] : -1;
    }
  }
  Decision d[2];
  float dv[2];
  PixelIn pin[2];
#pragma unroll
  for (int j = 0; j < 2; j++)
    if (ok[j]) pin[j] = tps_fetch_pixel<DISP>(a, (size_t)y * a.W + xs[j]);
#pragma unroll
  for (int j = 0; j < 2; j++)
    if (ok[j]) tps_decide<DISP>(a, spsrc, pin[j], xs[j], y, L, 1 + j, d[j], dv[j]);

  // apply (reads above are all pass-start values: only this thread writes them)
  int partner_delta[2] = {0, 0};
#pragma unroll
  for (int j = 0; j < 2; j++) {
    if (!ok[j]) continue;
    const int x = xs[j], c = 1 + j;
    const size_t p = (size_t)y * a.W + x;
    const bool moved = d[j].new_index != d[j].index;
    if (moved) {
      const int nxs[4] = {0, -1, 1, 0}, nys[4] = {-1, 0, 0, 1};
      const int nl[4] = {L[0][c], L[1][c - 1], L[1][c + 1], L[2][c]};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int i_n = nl[k];
        int delta = 0;
        if (i_n == d[j].new_index) delta = -1;
        else if (i_n == d[j].index) delta = 1;
        if (delta == 0 || i_n == -1) continue;
        const bool is_partner = (j == 0 && k == 2 && ok[1]) || (j == 1 && k == 1 && ok[0]);
        if (is_partner) partner_delta[1 - j] += delta;
        else atomicAdd(&a.bound[(size_t)(y + nys[k]) * a.W + (x + nxs[k])], delta);
      }
      const uchar4 col = pin[j].col;
      SpSums* o = &a.sums[d[j].index];
      SpSums* n = &a.sums[d[j].new_index];
      add64(&o->x, -x); add64(&o->y, -y); add64(&o->r, -(int)col.x); add64(&o->g, -(int)col.y);
      add64(&o->b, -(int)col.z); add64(&o->n, -1);
      add64(&n->x, x); add64(&n->y, y); add64(&n->r, col.x); add64(&n->g, col.y); add64(&n->b, col.z);
      add64(&n->n, 1);
      a.labels[p] = d[j].new_index;
    }
    if (DISP) {
      const unsigned char inl = d[j].inlier, pin = d[j].prev_inlier;
      if (inl && (!pin || moved)) {
        SpSums* s = &a.sums[d[j].new_index];
        const long long qd = quantize(dv[j], kDispFix, kDispClamp);
        add64(&s->dx, x); add64(&s->dy, y); add64(&s->dxx, (long long)x * x); add64(&s->dyy, (long long)y * y);
        add64(&s->dxy, (long long)x * y); add64(&s->dxd, (long long)x * qd); add64(&s->dyd, (long long)y * qd);
        add64(&s->dd, qd); add64(&s->dn, 1);
      }
      if (pin && (!inl || moved)) {
        SpSums* s = &a.sums[d[j].index];
        const long long qd = quantize(dv[j], kDispFix, kDispClamp);
        add64(&s->dx, -x); add64(&s->dy, -y); add64(&s->dxx, -(long long)x * x); add64(&s->dyy, -(long long)y * y);
        add64(&s->dxy, -(long long)x * y); add64(&s->dxd, -(long long)x * qd); add64(&s->dyd, -(long long)y * qd);
        add64(&s->dd, -qd); add64(&s->dn, -1);
      }
      if (inl != pin) a.inliers[p] = inl;
    }
  }
#pragma unroll
  for (int j = 0; j < 2; j++) {
    if (!ok[j]) continue;
    const size_t p = (size_t)y * a.W + xs[j];
    if (d[j].new_index != d[j].index) a.bound[p] = d[j].b;            // own "= b" wins
    else if (partner_delta[j] != 0) a.bound[p] += partner_delta[j];   // only this thread touches it
  }
}

template <bool DISP>
__global__ void __launch_bounds__(128) tps_pass_kernel(TpsArgs a, int OX, int OY) {
  pdl_sync();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;   // pair index along the row
  const int ry = blockIdx.y * blockDim.y + threadIdx.y;
  const SpGlobal src = {a.sp};
  tps_pass_item<DISP>(a, src, q, ry, OX, OY);
}

// ---- one relabelling pass, merge fused in: no separate "sums -> means" launch ----------------
// The multi-kernel form above needs a merge launch between two passes because a pass both reads
// the per-superpixel means and changes the sums they come from.  Here a pass reads the sums of
// the PREVIOUS pass from a quiescent buffer and accumulates its own changes into another one:
//   before pass p:  cur = buf[p % 3] holds the true sums, nxt = buf[(p+1) % 3] is all zero;
//   pass p:         every CTA derives the means it needs from cur (shared-memory window around
//                   its tile; identical arithmetic to the merge kernel, so identical bits);
//                   the CTA that owns superpixel s adds cur[s] into nxt[s] and clears
//                   zero[s] = buf[(p+2) % 3][s] (nobody reads that buffer during pass p);
//                   relabelled pixels add their +-deltas into nxt;
//   after pass p:   nxt = cur + deltas = true sums.  All sums are integers: order-free, exact.
// One thread per ACTIVE PIXEL (the pair-per-thread form halves the threads and doubles the
// dependent chain); the two pixels of an adjacent pair are neighbouring lanes and exchange the
// only cross-pixel quantity (the boundary-count delta each owes the other) with one shuffle.
// The 3-row label neighbourhood of the tile is staged in shared memory with coalesced 8-byte
// loads while the window means are being computed: one L2 round trip per pass.
constexpr int TILE_LANES = 32;                 // active pixels per tile row (64 image columns)
constexpr int TILE_ROWS = 8;                   // active rows per tile (16 image rows)
constexpr int TILE_THREADS = TILE_LANES * TILE_ROWS;
constexpr int TILE_COLS = 2 * TILE_LANES;      // image columns of a tile
constexpr int TILE_LROWS = 2 * TILE_ROWS + 1;  // label rows staged: y0 - 1 .. y0 + 2 * TILE_ROWS - 1
constexpr int TILE_SCOLS = TILE_COLS + 4;      // staged columns: a TMA box must start on a 16-byte boundary (x % 4 == 0)
constexpr int TILE_WIN = 64;                   // capacity of the means window (superpixels)

struct SpWindow {
  const Superpixel* cache;
  const SpSums* cur;
  int wx0, wy0, ww, wh;      // window of grid cells [wx0, wx0 + ww) x [wy0, wy0 + wh)
  int gx;
  unsigned long long gx_magic;
};
template <bool DISP>
struct SpTile {
  SpWindow w;
  __device__ __forceinline__ Superpixel get(int k) const {
    const int cy = (int)(((unsigned long long)(unsigned)k * w.gx_magic) >> 32);
    const int cx = k - cy * w.gx;
    const int ux = cx - w.wx0, uy = cy - w.wy0;
    if ((unsigned)ux < (unsigned)w.ww && (unsigned)uy < (unsigned)w.wh) return w.cache[uy * w.ww + ux];
    return tps_superpixel_from_sums<DISP>(w.cur[k]);      // wandered out of the window: exact, just slower
  }
};

// a / d for 0 <= a < 2^16 by the multiplier ceil(2^32 / d) (exact while a * d < 2^32)
__device__ __forceinline__ int div_magic(int a, unsigned long long magic) {
  return (int)(((unsigned long long)(unsigned)a * magic) >> 32);
}

// TMA: the label rows of the tile arrive as ONE 2-D tensor-map copy (cp.async.bulk.tensor, box 68 x 17
// int32) issued by one thread and counted on an mbarrier, instead of three 8-byte loads + stores per
// thread; out-of-image texels come back as 0 and are patched to -1 by the (border) tiles that have any.
// The innermost box coordinate must be 16-byte aligned (measured: x = -2 raises an illegal-instruction
// fault), so the box starts at the multiple of four at or below the tile's first column (the tile starts at
// 4 q0 - 2 in the OX = 0 passes) and is four columns wider than the tile; `co` is where the tile begins
// inside the staged rows.  Needs a 16-byte row pitch (W % 4 == 0), otherwise the loads below do the job.
// MINB = resident CTAs per SM the kernel is compiled for (SSF_TPS_OCC): 4 caps the colour + disparity variant at
// 64 registers (a few spilled words) so that the passes of more frames in flight fit on an SM side by side.
template <bool DISP, bool TMA, int MINB>
__global__ void __launch_bounds__(TILE_THREADS, MINB) tps_pass_tile_kernel(TpsArgs a, int OX, int OY,
                                                                     const __grid_constant__ CUtensorMap label_map) {
  // SSF_PDL=2: the successor may launch once this grid has decided (below), and a grid waits for its
  // predecessor only here, after its CTAs are resident
  if (!a.pdl_late) {
    pdl_trigger();
    pdl_wait();
  } else if (a.pdl_trig == 1) {
    pdl_trigger();
  }
  __shared__ __align__(128) int lab[TILE_LROWS][TILE_SCOLS];
  __shared__ __align__(8) uint64_t lab_bar;
  __shared__ Superpixel win[TILE_WIN];
  const int tid = threadIdx.x;
  const int lane = tid & 31, wrp = tid >> 5;
  const SpSums* cur = a.sums;
  long long* tr = (a.trace && tid == 0) ? a.trace + 8 * (size_t)(blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  if (tr) tr[0] = clock64();

  // The kernel is a chain of dependent latencies (L2 round trip -> means -> decisions -> atomics), so
  // it is written to have every global load of the pass in flight before anything waits: first the
  // sums behind the means window, then the label rows, then the pixel's own inputs.

  // ---- geometry of the tile
  const int q0 = blockIdx.x * (TILE_LANES / 2);          // first pair index of the tile
  const int xs0 = 4 * q0 + (OX ? 0 : -2);                // first image column of the staged labels
  const int ry0 = blockIdx.y * TILE_ROWS;
  const int y0 = 2 * ry0 + OY;                           // first active row; staged rows start at y0 - 1
  const int co = OX ? 0 : 2;                             // column of the staged rows that holds image column xs0

  auto stage_labels = [&]() {
    if (TMA && tid == 0) {
      mbar_init(&lab_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&lab_bar, TILE_LROWS * TILE_SCOLS * (int)sizeof(int));
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
              smem_u32(&lab[0][0])),
          "l"(&label_map), "r"(xs0 - co), "r"(y0 - 1), "r"(smem_u32(&lab_bar))
          : "memory");
    }
  };
  if (!a.pdl_late) stage_labels();

  // ---- window of superpixels whose means this tile may need: the grid cells under the tile plus a
  // margin of one cell (boundaries drift a few pixels over the 4 * seg_iter passes); with small
  // cells the margin, then the window itself, is cut to the capacity -- a label outside the window
  // takes the exact fallback in SpTile::get
  SpWindow w;
  {
    const int px0 = max(0, xs0), px1 = min(a.W - 1, xs0 + TILE_COLS - 1);
    const int py0 = max(0, y0 - 1), py1 = min(a.H - 1, y0 - 1 + TILE_LROWS - 1);
    const int cx0 = div_magic(px0, a.cell_magic), cx1 = div_magic(px1, a.cell_magic);
    const int cy0 = div_magic(py0, a.cell_magic), cy1 = div_magic(py1, a.cell_magic);
#pragma unroll
    for (int m = 1; m >= 0; m--) {
      w.wx0 = max(0, cx0 - m);
      w.wy0 = max(0, cy0 - m);
      w.ww = max(0, min(a.gx - 1, cx1 + m) - w.wx0 + 1);
      w.wh = max(0, min(a.gy - 1, cy1 + m) - w.wy0 + 1);
      if (w.ww * w.wh <= TILE_WIN) break;
    }
    if (w.ww * w.wh > TILE_WIN) {
      w.wh = min(w.wh, 8);
      w.ww = min(w.ww, TILE_WIN / w.wh);
    }
    w.cache = win; w.cur = cur; w.gx = a.gx; w.gx_magic = a.gx_magic;
  }
  // thread i < n computes the means of window slot i (with DISP, thread TILE_WIN + i its plane)
  const int nwin = w.ww * w.wh;
  const int slot = tid & (TILE_WIN - 1);
  const bool does_mean = tid < TILE_WIN && slot < nwin;
  const bool does_plane = DISP && tid >= TILE_WIN && tid < 2 * TILE_WIN && slot < nwin;
  int wk = 0;
  if (does_mean || does_plane) {
    const int wr = (int)(((float)slot + 0.5f) * __frcp_rn((float)w.ww));     // slot / ww, exact for these sizes
    wk = (w.wy0 + wr) * a.gx + (w.wx0 + (slot - wr * w.ww));
  }
  // late-trigger mode: everything above is arithmetic on kernel arguments and ran while the previous
  // pass was still applying its decisions; its labels and sums are read from here on
  if (a.pdl_late) {
    pdl_wait();
    stage_labels();
  }
  // 16-byte loads of the half of the record this thread needs
  longlong2 sv[5];
#pragma unroll
  for (int k = 0; k < 5; k++) sv[k] = make_longlong2(0, 0);
  {
    const longlong2* rec = reinterpret_cast<const longlong2*>(cur + wk);
    if (does_mean) {                       // x y | r g | b n
#pragma unroll
      for (int k = 0; k < 3; k++) sv[k] = rec[k];
    } else if (does_plane) {               // dx dy | dxx dyy | dxy dn | dxd dyd | dd pad
#pragma unroll
      for (int k = 0; k < 5; k++) sv[k] = rec[3 + k];
    }
  }

  // ---- buffer rotation, first half: the owner's words of `cur` are fetched now by the warps that compute no
  // means (their L2 round trip overlaps everything else); carried into `nxt` at the very end
  constexpr int ROT_FIRST = TILE_THREADS - 64;             // threads 192 .. 255 hold one word each per round
  const int rot_nb = gridDim.x * gridDim.y;
  const int rot_per = (a.S + rot_nb - 1) / rot_nb;
  const int rot_s0 = (blockIdx.y * gridDim.x + blockIdx.x) * rot_per;
  const int rot_words = max(0, min(a.S, rot_s0 + rot_per) - rot_s0) * 16;
  long long rot_v = 0;
  const bool rot_mine = tid >= ROT_FIRST && tid - ROT_FIRST < rot_words;
  if (rot_mine) rot_v = reinterpret_cast<const long long*>(cur)[(size_t)rot_s0 * 16 + (tid - ROT_FIRST)];

  // ---- label rows (coalesced 8-byte loads; xs0 is even and W need not be): all in flight, stored later
  int2 lv[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int r = wrp + k * TILE_ROWS;
    const int yy = y0 - 1 + r;
    const int xx = xs0 + 2 * lane;
    int2 v = make_int2(-1, -1);
    if (!TMA && r < TILE_LROWS && yy >= 0 && yy < a.H) {
      const size_t row = (size_t)yy * a.W;
      if (xx >= 0 && xx + 1 < a.W && (((row + xx) & 1) == 0)) {
        v = *reinterpret_cast<const int2*>(a.labels + row + xx);
      } else {
        if (xx >= 0 && xx < a.W) v.x = a.labels[row + xx];
        if (xx + 1 >= 0 && xx + 1 < a.W) v.y = a.labels[row + xx + 1];
      }
    }
    lv[k] = v;
  }

  // ---- this thread's pixel and its own inputs
  const int q = q0 + (lane >> 1), j = lane & 1;
  const int rx = (OX ? 2 * q : 2 * q - 1) + j;
  const int ry = ry0 + wrp;
  const int x = 2 * rx + ((rx + OX) & 1);
  const int y = 2 * ry + OY;
  const bool ok = ry < a.raw_h && y < a.H && 32 * (ry / 16) + OY < a.H &&
                  rx >= 0 && rx < a.raw_w && 32 * (rx / 16) < a.W && x < a.W;
  const size_t p = ok ? (size_t)y * a.W + x : 0;
  PixelIn pin;
  pin.bounds = 0; pin.col = make_uchar4(0, 0, 0, 0); pin.disp = 0.f; pin.inlier = 0;
  if (ok) pin = tps_fetch_pixel<DISP>(a, p);

  if (tr) tr[1] = clock64();
  if (a.pdl_late && a.pdl_trig == 2) pdl_trigger();
  // ---- means (and planes) of the window from the quiescent sums: the merge kernel's arithmetic,
  // split over two threads per superpixel so that the divisions of the means and the dependent
  // divisions of the plane solve run side by side
  if (does_mean) {
    const float n = (float)sv[2].y;
    Superpixel& s = win[slot];
    s.xy_rg = make_float4((float)sv[0].x / n, (float)sv[0].y / n, (float)sv[1].x / n, (float)sv[1].y / n);
    s.size = make_float4(n, 0.f, 0.f, 0.f);
    const float mb = (float)sv[2].x / n;
    if (DISP) s.theta_b.w = mb;
    else s.theta_b = make_float4(0.f, 0.f, 0.f, mb);
  } else if (does_plane) {
    const float dx = (float)sv[0].x, dy = (float)sv[0].y, dxx = (float)sv[1].x, dyy = (float)sv[1].y,
                dxy = (float)sv[2].x, dn = (float)sv[2].y;
    const float dxd = (float)((double)sv[3].x * (1.0 / kDispFix));
    const float dyd = (float)((double)sv[3].y * (1.0 / kDispFix));
    const float dd = (float)((double)sv[4].x * (1.0 / kDispFix));
    float tx = 0.f, ty = 0.f, tz = 0.f;
    if (!solve_plane(tx, ty, tz, dxx, dxy, dx, dxd, dxy, dyy, dy, dyd, dx, dy, dn, dd)) {
      tx = 0.f; ty = 0.f; tz = __int_as_float(0xFFE00000);
    }
    Superpixel& s = win[slot];
    s.theta_b.x = tx; s.theta_b.y = ty; s.theta_b.z = tz;
  }
  if (tr) tr[2] = clock64();
  if (TMA) {
    __syncthreads();                       // the barrier was initialised by thread 0: nobody polls it before this
    if (tr) tr[7] = clock64();
    mbar_wait(&lab_bar, 0);
    // the tensor map fills texels outside the image with 0, a valid label: patch them (border tiles only)
    const bool inside = xs0 >= 0 && xs0 + TILE_COLS <= a.W && y0 - 1 >= 0 && y0 - 1 + TILE_LROWS <= a.H;
    if (!inside) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int r = wrp + k * TILE_ROWS;
        const int yy = y0 - 1 + r, xx = xs0 + 2 * lane;
        if (r < TILE_LROWS) {
          const bool row_out = yy < 0 || yy >= a.H;
          if (row_out || xx < 0 || xx >= a.W) lab[r][co + 2 * lane] = -1;
          if (row_out || xx + 1 < 0 || xx + 1 >= a.W) lab[r][co + 2 * lane + 1] = -1;
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int r = wrp + k * TILE_ROWS;
      if (r < TILE_LROWS) *reinterpret_cast<int2*>(&lab[r][co + 2 * lane]) = lv[k];
    }
  }
  __syncthreads();

  if (tr) tr[3] = clock64();
  // ---- decide on the pass-start state
  const int c = x - xs0;                                 // column of the pixel in the staged rows (1 .. 62 when ok)
  Decision d;
  d.index = d.new_index = -1; d.b = 0; d.inlier = 0; d.prev_inlier = 0;
  float dv = 0.f;
  int nl[4] = {-1, -1, -1, -1};
  if (ok) {
    int L[3][4];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int k = 0; k < 3; k++) L[r][k] = lab[2 * wrp + r][co + c - 1 + k];
    L[0][3] = L[1][3] = L[2][3] = -1;
    const SpTile<DISP> src = {w};
    tps_decide<DISP>(a, src, pin, x, y, L, 1, d, dv);
    nl[0] = L[0][1]; nl[1] = L[1][0]; nl[2] = L[1][2]; nl[3] = L[2][1];   // up, left, right, down
  }
  const bool moved = ok && d.new_index != d.index;
  if (tr) tr[4] = clock64();
  if (a.pdl_late && a.pdl_trig == 0) pdl_trigger();

  // ---- apply.  Everything read above is pass-start state: the only pixels written in this pass
  // are active ones, each by its own thread, and the only active 4-neighbour of an active pixel is
  // its pair partner (the neighbouring lane).
  int owe_partner = 0;                                   // boundary-count delta this pixel owes its partner
  if (moved) {
    const int nxs[4] = {0, -1, 1, 0}, nys[4] = {-1, 0, 0, 1};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int i_n = nl[k];
      int delta = 0;
      if (i_n == d.new_index) delta = -1;
      else if (i_n == d.index) delta = 1;
      if (delta == 0 || i_n == -1) continue;
      const bool is_partner = (j == 0 && k == 2) || (j == 1 && k == 1);
      if (is_partner) owe_partner = delta;               // delivered by the shuffle below iff the partner is live
      else atomicAdd(&a.bound[(size_t)(y + nys[k]) * a.W + (x + nxs[k])], delta);
    }
    const uchar4 col = pin.col;
    SpSums* o = &a.sums_nxt[d.index];
    SpSums* n = &a.sums_nxt[d.new_index];
    add64(&o->x, -x); add64(&o->y, -y); add64(&o->r, -(int)col.x); add64(&o->g, -(int)col.y);
    add64(&o->b, -(int)col.z); add64(&o->n, -1);
    add64(&n->x, x); add64(&n->y, y); add64(&n->r, col.x); add64(&n->g, col.y); add64(&n->b, col.z);
    add64(&n->n, 1);
    a.labels[p] = d.new_index;
  }
  if (DISP && ok) {
    const unsigned char inl = d.inlier, was = d.prev_inlier;
    if (inl && (!was || moved)) {
      SpSums* s = &a.sums_nxt[d.new_index];
      const long long qd = quantize(dv, kDispFix, kDispClamp);
      add64(&s->dx, x); add64(&s->dy, y); add64(&s->dxx, (long long)x * x); add64(&s->dyy, (long long)y * y);
      add64(&s->dxy, (long long)x * y); add64(&s->dxd, (long long)x * qd); add64(&s->dyd, (long long)y * qd);
      add64(&s->dd, qd); add64(&s->dn, 1);
    }
    if (was && (!inl || moved)) {
      SpSums* s = &a.sums_nxt[d.index];
      const long long qd = quantize(dv, kDispFix, kDispClamp);
      add64(&s->dx, -x); add64(&s->dy, -y); add64(&s->dxx, -(long long)x * x); add64(&s->dyy, -(long long)y * y);
      add64(&s->dxy, -(long long)x * y); add64(&s->dxd, -(long long)x * qd); add64(&s->dyd, -(long long)y * qd);
      add64(&s->dd, -qd); add64(&s->dn, -1);
    }
    if (inl != was) a.inliers[p] = inl;
  }
  // the partner's debt; a dead partner (outside the reference's thread extent) is an inactive pixel
  // and gets its delta through the atomic path instead
  const int partner_ok = __shfl_xor_sync(0xffffffffu, ok ? 1 : 0, 1);
  const int from_partner = __shfl_xor_sync(0xffffffffu, owe_partner, 1);
  if (owe_partner != 0 && !partner_ok)
    atomicAdd(&a.bound[(size_t)y * a.W + (x + (j == 0 ? 1 : -1))], owe_partner);
  if (ok) {
    if (moved) a.bound[p] = d.b;                                        // own "= b" wins
    else if (from_partner != 0) a.bound[p] = pin.bounds + from_partner; // only this thread touches it
  }

  if (tr) tr[5] = clock64();
  if (a.pdl_late && a.pdl_trig == 3) pdl_trigger();
  // ---- buffer rotation, second half (last: nothing waits for it)
  {
    if (rot_mine) {
      const size_t off = (size_t)rot_s0 * 16 + (tid - ROT_FIRST);
      if (rot_v != 0) add64(reinterpret_cast<long long*>(a.sums_nxt) + off, rot_v);
      reinterpret_cast<long long*>(a.sums_zero)[off] = 0;
    }
    for (int i = 64 + tid; i < rot_words; i += TILE_THREADS) {     // more than 4 superpixels per CTA (coarse grids)
      const size_t off = (size_t)rot_s0 * 16 + i;
      const long long v = reinterpret_cast<const long long*>(cur)[off];
      if (v != 0) add64(reinterpret_cast<long long*>(a.sums_nxt) + off, v);
      reinterpret_cast<long long*>(a.sums_zero)[off] = 0;
    }
  }
  if (tr) tr[6] = clock64();
}

// ---- RANSAC plane initialisation (TPS_RGBD_kernels.cu:318-467, 112-190) --------------
__global__ void tps_rng_init_kernel(curandState* states, int n) {
  pdl_sync();
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < n) curand_init(1234, id, 0, &states[id]);
}

__device__ __forceinline__ int tex_label(const TpsArgs& a, float x, float y) {
  return a.labels[(size_t)tex_coord(y, a.H) * a.W + tex_coord(x, a.W)];
}
__device__ __forceinline__ float tex_disp(const TpsArgs& a, float x, float y) {
  return a.disp[(size_t)tex_coord(y, a.H) * a.W + tex_coord(x, a.W)];
}

__device__ __forceinline__ void tps_init_sample_item(const TpsArgs& a, float4* samples, int* votes,
                                                     curandState* states, int nbWalks, float radius, int index,
                                                     int idx) {
  curandState rnd = states[idx];
  // the centroid straight from the sums (what the merge kernel would have stored in a.sp: same
  // conversions, same divisions)
  const SpSums& sm = a.sums[index];
  const float cn = (float)sm.n;
  const float cx = (float)sm.x / cn, cy = (float)sm.y / cn;
  float x = cx, y = cy;
  int i = tex_label(a, x, y);
  int k = 0;
  while (i != index && k++ < 10) {
    x = (float)((double)cx + ((double)radius * 2.) * (double)(curand_uniform(&rnd) - 1.f));
    y = (float)((double)cy + ((double)radius * 2.) * (double)(curand_uniform(&rnd) - 1.f));
    i = tex_label(a, x, y);
  }
  float3 xyd[3];
  // The walk re-reads the label under the CURRENT position at every step (TPS_RGBD_kernels.cu:360); labels do
  // not change during this kernel, so that read is carried along and refreshed only when the walker moves --
  // together with the disparity of the new position, one round of loads per move instead of two in series.
  int here = i;                        // label under (x, y)
  const float d0 = tex_disp(a, x, y);
  xyd[0] = xyd[1] = xyd[2] = make_float3(x, y, d0);
#pragma unroll
  for (int j = 0; j < 3; j++) {
    for (int w = 0; w < nbWalks; w++) {
      const int dir = curand(&rnd) & 3;                       // 0 left, 1 up, 2 right, 3 down
      const float next_x = x + ((dir & 1) ? 0.f : (float)(dir - 1));
      const float next_y = y + ((dir & 1) ? (float)(dir - 2) : 0.f);
      if (here == index && next_x >= 0 && next_x < (float)a.W && next_y >= 0 && next_y < (float)a.H) {
        x = next_x; y = next_y;
        const size_t at = (size_t)tex_coord(y, a.H) * a.W + tex_coord(x, a.W);
        here = a.labels[at];
        const float dd = a.disp[at];
        if (isfinite(dd)) xyd[j] = make_float3(x, y, dd);
      }
    }
  }
  float tx = 0.f, ty = 0.f, tz = 0.f;
  if (!solve_plane(tx, ty, tz, xyd[0].x, xyd[0].y, 1.f, xyd[0].z, xyd[1].x, xyd[1].y, 1.f, xyd[1].z, xyd[2].x,
                   xyd[2].y, 1.f, xyd[2].z)) {
    tx = 0.f; ty = 0.f; tz = xyd[2].z;
  }
  samples[idx] = make_float4(tx, ty, tz, 0.f);
  votes[idx] = 0;
  states[idx] = rnd;
}

__global__ void tps_init_samples_kernel(TpsArgs a, float4* samples, int* votes, curandState* states, int nbWalks,
                                        float radius) {
  pdl_sync();
  tps_init_sample_item(a, samples, votes, states, nbWalks, radius, blockIdx.x, blockIdx.x * blockDim.x + threadIdx.x);
}

// evalSamples_kernel: integer votes, aggregated per warp over lanes that share a label.
// One warp handles 32 consecutive pixels of a row; all 32 lanes must call it.
__device__ __forceinline__ void tps_eval_row_item(const TpsArgs& a, const float4* samples, int* votes, int nbSamples,
                                                  int x, int y) {
  const bool in = (x < a.W && y < a.H);
  const size_t p = in ? (size_t)y * a.W + x : 0;
  const int index = in ? a.labels[p] : -1;
  const float d = in ? a.disp[p] : 0.f;
  const unsigned grp = __match_any_sync(0xffffffffu, index);
  const int lane = threadIdx.x & 31;
  const bool leader = (lane == (__ffs(grp) - 1));
  for (int k = 0; k < nbSamples; k++) {
    bool vote = false;
    if (in) {
      const float4 th = samples[(size_t)index * nbSamples + k];
      if (isfinite(th.z)) {
        const float dp = th.x * (float)x + th.y * (float)y + th.z;
        const float dd = (d - dp) * (d - dp);
        vote = dd < a.thresh_disp;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, vote);
    if (in && leader) {
      const int cnt = __popc(bal & grp);
      if (cnt) atomicAdd(&votes[(size_t)index * nbSamples + k], cnt);
    }
  }
}

__global__ void tps_eval_samples_kernel(TpsArgs a, const float4* samples, int* votes, int nbSamples) {
  pdl_sync();
  // blockDim.x == 32: a warp is one 32-pixel row segment
  tps_eval_row_item(a, samples, votes, nbSamples, blockIdx.x * blockDim.x + threadIdx.x,
                    blockIdx.y * blockDim.y + threadIdx.y);
}

__device__ __forceinline__ void tps_select_item(const TpsArgs& a, float4* samples, const int* votes, int nbSamples,
                                                int idx) {
  // first strict maximum of the votes (TPS_RGBD_kernels.cu:446-456).  The votes are fetched eight at a time so
  // that the loop is two L2 round trips for the usual 16 samples instead of one per sample; the winning plane is
  // read afterwards, the vote counts go back into the samples' fourth component for ssf_get_ransac_samples.
  const int* v = votes + (size_t)idx * nbSamples;
  float4* smp = samples + (size_t)idx * nbSamples;
  float best_w = 0.f;
  int best_k = -1;
  for (int k0 = 0; k0 < nbSamples; k0 += 8) {
    int cnt[8];
#pragma unroll
    for (int j = 0; j < 8; j++) cnt[j] = (k0 + j < nbSamples) ? v[k0 + j] : 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (k0 + j < nbSamples) {
        const float w = (float)cnt[j];
        smp[k0 + j].w = w;
        if (w > best_w) { best_w = w; best_k = k0 + j; }
      }
    }
  }
  float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
  if (best_k >= 0) best = smp[best_k];
  a.sp[idx].theta_b.x = best.x; a.sp[idx].theta_b.y = best.y; a.sp[idx].theta_b.z = best.z;
  SpSums* c = &a.sums[idx];
  c->dx = c->dy = c->dxx = c->dyy = c->dxy = c->dn = c->dxd = c->dyd = c->dd = 0;
}

__global__ void tps_select_samples_kernel(TpsArgs a, float4* samples, const int* votes, int nbSamples) {
  pdl_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < a.S) tps_select_item(a, samples, votes, nbSamples, idx);
}

// All 32 lanes of a warp (32 consecutive pixels of a row) call this; `in` is false for lanes
// beyond the image.  Moments are pre-reduced per run of equal labels.
__device__ __forceinline__ void tps_init_disp_item(const TpsArgs& a, int ransac, int x, int y, bool in) {
  const size_t p = in ? (size_t)y * a.W + x : 0;
  const int index = in ? a.labels[p] : -1;
  const float d = in ? a.disp[p] : 0.f;
  unsigned char inlier = 0;
  long long v[9];
#pragma unroll
  for (int k = 0; k < 9; k++) v[k] = 0;
  if (in && isfinite(d)) {
    bool okp = true;
    if (ransac) {
      const float4 th = a.sp[index].theta_b;
      const float dp = th.x * (float)x + th.y * (float)y + th.z;
      const float dd = (dp - d) * (dp - d);
      okp = isfinite(dd) && dd < a.thresh_disp && dp > 0.f;
    }
    if (okp) {
      inlier = 0xff;
      const long long qd = quantize(d, kDispFix, kDispClamp);
      v[0] = x; v[1] = y; v[2] = (long long)x * x; v[3] = (long long)y * y; v[4] = (long long)x * y;
      v[5] = (long long)x * qd; v[6] = (long long)y * qd; v[7] = qd; v[8] = 1;
    }
  }
  if (in) a.inliers[p] = inlier;
  const int head = run_head_lane(index);
  if (__ballot_sync(0xffffffffu, inlier != 0) == 0u) return;
  run_reduce<9>(v, head);
  if ((int)(threadIdx.x & 31) == head && index >= 0 && v[8] != 0) {
    SpSums* s = &a.sums[index];
    add64(&s->dx, v[0]); add64(&s->dy, v[1]); add64(&s->dxx, v[2]); add64(&s->dyy, v[3]); add64(&s->dxy, v[4]);
    add64(&s->dxd, v[5]); add64(&s->dyd, v[6]); add64(&s->dd, v[7]); add64(&s->dn, v[8]);
  }
}

__global__ void tps_init_disp_kernel(TpsArgs a, int ransac) {
  pdl_sync();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  tps_init_disp_item(a, ransac, x, y, x < a.W && y < a.H);   // blockDim.x == 32
}

// ---- plane smoothing (TPS_RGBD.cu:480-505; TPS_RGBD_kernels.cu:510-614) -----------
// Jacobi (double buffered).  Node record: X.xyz, Z.xyz, px, py.
__device__ __forceinline__ void tps_filter_init_item(const TpsArgs& a, float* buf, int i, bool merge = false) {
  Superpixel s;
  if (merge) {                                  // fused passes: the last merge happens here
    s = tps_superpixel_from_sums<true>(a.sums[i]);
    a.sp[i] = s;
  } else {
    s = a.sp[i];
  }
  const float X0 = s.xy_rg.x * s.theta_b.x + s.xy_rg.y * s.theta_b.y + s.theta_b.z;
  float* n = buf + 8 * (size_t)i;
  n[0] = X0; n[1] = s.theta_b.x; n[2] = s.theta_b.y;
  n[3] = X0; n[4] = s.theta_b.x; n[5] = s.theta_b.y;
  n[6] = s.xy_rg.x; n[7] = s.xy_rg.y;
}

__device__ __forceinline__ void tps_filter_iter_item(const TpsArgs& a, const float* cur, float* nxt, int idx,
                                                     float alpha, float beta, float threshold) {
  const int vv[4] = {-1, 0, 0, 1};
  const int uu[4] = {0, -1, 1, 0};
  const int x = idx % a.gx, y = idx / a.gx;
  const float* ni = cur + 8 * (size_t)idx;
  const V3 Xi = v3(ni[0], ni[1], ni[2]);
  const V3 Zi = v3(ni[3], ni[4], ni[5]);
  const float pxi = ni[6], pyi = ni[7];
  Sym3 A = sym3(alpha, 0.f, 0.f, alpha, 0.f, alpha);
  V3 R = alpha * Zi;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int yy = y + vv[j], xx = x + uu[j];
    if (yy >= 0 && yy < a.gy && xx >= 0 && x < a.gx) {   // sic: x, not xx (reference :582-583)
      const int nidx = yy * a.gx + xx;
      if (nidx >= a.S) continue;
      const float* nj = cur + 8 * (size_t)nidx;
      const V3 Xj = v3(nj[0], nj[1], nj[2]);
      const float dx = pxi - nj[6];
      const float dy = pyi - nj[7];
      const float dz = Xi.x - Xj.x;
      if (isfinite(dz) && dz * dz < threshold * threshold) {
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
  float* no = nxt + 8 * (size_t)idx;
  Sym3 Ai;
  V3 Xn = Xi;
  if (invert(A, Ai)) Xn = Ai * R;
  no[0] = Xn.x; no[1] = Xn.y; no[2] = Xn.z;
  no[3] = Zi.x; no[4] = Zi.y; no[5] = Zi.z; no[6] = pxi; no[7] = pyi;
}

__device__ __forceinline__ void tps_filter_finish_item(const TpsArgs& a, const float* cur, int i) {
  const float* n = cur + 8 * (size_t)i;
  Superpixel& s = a.sp[i];
  const float X0 = n[0], X1 = n[1], X2 = n[2];
  s.theta_b.x = X1;
  s.theta_b.y = X2;
  s.theta_b.z = X0 - s.xy_rg.x * X1 - s.xy_rg.y * X2;
}

// single-CTA version (multi-kernel path): the node state as planes -- the two Jacobi copies of X (3 floats
// each), the anchor Z (3) and the centroid (2): 44 bytes per node.  SMEM: the planes live in shared memory
// (up to 5200 superpixels: 640x480 and 1280x960 at the default cell size; larger frames run the same code
// on a global scratch buffer).  An iteration then costs a shared-memory round trip instead of an L2 round
// trip per neighbour, and the fused final merge has all its loads in flight at once.
struct FilterPlanes {
  float* xa;    // [3][S] current X
  float* xb;    // [3][S] next X
  float* z;     // [3][S] anchor
  float* p;     // [2][S] centroid
  int S;
};
constexpr int kFilterFloatsPerNode = 11;

__device__ __forceinline__ void tps_filter_iter_planes(const TpsArgs& a, const FilterPlanes& f, const float* cur,
                                                       float* nxt, int idx, float alpha, float beta, float threshold) {
  const int vv[4] = {-1, 0, 0, 1};
  const int uu[4] = {0, -1, 1, 0};
  const int S = f.S;
  const int x = idx % a.gx, y = idx / a.gx;
  const V3 Xi = v3(cur[idx], cur[S + idx], cur[2 * S + idx]);
  const V3 Zi = v3(f.z[idx], f.z[S + idx], f.z[2 * S + idx]);
  const float pxi = f.p[idx], pyi = f.p[S + idx];
  Sym3 A = sym3(alpha, 0.f, 0.f, alpha, 0.f, alpha);
  V3 R = alpha * Zi;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int yy = y + vv[j], xx = x + uu[j];
    if (yy >= 0 && yy < a.gy && xx >= 0 && x < a.gx) {   // sic: x, not xx (reference :582-583)
      const int nidx = yy * a.gx + xx;
      if (nidx >= a.S) continue;
      const V3 Xj = v3(cur[nidx], cur[S + nidx], cur[2 * S + nidx]);
      const float dx = pxi - f.p[nidx];
      const float dy = pyi - f.p[S + nidx];
      const float dz = Xi.x - Xj.x;
      if (isfinite(dz) && dz * dz < threshold * threshold) {
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
  Sym3 Ai;
  V3 Xn = Xi;
  if (invert(A, Ai)) Xn = Ai * R;
  nxt[idx] = Xn.x; nxt[S + idx] = Xn.y; nxt[2 * S + idx] = Xn.z;
}

template <bool SMEM>
__global__ void __launch_bounds__(1024) tps_filter_kernel(TpsArgs a, float* scratch, int iters, float alpha, float beta,
                                                          float threshold, int merge) {
  pdl_sync();
  extern __shared__ float filt_smem[];
  float* base = SMEM ? filt_smem : scratch;
  const int S = a.S;
  FilterPlanes f = {base, base + 3 * (size_t)S, base + 6 * (size_t)S, base + 9 * (size_t)S, S};
  // init (TPS_RGBD_kernels.cu:510-540), with the last merge of the fused passes folded in
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    Superpixel s;
    if (merge) {
      s = tps_superpixel_from_sums<true>(a.sums[i]);
      a.sp[i] = s;
    } else {
      s = a.sp[i];
    }
    const float X0 = s.xy_rg.x * s.theta_b.x + s.xy_rg.y * s.theta_b.y + s.theta_b.z;
    f.xa[i] = X0; f.xa[S + i] = s.theta_b.x; f.xa[2 * S + i] = s.theta_b.y;
    f.z[i] = X0; f.z[S + i] = s.theta_b.x; f.z[2 * S + i] = s.theta_b.y;
    f.p[i] = s.xy_rg.x; f.p[S + i] = s.xy_rg.y;
  }
  __syncthreads();
  float* cur = f.xa;
  float* nxt = f.xb;
  for (int it = 0; it < iters; it++) {
    for (int idx = threadIdx.x; idx < S; idx += blockDim.x) tps_filter_iter_planes(a, f, cur, nxt, idx, alpha, beta, threshold);
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  // finish (TPS_RGBD_kernels.cu:596-614)
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    Superpixel& s = a.sp[i];
    const float X0 = cur[i], X1 = cur[S + i], X2 = cur[2 * S + i];
    s.theta_b.x = X1;
    s.theta_b.y = X2;
    s.theta_b.z = X0 - f.p[i] * X1 - f.p[S + i] * X2;
  }
}

// ---- slanted-plane depth render (TPS_RGBD_kernels.cu:469-508) -> interleaved map -----
__device__ __forceinline__ void tps_render_item(const TpsArgs& a, int2* lmap, int x, int y) {
  const size_t p = (size_t)y * a.W + x;
  const int index = a.labels[p];
  const float4 th = a.sp[index].theta_b;
  const float disp = (float)x * th.x + (float)y * th.y + th.z;
  lmap[p] = make_int2(index, __float_as_int(1.f / disp));
}

__global__ void tps_render_kernel(TpsArgs a, int2* lmap) {
  pdl_sync();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < a.W && y < a.H) tps_render_item(a, lmap, x, y);
}

// ---- the whole segmentation as ONE persistent cooperative kernel ---------------------
// 40 relabelling passes, 41 merges, RANSAC, smoothing and the render are phases of a
// grid-stride kernel separated by grid-wide barriers (~1 us each) instead of ~90 dependent
// kernel launches (~4-7 us each at VGA, where every phase is far too small to fill the
// GPU).  One CTA per SM, co-resident by construction (cooperative launch).
struct TpsRun {
  int nb_iters, use_ransac, nb_samples, filter_iters;
  float alpha, beta, threshold, radius;
  float4* samples;
  int* votes;
  curandState* states;
  float* filt_a;
  float* filt_b;
  int2* lmap;
  unsigned int* barrier;
  int cache_slots;       // capacity of the shared-memory superpixel cache
  unsigned long long* trace;   // optional per-CTA phase timestamps (profiling aid), else NULL
};

// 768 threads: at VGA a CTA's share of a pass (77 040 adjacent-pixel pairs / 148 CTAs = 521) is
// ONE round of one thread per pair -- one dependent chain of L2 round trips per pass instead of
// the three rounds a 512-thread, lane-per-pixel CTA needed
constexpr int TPS_PERSIST_THREADS = 768;

// Grid-wide barrier for the co-resident CTAs of the persistent kernel: one release
// atomic per CTA on a monotonically increasing ticket and an acquire spin by thread 0
// (cooperative_groups::grid_group::sync costs ~5 us here; this is ~1 us).  The acquire
// at gpu scope also drops the SM's L1 lines, so the plain loads that follow observe
// what the other CTAs wrote before their release.
struct GridBarrier {
  unsigned int* ticket;
  unsigned int target;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += gridDim.x;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ticket) : "memory");
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ticket) : "memory");
      } while ((int)(seen - target) < 0);
    }
    __syncthreads();
  }
};

// One relabelling pass of the persistent kernel.  The CTA owns a band of active rows for
// the whole kernel; it first rebuilds its shared-memory cache of superpixel means from the
// (now quiescent) sums -- which replaces the separate merge phase and its barrier -- then
// relabels its band, then meets the other CTAs at the barrier.
__device__ __forceinline__ unsigned long long tps_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// trace layout: [cta][slot] with 4 stamps per pass (start, cache ready, items done, barrier passed)
constexpr int TPS_TRACE_SLOTS = 512;
struct TpsTrace {
  unsigned long long* base;
  int n;
  __device__ __forceinline__ void stamp() {
    if (base && threadIdx.x == 0 && n < TPS_TRACE_SLOTS) base[(size_t)blockIdx.x * TPS_TRACE_SLOTS + n++] = tps_now();
  }
};

struct TpsBand {
  int i0, i1;          // pair range [i0, i1) of this CTA; pair i = (active row i / pairs, pair i % pairs)
  int first, count;    // cached superpixel ids
};

template <bool DISP>
__device__ __forceinline__ void tps_phase_pass(const TpsArgs& a, GridBarrier& grid, Superpixel* cache,
                                               const TpsBand& band, int OX, int OY, TpsTrace& tr) {
  tr.stamp();
  for (int i = threadIdx.x; i < band.count; i += blockDim.x)
    cache[i] = tps_superpixel_from_sums<DISP>(a.sums[band.first + i]);
  __syncthreads();
  tr.stamp();
  const SpCached<DISP> src = {cache, a.sums, band.first, band.count};
  const int pairs = a.raw_w / 2 + 1;
  for (int i = band.i0 + threadIdx.x; i < band.i1; i += blockDim.x)     // one thread per adjacent pair
    tps_pass_item<DISP>(a, src, i % pairs, i / pairs, OX, OY);
  __syncthreads();
  tr.stamp();
  grid.sync();
  tr.stamp();
}

__device__ __forceinline__ void tps_phase_merge_global(const TpsArgs& a, GridBarrier& grid, int tid, int nth, bool disp) {
  for (int k = tid; k < a.S; k += nth) {
    if (disp) tps_merge_item<true>(a, k);
    else tps_merge_item<false>(a, k);
  }
  grid.sync();
}

__global__ void __launch_bounds__(TPS_PERSIST_THREADS, 1) tps_persistent_kernel(TpsArgs a, TpsRun r) {
  pdl_sync();
  GridBarrier grid;
  grid.ticket = r.barrier;
  grid.target = 0;   // the ticket is zeroed by a memset node ahead of this kernel
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  const int npix = a.W * a.H;
  // band of active rows owned by this CTA, and the grid-cell rows its cache covers
  extern __shared__ Superpixel sp_cache[];
  TpsTrace tr = {r.trace, 0};
  TpsBand band;
  {
    const int pairs = a.raw_w / 2 + 1;
    const int total = pairs * a.raw_h;
    const int per = (total + gridDim.x - 1) / gridDim.x;
    band.i0 = min(total, (int)blockIdx.x * per);
    band.i1 = min(total, band.i0 + per);
    const int ry0 = band.i0 / pairs, ry1 = (band.i1 > band.i0) ? (band.i1 - 1) / pairs + 1 : ry0;
    const int y0 = 2 * ry0, y1 = min(a.H - 1, 2 * ry1 + 1);
    // A boundary moves at most one pixel per pass, so after all 4*nb_iters passes a pixel's
    // label was seeded at most `margin` grid-cell rows away: the cache always hits.
    const int margin = (4 * r.nb_iters + 1 + a.cell - 1) / a.cell;
    const int c0 = max(0, y0 / a.cell - margin), c1 = min(a.gy - 1, y1 / a.cell + margin);
    band.first = c0 * a.gx;
    band.count = (band.i1 > band.i0) ? min(r.cache_slots, (c1 - c0 + 1) * a.gx) : 0;
  }
  // pass order per iteration: (0,0) (1,1) (0,1) (1,0) (TPS_RGBD.cu:190-272).  One copy of the
  // pass body per phase (not eight): the kernel must stay inside the instruction cache.
#pragma unroll 1
  for (int k = 0; k < 4 * (r.nb_iters / 2); k++) {
    const int s4 = k & 3;
    tps_phase_pass<false>(a, grid, sp_cache, band, (s4 == 1 || s4 == 3) ? 1 : 0, (s4 == 1 || s4 == 2) ? 1 : 0, tr);
  }
  if (r.nb_iters / 2 > 0) tps_phase_merge_global(a, grid, tid, nth, false);   // RANSAC reads the global means
  tr.stamp();
  if (r.use_ransac) {
    for (int i = tid; i < a.S * r.nb_samples; i += nth)
      tps_init_sample_item(a, r.samples, r.votes, r.states, 10, r.radius, i / r.nb_samples, i);
    grid.sync();
    {
      const int segs = (a.W + 31) / 32;
      const int rows = segs * a.H;
      const int lane = threadIdx.x & 31;
      for (int w = tid >> 5; w < rows; w += nth >> 5)
        tps_eval_row_item(a, r.samples, r.votes, r.nb_samples, (w % segs) * 32 + lane, w / segs);
    }
    grid.sync();
    for (int i = tid; i < a.S; i += nth) tps_select_item(a, r.samples, r.votes, r.nb_samples, i);
    grid.sync();
  }
  {
    const int segs = (a.W + 31) / 32;
    const int rows = segs * a.H;
    const int lane = threadIdx.x & 31;
    for (int w = tid >> 5; w < rows; w += nth >> 5) {
      const int x = (w % segs) * 32 + lane;
      tps_init_disp_item(a, r.use_ransac, x, w / segs, x < a.W);
    }
  }
  grid.sync();
  for (int k = tid; k < a.S; k += nth) tps_merge_item<true>(a, k);
  grid.sync();
  tr.stamp();
#pragma unroll 1
  for (int k = 0; k < 4 * (r.nb_iters - r.nb_iters / 2); k++) {
    const int s4 = k & 3;
    tps_phase_pass<true>(a, grid, sp_cache, band, (s4 == 1 || s4 == 3) ? 1 : 0, (s4 == 1 || s4 == 2) ? 1 : 0, tr);
  }
  if (r.nb_iters - r.nb_iters / 2 > 0) tps_phase_merge_global(a, grid, tid, nth, true);   // the filter reads the global means
  // plane smoothing
  for (int i = tid; i < a.S; i += nth) tps_filter_init_item(a, r.filt_a, i);
  grid.sync();
  float* cur = r.filt_a;
  float* nxt = r.filt_b;
  for (int it = 0; it < r.filter_iters; it++) {
    for (int i = tid; i < a.S; i += nth) tps_filter_iter_item(a, cur, nxt, i, r.alpha, r.beta, r.threshold);
    grid.sync();
    float* t = cur; cur = nxt; nxt = t;
  }
  for (int i = tid; i < a.S; i += nth) tps_filter_finish_item(a, cur, i);
  grid.sync();
  tr.stamp();
  for (int p = tid; p < npix; p += nth) tps_render_item(a, r.lmap, p % a.W, p / a.W);
  tr.stamp();
}

// ------------------------------------------------------------------- launchers
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

constexpr size_t kFilterSmemMax = 224 * 1024;
// opt the shared-memory form of the plane filter in to its dynamic shared memory (once per process)
void tps_configure() {
  cudaFuncSetAttribute(tps_filter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFilterSmemMax);
}

// The 2-D tensor map over a slot's label image (int32, H rows of W) with the tile box of the fused pass;
// cuTensorMapEncodeTiled is resolved through the runtime (no link-time dependency on libcuda).  Returns
// false when the image cannot be described (row pitch not a multiple of 16 bytes) or the driver lacks it.
bool tps_make_label_map(void* map128, const int* labels, int W, int H) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
  if (W % 4 != 0) return false;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
      q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
  const cuuint64_t strides[1] = {(cuuint64_t)W * sizeof(int)};
  const cuuint32_t box[2] = {(cuuint32_t)TILE_SCOLS, (cuuint32_t)TILE_LROWS};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = reinterpret_cast<EncodeFn>(fn)(reinterpret_cast<CUtensorMap*>(map128), CU_TENSOR_MAP_DATA_TYPE_INT32, 2,
                                                     const_cast<int*>(labels), dims, strides, box, estr,
                                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS;
}

void tps_init_rng(Engine* e) {
  const int n = e->S * e->cfg.nb_samples;
  launch_pdl(e, tps_rng_init_kernel, dim3(cdiv(n, 128)), dim3(128), 0, reinterpret_cast<curandState*>(e->rng), n);
  e->launches++;
}

size_t tps_rng_state_bytes() { return sizeof(curandState); }
size_t tps_trace_bytes(int grid) { return (size_t)grid * TPS_TRACE_SLOTS * sizeof(unsigned long long); }

// co-resident CTAs available to the persistent kernel (0 = cooperative launch unsupported)
int tps_persistent_grid(int device, int gx, int gy, int cell, int height, int nb_iters, int* cache_slots) {
  int coop = 0, sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  *cache_slots = 0;
  if (!coop) return 0;
  // rows of grid cells a band of ceil(H/2/sms) active rows touches, plus two on each side
  const int raw_h = 16 * ((height / 2 + 15) / 16);
  const int per = (raw_h + sms - 1) / sms + 1;      // a CTA's pair range may straddle one more active row
  const int margin = (4 * nb_iters + 1 + cell - 1) / cell;
  int rows = (2 * per + 1) / cell + 2 + 2 * margin;
  if (rows > gy) rows = gy;
  const int slots = rows * gx;
  const int max_slots = (200 * 1024) / (int)sizeof(Superpixel);
  if (slots > max_slots) return 0;   // image too large for the cached scheme: multi-kernel path
  const size_t smem = (size_t)slots * sizeof(Superpixel);
  cudaFuncSetAttribute(tps_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tps_persistent_kernel, TPS_PERSIST_THREADS, smem);
  if (per_sm <= 0) return 0;
  *cache_slots = slots;
  return sms;
}

void launch_ingest(Engine* e, const uint8_t* rgb_dev, size_t rgb_stride, const float* depth_dev,
                   size_t depth_stride) {
  TpsArgs a = tps_args(e);
  if (e->tps_fused && !e->tps_persistent) a.sums_nxt = sums_buffer(e, 1);   // zeroed for the first fused pass
  launch_pdl(e, tps_seed_kernel, dim3(e->S), dim3(256), 0, a, rgb_dev, rgb_stride, depth_dev, depth_stride);
  e->launches++;
}

template <bool DISP>
static void launch_pass(Engine* e, const TpsArgs& a, int OX, int OY) {
  const int pairs = a.raw_w / 2 + 1;
  // (fusing the merge into the pass as "last CTA done" was measured 40 % slower per frame: one
  // CTA merging 1200 superpixels serialises five dependent L2 round trips)
  dim3 blk(32, 4), grd(cdiv(pairs, 32), cdiv(a.raw_h, 4));
  launch_pdl(e, tps_pass_kernel<DISP>, dim3(grd), dim3(blk), 0, a, OX, OY);
  launch_pdl(e, tps_merge_kernel<DISP>, dim3(cdiv(a.S, 128)), dim3(128), 0, a);
  e->launches += 2;
}

// pass number `p` of the frame (0 .. 4 * seg_iter - 1) in its fused form: reads buffer p % 3,
// accumulates into (p + 1) % 3, clears (p + 2) % 3
template <bool DISP>
static void launch_pass_fused(Engine* e, TpsArgs a, int p, int OX, int OY) {
  const int pairs = a.raw_w / 2 + 1;
  a.sums = sums_buffer(e, p);
  a.sums_nxt = sums_buffer(e, p + 1);
  a.sums_zero = sums_buffer(e, p + 2);
  dim3 grd(cdiv(pairs, TILE_LANES / 2), cdiv(a.raw_h, TILE_ROWS));
  const CUtensorMap& map = *reinterpret_cast<const CUtensorMap*>(e->label_map[e->cur_slot]);
  if (e->tps_occ >= 4) {
    if (e->tps_tma) launch_kernel(e, true, tps_pass_tile_kernel<DISP, true, 4>, dim3(grd), dim3(TILE_THREADS), 0, a, OX, OY, map);
    else launch_kernel(e, true, tps_pass_tile_kernel<DISP, false, 4>, dim3(grd), dim3(TILE_THREADS), 0, a, OX, OY, map);
  } else {
    if (e->tps_tma) launch_kernel(e, true, tps_pass_tile_kernel<DISP, true, 3>, dim3(grd), dim3(TILE_THREADS), 0, a, OX, OY, map);
    else launch_kernel(e, true, tps_pass_tile_kernel<DISP, false, 3>, dim3(grd), dim3(TILE_THREADS), 0, a, OX, OY, map);
  }
  e->launches += 1;
}

template <bool DISP>
static void launch_iteration(Engine* e, const TpsArgs& a, int first_pass) {
  // pass order per iteration: (0,0) (1,1) (0,1) (1,0) (TPS_RGBD.cu:190-272)
  static const int ox[4] = {0, 1, 0, 1}, oy[4] = {0, 1, 1, 0};
  for (int k = 0; k < 4; k++) {
    if (e->tps_fused) launch_pass_fused<DISP>(e, a, first_pass + k, ox[k], oy[k]);
    else launch_pass<DISP>(e, a, ox[k], oy[k]);
  }
}

// The segmentation as a sequence of steps, so that the pipelined mode can cut it anywhere between
// two iterations: steps [0, I/2) are the colour-only iterations, step I/2 is the disparity-plane
// initialisation (RANSAC + inlier moments + merge), steps (I/2, I] the colour + disparity
// iterations, step I + 1 smoothing + render.  launch_tps(e, first, last) enqueues steps [first, last).
int tps_step_count(const Engine* e) { return e->tps_persistent ? 1 : e->cfg.seg_iter + 2; }

void launch_tps(Engine* e, int first, int last) {
  TpsArgs a = tps_args(e);
  const int nbIters = e->cfg.seg_iter;
  if (last < 0) last = tps_step_count(e);
  if (e->tps_persistent) {
    if (first > 0 || last < 1) return;      // the one-kernel form is a single step
    TpsRun r;
    r.nb_iters = nbIters; r.use_ransac = e->cfg.seg_use_ransac; r.nb_samples = e->cfg.nb_samples;
    r.filter_iters = e->cfg.filter_iter;
    r.alpha = e->cfg.filter_alpha; r.beta = e->cfg.filter_beta; r.threshold = e->cfg.filter_threshold;
    r.radius = (float)e->cfg.cell_size / 2.f;
    r.samples = e->samples;
    r.votes = reinterpret_cast<int*>(e->samples + (size_t)e->S * e->cfg.nb_samples);
    r.states = reinterpret_cast<curandState*>(e->rng);
    r.filt_a = e->filt_a; r.filt_b = e->filt_b; r.lmap = e->lmap;
    r.barrier = e->tps_barrier;
    r.cache_slots = e->tps_cache_slots;
    r.trace = e->tps_trace;
    cudaMemsetAsync(e->tps_barrier, 0, sizeof(unsigned int), e->stream);
    void* params[] = {&a, &r};
    const cudaError_t rc = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(tps_persistent_kernel), dim3(e->tps_grid),
                                                       dim3(TPS_PERSIST_THREADS), params,
                                                       (size_t)e->tps_cache_slots * sizeof(Superpixel), e->stream);
    if (rc != cudaSuccess && e->launch_err == cudaSuccess) e->launch_err = rc;
    e->launches++;
    return;
  }
  const int half = nbIters / 2;
  const bool fused = e->tps_fused != 0;
  dim3 blk(32, 8), grd(cdiv(e->W, 32), cdiv(e->H, 8));
  for (int step = first; step < last; step++) {
    // passes already run on this frame when the step starts = which of the rotating buffers holds the sums
    const int done = step <= half ? 4 * step : 4 * (step - 1);
    if (fused) a.sums = sums_buffer(e, done < 4 * nbIters ? done : 4 * nbIters);
    if (step < half) {                       // colour-only iteration
      launch_iteration<false>(e, a, done);
    } else if (step == half) {               // disparity planes: RANSAC, inlier moments, first merge
      if (e->cfg.seg_use_ransac) {
        int* votes = reinterpret_cast<int*>(e->samples + (size_t)e->S * e->cfg.nb_samples);
        launch_pdl(e, tps_init_samples_kernel, dim3(e->S), dim3(e->cfg.nb_samples), 0,
            a, e->samples, votes, reinterpret_cast<curandState*>(e->rng), 10, (float)e->cfg.cell_size / 2.f);
        launch_pdl(e, tps_eval_samples_kernel, dim3(grd), dim3(blk), 0, a, e->samples, votes, e->cfg.nb_samples);
        launch_pdl(e, tps_select_samples_kernel, dim3(cdiv(e->S, 128)), dim3(128), 0, a, e->samples, votes, e->cfg.nb_samples);
        launch_pdl(e, tps_init_disp_kernel, dim3(grd), dim3(blk), 0, a, 1);
        e->launches += 4;
      } else {
        launch_pdl(e, tps_init_disp_kernel, dim3(grd), dim3(blk), 0, a, 0);
        e->launches += 1;
      }
      if (!fused) {                          // the fused passes derive means + plane from the sums themselves
        launch_pdl(e, tps_merge_kernel<true>, dim3(cdiv(a.S, 128)), dim3(128), 0, a);
        e->launches++;
      }
    } else if (step <= nbIters) {            // colour + disparity iteration
      launch_iteration<true>(e, a, done);
    } else {                                 // plane smoothing + slanted-depth render
      const size_t filt_bytes = (size_t)e->S * kFilterFloatsPerNode * sizeof(float);
      if (filt_bytes <= kFilterSmemMax)
        launch_pdl(e, tps_filter_kernel<true>, dim3(1), dim3(1024), filt_bytes, a, e->filt_a, e->cfg.filter_iter,
                   e->cfg.filter_alpha, e->cfg.filter_beta, e->cfg.filter_threshold, fused ? 1 : 0);
      else   // filt_a holds 16 floats per node: room for the 11 planes
        launch_pdl(e, tps_filter_kernel<false>, dim3(1), dim3(1024), 0, a, e->filt_a, e->cfg.filter_iter,
                   e->cfg.filter_alpha, e->cfg.filter_beta, e->cfg.filter_threshold, fused ? 1 : 0);
      launch_pdl(e, tps_render_kernel, dim3(grd), dim3(blk), 0, a, e->lmap);
      e->launches += 2;
    }
  }
}

}  // namespace ssf
