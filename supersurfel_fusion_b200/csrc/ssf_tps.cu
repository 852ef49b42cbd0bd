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

#include <curand_kernel.h>

namespace ssf {

constexpr double kDispFix = 1073741824.0;      // 2^30
constexpr double kDispClamp = 137438953472.0;  // 2^37

struct TpsArgs {
  int W, H, cell, gx, gy, S;
  int raw_w, raw_h;        // thread extents of the reference launch (TPS_RGBD.cu:185-186)
  int min_size;
  float lambda_pos, lambda_bound, lambda_size, lambda_disp, thresh_disp;
  uchar4* rgba;
  float* disp;
  int* labels;
  int* bound;
  unsigned char* inliers;
  Superpixel* sp;
  SpSums* sums;
};

static TpsArgs tps_args(const Engine* e) {
  TpsArgs a;
  a.W = e->W; a.H = e->H; a.cell = e->cfg.cell_size; a.gx = e->gx; a.gy = e->gy; a.S = e->S;
  a.raw_w = 16 * ((e->W / 2 + 15) / 16);
  a.raw_h = 16 * ((e->H / 2 + 15) / 16);
  a.min_size = (int)((float)(e->cfg.cell_size * e->cfg.cell_size) / 4.f);
  a.lambda_pos = e->cfg.lambda_pos; a.lambda_bound = e->cfg.lambda_bound; a.lambda_size = e->cfg.lambda_size;
  a.lambda_disp = e->cfg.lambda_disp; a.thresh_disp = e->cfg.thresh_disp;
  a.rgba = e->rgba; a.disp = e->disp; a.labels = e->labels; a.bound = e->bound; a.inliers = e->inliers;
  a.sp = e->sp; a.sums = e->sums;
  return a;
}

__device__ __forceinline__ void add64(long long* p, long long v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}

// ---- seeding: cvtColor + depth2disp + initSuperpixelsRGBD + first merge ------------
// (TPS_RGBD.cu:126-180; TPS_RGBD_kernels.cu:61-110, 224-242, 278-296).  One CTA per
// grid cell; the cell's sums come from an in-CTA reduction.
__global__ void __launch_bounds__(256) tps_seed_kernel(TpsArgs a, const uint8_t* __restrict__ rgb, size_t rgb_stride,
                                                       const float* __restrict__ depth, size_t depth_stride) {
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

template <bool DISP>
__global__ void tps_merge_kernel(TpsArgs a) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.S) return;
  const SpSums c = a.sums[k];
  const float n = (float)c.n;
  Superpixel& s = a.sp[k];
  s.xy_rg = make_float4((float)c.x / n, (float)c.y / n, (float)c.r / n, (float)c.g / n);
  s.size.x = n;
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
    s.theta_b.w = mb;
  }
}

// ---- one relabelling pass (TPS_RGBD_kernels.cuh:235-651) ---------------------------
struct Decision {
  int index, new_index, b;
  unsigned char inlier, prev_inlier;
};

template <bool DISP>
__device__ __forceinline__ void tps_decide(const TpsArgs& a, int x, int y, const int (&L)[3][4], int c, Decision& d,
                                           float& disp_v) {
  const size_t p = (size_t)y * a.W + x;
  const int bounds = a.bound[p];
  const int index = L[1][c];
  int new_index = index;
  const Superpixel prev = a.sp[index];
  unsigned char inlier = 0xff, prev_inlier = 0;
  float disp_energy = 0.f;
  disp_v = 0.f;
  if (DISP) {
    disp_v = a.disp[p];
    prev_inlier = a.inliers[p];
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
    const uchar4 col = a.rgba[p];
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
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int i_n = nl[k];
      if (i_n == -1 || i_n == index) continue;
      const Superpixel ns = a.sp[i_n];
      const float ex = px - ns.xy_rg.x, ey = py - ns.xy_rg.y;
      const float fx = cr - ns.xy_rg.z, fy = cg - ns.xy_rg.w, fz = cb - ns.theta_b.w;
      const float nsize = ns.size.x + 1.f - (float)a.min_size;
      float n_energy = 0.f;
      unsigned char n_inlier = 0xff;
      if (DISP) {
        const float dp = ns.theta_b.x * (float)x + ns.theta_b.y * (float)y + ns.theta_b.z;
        n_energy = (dp - disp_v) * (dp - disp_v);
        if (!isfinite(n_energy) || n_energy > a.thresh_disp || dp < 0.f) {
          n_energy = a.thresh_disp;
          n_inlier = 0;
        }
      }
      int b = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) b += (nl[q] != i_n);
      float energy = (fx * fx + fy * fy + fz * fz) + a.lambda_pos * (ex * ex + ey * ey);
      if (DISP) energy = energy + a.lambda_disp * n_energy;
      energy = energy - a.lambda_size * fminf(nsize, 0.f);
      energy = energy + a.lambda_bound * (float)b;
      if (energy < best) {
        best = energy;
        new_index = i_n;
        if (DISP) inlier = n_inlier;
      }
    }
    if (new_index != index) {
#pragma unroll
      for (int q = 0; q < 4; q++) newb += (nl[q] != new_index);
    }
  }
  d.index = index; d.new_index = new_index; d.b = newb; d.inlier = inlier; d.prev_inlier = prev_inlier;
}

template <bool DISP>
__global__ void __launch_bounds__(128) tps_pass_kernel(TpsArgs a, int OX, int OY) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;   // pair index along the row
  const int ry = blockIdx.y * blockDim.y + threadIdx.y;
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
#pragma unroll
  for (int j = 0; j < 2; j++)
    if (ok[j]) tps_decide<DISP>(a, xs[j], y, L, 1 + j, d[j], dv[j]);

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
      const uchar4 col = a.rgba[p];
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

// ---- RANSAC plane initialisation (TPS_RGBD_kernels.cu:318-467, 112-190) --------------
__global__ void tps_rng_init_kernel(curandState* states, int n) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < n) curand_init(1234, id, 0, &states[id]);
}

__device__ __forceinline__ int tex_label(const TpsArgs& a, float x, float y) {
  return a.labels[(size_t)tex_coord(y, a.H) * a.W + tex_coord(x, a.W)];
}
__device__ __forceinline__ float tex_disp(const TpsArgs& a, float x, float y) {
  return a.disp[(size_t)tex_coord(y, a.H) * a.W + tex_coord(x, a.W)];
}

__global__ void tps_init_samples_kernel(TpsArgs a, float4* samples, int* votes, curandState* states, int nbWalks,
                                        float radius) {
  const int index = blockIdx.x;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  curandState rnd = states[idx];
  const float cx = a.sp[index].xy_rg.x, cy = a.sp[index].xy_rg.y;
  float x = cx, y = cy;
  int i = tex_label(a, x, y);
  int k = 0;
  while (i != index && k++ < 10) {
    x = (float)((double)cx + ((double)radius * 2.) * (double)(curand_uniform(&rnd) - 1.f));
    y = (float)((double)cy + ((double)radius * 2.) * (double)(curand_uniform(&rnd) - 1.f));
    i = tex_label(a, x, y);
  }
  const float dxs[4] = {-1.f, 0.f, 1.f, 0.f};
  const float dys[4] = {0.f, -1.f, 0.f, 1.f};
  float3 xyd[3];
  const float d0 = tex_disp(a, x, y);
  xyd[0] = xyd[1] = xyd[2] = make_float3(x, y, d0);
#pragma unroll
  for (int j = 0; j < 3; j++) {
    for (int w = 0; w < nbWalks; w++) {
      const int dir = curand(&rnd) & 3;
      const float next_x = x + dxs[dir], next_y = y + dys[dir];
      i = tex_label(a, x, y);
      if (i == index && next_x >= 0 && next_x < (float)a.W && next_y >= 0 && next_y < (float)a.H) {
        x = next_x; y = next_y;
        const float dd = tex_disp(a, x, y);
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

// evalSamples_kernel: integer votes, aggregated per warp over lanes that share a label
__global__ void tps_eval_samples_kernel(TpsArgs a, const float4* __restrict__ samples, int* votes, int nbSamples) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const bool in = (x < a.W && y < a.H);
  const size_t p = in ? (size_t)y * a.W + x : 0;
  const int index = in ? a.labels[p] : -1;
  const float d = in ? a.disp[p] : 0.f;
  const unsigned grp = __match_any_sync(0xffffffffu, index);
  const int lane = threadIdx.x & 31;   // blockDim.x == 32
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

__global__ void tps_select_samples_kernel(TpsArgs a, float4* samples, const int* votes, int nbSamples) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.S) return;
  float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < nbSamples; k++) {
    float4 th = samples[(size_t)idx * nbSamples + k];
    th.w = (float)votes[(size_t)idx * nbSamples + k];
    samples[(size_t)idx * nbSamples + k].w = th.w;
    if (th.w > best.w) best = th;
  }
  a.sp[idx].theta_b.x = best.x; a.sp[idx].theta_b.y = best.y; a.sp[idx].theta_b.z = best.z;
  SpSums* c = &a.sums[idx];
  c->dx = c->dy = c->dxx = c->dyy = c->dxy = c->dn = c->dxd = c->dyd = c->dd = 0;
}

__global__ void tps_init_disp_kernel(TpsArgs a, int ransac) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= a.W || y >= a.H) return;
  const size_t p = (size_t)y * a.W + x;
  const int index = a.labels[p];
  const float d = a.disp[p];
  unsigned char inlier = 0;
  if (isfinite(d)) {
    bool okp = true;
    if (ransac) {
      const float4 th = a.sp[index].theta_b;
      const float dp = th.x * (float)x + th.y * (float)y + th.z;
      const float dd = (dp - d) * (dp - d);
      okp = isfinite(dd) && dd < a.thresh_disp && dp > 0.f;
    }
    if (okp) {
      inlier = 0xff;
      SpSums* s = &a.sums[index];
      const long long qd = quantize(d, kDispFix, kDispClamp);
      add64(&s->dx, x); add64(&s->dy, y); add64(&s->dxx, (long long)x * x); add64(&s->dyy, (long long)y * y);
      add64(&s->dxy, (long long)x * y); add64(&s->dxd, (long long)x * qd); add64(&s->dyd, (long long)y * qd);
      add64(&s->dd, qd); add64(&s->dn, 1);
    }
  }
  a.inliers[p] = inlier;
}

// ---- plane smoothing (TPS_RGBD.cu:480-505; TPS_RGBD_kernels.cu:510-614) -----------
// Single CTA, Jacobi (double buffered).  Node record: X.xyz, Z.xyz, px, py.
__global__ void __launch_bounds__(1024) tps_filter_kernel(TpsArgs a, float* bufA, float* bufB, int iters, float alpha,
                                                          float beta, float threshold) {
  const int S = a.S;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const Superpixel s = a.sp[i];
    const float X0 = s.xy_rg.x * s.theta_b.x + s.xy_rg.y * s.theta_b.y + s.theta_b.z;
    float* n = bufA + 8 * (size_t)i;
    n[0] = X0; n[1] = s.theta_b.x; n[2] = s.theta_b.y;
    n[3] = X0; n[4] = s.theta_b.x; n[5] = s.theta_b.y;
    n[6] = s.xy_rg.x; n[7] = s.xy_rg.y;
  }
  __syncthreads();
  float* cur = bufA;
  float* nxt = bufB;
  const int vv[4] = {-1, 0, 0, 1};
  const int uu[4] = {0, -1, 1, 0};
  for (int it = 0; it < iters; it++) {
    for (int idx = threadIdx.x; idx < S; idx += blockDim.x) {
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
          if (nidx >= S) continue;
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
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float* n = cur + 8 * (size_t)i;
    Superpixel& s = a.sp[i];
    const float X0 = n[0], X1 = n[1], X2 = n[2];
    s.theta_b.x = X1;
    s.theta_b.y = X2;
    s.theta_b.z = X0 - s.xy_rg.x * X1 - s.xy_rg.y * X2;
  }
}

// ---- slanted-plane depth render (TPS_RGBD_kernels.cu:469-508) -> interleaved map -----
__global__ void tps_render_kernel(TpsArgs a, int2* lmap) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= a.W || y >= a.H) return;
  const size_t p = (size_t)y * a.W + x;
  const int index = a.labels[p];
  const float4 th = a.sp[index].theta_b;
  const float disp = (float)x * th.x + (float)y * th.y + th.z;
  lmap[p] = make_int2(index, __float_as_int(1.f / disp));
}

// ------------------------------------------------------------------- launchers
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

void tps_init_rng(Engine* e) {
  const int n = e->S * e->cfg.nb_samples;
  tps_rng_init_kernel<<<cdiv(n, 128), 128, 0, e->stream>>>(reinterpret_cast<curandState*>(e->rng), n);
  e->launches++;
}

size_t tps_rng_state_bytes() { return sizeof(curandState); }

void launch_ingest(Engine* e, const uint8_t* rgb_dev, size_t rgb_stride, const float* depth_dev,
                   size_t depth_stride) {
  TpsArgs a = tps_args(e);
  tps_seed_kernel<<<e->S, 256, 0, e->stream>>>(a, rgb_dev, rgb_stride, depth_dev, depth_stride);
  e->launches++;
}

template <bool DISP>
static void launch_pass(Engine* e, const TpsArgs& a, int OX, int OY) {
  const int pairs = a.raw_w / 2 + 1;
  dim3 blk(32, 4), grd(cdiv(pairs, 32), cdiv(a.raw_h, 4));
  tps_pass_kernel<DISP><<<grd, blk, 0, e->stream>>>(a, OX, OY);
  tps_merge_kernel<DISP><<<cdiv(a.S, 128), 128, 0, e->stream>>>(a);
  e->launches += 2;
}

void launch_tps(Engine* e) {
  TpsArgs a = tps_args(e);
  const int nbIters = e->cfg.seg_iter;
  for (int k = 0; k < nbIters / 2; k++) {
    launch_pass<false>(e, a, 0, 0);
    launch_pass<false>(e, a, 1, 1);
    launch_pass<false>(e, a, 0, 1);
    launch_pass<false>(e, a, 1, 0);
  }
  dim3 blk(32, 8), grd(cdiv(e->W, 32), cdiv(e->H, 8));
  if (e->cfg.seg_use_ransac) {
    int* votes = reinterpret_cast<int*>(e->samples + (size_t)e->S * e->cfg.nb_samples);
    tps_init_samples_kernel<<<e->S, e->cfg.nb_samples, 0, e->stream>>>(
        a, e->samples, votes, reinterpret_cast<curandState*>(e->rng), 10, (float)e->cfg.cell_size / 2.f);
    tps_eval_samples_kernel<<<grd, blk, 0, e->stream>>>(a, e->samples, votes, e->cfg.nb_samples);
    tps_select_samples_kernel<<<cdiv(e->S, 128), 128, 0, e->stream>>>(a, e->samples, votes, e->cfg.nb_samples);
    tps_init_disp_kernel<<<grd, blk, 0, e->stream>>>(a, 1);
    e->launches += 4;
  } else {
    tps_init_disp_kernel<<<grd, blk, 0, e->stream>>>(a, 0);
    e->launches += 1;
  }
  tps_merge_kernel<true><<<cdiv(a.S, 128), 128, 0, e->stream>>>(a);
  e->launches++;
  for (int k = nbIters / 2; k < nbIters; k++) {
    launch_pass<true>(e, a, 0, 0);
    launch_pass<true>(e, a, 1, 1);
    launch_pass<true>(e, a, 0, 1);
    launch_pass<true>(e, a, 1, 0);
  }
  tps_filter_kernel<<<1, 1024, 0, e->stream>>>(a, e->filt_a, e->filt_b, e->cfg.filter_iter, e->cfg.filter_alpha,
                                               e->cfg.filter_beta, e->cfg.filter_threshold);
  tps_render_kernel<<<grd, blk, 0, e->stream>>>(a, e->lmap);
  e->launches += 2;
}

}  // namespace ssf
