// Supersurfel extraction and model fusion / culling for sm_100a.
//
// Replaces computeSupersurfelCoeffs + computeSupersurfels
// (reference: core/src/supersurfel_fusion_kernels.cu:113-224, launched from
//  core/src/supersurfel_fusion.cu:551-593) and findBestMatches / updateSupersurfels /
// insertSupersurfels / filterModel + thrust::sort_by_key
// (supersurfel_fusion_kernels.cu:348-467,522-682; supersurfel_fusion.cu:351-483).
//
// Differences in mechanism (results follow the deterministic serialisation written
// down in oracle/oracle_surfels.cpp):
//  * per-superpixel moments are accumulated as 2^-32 fixed-point 64-bit integers
//    (RED.ADD.64, order-free) instead of fp32 atomicAdd;
//  * association is a packed 64-bit atomicMin (distance bits | model id): a true
//    arg-min with ties to the lowest id instead of a racy compare + two atomicExch;
//  * insertion is a block scan (ascending frame id), not a warp-aggregated counter;
//  * the 3-way reorder after culling is a stable partition by scan + scatter of the
//    planar attributes, not a sort of 104-byte tuples; no per-frame cudaMalloc.
#include "ssf_engine.h"
#include "ssf_math.cuh"

namespace ssf {

constexpr double kFix = 4294967296.0;            // 2^32
constexpr double kFixClamp = 1152921504606846976.0;  // 2^60

struct CamK { float fx, fy, cx, cy; int W, H; };
static CamK cam_of(const Engine* e) {
  CamK c; c.fx = e->cfg.cam.fx; c.fy = e->cfg.cam.fy; c.cx = e->cfg.cam.cx; c.cy = e->cfg.cam.cy;
  c.W = e->W; c.H = e->H; return c;
}

__device__ __forceinline__ V3 ldv(const SurfelSet& s, int plane, int i) {
  return v3(s.plane(plane)[i], s.plane(plane + 1)[i], s.plane(plane + 2)[i]);
}
__device__ __forceinline__ void stv(const SurfelSet& s, int plane, int i, V3 v) {
  s.plane(plane)[i] = v.x; s.plane(plane + 1)[i] = v.y; s.plane(plane + 2)[i] = v.z;
}
__device__ __forceinline__ Sym3 lds(const SurfelSet& s, int i) {
  return sym3(s.plane(P_SHAPE)[i], s.plane(P_SHAPE + 1)[i], s.plane(P_SHAPE + 2)[i], s.plane(P_SHAPE + 3)[i],
              s.plane(P_SHAPE + 4)[i], s.plane(P_SHAPE + 5)[i]);
}
__device__ __forceinline__ void sts(const SurfelSet& s, int i, const Sym3& c) {
  s.plane(P_SHAPE)[i] = c.xx; s.plane(P_SHAPE + 1)[i] = c.xy; s.plane(P_SHAPE + 2)[i] = c.xz;
  s.plane(P_SHAPE + 3)[i] = c.yy; s.plane(P_SHAPE + 4)[i] = c.yz; s.plane(P_SHAPE + 5)[i] = c.zz;
}
__device__ __forceinline__ M3 ldo(const SurfelSet& s, int i) {
  return m3(ldv(s, P_ORI, i), ldv(s, P_ORI + 3, i), ldv(s, P_ORI + 6, i));
}
__device__ __forceinline__ void sto(const SurfelSet& s, int i, const M3& m) {
  stv(s, P_ORI, i, m.r0); stv(s, P_ORI + 3, i, m.r1); stv(s, P_ORI + 6, i, m.r2);
}
__device__ __forceinline__ M3 pose_R(const DevicePose* p) {
  return m3(v3(p->R[0], p->R[1], p->R[2]), v3(p->R[3], p->R[4], p->R[5]), v3(p->R[6], p->R[7], p->R[8]));
}
__device__ __forceinline__ V3 pose_t(const DevicePose* p) { return v3(p->t[0], p->t[1], p->t[2]); }

// ------------------------------------------------------------------ extraction
// computeSupersurfelCoeffs (supersurfel_fusion_kernels.cu:113-167): one pixel per
// thread, 17 bytes read, 13 order-free reductions for a contributing pixel.
__global__ void extract_accumulate_kernel(const int2* __restrict__ lmap, const unsigned char* __restrict__ inliers,
                                          const int* __restrict__ bound, const uchar4* __restrict__ rgba,
                                          unsigned long long* __restrict__ xsums, CamK cam) {
  pdl_sync();
  // blockDim.x == 32: a warp is 32 consecutive pixels of one row
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const bool in = (x < cam.W && y < cam.H);
  const size_t p = in ? (size_t)y * cam.W + x : 0;
  const int2 lz = in ? lmap[p] : make_int2(-1, 0);
  const float depth = __int_as_float(lz.y);
  const bool use = in && inliers[p] && isfinite(depth) && depth > 0.0f && bound[p] == 0;
  long long v[13];
#pragma unroll
  for (int k = 0; k < 13; k++) v[k] = 0;
  if (use) {
    const uchar4 c = rgba[p];
    const V3 pos = v3(((float)x - cam.cx) * depth / cam.fx, ((float)y - cam.cy) * depth / cam.fy, depth);
    const V3 lab = rgb_to_lab(v3((float)c.x, (float)c.y, (float)c.z));
    const Sym3 o = outer(pos);
    const float vals[12] = {pos.x, pos.y, pos.z, lab.x, lab.y, lab.z, o.xx, o.xy, o.xz, o.yy, o.yz, o.zz};
#pragma unroll
    for (int k = 0; k < 12; k++) v[k] = quantize(vals[k], kFix, kFixClamp);
    v[12] = 1;
  }
  // one set of atomics per run of equal labels instead of one per pixel
  const int head = run_head_lane(lz.x);
  if (__ballot_sync(0xffffffffu, use) == 0u) return;
  run_reduce<13>(v, head);
  if ((int)(threadIdx.x & 31) == head && lz.x >= 0 && v[12] != 0) {
    unsigned long long* a = xsums + (size_t)lz.x * 16;
#pragma unroll
    for (int k = 0; k < 13; k++) atomicAdd(&a[k], (unsigned long long)v[k]);
  }
}

__device__ __forceinline__ float dequantize(unsigned long long v) {
  return (float)((double)(long long)v * (1.0 / kFix));
}

// Writes the derived per-supersurfel records the gather-side kernels read: CIELab
// plane and the 32-byte (Lab, conf | normal) record.
__device__ __forceinline__ void write_frame_tables(const SurfelSet& frame, float4* ftab, int k, V3 rgb, V3 nrm,
                                                   float conf) {
  const V3 lab = rgb_to_lab(rgb);
  stv(frame, P_LAB, k, lab);
  ftab[2 * k] = make_float4(lab.x, lab.y, lab.z, conf);
  ftab[2 * k + 1] = make_float4(nrm.x, nrm.y, nrm.z, 0.0f);
}

// computeSupersurfels (supersurfel_fusion_kernels.cu:169-224); also clears the
// accumulators and the per-frame association slots for the next use.
__global__ void extract_finalize_kernel(SurfelSet frame, float4* ftab, unsigned long long* xsums,
                                        unsigned char* matched, unsigned long long* best, float z_min,
                                        float z_max, const Counters* counters, int S) {
  pdl_sync();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  unsigned long long* a = xsums + (size_t)k * 16;
  V3 position = v3(dequantize(a[0]), dequantize(a[1]), dequantize(a[2]));
  V3 color = v3(dequantize(a[3]), dequantize(a[4]), dequantize(a[5]));
  Sym3 shape = sym3(dequantize(a[6]), dequantize(a[7]), dequantize(a[8]), dequantize(a[9]), dequantize(a[10]),
                    dequantize(a[11]));
  float conf = (float)(long long)a[12];
#pragma unroll
  for (int j = 0; j < 13; j++) a[j] = 0ull;
  matched[k] = 0;
  best[k] = ((unsigned long long)__float_as_uint(0.05f) << 32) | 0xFFFFFFFFull;

  M3 orient = m3(v3(0, 0, 0), v3(0, 0, 0), v3(0, 0, 0));
  float d0 = 0.f, d1 = 0.f;
  int s0 = 0, s1 = 0;
  const float z = position.z / conf;
  if (isfinite(z) && conf > 100.0f && z > z_min && z < z_max) {
    position = v3(position.x / conf, position.y / conf, z);
    color = lab_to_rgb(v3(color.x / conf, color.y / conf, color.z / conf));
    shape = shape / conf - outer(position);
    V3 vals;
    eigenframe(shape, orient, vals);
    d0 = vals.x; d1 = vals.y;
    s0 = s1 = counters->seg_stamp;   // == stamp, except in the pipelined mode where segmentation runs a frame ahead
    if (vals.x / vals.y > 50.0f) conf = -1.0f;
  } else {
    conf = -1.0f;
  }
  stv(frame, P_POS, k, position);
  stv(frame, P_COL, k, color);
  frame.plane(P_STAMP)[k] = __int_as_float(s0);
  frame.plane(P_STAMP + 1)[k] = __int_as_float(s1);
  sto(frame, k, orient);
  sts(frame, k, shape);
  frame.plane(P_DIM)[k] = d0;
  frame.plane(P_DIM + 1)[k] = d1;
  frame.plane(P_CONF)[k] = conf;
  write_frame_tables(frame, ftab, k, color, orient.r2, conf);
}

__global__ void frame_tables_kernel(SurfelSet frame, float4* ftab, unsigned char* matched,
                                    unsigned long long* best, int S) {
  pdl_sync();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  matched[k] = 0;
  best[k] = ((unsigned long long)__float_as_uint(0.05f) << 32) | 0xFFFFFFFFull;
  write_frame_tables(frame, ftab, k, ldv(frame, P_COL, k), ldv(frame, P_ORI + 6, k), frame.plane(P_CONF)[k]);
}

__global__ void model_lab_kernel(SurfelSet model, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  stv(model, P_LAB, i, rgb_to_lab(ldv(model, P_COL, i)));
}

__global__ void build_lmap_kernel(int2* lmap, const int* labels, const float* slanted, size_t n) {
  pdl_sync();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  lmap[i] = make_int2(labels[i], __float_as_int(slanted[i]));
}

__global__ void invalidate_kernel(SurfelSet frame, float4* ftab, const uint8_t* mask, int S) {
  pdl_sync();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S || !mask[k]) return;
  frame.plane(P_CONF)[k] = -1.0f;
  ftab[2 * k].w = -1.0f;
}

// ---------------------------------------------------------------------- fusion
// findBestMatches (supersurfel_fusion_kernels.cu:522-599)
__global__ void associate_kernel(SurfelSet model, SurfelSet frame, const float4* __restrict__ ftab,
                                 const int2* __restrict__ lmap, unsigned char* matched, unsigned long long* best,
                                 const DevicePose* pose, Counters* counters, CamK cam, float z_min,
                                 float z_max) {
  pdl_sync();
  const int n = counters->nb_supersurfels > 0 ? counters->nb_visible : 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {     // per-frame counters of the update start at zero
    counters->nb_matched = 0;
    counters->nb_inserted = 0;
    counters->nb_removed = 0;
  }
  const M3 R = pose_R(pose);
  const V3 t = pose_t(pose);
  const M3 Rview = transpose(R);
  const V3 tview = -(Rview * t);
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
    if (!(model.plane(P_CONF)[m] > 0.0f)) continue;
    const V3 mp = ldv(model, P_POS, m);
    const V3 pv = Rview * mp + tview;
    const int px = round_px(pv.x * cam.fx / pv.z + cam.cx);
    const int py = round_px(pv.y * cam.fy / pv.z + cam.cy);
    if (!(pv.z > z_min && pv.z < z_max && px >= 0 && px < cam.W && py >= 0 && py < cam.H)) continue;
    const int f = lmap[(size_t)py * cam.W + px].x;
    matched[f] = 1;
    const float4 f0 = ftab[2 * f];
    if (!(f0.w > 0.0f)) continue;
    const V3 fp = R * ldv(frame, P_POS, f) + t;
    const V3 fn = normalize(row_times(ldv(frame, P_ORI + 6, f), Rview));
    const V3 mn = normalize(ldv(model, P_ORI + 6, m));
    const float dist = length(mp - fp);
    const float lab_dist = length(ldv(model, P_LAB, m) - v3(f0.x, f0.y, f0.z));
    const float delta_norm = fabsf(dot(mn, fn));
    if (lab_dist < 15.0f && delta_norm > 0.8f && dist < 0.05f)
      atomicMin(&best[f], ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)m);
  }
}

// updateSupersurfels (supersurfel_fusion_kernels.cu:601-682)
__global__ void update_kernel(SurfelSet model, SurfelSet frame, const unsigned char* matched,
                              const unsigned long long* best, const DevicePose* pose, Counters* counters, int S) {
  pdl_sync();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= S) return;
  if (counters->nb_supersurfels <= 0 || counters->nb_visible <= 0) return;
  const unsigned id = (unsigned)(best[f] & 0xFFFFFFFFull);
  if (!(matched[f] && id != 0xFFFFFFFFu)) return;
  const int m = (int)id;
  atomicAdd(&counters->nb_matched, 1);
  const M3 R = pose_R(pose);
  const V3 t = pose_t(pose);
  const V3 mp = ldv(model, P_POS, m);
  const V3 fp = R * ldv(frame, P_POS, f) + t;
  const Sym3 fshape = rotate_sym(R, lds(frame, f));
  const Sym3 mshape = lds(model, m);
  const V3 flab = ldv(frame, P_LAB, f);
  const V3 mlab = ldv(model, P_LAB, m);
  const float m_conf = model.plane(P_CONF)[m];
  const float f_conf = frame.plane(P_CONF)[f];
  const float ratio = 1.0f / (m_conf + f_conf);
  model.plane(P_STAMP + 1)[m] = __int_as_float(counters->stamp);
  const V3 fused_color = lab_to_rgb(ratio * (f_conf * flab + m_conf * mlab));
  Sym3 f1, m1, fused_shape, fused1;
  V3 fused_pos;
  const float w = ratio * f_conf;
  bool info = false;
  if (invert(fshape, f1) && invert(mshape, m1)) {
    fused1 = w * f1 + (1.0f - w) * m1;
    if (invert(fused1, fused_shape)) {
      fused_pos = fused_shape * ((w * f1) * fp + ((1.0f - w) * m1) * mp);
      info = true;
    }
  }
  if (!info) {
    fused_shape = ratio * (f_conf * fshape + m_conf * mshape);
    fused_pos = ratio * (f_conf * fp + m_conf * mp);
  }
  stv(model, P_POS, m, fused_pos);
  model.plane(P_CONF)[m] = m_conf + f_conf;
  sts(model, m, fused_shape);
  M3 vecs;
  V3 vals;
  eigenframe(fused_shape, vecs, vals);
  sto(model, m, vecs);
  stv(model, P_COL, m, fused_color);
  stv(model, P_LAB, m, rgb_to_lab(fused_color));
  model.plane(P_DIM)[m] = vals.x;
  model.plane(P_DIM + 1)[m] = vals.y;
}

// insertSupersurfels (supersurfel_fusion_kernels.cu:348-395) in ascending frame id, or
// the first-frame bootstrap model <- frame (supersurfel_fusion.cu:477-483).  One CTA.
constexpr int INS_THREADS = 1024;
__global__ void __launch_bounds__(INS_THREADS) insert_kernel(SurfelSet model, SurfelSet frame,
                                                              const unsigned char* matched, const DevicePose* pose,
                                                              Counters* counters, int* skip_filter, int S, int cap) {
  pdl_sync();
  __shared__ int warp_sums[INS_THREADS / 32];
  __shared__ int running;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nb0 = counters->nb_supersurfels;
  if (nb0 <= 0) {
    for (int f = tid; f < S; f += INS_THREADS)
      for (int p = 0; p < P_COUNT; p++) model.plane(p)[f] = frame.plane(p)[f];
    if (tid == 0) {
      counters->nb_supersurfels = S;
      counters->nb_visible = S;
      counters->nb_removed = 0;
      counters->nb_inserted = S;
      *skip_filter = 1;
    }
    return;
  }
  const M3 R = pose_R(pose);
  const V3 t = pose_t(pose);
  const M3 Rt = transpose(R);
  const int stamp = counters->stamp;
  if (tid == 0) running = 0;
  __syncthreads();
  // four consecutive frame supersurfels per thread and round: 640x480 (S = 1200) is ONE round of flags ->
  // block scan -> writes instead of two
  constexpr int INS_ITEMS = 4;
  for (int base = 0; base < S; base += INS_THREADS * INS_ITEMS) {
    const int f0 = base + tid * INS_ITEMS;
    int flag[INS_ITEMS];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < INS_ITEMS; j++) {
      const int f = f0 + j;
      flag[j] = (f < S && frame.plane(P_CONF)[f] > 0.0f && !matched[f]) ? 1 : 0;
      mine += flag[j];
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int v = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      warp_sums[lane] = v;  // inclusive
    }
    __syncthreads();
    int before = running + (wid > 0 ? warp_sums[wid - 1] : 0) + incl - mine;
    const int total = warp_sums[INS_THREADS / 32 - 1];
#pragma unroll
    for (int j = 0; j < INS_ITEMS; j++) {
      if (!flag[j]) continue;
      const int f = f0 + j;
      const int k = nb0 + before;
      before++;
      if (k < cap) {
        stv(model, P_POS, k, R * ldv(frame, P_POS, f) + t);
        sts(model, k, rotate_sym(R, lds(frame, f)));
        sto(model, k, ldo(frame, f) * Rt);
        model.plane(P_CONF)[k] = frame.plane(P_CONF)[f];
        stv(model, P_COL, k, ldv(frame, P_COL, f));
        stv(model, P_LAB, k, ldv(frame, P_LAB, f));
        model.plane(P_STAMP)[k] = __int_as_float(stamp);
        model.plane(P_STAMP + 1)[k] = __int_as_float(stamp);
        model.plane(P_DIM)[k] = frame.plane(P_DIM)[f];
        model.plane(P_DIM + 1)[k] = frame.plane(P_DIM + 1)[f];
      }
    }
    __syncthreads();
    if (tid == 0) running += total;
    __syncthreads();
  }
  if (tid == 0) {
    const int want = nb0 + running;
    const int nb = want < cap ? want : cap;
    counters->nb_supersurfels = nb;
    counters->nb_inserted = nb - nb0;
    *skip_filter = 0;
  }
}

// filterModel (supersurfel_fusion_kernels.cu:397-467) + per-CTA state histogram for
// the partition.  1024 model supersurfels per CTA.
constexpr int PART_THREADS = 256;
constexpr int PART_ITEMS = 4;
constexpr int PART_CHUNK = PART_THREADS * PART_ITEMS;

__device__ __forceinline__ int cull_state(SurfelSet& model, int i, const M3& Rv, V3 tv, const int2* lmap, CamK cam,
                                          int stamp, int delta_t, float conf_thresh, float z_min, float z_max) {
  const int last = __float_as_int(model.plane(P_STAMP + 1)[i]);
  const int age = stamp - last;
  const float conf = model.plane(P_CONF)[i];
  if ((age > delta_t && conf < conf_thresh && stamp > delta_t) || conf <= 0.0f) {
    model.plane(P_CONF)[i] = -1.0f;
    return 2;
  }
  const V3 p = Rv * ldv(model, P_POS, i) + tv;
  if (!(p.z > z_min && p.z < z_max)) return 1;
  const float u = cam.fx * p.x / p.z + cam.cx;
  const float v = cam.fy * p.y / p.z + cam.cy;
  if (!(u >= 0.0f && u < (float)cam.W && v >= 0.0f && v < (float)cam.H)) return 1;
  const float z = __int_as_float(lmap[(size_t)tex_coord(v, cam.H) * cam.W + tex_coord(u, cam.W)].y);
  if (p.z < 0.8f * z) {
    model.plane(P_CONF)[i] = -1.0f;
    return 2;
  }
  return 0;
}

__global__ void __launch_bounds__(PART_THREADS) cull_kernel(SurfelSet model, int* states, int* block_hist,
                                                            const int2* lmap, const DevicePose* pose,
                                                            const Counters* counters, const int* skip_filter,
                                                            CamK cam, int delta_t, float conf_thresh, float z_min,
                                                            float z_max) {
  pdl_sync();
  if (*skip_filter) return;
  const int n = counters->nb_supersurfels;
  const int base = blockIdx.x * PART_CHUNK;
  if (base >= n) return;
  const M3 R = pose_R(pose);
  const M3 Rv = transpose(R);
  const V3 tv = -(Rv * pose_t(pose));
  const int stamp = counters->stamp;
  int c0 = 0, c1 = 0, c2 = 0;
  for (int j = 0; j < PART_ITEMS; j++) {
    const int i = base + j * PART_THREADS + threadIdx.x;
    if (i < n) {
      const int s = cull_state(model, i, Rv, tv, lmap, cam, stamp, delta_t, conf_thresh, z_min, z_max);
      states[i] = s;
      c0 += (s == 0); c1 += (s == 1); c2 += (s == 2);
    }
  }
  __shared__ int h[3];
  if (threadIdx.x < 3) h[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c2 += __shfl_xor_sync(0xffffffffu, c2, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&h[0], c0); atomicAdd(&h[1], c1); atomicAdd(&h[2], c2); }
  __syncthreads();
  if (threadIdx.x < 3) block_hist[blockIdx.x * 4 + threadIdx.x] = h[threadIdx.x];
}

// Exclusive scan of the per-CTA histograms -> per-CTA output offsets for each state,
// and the new counters (supersurfel_fusion.cu:463-475).  One CTA.
__global__ void __launch_bounds__(1024) partition_scan_kernel(int* block_hist, Counters* counters,
                                                               const int* skip_filter) {
  pdl_sync();
  if (*skip_filter) return;
  const int n = counters->nb_supersurfels;
  const int nblocks = (n + PART_CHUNK - 1) / PART_CHUNK;
  __shared__ int tot[3];
  __shared__ int wsum[32][3];
  __shared__ int carry[3];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < 3) { tot[tid] = 0; carry[tid] = 0; }
  __syncthreads();
  // pass 1: totals
  int t0 = 0, t1 = 0, t2 = 0;
  for (int b = tid; b < nblocks; b += 1024) { t0 += block_hist[4 * b]; t1 += block_hist[4 * b + 1]; t2 += block_hist[4 * b + 2]; }
  atomicAdd(&tot[0], t0); atomicAdd(&tot[1], t1); atomicAdd(&tot[2], t2);
  __syncthreads();
  const int start[3] = {0, tot[0], tot[0] + tot[1]};
  // pass 2: exclusive scan per state, chunk by chunk
  for (int base = 0; base < nblocks; base += 1024) {
    const int b = base + tid;
    int v[3] = {0, 0, 0};
    if (b < nblocks) { v[0] = block_hist[4 * b]; v[1] = block_hist[4 * b + 1]; v[2] = block_hist[4 * b + 2]; }
    int inc[3] = {v[0], v[1], v[2]};
#pragma unroll
    for (int s = 0; s < 3; s++) {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc[s], o);
        if (lane >= o) inc[s] += u;
      }
      if (lane == 31) wsum[wid][s] = inc[s];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
      for (int s = 0; s < 3; s++) {
        int w = wsum[lane][s];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, w, o);
          if (lane >= o) w += u;
        }
        wsum[lane][s] = w;
      }
    }
    __syncthreads();
    if (b < nblocks) {
#pragma unroll
      for (int s = 0; s < 3; s++) {
        const int excl = carry[s] + (wid > 0 ? wsum[wid - 1][s] : 0) + inc[s] - v[s];
        block_hist[4 * b + s] = start[s] + excl;
      }
    }
    __syncthreads();
    if (tid < 3) carry[tid] += wsum[31][tid];
    __syncthreads();
  }
  if (tid == 0) {
    counters->nb_visible = tot[0];
    counters->nb_removed = tot[2];
    counters->nb_supersurfels = n - tot[2];
    counters->pad = n;  // pre-compaction length, consumed by the scatter/copy kernels
  }
}

// Stable scatter into the alternate buffer.
__global__ void __launch_bounds__(PART_THREADS) partition_scatter_kernel(SurfelSet src, SurfelSet dst,
                                                                         const int* states, const int* block_off,
                                                                         const Counters* counters,
                                                                         const int* skip_filter) {
  pdl_sync();
  if (*skip_filter) return;
  const int n = counters->pad;
  const int base = blockIdx.x * PART_CHUNK;
  if (base >= n) return;
  __shared__ int wcount[PART_THREADS / 32][3];
  __shared__ int run[3];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < 3) run[tid] = block_off[blockIdx.x * 4 + tid];
  __syncthreads();
  for (int j = 0; j < PART_ITEMS; j++) {
    const int i = base + j * PART_THREADS + tid;
    const int s = (i < n) ? states[i] : -1;
    unsigned m[3];
#pragma unroll
    for (int q = 0; q < 3; q++) m[q] = __ballot_sync(0xffffffffu, s == q);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 3; q++) wcount[wid][q] = __popc(m[q]);
    }
    __syncthreads();
    if (s >= 0) {
      int off = run[s];
      for (int w = 0; w < wid; w++) off += wcount[w][s];
      off += __popc(m[s] & ((1u << lane) - 1u));
      float row[P_COUNT];
#pragma unroll
      for (int p = 0; p < P_COUNT; p++) row[p] = src.plane(p)[i];
#pragma unroll
      for (int p = 0; p < P_COUNT; p++) dst.plane(p)[off] = row[p];
    }
    __syncthreads();
    if (tid < 3) {
      int add = 0;
      for (int w = 0; w < PART_THREADS / 32; w++) add += wcount[w][tid];
      run[tid] += add;
    }
    __syncthreads();
  }
}

__global__ void partition_copyback_kernel(SurfelSet src, SurfelSet dst, Counters* counters, const int* skip_filter,
                                          const DevicePose* pose, const IcpState* icp, FrameReport* report, int advance) {
  pdl_sync();
  // the counters are final since partition_scan_kernel: the frame's report rides on this launch
  if (report && blockIdx.x == 0 && threadIdx.x == 0) frame_report(counters, pose, icp, report, advance);
  if (*skip_filter) return;
  const int n = counters->pad;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float row[P_COUNT];
#pragma unroll
  for (int p = 0; p < P_COUNT; p++) row[p] = src.plane(p)[i];
#pragma unroll
  for (int p = 0; p < P_COUNT; p++) dst.plane(p)[i] = row[p];
}

// ------------------------------------------------------- layout conversion etc.
struct Members { float* pos; float* col; int* stamps; float* ori; float* shape; float* dims; float* conf; };

__global__ void pack_kernel(SurfelSet set, Members d, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (d.pos) for (int k = 0; k < 3; k++) d.pos[3 * (size_t)i + k] = set.plane(P_POS + k)[i];
  if (d.col) for (int k = 0; k < 3; k++) d.col[3 * (size_t)i + k] = set.plane(P_COL + k)[i];
  if (d.stamps) for (int k = 0; k < 2; k++) d.stamps[2 * (size_t)i + k] = __float_as_int(set.plane(P_STAMP + k)[i]);
  if (d.ori) for (int k = 0; k < 9; k++) d.ori[9 * (size_t)i + k] = set.plane(P_ORI + k)[i];
  if (d.shape) for (int k = 0; k < 6; k++) d.shape[6 * (size_t)i + k] = set.plane(P_SHAPE + k)[i];
  if (d.dims) for (int k = 0; k < 2; k++) d.dims[2 * (size_t)i + k] = set.plane(P_DIM + k)[i];
  if (d.conf) d.conf[i] = set.plane(P_CONF)[i];
}

__global__ void unpack_kernel(Members s, SurfelSet set, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (s.pos) for (int k = 0; k < 3; k++) set.plane(P_POS + k)[i] = s.pos[3 * (size_t)i + k];
  if (s.col) for (int k = 0; k < 3; k++) set.plane(P_COL + k)[i] = s.col[3 * (size_t)i + k];
  if (s.stamps) for (int k = 0; k < 2; k++) set.plane(P_STAMP + k)[i] = __int_as_float(s.stamps[2 * (size_t)i + k]);
  if (s.ori) for (int k = 0; k < 9; k++) set.plane(P_ORI + k)[i] = s.ori[9 * (size_t)i + k];
  if (s.shape) for (int k = 0; k < 6; k++) set.plane(P_SHAPE + k)[i] = s.shape[6 * (size_t)i + k];
  if (s.dims) for (int k = 0; k < 2; k++) set.plane(P_DIM + k)[i] = s.dims[2 * (size_t)i + k];
  if (s.conf) set.plane(P_CONF)[i] = s.conf[i];
}

// applyTransformSuperSurfel (supersurfel_fusion_kernels.cu:469-488)
__global__ void transform_model_kernel(SurfelSet model, const Counters* counters, DevicePose tf) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counters->nb_supersurfels) return;
  if (model.plane(P_CONF)[i] <= 0.0f) return;      // supersurfel_fusion_kernels.cu:481
  const M3 R = m3(v3(tf.R[0], tf.R[1], tf.R[2]), v3(tf.R[3], tf.R[4], tf.R[5]), v3(tf.R[6], tf.R[7], tf.R[8]));
  const V3 t = v3(tf.t[0], tf.t[1], tf.t[2]);
  stv(model, P_POS, i, R * ldv(model, P_POS, i) + t);
  sto(model, i, ldo(model, i) * transpose(R));
  sts(model, i, rotate_sym(R, lds(model, i)));
}

// ---- consumers of the model ---------------------------
// applyDeformation (core/src/deformation_graph_kernels.cu:27-73): warps every model
// supersurfel by the embedded deformation graph -- blend of its four nearest nodes' rigid
// motions, rotations blended as weighted quaternions -- position, orientation and shape.
// quatToRotMat / rotMatToQuat are restated with the reference's own quirks
// (matrix_math.cuh:512-585: `wy` is computed as w*z, and the non-positive-trace branch picks
// index 2 when m22 exceeds EITHER other diagonal entry).
__device__ __forceinline__ float4 rot_to_quat(const M3& m) {
  float4 q;
  float s;
  const float trace = m.r0.x + m.r1.y + m.r2.z;
  if (trace > 0) {
    s = sqrtf(trace + 1);
    q.w = 0.5f * s;
    s = 0.5f / s;
    q.x = (m.r2.y - m.r1.z) * s;
    q.y = (m.r0.z - m.r2.x) * s;
    q.z = (m.r1.x - m.r0.y) * s;
  } else {
    int i = 0;
    if (m.r1.y > m.r0.x) i = 1;
    if (m.r2.z > m.r0.x || m.r2.z > m.r1.y) i = 2;
    if (i == 0) {
      s = sqrtf(1.0f + m.r0.x - m.r1.y - m.r2.z);
      q.x = 0.5f * s; s = 0.5f / s;
      q.w = (m.r2.y - m.r1.z) * s; q.y = (m.r0.y + m.r1.x) * s; q.z = (m.r0.z + m.r2.x) * s;
    } else if (i == 1) {
      s = sqrtf(1.0f + m.r1.y - m.r0.x - m.r2.z);
      q.y = 0.5f * s; s = 0.5f / s;
      q.w = (m.r0.z - m.r2.x) * s; q.x = (m.r0.y + m.r1.x) * s; q.z = (m.r1.z + m.r2.y) * s;
    } else {
      s = sqrtf(1.0f + m.r2.z - m.r0.x - m.r1.y);
      q.z = 0.5f * s; s = 0.5f / s;
      q.w = (m.r1.x - m.r0.y) * s; q.x = (m.r0.z + m.r2.x) * s; q.y = (m.r1.z + m.r2.y) * s;
    }
  }
  return q;
}
__device__ __forceinline__ M3 quat_to_rot(float4 q) {
  const float x2 = q.x * q.x, y2 = q.y * q.y, z2 = q.z * q.z;
  const float xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z;
  const float wx = q.w * q.x, wy = q.w * q.z /* sic, matrix_math.cuh:521 */, wz = q.w * q.z;
  return m3(v3(1.0f - 2.0f * (y2 + z2), 2.0f * (xy - wz), 2.0f * (xz + wy)),
            v3(2.0f * (xy + wz), 1.0f - 2.0f * (x2 + z2), 2.0f * (yz - wx)),
            v3(2.0f * (xz - wy), 2.0f * (yz + wx), 1.0f - 2.0f * (x2 + y2)));
}

__global__ void apply_deformation_kernel(SurfelSet model, const float* __restrict__ node_pos,
                                         const float* __restrict__ node_rot, const float* __restrict__ node_trans,
                                         const float4* __restrict__ weights, const int4* __restrict__ nn, int n, int n_nodes) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 w4 = weights[i];
  const int4 id4 = nn[i];
  const float w[4] = {w4.x, w4.y, w4.z, w4.w};
  const int id[4] = {id4.x, id4.y, id4.z, id4.w};
  const V3 pi = ldv(model, P_POS, i);
  V3 po = v3(0.f, 0.f, 0.f);
  float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int node = min(max(id[k], 0), n_nodes - 1);      // the reference indexes unchecked
    const V3 gk = v3(node_pos[3 * node], node_pos[3 * node + 1], node_pos[3 * node + 2]);
    const V3 tk = v3(node_trans[3 * node], node_trans[3 * node + 1], node_trans[3 * node + 2]);
    const float* r = node_rot + 9 * node;
    const M3 Rk = m3(v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), v3(r[6], r[7], r[8]));
    const float4 qk = rot_to_quat(Rk);
    po = po + w[k] * (Rk * (pi - gk) + gk + tk);
    bq.x += w[k] * qk.x; bq.y += w[k] * qk.y; bq.z += w[k] * qk.z; bq.w += w[k] * qk.w;
  }
  const float len = sqrtf(bq.x * bq.x + bq.y * bq.y + bq.z * bq.z + bq.w * bq.w);
  bq.x /= len; bq.y /= len; bq.z /= len; bq.w /= len;
  const M3 av = quat_to_rot(bq);
  sto(model, i, ldo(model, i) * transpose(av));
  sts(model, i, rotate_sym(av, lds(model, i)));
  stv(model, P_POS, i, po);
}

// The triangle list the node publishes for rviz (node/supersurfel_fusion_node.cpp:303-413):
// per supersurfel a quad of half-extents 3 sqrt(dims) along e1 / e2 as two triangles
// (p0 p1 p2, p0 p2 p3) and its colour / 255; below the confidence threshold six zero points and
// a black colour.  The node copies five arrays to the host and loops with OpenMP; here the
// geometry is produced next to the model.
__global__ void marker_kernel(SurfelSet set, int n, float conf_thresh, float* __restrict__ points, float* __restrict__ colors) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* P = points + (size_t)i * 18;
  float* C = colors + (size_t)i * 24;
  if (set.plane(P_CONF)[i] > conf_thresh) {
    float v0 = 3.0f * sqrtf(set.plane(P_DIM)[i]);
    float v1 = 3.0f * sqrtf(set.plane(P_DIM + 1)[i]);
    const V3 e0 = ldv(set, P_ORI, i), e1 = ldv(set, P_ORI + 3, i);
    V3 pos = ldv(set, P_POS, i);
    if (!isfinite(v0)) v0 = 0;
    if (!isfinite(v1)) v1 = 0;
    if (!isfinite(pos.x) || !isfinite(pos.y) || !isfinite(pos.z)) pos = v3(0.f, 0.f, 0.f);
    const V3 a = v0 * e0, b = v1 * e1;
    const V3 p0 = v3(pos.x + a.x + b.x, pos.y + a.y + b.y, pos.z + a.z + b.z);
    const V3 p1 = v3(pos.x + a.x - b.x, pos.y + a.y - b.y, pos.z + a.z - b.z);
    const V3 p2 = v3(pos.x - a.x - b.x, pos.y - a.y - b.y, pos.z - a.z - b.z);
    const V3 p3 = v3(pos.x - a.x + b.x, pos.y - a.y + b.y, pos.z - a.z + b.z);
    const V3 tri[6] = {p0, p1, p2, p0, p2, p3};
    const V3 col = ldv(set, P_COL, i);
#pragma unroll
    for (int k = 0; k < 6; k++) {
      P[3 * k] = tri[k].x; P[3 * k + 1] = tri[k].y; P[3 * k + 2] = tri[k].z;
      C[4 * k] = col.x / 255; C[4 * k + 1] = col.y / 255; C[4 * k + 2] = col.z / 255; C[4 * k + 3] = 1.f;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 18; k++) P[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) { C[4 * k] = 0.f; C[4 * k + 1] = 0.f; C[4 * k + 2] = 0.f; C[4 * k + 3] = 1.f; }
  }
}

// extractLocalPointCloudKernel (supersurfel_fusion_kernels.cu:490-520); output order is
// by atomic ticket, as in the reference.
__global__ void local_cloud_kernel(SurfelSet model, Counters* counters, const DevicePose* pose, float conf_thresh,
                                   float radius, float* out_pos, float* out_nrm, int capacity) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counters->nb_supersurfels) return;
  if (!(model.plane(P_CONF)[i] >= conf_thresh)) return;
  const M3 Rv = transpose(pose_R(pose));
  const V3 tv = -(Rv * pose_t(pose));
  const V3 p = Rv * ldv(model, P_POS, i) + tv;
  if (!(length(p) < radius)) return;
  const int id = atomicAdd(&counters->cloud_count, 1);
  if (id >= capacity) return;
  const V3 nrm = normalize(Rv * ldv(model, P_ORI + 6, i));
  out_pos[3 * id] = p.x; out_pos[3 * id + 1] = p.y; out_pos[3 * id + 2] = p.z;
  out_nrm[3 * id] = nrm.x; out_nrm[3 * id + 1] = nrm.y; out_nrm[3 * id + 2] = nrm.z;
}

// renderBoundaryImage_kernel (TPS_RGBD_kernels.cu:616-643)
__global__ void preview_kernel(uint8_t* out, const uchar4* rgba, const int* labels, int W, int H) {
  pdl_sync();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const size_t p = (size_t)y * W + x;
  const int index = labels[p];
  if (x < W - 1 && y < H - 1 && (labels[p + 1] != index || labels[p + W + 1] != index)) {
    out[3 * p] = 255; out[3 * p + 1] = 255; out[3 * p + 2] = 255;
  } else {
    const uchar4 c = rgba[p];
    out[3 * p] = (uint8_t)(0.8f * c.z); out[3 * p + 1] = (uint8_t)(0.8f * c.y); out[3 * p + 2] = (uint8_t)(0.8f * c.x);
  }
}

// ------------------------------------------------------------------- launchers
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

void launch_extract(Engine* e) {
  const CamK cam = cam_of(e);
  dim3 blk(32, 8), grd(cdiv(e->W, 32), cdiv(e->H, 8));
  launch_pdl(e, extract_accumulate_kernel, dim3(grd), dim3(blk), 0, e->lmap, e->inliers, e->bound, e->rgba, e->xsums, cam);
  launch_pdl(e, extract_finalize_kernel, dim3(cdiv(e->S, 128)), dim3(128), 0, e->frame, e->ftab, e->xsums, e->matched, e->best,
                                                                 e->cfg.range_min, e->cfg.range_max, e->counters,
                                                                 e->S);
  e->launches += 2;
}

void launch_frame_tables(Engine* e) {
  launch_pdl(e, frame_tables_kernel, dim3(cdiv(e->S, 128)), dim3(128), 0, e->frame, e->ftab, e->matched, e->best, e->S);
  e->launches++;
}

void launch_model_lab(Engine* e, int n) {
  if (n <= 0) return;
  launch_pdl(e, model_lab_kernel, dim3(cdiv(n, 256)), dim3(256), 0, e->model, n);
  e->launches++;
}

void launch_build_lmap(Engine* e, const float* slanted_dev) {
  launch_pdl(e, build_lmap_kernel, dim3((unsigned)((e->npix + 255) / 256)), dim3(256), 0, e->lmap, e->labels, slanted_dev, e->npix);
  e->launches++;
}

void launch_invalidate(Engine* e, const uint8_t* mask_dev) {
  launch_pdl(e, invalidate_kernel, dim3(cdiv(e->S, 128)), dim3(128), 0, e->frame, e->ftab, mask_dev, e->S);
  e->launches++;
}

void launch_fuse(Engine* e, FrameReport* report, int advance) {
  const CamK cam = cam_of(e);
  int* skip = e->scan_tmp;            // [0] skip flag, block histograms from [4]
  int* hist = e->scan_tmp + 4;
  const int cap_blocks = cdiv(e->cap, PART_CHUNK);
  launch_pdl(e, associate_kernel, dim3(cdiv(e->cap, 256) < 1184 ? cdiv(e->cap, 256) : 1184), dim3(256), 0, 
      e->model, e->frame, e->ftab, e->lmap, e->matched, e->best, e->pose, e->counters, cam, e->cfg.range_min,
      e->cfg.range_max);
  launch_pdl(e, update_kernel, dim3(cdiv(e->S, 128)), dim3(128), 0, e->model, e->frame, e->matched, e->best, e->pose,
                                                        e->counters, e->S);
  launch_pdl(e, insert_kernel, dim3(1), dim3(INS_THREADS), 0, e->model, e->frame, e->matched, e->pose, e->counters, skip, e->S,
                                                  e->cap);
  launch_pdl(e, cull_kernel, dim3(cap_blocks), dim3(PART_THREADS), 0, e->model, e->states, hist, e->lmap, e->pose, e->counters,
                                                          skip, cam, e->cfg.delta_t, e->cfg.conf_thresh,
                                                          e->cfg.range_min, e->cfg.range_max);
  launch_pdl(e, partition_scan_kernel, dim3(1), dim3(1024), 0, hist, e->counters, skip);
  launch_pdl(e, partition_scatter_kernel, dim3(cap_blocks), dim3(PART_THREADS), 0, e->model, e->model_alt, e->states, hist,
                                                                       e->counters, skip);
  launch_pdl(e, partition_copyback_kernel, dim3(cdiv(e->cap, 256)), dim3(256), 0, e->model_alt, e->model, e->counters, skip,
             e->pose, e->icp, report, advance);
  e->launches += 7;
}

static Members members_of(const SsfSurfels& s) {
  Members m; m.pos = s.positions; m.col = s.colors; m.stamps = s.stamps; m.ori = s.orientations;
  m.shape = s.shapes; m.dims = s.dims; m.conf = s.confidences; return m;
}

void launch_pack(Engine* e, const SurfelSet& set, int n, const SsfSurfels& dst_dev) {
  if (n <= 0) return;
  launch_pdl(e, pack_kernel, dim3(cdiv(n, 256)), dim3(256), 0, set, members_of(dst_dev), n);
  e->launches++;
}

void launch_unpack(Engine* e, const SsfSurfels& src_dev, int n, const SurfelSet& set) {
  if (n <= 0) return;
  launch_pdl(e, unpack_kernel, dim3(cdiv(n, 256)), dim3(256), 0, members_of(src_dev), set, n);
  e->launches++;
}

void launch_transform_model(Engine* e, const float* R, const float* t) {
  DevicePose tf;
  for (int i = 0; i < 9; i++) tf.R[i] = R[i];
  for (int i = 0; i < 3; i++) tf.t[i] = t[i];
  launch_pdl(e, transform_model_kernel, dim3(cdiv(e->cap, 256)), dim3(256), 0, e->model, e->counters, tf);
  e->launches++;
}

void launch_apply_deformation(Engine* e, const float* node_pos, const float* node_rot, const float* node_trans,
                              const float* weights, const int* nn, int n, int n_nodes) {
  launch_pdl(e, apply_deformation_kernel, dim3(cdiv(n, 128)), dim3(128), 0, e->model, node_pos, node_rot, node_trans,
             reinterpret_cast<const float4*>(weights), reinterpret_cast<const int4*>(nn), n, n_nodes);
  e->launches++;
}

void launch_markers(Engine* e, bool frame, int n, float conf_thresh, float* points_dev, float* colors_dev) {
  launch_pdl(e, marker_kernel, dim3(cdiv(n, 128)), dim3(128), 0, frame ? e->frame : e->model, n, conf_thresh, points_dev,
             colors_dev);
  e->launches++;
}

void launch_local_cloud(Engine* e, float radius, float* pos_dev, float* nrm_dev, int capacity) {
  launch_pdl(e, local_cloud_kernel, dim3(cdiv(e->cap, 256)), dim3(256), 0, e->model, e->counters, e->pose, e->cfg.conf_thresh,
                                                               radius, pos_dev, nrm_dev, capacity);
  e->launches++;
}

void launch_preview(Engine* e, uint8_t* bgr_dev) {
  dim3 blk(32, 8), grd(cdiv(e->W, 32), cdiv(e->H, 8));
  launch_pdl(e, preview_kernel, dim3(grd), dim3(blk), 0, bgr_dev, e->rgba, e->labels, e->W, e->H);
  e->launches++;
}

}  // namespace ssf
