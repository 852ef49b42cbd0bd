// libssf C-ABI (include/ssf.h): engine lifetime, the per-frame sequence and the stage
// entry points.  Host logic mirrors SupersurfelFusion::initialize / processFrame
// (reference: core/src/supersurfel_fusion.cu:49-164, 166-530) with the whole frame
// enqueued on one stream and replayed as a CUDA graph: no per-frame allocation, no
// intermediate device synchronisation, one small read-back at the end of the frame.
#include "ssf_engine.h"
#include "ssf_math.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>

namespace ssf {

void launch_icp_set_transform(Engine* e, const float* R, const float* t);
size_t tps_rng_state_bytes();
bool tps_make_label_map(void* map128, const int* labels, int W, int H);
void tps_configure();
size_t tps_trace_bytes(int grid);
int tps_persistent_grid(int device, int gx, int gy, int cell, int height, int nb_iters, int* cache_slots);

__global__ void frame_end_kernel(Counters* counters, const DevicePose* pose, const IcpState* icp, FrameReport* rep,
                                 int advance) {
  pdl_sync();
  frame_report(counters, pose, icp, rep, advance);
}

// end of the segmentation stage of the pipelined mode: the next frame to be segmented is stamp + 1
__global__ void seg_end_kernel(Counters* counters) {
  pdl_sync();
  counters->seg_stamp += 1;
}

__global__ void split_lmap_kernel(const int2* lmap, float* slanted, size_t n) {
  pdl_sync();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) slanted[i] = __int_as_float(lmap[i].y);
}

struct EngineImpl : public SsfEngine {
  // pipelined mode (ssf_submit_frame / ssf_wait_frame): the frame's launch sequence is cut into
  // nb_stages contiguous stages (stage_first[p] .. stage_first[p+1]); per slot and stage one CUDA
  // graph (two for the stage that holds the ingest: with / without the bilateral filter)
  int stage_first[SSF_SLOTS + 1];
  // variant of a stage graph: bit 0 = bilateral ingest (first stage), bit 1 = one-launch registration (last stage)
  cudaGraphExec_t stage_graph[SSF_SLOTS][SSF_SLOTS][4];
  bool stage_ready[SSF_SLOTS][SSF_SLOTS][4];
  uint64_t stage_launches[SSF_SLOTS][SSF_SLOTS][4];
  cudaEvent_t ev_stage[SSF_SLOTS][SSF_SLOTS], ev_done[SSF_SLOTS], ev_t0[SSF_SLOTS], ev_t1[SSF_SLOTS], ev_copy[SSF_SLOTS];
  FrameReport* d_report2[SSF_SLOTS];
  FrameReport* h_report2[SSF_SLOTS];
  float* h_prior2[SSF_SLOTS];
  int pipe_next;     // slot of the next submitted frame
  int pipe_oldest;   // slot of the oldest frame in flight
  int in_flight;
  uint64_t launches_per_frame[8];
  cudaEvent_t ev_tm[6];            // stage boundaries of a frame run with SSF_FLAG_STAGE_TIMING
  FrameReport* d_report;
  FrameReport* h_report;
  float* h_prior;          // pinned 12 floats
  // bit 0: SSF_FLAG_BILATERAL, bit 1: SSF_FLAG_STAGE_TIMING (event nodes at stage boundaries), bit 2: one-launch registration
  cudaGraphExec_t graph_exec[8];
  bool graph_ready[8];
  bool use_graph;
  bool created;
};

static inline int round4(int n) { return (n + 3) & ~3; }

template <typename T>
static cudaError_t dalloc(T** p, size_t count) {
  cudaError_t err = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
  if (err == cudaSuccess) err = cudaMemset(*p, 0, count * sizeof(T));
  return err;
}

// the reference's filter arguments (supersurfel_fusion.cu:180)
static const int kBilateralKernel = -1;
static const float kBilateralSigmaColor = 0.03f, kBilateralSigmaSpatial = 4.5f;

// point the engine at one of the frame slots
static void select_slot(EngineImpl* e, int s) {
  const FrameSlot& f = e->slot[s];
  e->cur_slot = s;
  e->in_rgb = f.in_rgb; e->in_depth = f.in_depth;
  e->rgba = f.rgba; e->disp = f.disp; e->labels = f.labels; e->bound = f.bound; e->inliers = f.inliers;
  e->sp = f.sp; e->sums = f.sums;
  e->lmap = f.lmap; e->frame = f.frame; e->ftab = f.ftab; e->matched = f.matched; e->best = f.best;
}

// The frame as a sequence of steps: 0 = ingest, 1 .. T = the segmentation steps (tps_step_count),
// T + 1 = extraction, T + 2 = registration + fusion.  enqueue_steps enqueues [g0, g1).
static void enqueue_track(EngineImpl* e, FrameReport* report, int advance, bool marks, bool small);

// NVTX range on the host side of the enqueue (SSF_NVTX=1): shows the stage structure of a frame
// in a timeline next to the kernels; under graph replay only the capture carries the ranges.
struct StageRange {
  bool on;
  StageRange(const EngineImpl* e, const char* name) : on(e->nvtx != 0) { if (on) nvtxRangePushA(name); }
  ~StageRange() { if (on) nvtxRangePop(); }
};
// stage boundary k of a frame run with SSF_FLAG_STAGE_TIMING (an event-record node under capture)
static void stage_mark(EngineImpl* e, bool marks, int k) {
  if (!marks) return;
  // under stream capture a plain cudaEventRecord only captures a dependency; the External flag makes
  // it an event-record NODE of the graph, which is what gives a timestamp at replay
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(e->stream, &st);
  cudaEventRecordWithFlags(e->ev_tm[k], e->stream,
                           st == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
}

static void enqueue_steps(EngineImpl* e, int g0, int g1, bool bilateral, bool pipelined, FrameReport* report,
                          bool marks = false, bool small = false) {
  const int T = tps_step_count(e);
  if (g0 <= 0) stage_mark(e, marks, 0);
  if (g0 <= 0 && g1 > 0) {
    StageRange r(e, "ssf:ingest");
    const float* depth = e->in_depth;
    if (bilateral) {
      launch_bilateral(e, e->in_depth, e->depth_f, kBilateralKernel, kBilateralSigmaColor, kBilateralSigmaSpatial);
      depth = e->depth_f;
    }
    launch_ingest(e, e->in_rgb, (size_t)e->W * 3, depth, (size_t)e->W * 4);
    stage_mark(e, marks, 1);
  }
  const int t0 = (g0 > 1 ? g0 : 1) - 1, t1 = (g1 < T + 1 ? g1 : T + 1) - 1;
  if (t1 > t0) {
    StageRange r(e, "ssf:segmentation");
    launch_tps(e, t0, t1);
    if (t1 == T) stage_mark(e, marks, 2);
  }
  if (g0 <= T + 1 && g1 > T + 1) {
    StageRange r(e, "ssf:extraction");
    launch_extract(e);
    stage_mark(e, marks, 3);
    if (pipelined) {     // the next frame to be extracted carries the next stamp
      launch_pdl(e, seg_end_kernel, dim3(1), dim3(1), 0, e->counters);
      e->launches++;
    }
  }
  if (g0 <= T + 2 && g1 > T + 2) enqueue_track(e, report, pipelined ? 1 : 3, marks, small);
}

// Cut the frame's steps (0 = ingest, 1 .. T = segmentation steps, T + 1 = extraction, T + 2 =
// registration + fusion) into `stages` contiguous groups minimising the heaviest group; weights ~
// kernel launches of a step, the quantity that sets a stage's duration at VGA.  Pure function of
// the configuration (also behind ssf_plan_pipeline for the CPU tests); writes first[0 .. used]
// and returns the number of stages used.
// Cost of frame step g in microseconds on a B200 at 640x480 (round-2 measurements: SSF_FLAG_STAGE_TIMING gives
// ingest 13, segmentation 303, extraction 29, registration 49, fusion 47 us per frame; inside the segmentation the
// fused relabelling passes run ~5 us (colour) / ~6.5 us (colour + disparity) each, four per iteration,
// profiles/launches_r2_*.txt).  Only the ratios matter; at other frame sizes they stay roughly the same.
static int step_weight(int g, int seg_iter, int icp_iter, int persistent) {
  const int T = persistent ? 1 : seg_iter + 2, half = seg_iter / 2;
  if (g == 0) return 13;                                  // ingest
  if (g <= T) {
    const int t = g - 1;
    if (persistent) return 300;
    if (t < half) return 20;                              // colour-only iteration
    if (t == half) return 40;                             // RANSAC + inlier moments
    if (t <= seg_iter) return 26;                         // colour + disparity iteration
    return 21;                                            // smoothing + render
  }
  if (g == T + 1) return 29;                              // extraction
  return 66 + 3 * icp_iter;                               // registration + fusion
}

static int plan_stages_impl(int seg_iter, int icp_iter, int persistent, int stages, int* first) {
  const int T = persistent ? 1 : seg_iter + 2, G = T + 3;
  if (stages < 1) stages = 1;
  if (stages > SSF_SLOTS) stages = SSF_SLOTS;
  if (stages > G) stages = G;
  if (G > 64) { first[0] = 0; first[1] = G; return 1; }      // absurd iteration counts: no pipelining
  int w[64];
  for (int g = 0; g < G; g++) w[g] = step_weight(g, seg_iter, icp_iter, persistent);
  // dynamic programme: best[p][g] = minimal heaviest group when the first g steps form p groups
  static const int INF = 1 << 28;
  int best[SSF_SLOTS + 1][65], cut[SSF_SLOTS + 1][65];
  for (int p = 0; p <= stages; p++)
    for (int g = 0; g <= G; g++) { best[p][g] = INF; cut[p][g] = 0; }
  best[0][0] = 0;
  for (int p = 1; p <= stages; p++)
    for (int g = p; g <= G; g++) {
      int sum = 0;
      for (int k = g - 1; k >= p - 1; k--) {
        sum += w[k];
        const int cand = best[p - 1][k] > sum ? best[p - 1][k] : sum;
        if (cand < best[p][g]) { best[p][g] = cand; cut[p][g] = k; }
      }
    }
  int g = G;
  for (int p = stages; p >= 1; p--) { first[p] = g; g = cut[p][g]; }
  first[0] = 0;
  return stages;
}

static void plan_stages(EngineImpl* e, int stages) {
  e->nb_stages = plan_stages_impl(e->cfg.seg_iter, e->cfg.icp_iter, e->tps_persistent, stages, e->stage_first);
}

static void enqueue_seg(EngineImpl* e, bool bilateral, bool marks) {
  enqueue_steps(e, 0, tps_step_count(e) + 2, bilateral, false, nullptr, marks);
}

// stage C: registration + fusion (reads the slot's hand-over set, owns pose / model / counters)
static void enqueue_track(EngineImpl* e, FrameReport* report, int advance, bool marks, bool small) {
  {
    StageRange r(e, "ssf:registration");
    if (small) {
      launch_icp_registration_loop(e, true);      // begin + Gauss-Newton loop + finish: one cluster launch
    } else {
      launch_icp_begin_from_pose(e);
      launch_icp_loop(e);
      launch_icp_finish(e, true);
    }
    stage_mark(e, marks, 4);
  }
  StageRange r(e, "ssf:fusion");
  launch_fuse(e, report, advance);          // the last kernel of the update also writes the frame's report
  stage_mark(e, marks, 5);
}

static void enqueue_frame(EngineImpl* e, bool bilateral, bool marks, bool small) {
  enqueue_seg(e, bilateral, marks);
  enqueue_track(e, e->d_report, 3, marks, small);
}

// Whether a registration over about `n_visible` model supersurfels should take the one-launch path.
// Purely a performance choice: where it is allowed at all (icp_loop_equivalent) the two paths give
// identical bits, so a stale estimate only costs time.
static bool pick_small_icp(const EngineImpl* e, long long n_visible) {
  return e->icp_loop && icp_loop_equivalent(e) && n_visible <= (long long)icp_loop_max_sources();
}

// the synchronous frame graph of variant gi (bit 0 bilateral, bit 1 stage timing, bit 2 one-launch registration)
static int ensure_frame_graph(EngineImpl* e, int gi, bool upload) {
  if (e->graph_ready[gi]) return SSF_OK;
  if (e->cur_slot != 0) select_slot(e, 0);
  cudaGraph_t g;
  const uint64_t before = e->launches;
  SSF_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
  e->pdl_now = e->pdl;
  enqueue_frame(e, (gi & 1) != 0, (gi & 2) != 0, (gi & 4) != 0);
  e->pdl_now = 0;
  SSF_CUDA(e, cudaStreamEndCapture(e->stream, &g));
  e->launches_per_frame[gi] = e->launches - before;
  e->launches = before;
  SSF_CUDA(e, cudaGraphInstantiate(&e->graph_exec[gi], g, 0));
  cudaGraphDestroy(g);
  e->graph_ready[gi] = true;
  if (upload) SSF_CUDA(e, cudaGraphUpload(e->graph_exec[gi], e->stream));
  return SSF_OK;
}

static int ensure_scratch(EngineImpl* e, size_t bytes) {
  if (e->scratch_bytes >= bytes) return SSF_OK;
  if (e->scratch) cudaFree(e->scratch);
  e->scratch = nullptr;
  e->scratch_bytes = 0;
  SSF_CUDA(e, cudaMalloc(&e->scratch, bytes));
  e->scratch_bytes = bytes;
  return SSF_OK;
}

// member-layout view carved out of the scratch buffer
static SsfSurfels scratch_view(void* base, int n) {
  char* p = reinterpret_cast<char*>(base);
  SsfSurfels v;
  v.positions = reinterpret_cast<float*>(p); p += (size_t)n * 12;
  v.colors = reinterpret_cast<float*>(p); p += (size_t)n * 12;
  v.stamps = reinterpret_cast<int32_t*>(p); p += (size_t)n * 8;
  v.orientations = reinterpret_cast<float*>(p); p += (size_t)n * 36;
  v.shapes = reinterpret_cast<float*>(p); p += (size_t)n * 24;
  v.dims = reinterpret_cast<float*>(p); p += (size_t)n * 8;
  v.confidences = reinterpret_cast<float*>(p);
  return v;
}

static int copy_members(EngineImpl* e, const SsfSurfels& dst, const SsfSurfels& src, int n) {
  if (dst.positions) SSF_CUDA(e, cudaMemcpyAsync(dst.positions, src.positions, (size_t)n * 12, cudaMemcpyDefault, e->stream));
  if (dst.colors) SSF_CUDA(e, cudaMemcpyAsync(dst.colors, src.colors, (size_t)n * 12, cudaMemcpyDefault, e->stream));
  if (dst.stamps) SSF_CUDA(e, cudaMemcpyAsync(dst.stamps, src.stamps, (size_t)n * 8, cudaMemcpyDefault, e->stream));
  if (dst.orientations) SSF_CUDA(e, cudaMemcpyAsync(dst.orientations, src.orientations, (size_t)n * 36, cudaMemcpyDefault, e->stream));
  if (dst.shapes) SSF_CUDA(e, cudaMemcpyAsync(dst.shapes, src.shapes, (size_t)n * 24, cudaMemcpyDefault, e->stream));
  if (dst.dims) SSF_CUDA(e, cudaMemcpyAsync(dst.dims, src.dims, (size_t)n * 8, cudaMemcpyDefault, e->stream));
  if (dst.confidences) SSF_CUDA(e, cudaMemcpyAsync(dst.confidences, src.confidences, (size_t)n * 4, cudaMemcpyDefault, e->stream));
  return SSF_OK;
}

static int read_report(EngineImpl* e, bool advance) {
  if (e->in_flight) { e->err = "pipelined frames in flight: ssf_wait_frame first"; return SSF_ERR_STATE; }
  launch_pdl(e, frame_end_kernel, dim3(1), dim3(1), 0, e->counters, e->pose, e->icp, e->d_report, advance ? 1 : 0);
  e->launches++;
  SSF_CUDA(e, cudaMemcpyAsync(e->h_report, e->d_report, sizeof(FrameReport), cudaMemcpyDeviceToHost, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  return SSF_OK;
}

static void fill_stats(EngineImpl* e, float gpu_ms) {
  const FrameReport& r = *e->h_report;
  SsfFrameStats& s = e->stats;
  s.stamp = r.counters.stamp;
  s.nb_supersurfels = r.counters.nb_supersurfels;
  s.nb_visible = r.counters.nb_visible;
  s.nb_removed = r.counters.nb_removed;
  s.nb_matched = r.counters.nb_matched;
  s.nb_inserted = r.counters.nb_inserted;
  s.icp_ran = r.icp_active;
  s.icp_valid = r.icp_valid;
  s.icp_iters = r.icp_iters;
  s.icp_inliers = r.icp_inliers;
  s.icp_error = r.icp_error;
  s.gpu_ms = gpu_ms;
  s.ms_ingest = s.ms_segmentation = s.ms_extraction = s.ms_registration = s.ms_fusion = 0.f;
}

// per-stage device times of the frame just run with SSF_FLAG_STAGE_TIMING (what the reference prints
// per frame, supersurfel_fusion.cu:516-528)
static void fill_stage_ms(EngineImpl* e) {
  float* dst[5] = {&e->stats.ms_ingest, &e->stats.ms_segmentation, &e->stats.ms_extraction, &e->stats.ms_registration,
                   &e->stats.ms_fusion};
  for (int k = 0; k < 5; k++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->ev_tm[k], e->ev_tm[k + 1]) != cudaSuccess) { ms = 0.f; cudaGetLastError(); }
    *dst[k] = ms;
  }
}

}  // namespace ssf

using namespace ssf;

#define H_CHECK(h)                                  \
  if (!(h)) return SSF_ERR_INVALID_ARG;             \
  EngineImpl* e = static_cast<EngineImpl*>(h);      \
  if (e->failed) return SSF_ERR_STATE;              \
  cudaSetDevice(e->device)

// Every entry point that touches device state outside the pipeline (everything except
// ssf_submit_frame / ssf_wait_frame / the pure getters) refuses while frames are in flight:
// the last-stage stream owns pose, model, counters and the registration state then.
#define H_CHECK_IDLE(h)                                                          \
  H_CHECK(h);                                                                    \
  if (e->in_flight) {                                                            \
    e->err = "pipelined frames in flight: ssf_wait_frame first";                 \
    return SSF_ERR_STATE;                                                        \
  }

// first kernel-launch failure since the last check (launch_pdl records it), then the sticky error
static int launch_status(EngineImpl* e) {
  cudaError_t rc = e->launch_err;
  e->launch_err = cudaSuccess;
  if (rc == cudaSuccess) rc = cudaGetLastError();
  if (rc != cudaSuccess) {
    e->err = std::string("kernel launch: ") + cudaGetErrorString(rc);
    return SSF_ERR_CUDA;
  }
  return SSF_OK;
}
#define SSF_LAUNCH_OK(e)                 \
  do {                                   \
    const int _rc = launch_status(e);    \
    if (_rc) return _rc;                 \
  } while (0)

extern "C" {

int ssf_config_default(SsfConfig* c) {
  if (!c) return SSF_ERR_INVALID_ARG;
  memset(c, 0, sizeof(*c));
  c->cam.fx = 525.0f; c->cam.fy = 525.0f; c->cam.cx = 319.5f; c->cam.cy = 239.5f;
  c->cam.height = 480; c->cam.width = 640;
  c->cell_size = 16;
  c->lambda_pos = 50.0f; c->lambda_bound = 1000.0f; c->lambda_size = 10000.0f; c->lambda_disp = 1000000.0f;
  c->thresh_disp = 0.0001f;
  c->seg_iter = 10; c->seg_use_ransac = 1; c->nb_samples = 16;
  c->filter_iter = 4; c->filter_alpha = 0.1f; c->filter_beta = 1.0f; c->filter_threshold = 0.05f;
  c->range_min = 0.2f; c->range_max = 5.0f;
  c->delta_t = 20; c->conf_thresh = 2500.0f; c->nb_supersurfels_max = 50000;
  c->icp_iter = 10; c->icp_cov_thresh = 0.04;
  c->enable_loop_closure = 1; c->enable_mod = 1;   /* supersurfel_fusion.hpp:72-73; recorded, see ssf.h */
  return SSF_OK;
}

int ssf_create(const SsfConfig* cfg, int device, SsfHandle* out) {
  if (!cfg || !out) return SSF_ERR_INVALID_ARG;
  *out = nullptr;
  if (cfg->cam.width <= 0 || cfg->cam.height <= 0 || cfg->cell_size < 2 || cfg->nb_supersurfels_max <= 0 ||
      cfg->nb_samples <= 0 || cfg->nb_samples > 1024 || cfg->seg_iter < 0 || cfg->icp_iter < 1)
    return SSF_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return SSF_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return SSF_ERR_NO_DEVICE;
  EngineImpl* e = new (std::nothrow) EngineImpl();
  if (!e) return SSF_ERR_INVALID_ARG;
  e->cfg = *cfg;
  e->device = device;
  e->launches = 0;
  e->launch_err = cudaSuccess;
  e->failed = 0;
  e->nvtx = 0;
  if (const char* v = getenv("SSF_NVTX")) e->nvtx = atoi(v) != 0;
  for (int k = 0; k < 8; k++) e->graph_ready[k] = false;
  e->use_graph = true;
  e->created = false;
  e->W = cfg->cam.width; e->H = cfg->cam.height;
  e->npix = (size_t)e->W * e->H;
  e->gx = (e->W + cfg->cell_size - 1) / cfg->cell_size;
  e->gy = (e->H + cfg->cell_size - 1) / cfg->cell_size;
  e->S = e->gx * e->gy;
  e->cap = cfg->nb_supersurfels_max > e->S ? cfg->nb_supersurfels_max : e->S;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  const int sms = prop.multiProcessorCount;
  const int need_blocks = (e->cap + icp_chunk_size() - 1) / icp_chunk_size();
  // measured on B200 at VGA (300 frames): 0.563 ms/frame without, 0.582 ms/frame with PDL edges in the
  // frame graph -- the chain is bound by each kernel's own dependent L2 round trips, not by launch gaps
  e->pdl = 3;
  e->pdl_pipe = 0;
  e->pdl_now = 0;
  if (const char* v = getenv("SSF_PDL")) e->pdl = atoi(v);
  if (const char* v = getenv("SSF_PDL_PIPE")) e->pdl_pipe = atoi(v);
  e->icp_occ = 3;
  if (const char* v = getenv("SSF_ICP_OCC")) e->icp_occ = atoi(v);   // tuning knob: 2 .. 5
  if (e->icp_occ < 2 || e->icp_occ > 5) e->icp_occ = 3;
  e->icp_stages = 1;
  if (const char* v = getenv("SSF_ICP_STAGES")) e->icp_stages = atoi(v);   // tuning knob: 1 (no ring) .. 4 TMA ring per CTA; -2 .. -4 thread-private cp.async ring
  e->icp_stages = icp_configure(e->icp_stages);
  tps_configure();
  e->tps_grid = tps_persistent_grid(device, e->gx, e->gy, cfg->cell_size, e->H, cfg->seg_iter, &e->tps_cache_slots);
  // The one-kernel (cooperative, band-owned) form of the segmentation is kept as an option:
  // measured on B200 at VGA it is ~6 % slower per frame than the graph of small kernels
  // (0.602 vs 0.559 ms; per pass ~1.4 us cache fill + ~3.7 us relabel + ~1.7 us barrier for the
  // colour passes, 3.2 + 4.1 + 2.4 us with the disparity plane), see DESIGN.md section 3.
  e->tps_occ = 3;   // 80 registers, no spills; 4 (64 registers) measured the same on B200: 5790 vs 5780 frames/s, 0.466 vs 0.467 ms
  if (const char* v = getenv("SSF_TPS_OCC")) e->tps_occ = atoi(v);
  e->tps_fused = 1;
  if (const char* v = getenv("SSF_TPS_FUSED")) e->tps_fused = atoi(v) != 0;   // 0: round-1 pass + merge launches (A/B)
  e->tps_persistent = 0;
  if (const char* v = getenv("SSF_TPS_PERSISTENT")) e->tps_persistent = (atoi(v) != 0 && e->tps_grid > 0) ? 1 : 0;
  e->icp_debug = 0;
  if (const char* v = getenv("SSF_ICP_DEBUG")) e->icp_debug = atoi(v);
  e->icp_loop = 1;
  if (const char* v = getenv("SSF_ICP_LOOP")) e->icp_loop = atoi(v) != 0;   // 0: always the multi-launch registration (A/B)
  e->icp_grid = need_blocks < e->icp_occ * sms ? need_blocks : e->icp_occ * sms;

  cudaError_t err = cudaSuccess;
#define A(call) if (err == cudaSuccess) err = (call)
  A(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
  e->stream = e->own_stream;
  A(cudaEventCreate(&e->ev0)); A(cudaEventCreate(&e->ev1)); A(cudaEventCreate(&e->evf0)); A(cudaEventCreate(&e->evf1));
  for (int k = 0; k < 6; k++) A(cudaEventCreate(&e->ev_tm[k]));
  const size_t N = e->npix;
  const int S = e->S, nbs = cfg->nb_samples;
  A(dalloc(&e->rgba, N)); A(dalloc(&e->disp, N)); A(dalloc(&e->labels, N)); A(dalloc(&e->bound, N));
  A(dalloc(&e->inliers, N)); A(dalloc(&e->lmap, N)); A(dalloc(&e->in_rgb, N * 3)); A(dalloc(&e->in_depth, N));
  A(dalloc(&e->depth_f, N)); A(dalloc(&e->in_depth16, N));
  A(dalloc(&e->sp, (size_t)S)); A(dalloc(&e->sums, (size_t)3 * S));   // three rotating sum buffers, see ssf_tps.cu
  {
    char* p = nullptr;
    A(dalloc(&p, (size_t)S * nbs * (sizeof(float4) + sizeof(int))));
    e->samples = reinterpret_cast<float4*>(p);
    char* r = nullptr;
    A(dalloc(&r, (size_t)S * nbs * tps_rng_state_bytes()));
    e->rng = r;
  }
  A(dalloc(&e->filt_a, (size_t)S * 16)); A(dalloc(&e->filt_b, (size_t)S * 8));   // filt_a also serves as the 11-plane scratch of tps_filter_kernel<false>
  A(dalloc(&e->xsums, (size_t)S * 16));
  A(dalloc(&e->tps_barrier, (size_t)32));
  if (getenv("SSF_TPS_TRACE")) {
    char* t = nullptr;
    A(dalloc(&t, tps_trace_bytes(e->tps_grid > 148 ? e->tps_grid : 148)));   // also holds 8 stamps per CTA of a fused pass
    e->tps_trace = reinterpret_cast<unsigned long long*>(t);
  }
  e->frame.stride = round4(S);
  e->model.stride = e->model_alt.stride = round4(e->cap);
  A(dalloc(&e->frame.base, (size_t)P_COUNT * e->frame.stride));
  A(dalloc(&e->model.base, (size_t)P_COUNT * e->model.stride));
  A(dalloc(&e->model_alt.base, (size_t)P_COUNT * e->model_alt.stride));
  A(dalloc(&e->ftab, (size_t)2 * S)); A(dalloc(&e->matched, (size_t)S)); A(dalloc(&e->best, (size_t)S));
  // slot 0 = the buffers above; two more frame slots, two more streams for the pipelined mode
  {
    FrameSlot& f = e->slot[0];
    f.in_rgb = e->in_rgb; f.in_depth = e->in_depth;
    f.rgba = e->rgba; f.disp = e->disp; f.labels = e->labels; f.bound = e->bound; f.inliers = e->inliers;
    f.sp = e->sp; f.sums = e->sums;
    f.lmap = e->lmap; f.frame = e->frame; f.ftab = e->ftab; f.matched = e->matched; f.best = e->best;
  }
  for (int k = 1; k < SSF_SLOTS; k++) {
    FrameSlot& f = e->slot[k];
    A(dalloc(&f.in_rgb, N * 3)); A(dalloc(&f.in_depth, N));
    A(dalloc(&f.rgba, N)); A(dalloc(&f.disp, N)); A(dalloc(&f.labels, N)); A(dalloc(&f.bound, N)); A(dalloc(&f.inliers, N));
    A(dalloc(&f.sp, (size_t)S)); A(dalloc(&f.sums, (size_t)3 * S));
    f.frame.stride = e->frame.stride;
    A(dalloc(&f.lmap, N)); A(dalloc(&f.frame.base, (size_t)P_COUNT * e->frame.stride));
    A(dalloc(&f.ftab, (size_t)2 * S)); A(dalloc(&f.matched, (size_t)S)); A(dalloc(&f.best, (size_t)S));
  }
  if (err == cudaSuccess) {
    // tensor maps of the label images (fused pass, TMA staging); any failure falls back to plain loads
    e->tps_tma = 1;
    if (const char* v = getenv("SSF_TPS_TMA")) e->tps_tma = atoi(v) != 0;
    for (int k = 0; k < SSF_SLOTS && e->tps_tma; k++)
      if (!tps_make_label_map(e->label_map[k], e->slot[k].labels, e->W, e->H)) e->tps_tma = 0;
  }
  for (int p = 1; p < SSF_SLOTS; p++) A(cudaStreamCreateWithFlags(&e->stage_stream[p], cudaStreamNonBlocking));
  A(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < SSF_SLOTS; k++) {
    for (int p = 0; p < SSF_SLOTS; p++) A(cudaEventCreateWithFlags(&e->ev_stage[k][p], cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&e->ev_done[k], cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&e->ev_copy[k], cudaEventDisableTiming));
    A(cudaEventCreate(&e->ev_t0[k])); A(cudaEventCreate(&e->ev_t1[k]));
    A(dalloc(&e->d_report2[k], (size_t)1));
    A(cudaMallocHost(reinterpret_cast<void**>(&e->h_report2[k]), sizeof(FrameReport)));
    A(cudaMallocHost(reinterpret_cast<void**>(&e->h_prior2[k]), 12 * sizeof(float)));
  }
  {
    // measured on B200 at VGA (200 frames, resident / from pinned host): 4 stages 5160 / 4740, 5 stages 5510 / 4910,
    // 6 stages 5840 / 5790 frames/s
    int stages = SSF_SLOTS;
    if (const char* v = getenv("SSF_PIPELINE_STAGES")) stages = atoi(v);   // 1 .. 6 frames in flight
    if (stages < 1) stages = 1;
    if (stages > SSF_SLOTS) stages = SSF_SLOTS;
    plan_stages(e, stages);
  }
  A(dalloc(&e->states, (size_t)e->cap));
  A(dalloc(&e->scan_tmp, (size_t)8 + 4 * ((size_t)(e->cap + 1023) / 1024)));
  A(dalloc(&e->icp, (size_t)1)); A(dalloc(&e->icp_partials, (size_t)e->icp_grid * 32));
  A(dalloc(&e->xbuf, (size_t)2 * SSF_MAX_PEERS * 64)); A(dalloc(&e->xpeers_dev, (size_t)SSF_MAX_PEERS));
  e->xrank = 0; e->xworld = 1;
  A(dalloc(&e->counters, (size_t)1)); A(dalloc(&e->pose, (size_t)1));
  A(dalloc(&e->d_report, (size_t)1));
  A(cudaMallocHost(reinterpret_cast<void**>(&e->h_report), sizeof(FrameReport)));
  A(cudaMallocHost(reinterpret_cast<void**>(&e->h_prior), 12 * sizeof(float)));
  A(cudaMallocHost(reinterpret_cast<void**>(&e->h_icp), sizeof(IcpState)));
#undef A
  if (err != cudaSuccess) {
    fprintf(stderr, "ssf_create: %s\n", cudaGetErrorString(err));
    ssf_destroy(e);
    return SSF_ERR_CUDA;
  }
  // identity pose (supersurfel_fusion.cu:133-136)
  DevicePose ident = {{1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 0, 0}};
  cudaMemcpy(e->pose, &ident, sizeof(ident), cudaMemcpyHostToDevice);
  memset(e->h_report, 0, sizeof(FrameReport));
  e->h_report->pose = ident;
  tps_init_rng(e);   // initRandStates_kernel, once (TPS_RGBD.cu:123)
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) { ssf_destroy(e); return SSF_ERR_CUDA; }
  e->created = true;
  *out = e;
  return SSF_OK;
}

int ssf_destroy(SsfHandle h) {
  if (!h) return SSF_ERR_INVALID_ARG;
  EngineImpl* e = static_cast<EngineImpl*>(h);
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (int k = 0; k < 8; k++)
    if (e->graph_ready[k]) cudaGraphExecDestroy(e->graph_exec[k]);
  for (int k = 0; k < 6; k++)
    if (e->ev_tm[k]) cudaEventDestroy(e->ev_tm[k]);
  if (e->slot[0].lmap) select_slot(e, 0);
  for (int k = 0; k < SSF_SLOTS; k++) {
    for (int p = 0; p < SSF_SLOTS; p++) {
      for (int b = 0; b < 4; b++)
        if (e->stage_ready[k][p][b]) cudaGraphExecDestroy(e->stage_graph[k][p][b]);
      if (e->ev_stage[k][p]) cudaEventDestroy(e->ev_stage[k][p]);
    }
    if (e->ev_done[k]) cudaEventDestroy(e->ev_done[k]);
    if (e->ev_copy[k]) cudaEventDestroy(e->ev_copy[k]);
    if (e->ev_t0[k]) cudaEventDestroy(e->ev_t0[k]);
    if (e->ev_t1[k]) cudaEventDestroy(e->ev_t1[k]);
    if (e->d_report2[k]) cudaFree(e->d_report2[k]);
    if (e->h_report2[k]) cudaFreeHost(e->h_report2[k]);
    if (e->h_prior2[k]) cudaFreeHost(e->h_prior2[k]);
  }
  for (int k = 1; k < SSF_SLOTS; k++) {
    FrameSlot& f = e->slot[k];
    void* own[] = {f.in_rgb, f.in_depth, f.rgba, f.disp, f.labels, f.bound, f.inliers, f.sp, f.sums, f.lmap, f.frame.base, f.ftab, f.matched, f.best};
    for (void* b : own)
      if (b) cudaFree(b);
  }
  for (int p = 1; p < SSF_SLOTS; p++)
    if (e->stage_stream[p]) cudaStreamDestroy(e->stage_stream[p]);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  for (int g = 0; g < SSF_MAX_PEERS; g++)
    if (e->xpeer_open[g]) cudaIpcCloseMemHandle(e->xpeer_open[g]);
  void* bufs[] = {e->rgba, e->disp, e->labels, e->bound, e->inliers, e->lmap, e->in_rgb, e->in_depth, e->depth_f, e->in_depth16, e->sp, e->sums,
                  e->samples, e->rng, e->filt_a, e->filt_b, e->xsums, e->tps_barrier, e->tps_trace, e->xbuf, e->xpeers_dev, e->frame.base, e->model.base,
                  e->model_alt.base, e->ftab, e->matched, e->best, e->states, e->scan_tmp, e->icp, e->icp_partials,
                  e->counters, e->pose, e->d_report, e->scratch};
  for (void* b : bufs)
    if (b) cudaFree(b);
  if (e->h_report) cudaFreeHost(e->h_report);
  if (e->h_prior) cudaFreeHost(e->h_prior);
  if (e->h_icp) cudaFreeHost(e->h_icp);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->evf0) cudaEventDestroy(e->evf0);
  if (e->evf1) cudaEventDestroy(e->evf1);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  delete e;
  return SSF_OK;
}

int ssf_set_stream(SsfHandle h, void* cuda_stream) {
  H_CHECK_IDLE(h);
  if (e->in_flight) { e->err = "pipelined frames in flight: ssf_wait_frame first"; return SSF_ERR_STATE; }
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  e->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : e->own_stream;
  for (int k = 0; k < 8; k++)
    if (e->graph_ready[k]) { cudaGraphExecDestroy(e->graph_exec[k]); e->graph_ready[k] = false; }
  for (int k = 0; k < SSF_SLOTS; k++)
    for (int p = 0; p < SSF_SLOTS; p++)
      for (int b = 0; b < 4; b++)
        if (e->stage_ready[k][p][b]) { cudaGraphExecDestroy(e->stage_graph[k][p][b]); e->stage_ready[k][p][b] = false; }
  return SSF_OK;
}

const char* ssf_last_error(SsfHandle h) {
  if (!h) return "invalid handle";
  return static_cast<EngineImpl*>(h)->err.c_str();
}

int ssf_is_initialized(SsfHandle h) { return (h && static_cast<EngineImpl*>(h)->created) ? 1 : 0; }

static int run_frame(EngineImpl* e, const float* prior, uint32_t flags) {
  const bool marks = (flags & SSF_FLAG_STAGE_TIMING) != 0;
  const bool small = pick_small_icp(e, e->h_report->counters.nb_visible);   // what the last report saw
  const int gi = ((flags & SSF_FLAG_BILATERAL) ? 1 : 0) | (marks ? 2 : 0) | (small ? 4 : 0);
  if (e->in_flight) { e->err = "synchronous frame while pipelined frames are in flight: ssf_wait_frame first"; return SSF_ERR_STATE; }
  if (e->cur_slot != 0) select_slot(e, 0);
  if (prior) {
    memcpy(e->h_prior, prior, 12 * sizeof(float));
    SSF_CUDA(e, cudaMemcpyAsync(e->pose, e->h_prior, 12 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  }
  SSF_CUDA(e, cudaEventRecord(e->evf0, e->stream));
  if (e->use_graph) {
    {
      const int rc = ensure_frame_graph(e, gi, false);
      if (rc) return rc;
    }
    SSF_CUDA(e, cudaGraphLaunch(e->graph_exec[gi], e->stream));
    e->launches += e->launches_per_frame[gi];
  } else {
    enqueue_frame(e, (gi & 1) != 0, marks, small);
  }
  SSF_CUDA(e, cudaEventRecord(e->evf1, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(e->h_report, e->d_report, sizeof(FrameReport), cudaMemcpyDeviceToHost, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e->evf0, e->evf1);
  fill_stats(e, ms);
  if (marks) fill_stage_ms(e);
  return SSF_OK;
}

int ssf_process_frame(SsfHandle h, const uint8_t* rgb, size_t rgb_stride, const float* depth, size_t depth_stride,
                      const float* pose_prior_Rt12, uint32_t flags) {
  H_CHECK_IDLE(h);
  if (!rgb || !depth) return SSF_ERR_INVALID_ARG;
  if (rgb_stride == 0) rgb_stride = (size_t)e->W * 3;
  if (depth_stride == 0) depth_stride = (size_t)e->W * 4;
  if (rgb_stride < (size_t)e->W * 3 || depth_stride < (size_t)e->W * 4) return SSF_ERR_INVALID_ARG;
  // rgb.upload / depth.upload (supersurfel_fusion.cu:173-174)
  SSF_CUDA(e, cudaMemcpy2DAsync(e->in_rgb, (size_t)e->W * 3, rgb, rgb_stride, (size_t)e->W * 3, e->H, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaMemcpy2DAsync(e->in_depth, (size_t)e->W * 4, depth, depth_stride, (size_t)e->W * 4, e->H, cudaMemcpyDefault, e->stream));
  return run_frame(e, pose_prior_Rt12, flags);
}

int ssf_process_frame_device(SsfHandle h, const uint8_t* rgb_dev, const float* depth_dev, const float* pose_prior_Rt12,
                             uint32_t flags) {
  H_CHECK_IDLE(h);
  if (!rgb_dev || !depth_dev) return SSF_ERR_INVALID_ARG;
  SSF_CUDA(e, cudaMemcpyAsync(e->in_rgb, rgb_dev, e->npix * 3, cudaMemcpyDeviceToDevice, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(e->in_depth, depth_dev, e->npix * 4, cudaMemcpyDeviceToDevice, e->stream));
  return run_frame(e, pose_prior_Rt12, flags);
}

int ssf_process_frame_depth16(SsfHandle h, const uint8_t* rgb, size_t rgb_stride, const uint16_t* depth16,
                              size_t depth_stride, float depth_scale, const float* pose_prior_Rt12, uint32_t flags) {
  H_CHECK_IDLE(h);
  if (!rgb || !depth16) return SSF_ERR_INVALID_ARG;
  if (rgb_stride == 0) rgb_stride = (size_t)e->W * 3;
  if (depth_stride == 0) depth_stride = (size_t)e->W * 2;
  if (rgb_stride < (size_t)e->W * 3 || depth_stride < (size_t)e->W * 2) return SSF_ERR_INVALID_ARG;
  SSF_CUDA(e, cudaMemcpy2DAsync(e->in_rgb, (size_t)e->W * 3, rgb, rgb_stride, (size_t)e->W * 3, e->H, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaMemcpy2DAsync(e->in_depth16, (size_t)e->W * 2, depth16, depth_stride, (size_t)e->W * 2, e->H, cudaMemcpyDefault, e->stream));
  // depth.convertTo(CV_32FC1, depth_scale) (node/supersurfel_fusion_rgbd_benchmark_node.cpp:609-610)
  launch_depth16(e, e->in_depth16, e->in_depth, depth_scale);
  return run_frame(e, pose_prior_Rt12, flags);
}

int ssf_bilateral_filter(SsfHandle h, const float* depth, size_t depth_stride, int kernel_size, float sigma_color,
                         float sigma_spatial, float* out) {
  H_CHECK_IDLE(h);
  if (!depth || !out) return SSF_ERR_INVALID_ARG;
  if (depth_stride == 0) depth_stride = (size_t)e->W * 4;
  if (depth_stride < (size_t)e->W * 4) return SSF_ERR_INVALID_ARG;
  SSF_CUDA(e, cudaMemcpy2DAsync(e->in_depth, (size_t)e->W * 4, depth, depth_stride, (size_t)e->W * 4, e->H, cudaMemcpyDefault, e->stream));
  int rc = launch_bilateral(e, e->in_depth, e->depth_f, kernel_size, sigma_color, sigma_spatial);
  if (rc) { e->err = "ssf_bilateral_filter: kernel radius above the supported 12"; return rc; }
  SSF_CUDA(e, cudaMemcpyAsync(out, e->depth_f, e->npix * 4, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_get_gray(SsfHandle h, uint8_t* gray) {
  H_CHECK_IDLE(h);
  if (!gray) return SSF_ERR_INVALID_ARG;
  int rc = ensure_scratch(e, e->npix);
  if (rc) return rc;
  launch_gray(e, e->in_rgb, reinterpret_cast<uint8_t*>(e->scratch));
  SSF_CUDA(e, cudaMemcpyAsync(gray, e->scratch, e->npix, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_get_filtered_depth(SsfHandle h, float* depth) {
  H_CHECK_IDLE(h);
  if (!depth) return SSF_ERR_INVALID_ARG;
  SSF_CUDA(e, cudaMemcpyAsync(depth, e->depth_f, e->npix * 4, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  return SSF_OK;
}

// ---- pipelined mode ---------------------------------------------------------------------
// A frame's launch sequence is a chain: ingest -> segmentation iterations -> extraction ->
// registration + fusion.  Only the last link touches the model, so the chain of frame k+1 can run
// behind that of frame k as soon as it keeps its own copy of the per-frame state.  The sequence is
// cut into nb_stages contiguous stages of about equal cost; each stage is a CUDA graph on its own
// stream, the stages of one frame are chained by events, a stage of consecutive frames is
// serialised by its stream, and every frame in flight owns one FrameSlot.  Same kernels in the
// same order per frame: results are identical to the synchronous path; the frame rate is set by
// the longest stage instead of by the whole chain.
static int capture_stage(EngineImpl* e, int slot, int stage, int gi) {
  cudaGraph_t g;
  const uint64_t before = e->launches;
  SSF_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
  e->pdl_now = e->pdl_pipe;
  enqueue_steps(e, e->stage_first[stage], e->stage_first[stage + 1], (gi & 1) != 0, true, e->d_report2[slot], false, (gi & 2) != 0);
  e->pdl_now = 0;
  SSF_CUDA(e, cudaStreamEndCapture(e->stream, &g));
  e->stage_launches[slot][stage][gi] = e->launches - before;
  e->launches = before;
  SSF_CUDA(e, cudaGraphInstantiate(&e->stage_graph[slot][stage][gi], g, 0));
  cudaGraphDestroy(g);
  e->stage_ready[slot][stage][gi] = true;
  return SSF_OK;
}

// variant of the graph of stage p for these frame flags / registration path
static int stage_variant(const EngineImpl* e, int p, uint32_t flags, bool small) {
  return ((p == 0 && (flags & SSF_FLAG_BILATERAL)) ? 1 : 0) | ((p == e->nb_stages - 1 && small) ? 2 : 0);
}

// Everything of ssf_submit_frame that can fail after work was enqueued; see the caller.
static int submit_enqueue(EngineImpl* e, int s, const uint8_t* rgb, size_t rgb_stride, const float* depth,
                          size_t depth_stride, const float* pose_prior_Rt12, uint32_t flags, bool small) {
  const int P = e->nb_stages;
  for (int p = 0; p < P; p++) {
    cudaStream_t st = p == 0 ? e->stream : e->stage_stream[p];
    const int gi = stage_variant(e, p, flags, small);
    if (p == 0) {
      // The slot (its staging buffers included) is free once the frame that last used it has left the last
      // stage.  The upload runs on the copy stream into the SLOT's staging buffers, so the DMA of this frame
      // overlaps the first-stage kernels of the previous one; the first stage waits for the upload only.
      cudaStream_t cs = e->copy_stream;
      SSF_CUDA(e, cudaStreamWaitEvent(cs, e->ev_done[s], 0));
      SSF_CUDA(e, cudaEventRecord(e->ev_t0[s], cs));
      SSF_CUDA(e, cudaMemcpy2DAsync(e->in_rgb, (size_t)e->W * 3, rgb, rgb_stride, (size_t)e->W * 3, e->H, cudaMemcpyDefault, cs));
      SSF_CUDA(e, cudaMemcpy2DAsync(e->in_depth, (size_t)e->W * 4, depth, depth_stride, (size_t)e->W * 4, e->H, cudaMemcpyDefault, cs));
      SSF_CUDA(e, cudaEventRecord(e->ev_copy[s], cs));
      SSF_CUDA(e, cudaStreamWaitEvent(st, e->ev_copy[s], 0));
    } else {
      SSF_CUDA(e, cudaStreamWaitEvent(st, e->ev_stage[s][p - 1], 0));
    }
    if (p == P - 1 && pose_prior_Rt12) {
      memcpy(e->h_prior2[s], pose_prior_Rt12, 12 * sizeof(float));
      SSF_CUDA(e, cudaMemcpyAsync(e->pose, e->h_prior2[s], 12 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    SSF_CUDA(e, cudaGraphLaunch(e->stage_graph[s][p][gi], st));
    e->launches += e->stage_launches[s][p][gi];
    if (p < P - 1) {
      SSF_CUDA(e, cudaEventRecord(e->ev_stage[s][p], st));
    } else {
      SSF_CUDA(e, cudaEventRecord(e->ev_t1[s], st));
      SSF_CUDA(e, cudaMemcpyAsync(e->h_report2[s], e->d_report2[s], sizeof(FrameReport), cudaMemcpyDeviceToHost, st));
      SSF_CUDA(e, cudaEventRecord(e->ev_done[s], st));
    }
  }
  return SSF_OK;
}

// capture + instantiate + upload whatever graphs of the pipelined mode are still missing
static int prepare_pipeline(EngineImpl* e, uint32_t flags) {
  const int P = e->nb_stages;
  const int keep = e->cur_slot;
  for (int s = 0; s < P; s++) {
    select_slot(e, s);
    for (int p = 0; p < P; p++) {
      for (int small = 0; small < 2; small++) {       // both registration paths: the pick may change from frame to frame
        const int gi = stage_variant(e, p, flags, small != 0);   // the ingest is always in stage 0
        if (e->stage_ready[s][p][gi]) continue;
        int rc = capture_stage(e, s, p, gi);
        if (rc) { select_slot(e, keep); return rc; }
        SSF_CUDA(e, cudaGraphUpload(e->stage_graph[s][p][gi], p == 0 ? e->stream : e->stage_stream[p]));
      }
    }
  }
  select_slot(e, keep);
  return SSF_OK;
}

int ssf_prepare(SsfHandle h, uint32_t flags) {
  H_CHECK_IDLE(h);
  int rc = prepare_pipeline(e, flags);
  if (rc) return rc;
  // the synchronous frame graph too
  const bool marks = (flags & SSF_FLAG_STAGE_TIMING) != 0;
  if (e->use_graph) {
    for (int small = 0; small < 2; small++) {
      rc = ensure_frame_graph(e, ((flags & SSF_FLAG_BILATERAL) ? 1 : 0) | (marks ? 2 : 0) | (small ? 4 : 0), true);
      if (rc) return rc;
    }
  }
  for (int p = 1; p < e->nb_stages; p++) SSF_CUDA(e, cudaStreamSynchronize(e->stage_stream[p]));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_submit_frame(SsfHandle h, const uint8_t* rgb, size_t rgb_stride, const float* depth, size_t depth_stride,
                     const float* pose_prior_Rt12, uint32_t flags) {
  H_CHECK(h);
  if (!rgb || !depth) return SSF_ERR_INVALID_ARG;
  if (rgb_stride == 0) rgb_stride = (size_t)e->W * 3;
  if (depth_stride == 0) depth_stride = (size_t)e->W * 4;
  if (rgb_stride < (size_t)e->W * 3 || depth_stride < (size_t)e->W * 4) return SSF_ERR_INVALID_ARG;
  const int P = e->nb_stages;
  if (e->in_flight >= P) { e->err = "every pipeline stage is occupied: ssf_wait_frame first"; return SSF_ERR_STATE; }
  const int s = e->pipe_next;
  // the model this frame registers against is the last reported one plus at most S insertions per frame
  // that is still in flight or about to be
  const bool small = pick_small_icp(e, (long long)e->h_report->counters.nb_visible + (long long)(e->in_flight + 1) * e->S);
  // graphs first (nothing enqueued yet: a failure here leaves the pipeline as it was)
  {
    const int keep = e->cur_slot;
    select_slot(e, s);
    for (int p = 0; p < P; p++) {
      const int gi = stage_variant(e, p, flags, small);
      if (!e->stage_ready[s][p][gi]) {
        int rc = capture_stage(e, s, p, gi);
        if (rc) { select_slot(e, keep); return rc; }
      }
    }
  }
  if (e->in_flight == 0) e->pipe_oldest = s;
  const int rc = submit_enqueue(e, s, rgb, rgb_stride, depth, depth_stride, pose_prior_Rt12, flags, small);
  if (rc) {
    // Part of the frame may be running and its completion event was never recorded: the slot, the
    // shared staging buffers and the prior buffer cannot be reused safely.  Drain everything and
    // retire the handle -- every later call returns SSF_ERR_STATE (ssf_last_error keeps the cause).
    for (int p = 1; p < P; p++) cudaStreamSynchronize(e->stage_stream[p]);
    cudaStreamSynchronize(e->copy_stream);
    cudaStreamSynchronize(e->stream);
    cudaGetLastError();
    e->failed = 1;
    e->err = "ssf_submit_frame failed half-way, handle retired: " + e->err;
    return rc;
  }
  e->in_flight++;
  e->pipe_next = (s + 1) % P;
  return SSF_OK;
}

int ssf_wait_frame(SsfHandle h, SsfFrameStats* out, float R[9], float t[3]) {
  H_CHECK(h);
  if (e->in_flight <= 0) { e->err = "no frame in flight"; return SSF_ERR_STATE; }
  const int s = e->pipe_oldest;
  SSF_CUDA(e, cudaEventSynchronize(e->ev_done[s]));
  SSF_LAUNCH_OK(e);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e->ev_t0[s], e->ev_t1[s]);
  *e->h_report = *e->h_report2[s];
  fill_stats(e, ms);
  if (out) *out = e->stats;
  if (R) memcpy(R, e->h_report->pose.R, 36);
  if (t) memcpy(t, e->h_report->pose.t, 12);
  e->in_flight--;
  e->pipe_oldest = (s + 1) % e->nb_stages;
  if (e->in_flight == 0) {
    // back to a quiescent state: every getter / stage entry point works on `stream` and on the
    // slot of the frame just returned
    for (int p = 1; p < e->nb_stages; p++) SSF_CUDA(e, cudaStreamSynchronize(e->stage_stream[p]));
    SSF_CUDA(e, cudaStreamSynchronize(e->copy_stream));
    select_slot(e, s);
  }
  return SSF_OK;
}

int ssf_plan_pipeline(const SsfConfig* cfg, int stages, int persistent_segmentation, int* first, int* nb_steps) {
  if (!cfg || !first) return SSF_ERR_INVALID_ARG;
  const int used = plan_stages_impl(cfg->seg_iter, cfg->icp_iter, persistent_segmentation, stages, first);
  if (nb_steps) *nb_steps = (persistent_segmentation ? 1 : cfg->seg_iter + 2) + 3;
  return used;
}

int ssf_plan_weights(const SsfConfig* cfg, int persistent_segmentation, int* weights, int capacity) {
  if (!cfg || !weights) return SSF_ERR_INVALID_ARG;
  const int G = (persistent_segmentation ? 1 : cfg->seg_iter + 2) + 3;
  if (capacity < G) return SSF_ERR_INVALID_ARG;
  for (int g = 0; g < G; g++) weights[g] = step_weight(g, cfg->seg_iter, cfg->icp_iter, persistent_segmentation);
  return G;
}

int ssf_get_pipeline_depth(SsfHandle h, int* stages) {
  H_CHECK(h);
  if (!stages) return SSF_ERR_INVALID_ARG;
  *stages = e->nb_stages;
  return SSF_OK;
}

int ssf_get_frame_stats(SsfHandle h, SsfFrameStats* out) {
  H_CHECK(h);
  if (!out) return SSF_ERR_INVALID_ARG;
  *out = e->stats;
  return SSF_OK;
}

int ssf_get_pose(SsfHandle h, float R[9], float t[3]) {
  H_CHECK_IDLE(h);
  if (!R || !t) return SSF_ERR_INVALID_ARG;
  int rc = read_report(e, false);
  if (rc) return rc;
  memcpy(R, e->h_report->pose.R, 36);
  memcpy(t, e->h_report->pose.t, 12);
  return SSF_OK;
}

int ssf_set_pose(SsfHandle h, const float R[9], const float t[3]) {
  H_CHECK_IDLE(h);
  if (!R || !t) return SSF_ERR_INVALID_ARG;
  memcpy(e->h_prior, R, 36);
  memcpy(e->h_prior + 9, t, 12);
  SSF_CUDA(e, cudaMemcpyAsync(e->pose, e->h_prior, 48, cudaMemcpyHostToDevice, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  return SSF_OK;
}

int ssf_get_stamp(SsfHandle h, int* stamp) {
  H_CHECK_IDLE(h);
  if (!stamp) return SSF_ERR_INVALID_ARG;
  int rc = read_report(e, false);
  if (rc) return rc;
  *stamp = e->h_report->counters.stamp;
  return SSF_OK;
}

int ssf_set_stamp(SsfHandle h, int stamp) {
  H_CHECK_IDLE(h);
  SSF_CUDA(e, cudaMemcpyAsync(&e->counters->stamp, &stamp, sizeof(int), cudaMemcpyHostToDevice, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(&e->counters->seg_stamp, &stamp, sizeof(int), cudaMemcpyHostToDevice, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  return SSF_OK;
}

int ssf_get_counts(SsfHandle h, int* nb_supersurfels, int* nb_visible, int* nb_removed) {
  H_CHECK_IDLE(h);
  int rc = read_report(e, false);
  if (rc) return rc;
  if (nb_supersurfels) *nb_supersurfels = e->h_report->counters.nb_supersurfels;
  if (nb_visible) *nb_visible = e->h_report->counters.nb_visible;
  if (nb_removed) *nb_removed = e->h_report->counters.nb_removed;
  return SSF_OK;
}

int ssf_get_nb_superpixels(SsfHandle h, int* n) {
  H_CHECK(h);
  if (!n) return SSF_ERR_INVALID_ARG;
  *n = e->S;
  return SSF_OK;
}

static int copy_set_out(EngineImpl* e, const SurfelSet& set, const SsfSurfels* dst, int n) {
  if (!dst || n < 0) return SSF_ERR_INVALID_ARG;
  if (n == 0) return SSF_OK;
  int rc = ensure_scratch(e, (size_t)n * 104);
  if (rc) return rc;
  SsfSurfels sv = scratch_view(e->scratch, n);
  launch_pack(e, set, n, sv);
  rc = copy_members(e, *dst, sv, n);
  if (rc) return rc;
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  return SSF_OK;
}

static int copy_set_in(EngineImpl* e, const SsfSurfels* src, int n, const SurfelSet& set) {
  if (!src || n < 0) return SSF_ERR_INVALID_ARG;
  if (n == 0) return SSF_OK;
  int rc = ensure_scratch(e, (size_t)n * 104);
  if (rc) return rc;
  SsfSurfels sv = scratch_view(e->scratch, n);
  SsfSurfels present = sv;
  if (!src->positions) present.positions = nullptr;
  if (!src->colors) present.colors = nullptr;
  if (!src->stamps) present.stamps = nullptr;
  if (!src->orientations) present.orientations = nullptr;
  if (!src->shapes) present.shapes = nullptr;
  if (!src->dims) present.dims = nullptr;
  if (!src->confidences) present.confidences = nullptr;
  rc = copy_members(e, present, *src, n);
  if (rc) return rc;
  launch_unpack(e, present, n, set);
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  return SSF_OK;
}

int ssf_copy_model(SsfHandle h, const SsfSurfels* dst, int n) {
  H_CHECK_IDLE(h);
  if (n > e->cap) return SSF_ERR_INVALID_ARG;
  return copy_set_out(e, e->model, dst, n);
}

int ssf_copy_frame(SsfHandle h, const SsfSurfels* dst) {
  H_CHECK_IDLE(h);
  return copy_set_out(e, e->frame, dst, e->S);
}

int ssf_get_segmentation(SsfHandle h, int32_t* labels, int32_t* bound, uint8_t* inliers, float* disp,
                         float* slanted_depth, float* superpixels, uint8_t* rgba) {
  H_CHECK_IDLE(h);
  const size_t N = e->npix;
  if (labels) SSF_CUDA(e, cudaMemcpyAsync(labels, e->labels, N * 4, cudaMemcpyDefault, e->stream));
  if (bound) SSF_CUDA(e, cudaMemcpyAsync(bound, e->bound, N * 4, cudaMemcpyDefault, e->stream));
  if (inliers) SSF_CUDA(e, cudaMemcpyAsync(inliers, e->inliers, N, cudaMemcpyDefault, e->stream));
  if (disp) SSF_CUDA(e, cudaMemcpyAsync(disp, e->disp, N * 4, cudaMemcpyDefault, e->stream));
  if (rgba) SSF_CUDA(e, cudaMemcpyAsync(rgba, e->rgba, N * 4, cudaMemcpyDefault, e->stream));
  if (superpixels) SSF_CUDA(e, cudaMemcpyAsync(superpixels, e->sp, (size_t)e->S * sizeof(Superpixel), cudaMemcpyDefault, e->stream));
  if (slanted_depth) {
    int rc = ensure_scratch(e, N * 4);
    if (rc) return rc;
    launch_pdl(e, split_lmap_kernel, dim3((unsigned)((N + 255) / 256)), dim3(256), 0, e->lmap, reinterpret_cast<float*>(e->scratch), N);
    e->launches++;
    SSF_CUDA(e, cudaMemcpyAsync(slanted_depth, e->scratch, N * 4, cudaMemcpyDefault, e->stream));
  }
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_get_slanted_depth(SsfHandle h, float* depth) {
  return ssf_get_segmentation(h, nullptr, nullptr, nullptr, nullptr, depth, nullptr, nullptr);
}

int ssf_render_preview(SsfHandle h, uint8_t* bgr) {
  H_CHECK_IDLE(h);
  if (!bgr) return SSF_ERR_INVALID_ARG;
  int rc = ensure_scratch(e, e->npix * 3);
  if (rc) return rc;
  launch_preview(e, reinterpret_cast<uint8_t*>(e->scratch));
  SSF_CUDA(e, cudaMemcpyAsync(bgr, e->scratch, e->npix * 3, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  return SSF_OK;
}

int ssf_get_model_view(SsfHandle h, SsfPlanarView* out) {
  H_CHECK_IDLE(h);
  if (!out) return SSF_ERR_INVALID_ARG;
  int rc = read_report(e, false);
  if (rc) return rc;
  out->base = e->model.base;
  out->stride = e->model.stride;
  out->count = e->h_report->counters.nb_supersurfels;
  out->planes = P_COUNT;
  return SSF_OK;
}

int ssf_get_frame_view(SsfHandle h, SsfPlanarView* out) {
  H_CHECK_IDLE(h);
  if (!out) return SSF_ERR_INVALID_ARG;
  if (e->in_flight) { e->err = "pipelined frames in flight: ssf_wait_frame first"; return SSF_ERR_STATE; }
  out->base = e->frame.base;
  out->stride = e->frame.stride;
  out->count = e->S;
  out->planes = P_COUNT;
  return SSF_OK;
}

int ssf_export_model(SsfHandle h, const char* path) {
  H_CHECK_IDLE(h);
  if (!path) return SSF_ERR_INVALID_ARG;
  int rc = read_report(e, false);
  if (rc) return rc;
  const int n = e->h_report->counters.nb_supersurfels;
  std::vector<float> pos((size_t)n * 3), col((size_t)n * 3), ori((size_t)n * 9), shp((size_t)n * 6), dms((size_t)n * 2),
      cnf((size_t)n);
  std::vector<int32_t> stp((size_t)n * 2);
  SsfSurfels dst = {pos.data(), col.data(), stp.data(), ori.data(), shp.data(), dms.data(), cnf.data()};
  rc = copy_set_out(e, e->model, &dst, n);
  if (rc) return rc;
  FILE* f = fopen(path, "w");
  if (!f) { e->err = std::string("cannot open ") + path; return SSF_ERR_IO; }
  // std::to_string(float) == "%f" (supersurfel_fusion.cu:616-630)
  for (int i = 0; i < n; i++) {
    if (!(cnf[i] > e->cfg.conf_thresh)) continue;
    fprintf(f, "%d %d %f\n", stp[2 * i], stp[2 * i + 1], cnf[i]);
    fprintf(f, "%f %f %f\n", pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    fprintf(f, "%f %f %f\n", col[3 * i], col[3 * i + 1], col[3 * i + 2]);
    fprintf(f, "%f %f\n", dms[2 * i], dms[2 * i + 1]);
    fprintf(f, "%f %f %f %f %f %f %f %f %f\n", ori[9 * i], ori[9 * i + 1], ori[9 * i + 2], ori[9 * i + 3],
            ori[9 * i + 4], ori[9 * i + 5], ori[9 * i + 6], ori[9 * i + 7], ori[9 * i + 8]);
    fprintf(f, "%f %f %f %f %f %f\n", shp[6 * i], shp[6 * i + 1], shp[6 * i + 2], shp[6 * i + 3], shp[6 * i + 4],
            shp[6 * i + 5]);
    fprintf(f, "\n");
  }
  fclose(f);
  return SSF_OK;
}

int ssf_extract_local_point_cloud(SsfHandle h, float radius, float* positions, float* normals, int capacity,
                                  int* count) {
  H_CHECK_IDLE(h);
  if (!positions || !normals || capacity < 0 || !count) return SSF_ERR_INVALID_ARG;
  int rc = ensure_scratch(e, (size_t)capacity * 24 + 16);
  if (rc) return rc;
  float* dpos = reinterpret_cast<float*>(e->scratch);
  float* dnrm = dpos + (size_t)capacity * 3;
  SSF_CUDA(e, cudaMemsetAsync(&e->counters->cloud_count, 0, sizeof(int), e->stream));
  launch_local_cloud(e, radius, dpos, dnrm, capacity);
  int n = 0;
  SSF_CUDA(e, cudaMemcpyAsync(&n, &e->counters->cloud_count, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  if (n > capacity) n = capacity;
  if (n > 0) {
    SSF_CUDA(e, cudaMemcpyAsync(positions, dpos, (size_t)n * 12, cudaMemcpyDefault, e->stream));
    SSF_CUDA(e, cudaMemcpyAsync(normals, dnrm, (size_t)n * 12, cudaMemcpyDefault, e->stream));
    SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  }
  *count = n;
  return SSF_OK;
}

int ssf_invalidate_frame_supersurfels(SsfHandle h, const uint8_t* mask) {
  H_CHECK_IDLE(h);
  if (!mask) return SSF_ERR_INVALID_ARG;
  int rc = ensure_scratch(e, (size_t)e->S);
  if (rc) return rc;
  SSF_CUDA(e, cudaMemcpyAsync(e->scratch, mask, (size_t)e->S, cudaMemcpyDefault, e->stream));
  launch_invalidate(e, reinterpret_cast<const uint8_t*>(e->scratch));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_transform_model(SsfHandle h, const float R[9], const float t[3]) {
  H_CHECK_IDLE(h);
  if (!R || !t) return SSF_ERR_INVALID_ARG;
  launch_transform_model(e, R, t);
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_set_model(SsfHandle h, const SsfSurfels* src, int nb_supersurfels, int nb_visible) {
  H_CHECK_IDLE(h);
  if (nb_supersurfels < 0 || nb_supersurfels > e->cap || nb_visible < 0 || nb_visible > nb_supersurfels)
    return SSF_ERR_INVALID_ARG;
  int rc = copy_set_in(e, src, nb_supersurfels, e->model);
  if (rc) return rc;
  launch_model_lab(e, nb_supersurfels);
  int c[2] = {nb_supersurfels, nb_visible};
  SSF_CUDA(e, cudaMemcpyAsync(&e->counters->nb_supersurfels, c, 2 * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_set_frame(SsfHandle h, const SsfSurfels* src) {
  H_CHECK_IDLE(h);
  int rc = copy_set_in(e, src, e->S, e->frame);
  if (rc) return rc;
  launch_frame_tables(e);
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_set_segmentation(SsfHandle h, const int32_t* labels, const int32_t* bound, const uint8_t* inliers,
                         const float* slanted_depth, const uint8_t* rgba) {
  H_CHECK_IDLE(h);
  const size_t N = e->npix;
  if (labels) SSF_CUDA(e, cudaMemcpyAsync(e->labels, labels, N * 4, cudaMemcpyDefault, e->stream));
  if (bound) SSF_CUDA(e, cudaMemcpyAsync(e->bound, bound, N * 4, cudaMemcpyDefault, e->stream));
  if (inliers) SSF_CUDA(e, cudaMemcpyAsync(e->inliers, inliers, N, cudaMemcpyDefault, e->stream));
  if (rgba) SSF_CUDA(e, cudaMemcpyAsync(e->rgba, rgba, N * 4, cudaMemcpyDefault, e->stream));
  if (slanted_depth) {
    int rc = ensure_scratch(e, N * 4);
    if (rc) return rc;
    SSF_CUDA(e, cudaMemcpyAsync(e->scratch, slanted_depth, N * 4, cudaMemcpyDefault, e->stream));
    launch_build_lmap(e, reinterpret_cast<const float*>(e->scratch));
  }
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_tps_segment(SsfHandle h, const uint8_t* rgb, size_t rgb_stride, const float* depth, size_t depth_stride) {
  H_CHECK_IDLE(h);
  if (!rgb || !depth) return SSF_ERR_INVALID_ARG;
  if (rgb_stride == 0) rgb_stride = (size_t)e->W * 3;
  if (depth_stride == 0) depth_stride = (size_t)e->W * 4;
  if (rgb_stride < (size_t)e->W * 3 || depth_stride < (size_t)e->W * 4) return SSF_ERR_INVALID_ARG;
  SSF_CUDA(e, cudaMemcpy2DAsync(e->in_rgb, (size_t)e->W * 3, rgb, rgb_stride, (size_t)e->W * 3, e->H, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaMemcpy2DAsync(e->in_depth, (size_t)e->W * 4, depth, depth_stride, (size_t)e->W * 4, e->H, cudaMemcpyDefault, e->stream));
  launch_ingest(e, e->in_rgb, (size_t)e->W * 3, e->in_depth, (size_t)e->W * 4);
  e->pdl_now = (e->tps_trace && getenv("SSF_TPS_TRACE_PDL")) ? e->pdl : 0;   // profiling aid: trace the passes as the frame graph chains them
  launch_tps(e);
  e->pdl_now = 0;
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  if (e->tps_trace) {   // profiling aid: dump the per-CTA phase timestamps of this call
    std::vector<unsigned long long> host(tps_trace_bytes(e->tps_grid > 148 ? e->tps_grid : 148) / 8);
    cudaMemcpy(host.data(), e->tps_trace, host.size() * 8, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(getenv("SSF_TPS_TRACE"), "wb")) { fwrite(host.data(), 8, host.size(), f); fclose(f); }
    cudaMemset(e->tps_trace, 0, host.size() * 8);
  }
  return SSF_OK;
}

int ssf_get_ransac_samples(SsfHandle h, float* samples) {
  H_CHECK_IDLE(h);
  if (!samples) return SSF_ERR_INVALID_ARG;
  SSF_CUDA(e, cudaMemcpyAsync(samples, e->samples, (size_t)e->S * e->cfg.nb_samples * sizeof(float4), cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  return SSF_OK;
}

int ssf_generate_supersurfels(SsfHandle h) {
  H_CHECK_IDLE(h);
  launch_extract(e);
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

static int resolve_n(EngineImpl* e, int n_src, int* out) {
  if (n_src <= 0) {
    int rc = read_report(e, false);
    if (rc) return rc;
    n_src = e->h_report->counters.nb_visible;
  }
  if (n_src > e->cap) return SSF_ERR_INVALID_ARG;
  *out = n_src;
  return SSF_OK;
}

int ssf_icp_system(SsfHandle h, const float R[9], const float t[3], int n_src, float out29[29]) {
  H_CHECK_IDLE(h);
  if (!R || !t || !out29) return SSF_ERR_INVALID_ARG;
  int n = 0;
  int rc = resolve_n(e, n_src, &n);
  if (rc) return rc;
  launch_icp_set_transform(e, R, t);
  if (n > 0) launch_icp_system(e, e->model, nullptr, n, false);
  else SSF_CUDA(e, cudaMemsetAsync(e->icp->sys, 0, sizeof(float) * 32, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(out29, e->icp->sys, 29 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_icp_system_enqueue(SsfHandle h, const float R[9], const float t[3], int n_src, int launches) {
  H_CHECK_IDLE(h);
  if (!R || !t || launches < 0) return SSF_ERR_INVALID_ARG;
  int n = 0;
  int rc = resolve_n(e, n_src, &n);
  if (rc) return rc;
  if (n <= 0) return SSF_ERR_STATE;
  launch_icp_set_transform(e, R, t);
  for (int i = 0; i < launches; i++) launch_icp_system(e, e->model, nullptr, n, false);
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_icp(SsfHandle h, const float* R_init, const float* t_init, float out29[29], float R_rel[9], float t_rel[3],
            int* iters, int* valid) {
  H_CHECK_IDLE(h);
  if ((R_init == nullptr) != (t_init == nullptr)) return SSF_ERR_INVALID_ARG;
  if (R_init) launch_icp_begin(e, R_init, t_init);
  else launch_icp_begin_from_pose(e);
  launch_icp_loop(e);
  launch_icp_finish(e, false);
  IcpState* hs = e->h_icp;
  SSF_CUDA(e, cudaMemcpyAsync(hs, e->icp, sizeof(IcpState), cudaMemcpyDeviceToHost, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  if (out29) memcpy(out29, hs->sys, 29 * sizeof(float));
  if (R_rel) memcpy(R_rel, hs->Rrel, 36);
  if (t_rel) memcpy(t_rel, hs->trel, 12);
  if (iters) *iters = hs->active ? hs->iter : 0;
  if (valid) *valid = hs->active ? hs->valid : 0;
  return SSF_OK;
}

int ssf_icp_begin(SsfHandle h, const float* R_init, const float* t_init) {
  H_CHECK_IDLE(h);
  if ((R_init == nullptr) != (t_init == nullptr)) return SSF_ERR_INVALID_ARG;
  if (R_init) launch_icp_begin(e, R_init, t_init);
  else launch_icp_begin_from_pose(e);
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_icp_build(SsfHandle h, int src_begin, int src_count, float out29[29]) {
  H_CHECK_IDLE(h);
  if (!out29 || src_begin < 0 || src_count < 0 || (src_begin & 3) || src_begin + src_count > e->cap)
    return SSF_ERR_INVALID_ARG;
  if (src_count > 0) launch_icp_build_range(e, src_begin, src_count);
  else SSF_CUDA(e, cudaMemsetAsync(e->icp->sys, 0, sizeof(float) * 32, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(out29, e->icp->sys, 29 * sizeof(float), cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_icp_solve(SsfHandle h, const float sys29[29], int* done) {
  H_CHECK_IDLE(h);
  if (!sys29) return SSF_ERR_INVALID_ARG;
  int rc = ensure_scratch(e, 256);
  if (rc) return rc;
  SSF_CUDA(e, cudaMemcpyAsync(e->scratch, sys29, 29 * sizeof(float), cudaMemcpyDefault, e->stream));
  launch_icp_solve(e, reinterpret_cast<const float*>(e->scratch));
  int d = 0;
  SSF_CUDA(e, cudaMemcpyAsync(&d, &e->icp->done, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  if (done) *done = d;
  return SSF_OK;
}

int ssf_icp_finish(SsfHandle h, int apply_to_pose, float R_rel[9], float t_rel[3], int* iters, int* valid) {
  H_CHECK_IDLE(h);
  launch_icp_finish(e, apply_to_pose != 0);
  IcpState* hs = e->h_icp;
  SSF_CUDA(e, cudaMemcpyAsync(hs, e->icp, sizeof(IcpState), cudaMemcpyDeviceToHost, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  if (R_rel) memcpy(R_rel, hs->Rrel, 36);
  if (t_rel) memcpy(t_rel, hs->trel, 12);
  if (iters) *iters = hs->active ? hs->iter : 0;
  if (valid) *valid = hs->active ? hs->valid : 0;
  return SSF_OK;
}

int ssf_peer_handle(SsfHandle h, void* handle64) {
  H_CHECK(h);
  if (!handle64) return SSF_ERR_INVALID_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == SSF_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t hd;
  SSF_CUDA(e, cudaIpcGetMemHandle(&hd, e->xbuf));
  memcpy(handle64, &hd, sizeof(hd));
  return SSF_OK;
}

int ssf_connect_peers(SsfHandle h, int rank, int world, const void* handles) {
  H_CHECK_IDLE(h);
  if (!handles || world < 1 || world > SSF_MAX_PEERS || rank < 0 || rank >= world) return SSF_ERR_INVALID_ARG;
  float* ptrs[SSF_MAX_PEERS] = {nullptr};
  for (int g = 0; g < world; g++) {
    if (g == rank) { ptrs[g] = e->xbuf; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, static_cast<const char*>(handles) + (size_t)g * SSF_PEER_HANDLE_BYTES, sizeof(hd));
    void* p = nullptr;
    if (e->xpeer_open[g]) { cudaIpcCloseMemHandle(e->xpeer_open[g]); e->xpeer_open[g] = nullptr; }   // reconnect
    SSF_CUDA(e, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    e->xpeer_open[g] = p;
    ptrs[g] = static_cast<float*>(p);
  }
  SSF_CUDA(e, cudaMemcpy(e->xpeers_dev, ptrs, sizeof(ptrs), cudaMemcpyHostToDevice));
  // The exchange buffer and the sequence counter are zero from ssf_create and must NOT be
  // cleared here: a peer that connected earlier may already have stored its first slot and
  // flag into this rank's buffer (clearing it made the first exchange hang, intermittently).
  e->xrank = rank;
  e->xworld = world;
  return SSF_OK;
}

int ssf_icp_tiled(SsfHandle h, const float* R_init, const float* t_init, int src_begin, int src_count,
                  float out29[29], float R_rel[9], float t_rel[3], int* iters, int* valid) {
  H_CHECK_IDLE(h);
  if ((R_init == nullptr) != (t_init == nullptr)) return SSF_ERR_INVALID_ARG;
  if (src_begin < 0 || src_count < 0 || (src_begin & 3) || src_begin + src_count > e->cap) return SSF_ERR_INVALID_ARG;
  if (e->xworld < 2) return SSF_ERR_STATE;
  if (R_init) launch_icp_begin(e, R_init, t_init);
  else launch_icp_begin_from_pose(e);
  launch_icp_tiled_loop(e, src_begin, src_count);
  launch_icp_finish(e, false);
  IcpState* hs = e->h_icp;
  SSF_CUDA(e, cudaMemcpyAsync(hs, e->icp, sizeof(IcpState), cudaMemcpyDeviceToHost, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  if (out29) memcpy(out29, hs->sys, 29 * sizeof(float));
  if (R_rel) memcpy(R_rel, hs->Rrel, 36);
  if (t_rel) memcpy(t_rel, hs->trel, 12);
  if (iters) *iters = hs->active ? hs->iter : 0;
  if (valid) *valid = hs->active ? hs->valid : 0;
  return SSF_OK;
}

int ssf_align(SsfHandle h, const SsfSurfels* source, int source_size, const float R_init[9], const float t_init[3],
              float R[9], float t[3], int* valid, int* iters, int* pairs, float out29[29]) {
  H_CHECK_IDLE(h);
  if (!source || source_size <= 0 || !R_init || !t_init || !R || !t) return SSF_ERR_INVALID_ARG;
  if (!source->positions || !source->colors || !source->orientations || !source->confidences) return SSF_ERR_INVALID_ARG;
  const size_t n = (size_t)source_size;
  // scratch: the four source members, then Lab, the matched records, flags and the result block
  const size_t floats = n * (3 + 3 + 9 + 1 + 3 + 12);
  const size_t bytes = floats * sizeof(float) + ((n + 15) & ~(size_t)15) + sizeof(AlignResult) + 64;
  int rc = ensure_scratch(e, bytes);
  if (rc) return rc;
  float* pos = reinterpret_cast<float*>(e->scratch);
  float* col = pos + 3 * n;
  float* ori = col + 3 * n;
  float* conf = ori + 9 * n;
  float* lab = conf + n;
  float* rec = lab + 3 * n;
  unsigned char* ok = reinterpret_cast<unsigned char*>(rec + 12 * n);
  AlignResult* res = reinterpret_cast<AlignResult*>(ok + ((n + 15) & ~(size_t)15));
  SSF_CUDA(e, cudaMemcpyAsync(pos, source->positions, n * 12, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(col, source->colors, n * 12, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(ori, source->orientations, n * 36, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(conf, source->confidences, n * 4, cudaMemcpyDefault, e->stream));
  AlignResult init = {};
  init.lab_sq = icp_lab_gate_sq();
  init.dist_sq = icp_dist_gate_sq();
  SSF_CUDA(e, cudaMemcpyAsync(res, &init, sizeof(init), cudaMemcpyHostToDevice, e->stream));
  launch_align(e, pos, col, ori, conf, source_size, lab, rec, ok, R_init, t_init, res);
  AlignResult out;
  SSF_CUDA(e, cudaMemcpyAsync(&out, res, sizeof(out), cudaMemcpyDeviceToHost, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  memcpy(R, out.R, 36);
  memcpy(t, out.t, 12);
  if (valid) *valid = out.valid;
  if (iters) *iters = out.iters;
  if (pairs) *pairs = out.pairs;
  if (out29) memcpy(out29, out.sys, 29 * sizeof(float));
  return SSF_OK;
}

int ssf_apply_deformation(SsfHandle h, const float* nodes_positions, const float* nodes_rotations,
                          const float* nodes_translations, int nb_nodes, const float* neighbours_weights,
                          const int32_t* neighbours_idx, int model_size) {
  H_CHECK_IDLE(h);
  if (!nodes_positions || !nodes_rotations || !nodes_translations || !neighbours_weights || !neighbours_idx ||
      nb_nodes <= 0 || model_size < 0 || model_size > e->cap)
    return SSF_ERR_INVALID_ARG;
  if (model_size == 0) return SSF_OK;
  const size_t nn = (size_t)nb_nodes, n = (size_t)model_size;
  const size_t bytes = nn * (3 + 9 + 3) * 4 + n * 32 + 64;
  int rc = ensure_scratch(e, bytes);
  if (rc) return rc;
  // 16-byte aligned first: the float4 / int4 per-supersurfel tables
  float* w = reinterpret_cast<float*>(e->scratch);
  int* idx = reinterpret_cast<int*>(w + 4 * n);
  float* gp = reinterpret_cast<float*>(idx + 4 * n);
  float* gr = gp + 3 * nn;
  float* gt = gr + 9 * nn;
  SSF_CUDA(e, cudaMemcpyAsync(w, neighbours_weights, n * 16, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(idx, neighbours_idx, n * 16, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(gp, nodes_positions, nn * 12, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(gr, nodes_rotations, nn * 36, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(gt, nodes_translations, nn * 12, cudaMemcpyDefault, e->stream));
  launch_apply_deformation(e, gp, gr, gt, w, idx, model_size, nb_nodes);
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

int ssf_get_markers(SsfHandle h, int which, float conf_thresh, float* points, float* colors, int capacity, int* count) {
  H_CHECK_IDLE(h);
  if (!points || !colors || capacity < 0 || (which != 0 && which != 1)) return SSF_ERR_INVALID_ARG;
  int n = e->S;
  if (which == 0) {
    int rc = read_report(e, false);
    if (rc) return rc;
    n = e->h_report->counters.nb_supersurfels;
  }
  if (count) *count = n;
  if (n > capacity) n = capacity;
  if (n == 0) return SSF_OK;
  int rc = ensure_scratch(e, (size_t)n * (18 + 24) * 4);
  if (rc) return rc;
  float* dp = reinterpret_cast<float*>(e->scratch);
  float* dc = dp + (size_t)n * 18;
  launch_markers(e, which == 1, n, conf_thresh, dp, dc);
  SSF_CUDA(e, cudaMemcpyAsync(points, dp, (size_t)n * 72, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaMemcpyAsync(colors, dc, (size_t)n * 96, cudaMemcpyDefault, e->stream));
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  SSF_LAUNCH_OK(e);
  return SSF_OK;
}

// tf::Matrix3x3::getRotation (the quaternion the nodes publish and write): Shoemake's method in
// double, as in tf's LinearMath/Matrix3x3.h
static void rotation_to_tf_quaternion(const float R[9], double q[4]) {
  const double m[3][3] = {{R[0], R[1], R[2]}, {R[3], R[4], R[5]}, {R[6], R[7], R[8]}};
  const double trace = m[0][0] + m[1][1] + m[2][2];
  if (trace > 0.0) {
    double s = sqrt(trace + 1.0);
    q[3] = s * 0.5;
    s = 0.5 / s;
    q[0] = (m[2][1] - m[1][2]) * s;
    q[1] = (m[0][2] - m[2][0]) * s;
    q[2] = (m[1][0] - m[0][1]) * s;
  } else {
    const int i = m[0][0] < m[1][1] ? (m[1][1] < m[2][2] ? 2 : 1) : (m[0][0] < m[2][2] ? 2 : 0);
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[i] = s * 0.5;
    s = 0.5 / s;
    q[3] = (m[k][j] - m[j][k]) * s;
    q[j] = (m[j][i] + m[i][j]) * s;
    q[k] = (m[k][i] + m[i][k]) * s;
  }
}

int ssf_format_tum_pose(SsfHandle h, const char* timestamp, char* line, size_t line_size) {
  H_CHECK_IDLE(h);
  if (!timestamp || !line || line_size == 0) return SSF_ERR_INVALID_ARG;
  int rc = read_report(e, false);
  if (rc) return rc;
  double q[4];
  rotation_to_tf_quaternion(e->h_report->pose.R, q);
  const float* t = e->h_report->pose.t;
  // operator<< of doubles: "%g" (6 significant digits), as the benchmark node's trajectory_file
  // (node/supersurfel_fusion_rgbd_benchmark_node.cpp:727-729); tf stores the origin in double
  const int w = snprintf(line, line_size, "%s %g %g %g %g %g %g %g\n", timestamp, (double)t[0], (double)t[1], (double)t[2],
                         q[0], q[1], q[2], q[3]);
  return (w < 0 || (size_t)w >= line_size) ? SSF_ERR_INVALID_ARG : SSF_OK;
}

int ssf_fuse(SsfHandle h) {
  H_CHECK_IDLE(h);
  launch_fuse(e);
  int rc = read_report(e, false);
  if (rc) return rc;
  SSF_LAUNCH_OK(e);
  fill_stats(e, 0.f);
  return SSF_OK;
}

int ssf_timer_start(SsfHandle h) {
  H_CHECK(h);
  SSF_CUDA(e, cudaEventRecord(e->ev0, e->stream));
  return SSF_OK;
}

int ssf_timer_stop(SsfHandle h, float* ms) {
  H_CHECK(h);
  SSF_CUDA(e, cudaEventRecord(e->ev1, e->stream));
  SSF_CUDA(e, cudaEventSynchronize(e->ev1));
  if (ms) SSF_CUDA(e, cudaEventElapsedTime(ms, e->ev0, e->ev1));
  return SSF_OK;
}

int ssf_synchronize(SsfHandle h) {
  H_CHECK(h);
  SSF_CUDA(e, cudaStreamSynchronize(e->stream));
  return SSF_OK;
}

int ssf_get_launch_count(SsfHandle h, uint64_t* launches) {
  H_CHECK(h);
  if (!launches) return SSF_ERR_INVALID_ARG;
  *launches = e->launches;
  return SSF_OK;
}

}  // extern "C"
