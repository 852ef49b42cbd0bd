// Internal (not exported) declarations of libssf: the engine state that lives in HBM
// and the stage launchers.  Public surface is include/ssf.h only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/ssf.h"

namespace ssf {

// ---- HBM layout ---------------------------------------------------------------
// Supersurfel sets (frame: S rows, model: capacity rows) are stored PLANAR: one
// contiguous float plane per scalar attribute, `stride` elements apart, so that a
// warp reading attribute k of 32 consecutive supersurfels issues one fully coalesced
// 128-byte request and a thread reading 4 consecutive supersurfels issues one
// float4 load.  The reference keeps array-of-float3 / Mat33 / Cov3
// (supersurfels.hpp:32-41), which strides a warp over 12/36/24-byte records.
enum Plane {
  P_POS = 0,      // 3: position
  P_COL = 3,      // 3: colour, RGB 0..255
  P_STAMP = 6,    // 2: int stamps (t0, t) bit-cast
  P_ORI = 8,      // 9: e1, e2, normal
  P_SHAPE = 17,   // 6: xx xy xz yy yz zz
  P_DIM = 23,     // 2
  P_CONF = 25,    // 1
  P_LAB = 26,     // 3: CIELab of P_COL, derived (hoisted out of the ICP / association loops)
  P_COUNT = 29
};

struct SurfelSet {
  float* base;     // P_COUNT * stride floats
  int stride;      // elements per plane, multiple of 4 (float4 alignment)
  __host__ __device__ float* plane(int p) const { return base + (size_t)p * stride; }
};

// per-superpixel TPS state (reference: SuperpixelRGBD, TPS_RGBD.hpp:33-38)
struct Superpixel { float4 xy_rg, theta_b, size; };
// running sums (reference: SuperpixelRGBDCoeffs, TPS_RGBD.hpp:40-44), integer so that
// atomic accumulation is order-free
struct SpSums {
  long long x, y, r, g, b, n;
  long long dx, dy, dxx, dyy, dxy, dn;
  long long dxd, dyd, dd;   // 2^-30 fixed point
  long long pad;
};

// Gauss-Newton state of one frame-to-model registration, device resident
struct IcpState {
  double tf_inc[16];     // accumulated increment (row-major 4x4)
  double JtJ[36];        // last built system, full symmetric
  double prev_error, error;
  float Rinit[9], tinit[3];
  float Rc[9], tc[3];    // transform the next system build uses (R_corres, t_corres)
  float tinc_top[3];     // t_inc at the top of the last executed iteration
  float sys[32];         // last built system (29 used)
  float Rrel[9], trel[3];
  float inliers;
  int iter;              // iterations executed
  int done;              // loop finished (converged, starved, or budget spent)
  int valid;
  int active;            // 0 when there is nothing to register against (bootstrap)
  unsigned int ticket;   // last-block-done counter
  unsigned int xseq;     // sequence number of the cross-GPU exchange (never reset)
};

// result block of one loop-closure registration (align_kernel); lab_sq / dist_sq are inputs
struct AlignResult {
  float R[9], t[3];
  int valid, iters, pairs;
  float sys[29];
  float lab_sq, dist_sq;
};

struct Counters {
  int nb_supersurfels, nb_visible, nb_removed, nb_matched, nb_inserted;
  int stamp;
  int cloud_count;
  int pad;         // scratch of the partition kernels (pre-compaction length)
  int seg_stamp;   // stamp of the frame being segmented / extracted (runs ahead of `stamp` when pipelined)
  int pad2;
};

struct DevicePose { float R[9]; float t[3]; };

// what a frame hands back to the host (one small D2H copy per frame)
struct FrameReport {
  Counters counters;
  DevicePose pose;
  int icp_active, icp_valid, icp_iters;
  float icp_inliers;
  double icp_error;
};


// The frame's report + stamp bookkeeping (supersurfel_fusion.cu:521), one thread.
__device__ __forceinline__ void frame_report(Counters* counters, const DevicePose* pose, const IcpState* icp, FrameReport* rep,
                                             int advance) {
  rep->counters = *counters;
  rep->pose = *pose;
  rep->icp_active = icp->active;
  rep->icp_valid = icp->active ? icp->valid : 0;
  rep->icp_iters = icp->active ? icp->iter : 0;
  rep->icp_inliers = icp->active ? icp->inliers : 0.0f;
  rep->icp_error = icp->active ? icp->error : 0.0;
  if (advance & 1) counters->stamp += 1;                  // supersurfel_fusion.cu:521
  if (advance & 2) counters->seg_stamp = counters->stamp; // synchronous mode: the two stamps move together
}

// Everything that belongs to ONE frame on its way through the stages: the segmentation images and
// per-superpixel state, and what segmentation + extraction hand to registration + fusion.  One set
// per pipeline stage, so that consecutive frames can be in different stages (ssf_submit_frame).
constexpr int SSF_SLOTS = 6;          // upper bound of pipeline stages = frames in flight
struct FrameSlot {
  uint8_t* in_rgb;     // staging of the raw inputs: per slot, so that the upload of the next frame (own copy
  float* in_depth;     // stream) overlaps the first stage of the previous one
  uchar4* rgba;
  float* disp;
  int* labels;
  int* bound;
  unsigned char* inliers;
  Superpixel* sp;
  SpSums* sums;
  int2* lmap;
  SurfelSet frame;
  float4* ftab;
  unsigned char* matched;
  unsigned long long* best;
};

struct Engine {
  SsfConfig cfg;
  int device;
  cudaStream_t own_stream, stream;
  cudaStream_t stage_stream[SSF_SLOTS];   // pipelined mode: one stream per stage ([0] = `stream`)
  cudaStream_t copy_stream;               // pipelined mode: input uploads (H2D DMA runs beside the stage kernels)
  FrameSlot slot[SSF_SLOTS];              // slot[0] is what the synchronous entry points use
  int cur_slot;
  int nb_stages;                          // stages = frames in flight of the pipelined mode (SSF_PIPELINE_STAGES)
  cudaEvent_t ev0, ev1, evf0, evf1;
  std::string err;
  cudaError_t launch_err;                 // first kernel-launch failure since the last entry-point check
  int failed;                             // a pipelined submit failed half-way: the handle refuses further work
  int nvtx;                               // SSF_NVTX=1: NVTX ranges around the stages (host side of the enqueue)
  uint64_t launches;

  int W, H, S, gx, gy, cap;
  size_t npix;

  // images
  uchar4* rgba;
  float* disp;
  int* labels;
  int* bound;
  unsigned char* inliers;
  int2* lmap;            // (label, slanted depth bits) interleaved: one 8-byte gather per lookup
  uint8_t* in_rgb;       // staging for the raw inputs
  float* in_depth;
  float* depth_f;        // bilateral-filtered depth (optional ingest)
  uint16_t* in_depth16;  // staging of a 16-bit depth image (ssf_process_frame_depth16)

  // TPS
  Superpixel* sp;
  SpSums* sums;
  float4* samples;
  void* rng;             // curandState[S * nb_samples]
  float* filt_a;         // 8 floats per node (X, Z, px, py), double buffered
  float* filt_b;
  int tps_tma;           // 1: the fused pass stages its label tile with a 2-D tensor-map TMA copy (SSF_TPS_TMA; needs W % 4 == 0)
  alignas(64) unsigned char label_map[SSF_SLOTS][128];   // CUtensorMap over each slot's label image
  int tps_occ;           // resident CTAs per SM the fused pass is compiled for (SSF_TPS_OCC: 3 = unconstrained, 4)
  int tps_fused;         // 1 (default): relabelling passes derive the means themselves, no merge launches (SSF_TPS_FUSED)
  int tps_persistent;    // 1: whole segmentation is one cooperative kernel
  int tps_grid;          // its grid (one CTA per SM)
  unsigned int* tps_barrier;   // ticket of its grid-wide barrier
  int tps_cache_slots;   // superpixels its per-CTA shared-memory cache holds
  unsigned long long* tps_trace;   // profiling aid (SSF_TPS_TRACE=<file>), normally NULL
  unsigned long long* xsums;  // extraction accumulators, 16 per superpixel

  // supersurfels
  SurfelSet frame, model, model_alt;
  float4* ftab;          // per frame supersurfel: (L,a,b,conf) (nx,ny,nz,0)
  unsigned char* matched;
  unsigned long long* best;   // packed (dist bits << 32 | model id) arg-min keys
  int* states;
  int* scan_tmp;

  // registration
  IcpState* icp;
  float* icp_partials;   // [grid][32]
  int icp_grid;
  // Programmatic dependent launch (SSF_PDL) inside the SYNCHRONOUS frame graph: 0 plain graph edges; 1 every kernel
  // lets its successor launch at its top; 2 as 1, but the fused segmentation pass triggers only after its decisions;
  // 3 only the pass -> pass edges are programmatic (late trigger).  pdl_now is the mode of the launches being
  // enqueued right now (0 outside that graph's capture: the pipelined stage graphs measured slower with any of them)
  int pdl, pdl_pipe, pdl_now;
  int icp_occ;           // resident CTAs per SM the system kernel is compiled for
  int icp_stages;        // staging of the streamed planes (SSF_ICP_STAGES): 1 direct loads (default), 2..4 TMA ring, < 0 pipelined kernel
  int icp_debug;         // profiling knob, see IcpArgs::debug
  int icp_loop;          // 1 (default): small visible models register in one cluster launch (SSF_ICP_LOOP=0: never)
  // tile-parallel registration over peer memory
  float* xbuf;           // this rank's exchange buffer: [2 parities][SSF_MAX_PEERS][64 floats]
  float** xpeers_dev;    // device array: exchange buffer of every rank (own entry = xbuf)
  void* xpeer_open[SSF_MAX_PEERS];   // IPC mappings to close
  int xrank, xworld;

  Counters* counters;
  DevicePose* pose;
  // pinned host mirrors
  Counters* h_counters;
  DevicePose* h_pose;
  IcpState* h_icp;
  SsfFrameStats stats;

  void* scratch;         // AoS <-> planar conversion buffer
  size_t scratch_bytes;
};

// Every kernel of the library is launched through here.  With programmatic dependent launch on, the
// launch carries cudaLaunchAttributeProgrammaticStreamSerialization, which (also under stream capture,
// as a programmatic graph edge) lets the kernel start while its predecessor drains; the kernels order
// themselves with pdl_wait() / pdl_trigger() (ssf_math.cuh).  pass_edge: the launch is a fused
// segmentation pass (the only programmatic edges of mode 3).
template <typename... P, typename... A>
inline void launch_kernel(Engine* e, bool pass_edge, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = e->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (e->pdl_now == 1 || e->pdl_now == 2 || (e->pdl_now >= 3 && pass_edge)) ? 1 : 0;
  const cudaError_t rc = cudaLaunchKernelEx(&cfg, kernel, P(args)...);
  if (rc != cudaSuccess && e->launch_err == cudaSuccess) e->launch_err = rc;   // surfaced by the entry point (launch_status)
}
template <typename... P, typename... A>
inline void launch_pdl(Engine* e, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, A&&... args) {
  launch_kernel(e, false, kernel, grid, block, smem, static_cast<A&&>(args)...);
}

// ---- stage launchers (each enqueues on e->stream, no host sync) -----------------
void launch_apply_deformation(Engine* e, const float* node_pos, const float* node_rot, const float* node_trans,
                              const float* weights, const int* nn, int n, int n_nodes);
void launch_markers(Engine* e, bool frame, int n, float conf_thresh, float* points_dev, float* colors_dev);
// ingest (ssf_ingest.cu)
int launch_bilateral(Engine* e, const float* src_dev, float* dst_dev, int kernel_size, float sigma_color,
                     float sigma_spatial);
void launch_gray(Engine* e, const uint8_t* rgb_dev, uint8_t* gray_dev);
void launch_depth16(Engine* e, const uint16_t* d16_dev, float* out_dev, float scale);
int icp_chunk_size();     // supersurfels one CTA of the system kernel consumes per grid-stride step
int icp_ctas_per_sm();
int icp_configure(int stages);   // opts the system kernel in to its shared-memory ring; returns the clamped depth    // resident CTAs per SM the system kernel is compiled for
void launch_icp_system(Engine* e, const SurfelSet& src, const int* n_dev, int n_host, bool solve);
void launch_icp_begin(Engine* e, const float* Rinit_or_null, const float* tinit_or_null);  // host ptrs
void launch_icp_begin_from_pose(Engine* e);
void launch_icp_set_transform(Engine* e, const float* R, const float* t);
void launch_icp_loop(Engine* e);
void launch_icp_build_range(Engine* e, int begin, int count);
void launch_icp_solve(Engine* e, const float* sys29_dev);
void launch_icp_tiled_loop(Engine* e, int begin, int count);
void launch_icp_finish(Engine* e, bool apply_to_pose);
void launch_icp_registration_loop(Engine* e, bool apply_to_pose);   // begin + loop + finish in one launch
int icp_loop_max_sources();
bool icp_loop_equivalent(const Engine* e);
void launch_align(Engine* e, const float* pos, const float* col, const float* ori, const float* conf, int n,
                  float* lab, float* rec, unsigned char* ok, const float* Rinit, const float* tinit, AlignResult* out_dev);
float icp_lab_gate_sq();
float icp_dist_gate_sq();
void launch_ingest(Engine* e, const uint8_t* rgb_dev, size_t rgb_stride, const float* depth_dev,
                   size_t depth_stride);
int tps_step_count(const Engine* e);
void launch_tps(Engine* e, int first = 0, int last = -1);   // segmentation steps [first, last)
void launch_extract(Engine* e);
// model update; with `report` the frame's report is written (and the stamps advanced, see frame_report) by the
// last kernel of the update instead of by a launch of its own
void launch_fuse(Engine* e, FrameReport* report = nullptr, int advance = 0);
void launch_build_lmap(Engine* e, const float* slanted_dev);
void launch_frame_tables(Engine* e);
void launch_model_lab(Engine* e, int n);
void launch_pack(Engine* e, const SurfelSet& set, int n, const SsfSurfels& dst_dev);   // planar -> member layout
void launch_unpack(Engine* e, const SsfSurfels& src_dev, int n, const SurfelSet& set);
void launch_transform_model(Engine* e, const float* R, const float* t);
void launch_local_cloud(Engine* e, float radius, float* pos_dev, float* nrm_dev, int capacity);
void launch_preview(Engine* e, uint8_t* bgr_dev);
void launch_invalidate(Engine* e, const uint8_t* mask_dev);
void tps_init_rng(Engine* e);

#define SSF_CUDA(e, call)                                                            \
  do {                                                                               \
    cudaError_t _err = (call);                                                       \
    if (_err != cudaSuccess) {                                                       \
      (e)->err = std::string(#call) + ": " + cudaGetErrorString(_err);               \
      return SSF_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

}  // namespace ssf

struct SsfEngine : public ssf::Engine {};
