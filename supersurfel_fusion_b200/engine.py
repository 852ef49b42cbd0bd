"""Python host mirror of the reference's ``SupersurfelFusion`` class over the libssf C-ABI.

Method names, argument meaning and defaults follow
``core/include/supersurfel_fusion/supersurfel_fusion.hpp:40-143`` of the reference so that
code (and tests) written against the reference class read the same here.  All compute
happens in ``libssf.so`` (hand-written sm_100a CUDA); there is no CPU fallback: a missing
library or a machine without a CUDA device raises :class:`SsfError`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SSF_OK = 0
_ERR_NAMES = {-1: "SSF_ERR_INVALID_ARG", -2: "SSF_ERR_CUDA", -3: "SSF_ERR_NO_DEVICE", -4: "SSF_ERR_IO",
              -5: "SSF_ERR_STATE"}


class SsfError(RuntimeError):
    pass


def lib_path():
    """The in-tree library; SSF_LIB points at another build of the same sources (compile-time A/B experiments)."""
    return os.environ.get("SSF_LIB") or os.path.join(_HERE, "libssf.so")


class CamParam(C.Structure):
    """core/include/supersurfel_fusion/cam_param.hpp:27-31"""
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("height", C.c_int), ("width", C.c_int)]


class SsfConfig(C.Structure):
    _fields_ = [("cam", CamParam), ("cell_size", C.c_int),
                ("lambda_pos", C.c_float), ("lambda_bound", C.c_float), ("lambda_size", C.c_float),
                ("lambda_disp", C.c_float), ("thresh_disp", C.c_float),
                ("seg_iter", C.c_int), ("seg_use_ransac", C.c_int), ("nb_samples", C.c_int),
                ("filter_iter", C.c_int), ("filter_alpha", C.c_float), ("filter_beta", C.c_float),
                ("filter_threshold", C.c_float), ("range_min", C.c_float), ("range_max", C.c_float),
                ("delta_t", C.c_int), ("conf_thresh", C.c_float), ("nb_supersurfels_max", C.c_int),
                ("icp_iter", C.c_int), ("icp_cov_thresh", C.c_double),
                ("enable_loop_closure", C.c_int), ("enable_mod", C.c_int)]


class SsfSurfels(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("colors", C.c_void_p), ("stamps", C.c_void_p),
                ("orientations", C.c_void_p), ("shapes", C.c_void_p), ("dims", C.c_void_p),
                ("confidences", C.c_void_p)]


class SsfPlanarView(C.Structure):
    _fields_ = [("base", C.c_void_p), ("stride", C.c_int), ("count", C.c_int), ("planes", C.c_int)]


class SsfFrameStats(C.Structure):
    _fields_ = [("stamp", C.c_int32), ("nb_supersurfels", C.c_int32), ("nb_visible", C.c_int32),
                ("nb_removed", C.c_int32), ("nb_matched", C.c_int32), ("nb_inserted", C.c_int32),
                ("icp_ran", C.c_int32), ("icp_valid", C.c_int32), ("icp_iters", C.c_int32),
                ("icp_inliers", C.c_float), ("icp_error", C.c_double), ("gpu_ms", C.c_float),
                ("ms_ingest", C.c_float), ("ms_segmentation", C.c_float), ("ms_extraction", C.c_float),
                ("ms_registration", C.c_float), ("ms_fusion", C.c_float)]


SSF_FLAG_BILATERAL = 1      # include/ssf.h
SSF_FLAG_STAGE_TIMING = 2

# every symbol include/ssf.h declares
EXPORTS = [
    "ssf_config_default", "ssf_create", "ssf_destroy", "ssf_set_stream", "ssf_last_error", "ssf_is_initialized",
    "ssf_process_frame", "ssf_process_frame_depth16", "ssf_process_frame_device", "ssf_bilateral_filter",
    "ssf_get_filtered_depth", "ssf_get_gray", "ssf_get_frame_stats", "ssf_prepare", "ssf_submit_frame", "ssf_wait_frame", "ssf_get_pipeline_depth", "ssf_plan_pipeline", "ssf_plan_weights", "ssf_get_model_view", "ssf_get_frame_view", "ssf_get_pose", "ssf_set_pose",
    "ssf_get_stamp", "ssf_set_stamp", "ssf_get_counts", "ssf_get_nb_superpixels", "ssf_copy_model",
    "ssf_copy_frame", "ssf_get_segmentation", "ssf_render_preview", "ssf_get_slanted_depth", "ssf_export_model",
    "ssf_extract_local_point_cloud", "ssf_invalidate_frame_supersurfels", "ssf_transform_model", "ssf_set_model",
    "ssf_set_frame", "ssf_set_segmentation", "ssf_tps_segment", "ssf_get_ransac_samples",
    "ssf_generate_supersurfels", "ssf_icp_system", "ssf_icp_system_enqueue", "ssf_icp", "ssf_icp_begin",
    "ssf_icp_build", "ssf_icp_solve", "ssf_icp_finish", "ssf_peer_handle", "ssf_connect_peers", "ssf_icp_tiled",
    "ssf_fuse", "ssf_align", "ssf_apply_deformation", "ssf_get_markers", "ssf_format_tum_pose",
    "ssf_timer_start", "ssf_timer_stop", "ssf_synchronize", "ssf_get_launch_count",
]


def load_library():
    """dlopen libssf.so.  Raises SsfError when it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise SsfError("libssf.so is missing: build it with `python -m supersurfel_fusion_b200.build` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(path)
        lib.ssf_last_error.restype = C.c_char_p
        lib.ssf_last_error.argtypes = [C.c_void_p]
        for name in EXPORTS:
            fn = getattr(lib, name)
            if name != "ssf_last_error":
                fn.restype = C.c_int
        lib.ssf_process_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                          C.c_uint32]
        lib.ssf_process_frame_depth16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_float,
                                                  C.c_void_p, C.c_uint32]
        lib.ssf_bilateral_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_float,
                                             C.c_void_p]
        lib.ssf_get_filtered_depth.argtypes = [C.c_void_p, C.c_void_p]
        lib.ssf_get_gray.argtypes = [C.c_void_p, C.c_void_p]
        lib.ssf_submit_frame.argtypes = lib.ssf_process_frame.argtypes
        lib.ssf_prepare.argtypes = [C.c_void_p, C.c_uint32]
        lib.ssf_wait_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ssf_process_frame_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        lib.ssf_tps_segment.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.ssf_icp_system.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.ssf_icp_system_enqueue.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.ssf_icp.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        lib.ssf_apply_deformation.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_int]
        lib.ssf_get_markers.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.ssf_format_tum_pose.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t]
        lib.ssf_align.argtypes = [C.c_void_p] + [C.c_void_p, C.c_int] + [C.c_void_p] * 8
        lib.ssf_set_model.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.ssf_copy_model.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.ssf_copy_frame.argtypes = [C.c_void_p, C.c_void_p]
        lib.ssf_set_frame.argtypes = [C.c_void_p, C.c_void_p]
        lib.ssf_get_segmentation.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        lib.ssf_set_segmentation.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        lib.ssf_extract_local_point_cloud.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                                      C.c_void_p]
        lib.ssf_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.ssf_destroy.argtypes = [C.c_void_p]
        _LIB = lib
    return _LIB


def _ptr(a):
    """numpy array -> host pointer; torch tensor -> its data_ptr (host or device); int passthrough."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


class Supersurfels:
    """Host copy of a supersurfel set in the reference's member layout (supersurfels.hpp:32-41)."""
    FIELDS = (("positions", np.float32, 3), ("colors", np.float32, 3), ("stamps", np.int32, 2),
              ("orientations", np.float32, 9), ("shapes", np.float32, 6), ("dims", np.float32, 2),
              ("confidences", np.float32, 1))

    def __init__(self, n):
        self.n = int(n)
        for name, dt, w in self.FIELDS:
            setattr(self, name, np.zeros((self.n, w) if w > 1 else (self.n,), dtype=dt))

    @classmethod
    def from_arrays(cls, **arrays):
        n = len(arrays["positions"])
        out = cls(n)
        for name, dt, w in cls.FIELDS:
            if name in arrays and arrays[name] is not None:
                setattr(out, name, np.ascontiguousarray(arrays[name], dtype=dt).reshape((n, w) if w > 1 else (n,)))
        return out

    def view(self):
        return SsfSurfels(*[_ptr(getattr(self, name)) for name, _, _ in self.FIELDS])

    def as_dict(self):
        return {name: getattr(self, name) for name, _, _ in self.FIELDS}


class SupersurfelFusion:
    """Mirror of ``supersurfel_fusion::SupersurfelFusion`` (supersurfel_fusion.hpp:40-143)."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._h = None
        self._device = device
        self.cfg = None

    # -- lifecycle -----------------------------------------------------------------
    def _check(self, rc, what):
        if rc != SSF_OK:
            msg = self._lib.ssf_last_error(self._h).decode() if self._h else ""
            raise SsfError("%s failed: %s %s" % (what, _ERR_NAMES.get(rc, rc), msg))

    def initialize(self, cam_param, cell_size=16, lambda_pos=50.0, lambda_bound=1000.0, lambda_size=10000.0,
                   lambda_disp=1000000.0, thresh_disp=0.0001, seg_iter=10, seg_use_ransac=True, nb_samples=16,
                   filter_iter=4, filter_alpha=0.1, filter_beta=1.0, filter_threshold=0.05, range_min=0.2,
                   range_max=5.0, delta_t=20, conf_thresh=2500.0, nb_supersurfels_max=50000, icp_iter=10,
                   icp_cov_thresh=0.04, nb_features=2000, features_scale_factor=1.2, features_nb_levels=8,
                   ini_th_fast=20, min_th_fast=7, untracked_threshold=10, enable_loop_closure=True,
                   enable_mod=True):
        """initialize() of the reference (supersurfel_fusion.hpp:46-74), same defaults.  The six
        sparse-VO arguments are accepted and ignored (out-of-scope neighbour); enable_loop_closure
        and enable_mod are recorded in the configuration but not acted on: the loop detector and the
        moving-object detector are the caller's, their results enter through align /
        applyDeformation / setPose / transformModel / invalidateFrameSupersurfels (include/ssf.h)."""
        if self._h:
            self.close()
        cfg = SsfConfig()
        self._lib.ssf_config_default(C.byref(cfg))
        if not isinstance(cam_param, CamParam):
            cam_param = CamParam(*cam_param)
        cfg.cam = cam_param
        cfg.cell_size = cell_size
        cfg.lambda_pos, cfg.lambda_bound, cfg.lambda_size = lambda_pos, lambda_bound, lambda_size
        cfg.lambda_disp, cfg.thresh_disp = lambda_disp, thresh_disp
        cfg.seg_iter, cfg.seg_use_ransac, cfg.nb_samples = seg_iter, int(bool(seg_use_ransac)), nb_samples
        cfg.filter_iter, cfg.filter_alpha, cfg.filter_beta = filter_iter, filter_alpha, filter_beta
        cfg.filter_threshold = filter_threshold
        cfg.range_min, cfg.range_max = range_min, range_max
        cfg.delta_t, cfg.conf_thresh, cfg.nb_supersurfels_max = delta_t, conf_thresh, nb_supersurfels_max
        cfg.icp_iter, cfg.icp_cov_thresh = icp_iter, icp_cov_thresh
        cfg.enable_loop_closure, cfg.enable_mod = int(bool(enable_loop_closure)), int(bool(enable_mod))
        h = C.c_void_p()
        rc = self._lib.ssf_create(C.byref(cfg), C.c_int(self._device), C.byref(h))
        if rc != SSF_OK:
            raise SsfError("ssf_create failed: %s (libssf needs a CUDA device; there is no CPU fallback)"
                           % _ERR_NAMES.get(rc, rc))
        self._h = h
        self.cfg = cfg
        n = C.c_int()
        self._check(self._lib.ssf_get_nb_superpixels(self._h, C.byref(n)), "ssf_get_nb_superpixels")
        self.nbSuperpixels = n.value
        self.width, self.height = cam_param.width, cam_param.height
        self._pending = []          # input buffers of the frames in flight (submitFrame keeps them alive)
        return self

    def prepare(self, flags=0):
        """Build and upload every CUDA graph the given flags need before the first frame
        (ssf_prepare): the first processFrame / submitFrame calls then cost what the later ones do."""
        self._check(self._lib.ssf_prepare(self._h, C.c_uint32(flags)), "ssf_prepare")
        return self

    def close(self):
        if self._h:
            self._lib.ssf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def isInitialized(self):
        return bool(self._h) and bool(self._lib.ssf_is_initialized(self._h))

    def setStream(self, cuda_stream):
        self._check(self._lib.ssf_set_stream(self._h, C.c_void_p(cuda_stream or 0)), "ssf_set_stream")

    # -- per-frame -------------------------------------------------------------------
    def processFrame(self, rgb_h, depth_h, pose_prior=None, flags=0):
        """processFrame(rgb 8UC3 RGB, depth 32FC1 metres) (supersurfel_fusion.cu:166-530).
        Accepts numpy arrays (host) or torch tensors (pinned host / device)."""
        prior = None
        if pose_prior is not None:
            R, t = pose_prior
            prior = np.concatenate([np.asarray(R, np.float32).reshape(9), np.asarray(t, np.float32).reshape(3)])
        rgb_h, depth_h, rs, ds = self._check_images(rgb_h, depth_h, "processFrame")
        rc = self._lib.ssf_process_frame(self._h, _ptr(rgb_h), rs, _ptr(depth_h), ds, _ptr(prior), flags)
        self._check(rc, "ssf_process_frame")
        return self.getFrameStats()

    def processFrameDepth16(self, rgb_h, depth16_h, depth_scale, pose_prior=None, flags=0):
        """processFrame on the raw 16-bit depth image of the TUM / live drivers: the
        depth.convertTo(CV_32FC1, depth_scale) of the node runs on the device."""
        prior = None
        if pose_prior is not None:
            R, t = pose_prior
            prior = np.concatenate([np.asarray(R, np.float32).reshape(9), np.asarray(t, np.float32).reshape(3)])
        rgb_h = np.ascontiguousarray(rgb_h, np.uint8)
        depth16_h = np.ascontiguousarray(depth16_h, np.uint16)
        if rgb_h.shape[:2] != (self.height, self.width) or depth16_h.shape != (self.height, self.width):
            raise SsfError("image size does not match the camera")
        rc = self._lib.ssf_process_frame_depth16(self._h, _ptr(rgb_h), self.width * 3, _ptr(depth16_h), self.width * 2,
                                                 float(depth_scale), _ptr(prior), flags)
        self._check(rc, "ssf_process_frame_depth16")
        return self.getFrameStats()

    def processFrameStaged(self, rgb_h, depth_h, pose_prior=None, dynamic_mask=None):
        """processFrame driven stage by stage through the stage entry points, which is how a caller
        with a moving-object detector plugs it in: segmentation -> generateSupersurfels -> [MOD hook:
        ssf_invalidate_frame_supersurfels(mask), what detectMotion does to frame.confidences,
        supersurfel_fusion.cu:198-213 / motion_detection.cu:573] -> registration (+ pose composition)
        -> fusion -> stamp + 1.  Same kernels as processFrame; depth is taken as already filtered."""
        rgb_h = np.ascontiguousarray(rgb_h, np.uint8)
        depth_h = np.ascontiguousarray(depth_h, np.float32)
        self._check(self._lib.ssf_tps_segment(self._h, _ptr(rgb_h), 0, _ptr(depth_h), 0), "ssf_tps_segment")
        self._check(self._lib.ssf_generate_supersurfels(self._h), "ssf_generate_supersurfels")
        if dynamic_mask is not None:
            self.invalidateFrameSupersurfels(dynamic_mask)
        if pose_prior is not None:
            self.setPose(*pose_prior)
        valid, _, _, info = self.icp()
        self.icpFinish(apply_to_pose=True)          # supersurfel_fusion.cu:313-328
        stats = self.fuse()
        stamp = self.getStamp()
        self.setStamp(stamp + 1)                    # supersurfel_fusion.cu:521
        stats.update(stamp=stamp, icp_valid=int(valid), icp_iters=info["iters"])
        return stats

    # -- ingest (supersurfel_fusion.cu:171-181) -----------------------------------------
    def bilateralFilter(self, depth, kernel_size=-1, sigma_color=0.03, sigma_spatial=4.5):
        """cv::cuda::bilateralFilter(depth, depth, -1, 0.03, 4.5) (supersurfel_fusion.cu:180), out of place."""
        depth = np.ascontiguousarray(depth, np.float32)
        out = np.empty((self.height, self.width), np.float32)
        rc = self._lib.ssf_bilateral_filter(self._h, _ptr(depth), self.width * 4, int(kernel_size), float(sigma_color),
                                            float(sigma_spatial), _ptr(out))
        self._check(rc, "ssf_bilateral_filter")
        return out

    def getFilteredDepth(self):
        out = np.empty((self.height, self.width), np.float32)
        self._check(self._lib.ssf_get_filtered_depth(self._h, _ptr(out)), "ssf_get_filtered_depth")
        return out

    def getGray(self):
        """cv::cuda::cvtColor(rgb, gray, CV_RGB2GRAY) of the last frame (supersurfel_fusion.cu:175-177)."""
        out = np.empty((self.height, self.width), np.uint8)
        self._check(self._lib.ssf_get_gray(self._h, _ptr(out)), "ssf_get_gray")
        return out

    def _check_images(self, rgb, depth, what):
        """dtype / shape / stride checks shared by processFrame and submitFrame; returns the
        (possibly compacted) arrays and their row strides in bytes."""
        if isinstance(rgb, np.ndarray) != isinstance(depth, np.ndarray):
            raise SsfError("%s: rgb and depth must both be numpy arrays or both be torch tensors" % what)
        if isinstance(rgb, np.ndarray):
            if rgb.dtype != np.uint8 or depth.dtype != np.float32:
                raise SsfError("%s expects uint8 RGB and float32 depth" % what)
            if rgb.shape != (self.height, self.width, 3) or depth.shape != (self.height, self.width):
                raise SsfError("image size does not match the camera")
            if rgb.strides[1:] != (3, 1) or depth.strides[1] != 4:
                rgb, depth = np.ascontiguousarray(rgb), np.ascontiguousarray(depth)
            return rgb, depth, rgb.strides[0], depth.strides[0]
        if hasattr(rgb, "data_ptr"):      # torch tensors: pinned host or device memory
            import torch
            if rgb.dtype != torch.uint8 or depth.dtype != torch.float32:
                raise SsfError("%s expects uint8 RGB and float32 depth" % what)
            if tuple(rgb.shape) != (self.height, self.width, 3) or tuple(depth.shape) != (self.height, self.width):
                raise SsfError("image size does not match the camera")
            if not rgb.is_contiguous() or not depth.is_contiguous():
                raise SsfError("%s: torch inputs must be contiguous" % what)
            return rgb, depth, self.width * 3, self.width * 4
        raise SsfError("%s: unsupported image type %r" % (what, type(rgb)))

    def submitFrame(self, rgb, depth, pose_prior=None, flags=0):
        """Pipelined processFrame: enqueue and return.  Up to pipelineDepth() frames (= pipeline
        stages, default 4) may be in flight.  Inputs: numpy arrays or torch tensors (pinned host or
        device), same dtype / shape / stride rules as processFrame; they are kept referenced until
        waitFrame() has returned the frame.  Only pinned host or device memory gives an asynchronous
        copy: with pageable numpy arrays cudaMemcpy2DAsync blocks until the copy is staged and the
        stage-0 stream serialises behind it, so frames no longer overlap their own upload."""
        prior = None
        if pose_prior is not None:
            R, t = pose_prior
            prior = np.concatenate([np.asarray(R, np.float32).reshape(9), np.asarray(t, np.float32).reshape(3)])
        rgb, depth, rs, ds = self._check_images(rgb, depth, "submitFrame")
        rc = self._lib.ssf_submit_frame(self._h, _ptr(rgb), rs, _ptr(depth), ds, _ptr(prior), flags)
        self._check(rc, "ssf_submit_frame")
        self._pending.append((rgb, depth, prior))

    def pipelineDepth(self):
        """Frames that may be in flight through submitFrame (= pipeline stages)."""
        n = C.c_int(0)
        self._check(self._lib.ssf_get_pipeline_depth(self._h, C.byref(n)), "ssf_get_pipeline_depth")
        return n.value

    def waitFrame(self):
        """Blocks until the oldest submitted frame is done; returns (stats, R, t)."""
        st = SsfFrameStats()
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        self._check(self._lib.ssf_wait_frame(self._h, C.byref(st), _ptr(R), _ptr(t)), "ssf_wait_frame")
        if self._pending:
            self._pending.pop(0)
        return {k: getattr(st, k) for k, _ in SsfFrameStats._fields_}, R.reshape(3, 3), t

    def processFrameDevice(self, rgb_dev, depth_dev, pose_prior=None, flags=0):
        prior = None
        if pose_prior is not None:
            R, t = pose_prior
            prior = np.concatenate([np.asarray(R, np.float32).reshape(9), np.asarray(t, np.float32).reshape(3)])
        rc = self._lib.ssf_process_frame_device(self._h, _ptr(rgb_dev), _ptr(depth_dev), _ptr(prior), flags)
        self._check(rc, "ssf_process_frame_device")

    def getFrameStats(self):
        st = SsfFrameStats()
        self._check(self._lib.ssf_get_frame_stats(self._h, C.byref(st)), "ssf_get_frame_stats")
        return {k: getattr(st, k) for k, _ in SsfFrameStats._fields_}

    # -- getters (supersurfel_fusion.hpp:85-91) -----------------------------------------
    def getPose(self):
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        self._check(self._lib.ssf_get_pose(self._h, _ptr(R), _ptr(t)), "ssf_get_pose")
        return R.reshape(3, 3), t

    def setPose(self, R, t):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        t = np.ascontiguousarray(t, np.float32).reshape(3)
        self._check(self._lib.ssf_set_pose(self._h, _ptr(R), _ptr(t)), "ssf_set_pose")

    def getStamp(self):
        s = C.c_int()
        self._check(self._lib.ssf_get_stamp(self._h, C.byref(s)), "ssf_get_stamp")
        return s.value

    def setStamp(self, stamp):
        self._check(self._lib.ssf_set_stamp(self._h, C.c_int(stamp)), "ssf_set_stamp")

    def getCounts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.ssf_get_counts(self._h, C.byref(a), C.byref(b), C.byref(c)), "ssf_get_counts")
        return a.value, b.value, c.value

    def getnbSupersurfels(self):
        return self.getCounts()[0]

    def getModel(self, n=None):
        if n is None:
            n = self.getnbSupersurfels()
        m = Supersurfels(n)
        v = m.view()
        self._check(self._lib.ssf_copy_model(self._h, C.byref(v), n), "ssf_copy_model")
        return m

    def getModelView(self):
        """getModel() without a copy: (device pointer, stride, count, planes) of the planar model storage."""
        v = SsfPlanarView()
        self._check(self._lib.ssf_get_model_view(self._h, C.byref(v)), "ssf_get_model_view")
        return v.base, v.stride, v.count, v.planes

    def getFrameView(self):
        v = SsfPlanarView()
        self._check(self._lib.ssf_get_frame_view(self._h, C.byref(v)), "ssf_get_frame_view")
        return v.base, v.stride, v.count, v.planes

    def getFrame(self):
        f = Supersurfels(self.nbSuperpixels)
        v = f.view()
        self._check(self._lib.ssf_copy_frame(self._h, C.byref(v)), "ssf_copy_frame")
        return f

    def getSegmentation(self):
        H, W, S = self.height, self.width, self.nbSuperpixels
        out = dict(labels=np.zeros((H, W), np.int32), bound=np.zeros((H, W), np.int32),
                   inliers=np.zeros((H, W), np.uint8), disp=np.zeros((H, W), np.float32),
                   slanted=np.zeros((H, W), np.float32), superpixels=np.zeros((S, 12), np.float32),
                   rgba=np.zeros((H, W, 4), np.uint8))
        rc = self._lib.ssf_get_segmentation(self._h, _ptr(out["labels"]), _ptr(out["bound"]), _ptr(out["inliers"]),
                                            _ptr(out["disp"]), _ptr(out["slanted"]), _ptr(out["superpixels"]),
                                            _ptr(out["rgba"]))
        self._check(rc, "ssf_get_segmentation")
        return out

    def computeSuperpixelSegIm(self):
        im = np.zeros((self.height, self.width, 3), np.uint8)
        self._check(self._lib.ssf_render_preview(self._h, _ptr(im)), "ssf_render_preview")
        return im

    def computeSlantedPlaneIm(self):
        im = np.zeros((self.height, self.width), np.float32)
        self._check(self._lib.ssf_get_slanted_depth(self._h, _ptr(im)), "ssf_get_slanted_depth")
        return im

    def exportModel(self, filename):
        self._check(self._lib.ssf_export_model(self._h, filename.encode()), "ssf_export_model")

    def extractLocalPointCloud(self, radius=None):
        """extractLocalPointCloud (supersurfel_fusion.cu:884-927); radius defaults to range_max."""
        n = max(self.getnbSupersurfels(), 1)
        pos = np.zeros((n, 3), np.float32)
        nrm = np.zeros((n, 3), np.float32)
        cnt = C.c_int()
        r = self.cfg.range_max if radius is None else radius
        rc = self._lib.ssf_extract_local_point_cloud(self._h, C.c_float(r), _ptr(pos), _ptr(nrm), n, C.byref(cnt))
        self._check(rc, "ssf_extract_local_point_cloud")
        return pos[:cnt.value], nrm[:cnt.value]

    def invalidateFrameSupersurfels(self, mask):
        mask = np.ascontiguousarray(mask, np.uint8)
        self._check(self._lib.ssf_invalidate_frame_supersurfels(self._h, _ptr(mask)), "ssf_invalidate")

    def transformModel(self, R, t):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        t = np.ascontiguousarray(t, np.float32).reshape(3)
        self._check(self._lib.ssf_transform_model(self._h, _ptr(R), _ptr(t)), "ssf_transform_model")

    # -- stage entry points ---------------------------------------------------------------
    def setModel(self, surfels, nb_supersurfels=None, nb_visible=None):
        n = surfels.n if nb_supersurfels is None else nb_supersurfels
        v = surfels.view()
        self._check(self._lib.ssf_set_model(self._h, C.byref(v), n, n if nb_visible is None else nb_visible),
                    "ssf_set_model")

    def setModelPointers(self, view, nb_supersurfels, nb_visible):
        """view: SsfSurfels of host-or-device pointers (e.g. torch tensors' data_ptr)."""
        self._check(self._lib.ssf_set_model(self._h, C.byref(view), nb_supersurfels, nb_visible), "ssf_set_model")

    def setFrame(self, surfels):
        v = surfels.view()
        self._check(self._lib.ssf_set_frame(self._h, C.byref(v)), "ssf_set_frame")

    def setSegmentation(self, labels=None, bound=None, inliers=None, slanted=None, rgba=None):
        def prep(a, dt):
            return None if a is None else (np.ascontiguousarray(a, dt) if isinstance(a, np.ndarray) else a)
        keep = [prep(labels, np.int32), prep(bound, np.int32), prep(inliers, np.uint8), prep(slanted, np.float32),
                prep(rgba, np.uint8)]
        rc = self._lib.ssf_set_segmentation(self._h, *[_ptr(k) for k in keep])
        self._check(rc, "ssf_set_segmentation")

    def tpsSegment(self, rgb, depth):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        self._check(self._lib.ssf_tps_segment(self._h, _ptr(rgb), 0, _ptr(depth), 0), "ssf_tps_segment")
        return self.getSegmentation()

    def getRansacSamples(self):
        s = np.zeros((self.nbSuperpixels, self.cfg.nb_samples, 4), np.float32)
        self._check(self._lib.ssf_get_ransac_samples(self._h, _ptr(s)), "ssf_get_ransac_samples")
        return s

    def generateSupersurfels(self):
        self._check(self._lib.ssf_generate_supersurfels(self._h), "ssf_generate_supersurfels")
        return self.getFrame()

    def icpSystem(self, R, t, n_src=0):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        t = np.ascontiguousarray(t, np.float32).reshape(3)
        out = np.zeros(29, np.float32)
        self._check(self._lib.ssf_icp_system(self._h, _ptr(R), _ptr(t), n_src, _ptr(out)), "ssf_icp_system")
        return out

    def icpSystemEnqueue(self, R, t, n_src=0, launches=1):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        t = np.ascontiguousarray(t, np.float32).reshape(3)
        self._check(self._lib.ssf_icp_system_enqueue(self._h, _ptr(R), _ptr(t), n_src, launches),
                    "ssf_icp_system_enqueue")

    def icp(self, R_init=None, t_init=None):
        """featureConstrainedSymmetricICP (dense_registration.cu:245-424) on the device."""
        Ri = None if R_init is None else np.ascontiguousarray(R_init, np.float32).reshape(9)
        ti = None if t_init is None else np.ascontiguousarray(t_init, np.float32).reshape(3)
        sys29 = np.zeros(29, np.float32)
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        iters, valid = C.c_int(), C.c_int()
        rc = self._lib.ssf_icp(self._h, _ptr(Ri), _ptr(ti), _ptr(sys29), _ptr(R), _ptr(t), C.byref(iters),
                               C.byref(valid))
        self._check(rc, "ssf_icp")
        return bool(valid.value), R.reshape(3, 3), t, dict(iters=iters.value, valid=valid.value, system=sys29)

    # -- step-wise loop for tile-parallel registration (see multi.py) ---------------------------
    def icpBegin(self, R_init=None, t_init=None):
        Ri = None if R_init is None else np.ascontiguousarray(R_init, np.float32).reshape(9)
        ti = None if t_init is None else np.ascontiguousarray(t_init, np.float32).reshape(3)
        self._check(self._lib.ssf_icp_begin(self._h, _ptr(Ri), _ptr(ti)), "ssf_icp_begin")

    def icpBuild(self, src_begin, src_count):
        out = np.zeros(29, np.float32)
        self._check(self._lib.ssf_icp_build(self._h, C.c_int(src_begin), C.c_int(src_count), _ptr(out)),
                    "ssf_icp_build")
        return out

    def icpSolve(self, sys29):
        sys29 = np.ascontiguousarray(sys29, np.float32).reshape(29)
        done = C.c_int()
        self._check(self._lib.ssf_icp_solve(self._h, _ptr(sys29), C.byref(done)), "ssf_icp_solve")
        return bool(done.value)

    def icpFinish(self, apply_to_pose=False):
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        iters, valid = C.c_int(), C.c_int()
        rc = self._lib.ssf_icp_finish(self._h, C.c_int(int(apply_to_pose)), _ptr(R), _ptr(t), C.byref(iters),
                                      C.byref(valid))
        self._check(rc, "ssf_icp_finish")
        return bool(valid.value), R.reshape(3, 3), t, dict(iters=iters.value, valid=valid.value)

    # -- fused build + exchange + solve over NVLink peer memory ------------------------------
    def peerHandle(self):
        buf = np.zeros(64, np.uint8)
        self._check(self._lib.ssf_peer_handle(self._h, _ptr(buf)), "ssf_peer_handle")
        return buf

    def connectPeers(self, rank, world, handles):
        handles = np.ascontiguousarray(handles, np.uint8).reshape(world * 64)
        self._check(self._lib.ssf_connect_peers(self._h, C.c_int(rank), C.c_int(world), _ptr(handles)),
                    "ssf_connect_peers")

    def icpTiled(self, src_begin, src_count, R_init=None, t_init=None):
        """Collective: every rank calls it with the same R_init / t_init and its own slice."""
        Ri = None if R_init is None else np.ascontiguousarray(R_init, np.float32).reshape(9)
        ti = None if t_init is None else np.ascontiguousarray(t_init, np.float32).reshape(3)
        sys29 = np.zeros(29, np.float32)
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        iters, valid = C.c_int(), C.c_int()
        rc = self._lib.ssf_icp_tiled(self._h, _ptr(Ri), _ptr(ti), C.c_int(src_begin), C.c_int(src_count), _ptr(sys29),
                                     _ptr(R), _ptr(t), C.byref(iters), C.byref(valid))
        self._check(rc, "ssf_icp_tiled")
        return bool(valid.value), R.reshape(3, 3), t, dict(iters=iters.value, valid=valid.value, system=sys29)

    def align(self, source, R_init, t_init):
        """DenseRegistration::align (dense_registration.cu:52-243): registers a keyframe's
        supersurfels (`source`, a Supersurfels in the keyframe's camera frame) against the
        current frame.  Returns (valid, R, t, stats)."""
        Ri = np.ascontiguousarray(R_init, np.float32).reshape(9)
        ti = np.ascontiguousarray(t_init, np.float32).reshape(3)
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        sys29 = np.zeros(29, np.float32)
        valid, iters, pairs = C.c_int(0), C.c_int(0), C.c_int(0)
        view = source.view()
        rc = self._lib.ssf_align(self._h, C.byref(view), len(source.positions), _ptr(Ri), _ptr(ti), _ptr(R), _ptr(t),
                                 C.byref(valid), C.byref(iters), C.byref(pairs), _ptr(sys29))
        self._check(rc, "ssf_align")
        return bool(valid.value), R.reshape(3, 3), t, dict(iters=iters.value, pairs=pairs.value, system=sys29)

    def applyDeformation(self, node_pos, node_rot, node_trans, weights, nn, model_size=None):
        """DeformationGraph::applyGraphToModel (deformation_graph.cu:840-861): warp the model by an
        embedded deformation graph supplied by the caller."""
        node_pos = np.ascontiguousarray(node_pos, np.float32)
        node_rot = np.ascontiguousarray(node_rot, np.float32)
        node_trans = np.ascontiguousarray(node_trans, np.float32)
        weights = np.ascontiguousarray(weights, np.float32)
        nn = np.ascontiguousarray(nn, np.int32)
        n = len(weights) if model_size is None else model_size
        rc = self._lib.ssf_apply_deformation(self._h, _ptr(node_pos), _ptr(node_rot), _ptr(node_trans), len(node_pos),
                                             _ptr(weights), _ptr(nn), n)
        self._check(rc, "ssf_apply_deformation")

    def getMarkers(self, which="model", conf_thresh=None):
        """Triangle-list geometry of publishModelMarker / publishFrameMarker
        (node/supersurfel_fusion_node.cpp:303-520): (points [n,6,3], colors [n,6,4])."""
        w = 0 if which == "model" else 1
        n = self.getCounts()[0] if w == 0 else self.nbSuperpixels
        thr = self.cfg.conf_thresh if conf_thresh is None else conf_thresh
        pts = np.zeros((n, 6, 3), np.float32)
        col = np.zeros((n, 6, 4), np.float32)
        cnt = C.c_int(0)
        rc = self._lib.ssf_get_markers(self._h, w, float(thr), _ptr(pts), _ptr(col), n, C.byref(cnt))
        self._check(rc, "ssf_get_markers")
        return pts, col

    def formatTumPose(self, timestamp):
        """One line of the benchmark node's estimated.txt (…rgbd_benchmark_node.cpp:727-729)."""
        buf = C.create_string_buffer(256)
        self._check(self._lib.ssf_format_tum_pose(self._h, str(timestamp).encode(), buf, 256), "ssf_format_tum_pose")
        return buf.value.decode()

    def fuse(self):
        self._check(self._lib.ssf_fuse(self._h), "ssf_fuse")
        return self.getFrameStats()

    # -- timing ---------------------------------------------------------------------------
    def timerStart(self):
        self._check(self._lib.ssf_timer_start(self._h), "ssf_timer_start")

    def timerStop(self):
        ms = C.c_float()
        self._check(self._lib.ssf_timer_stop(self._h, C.byref(ms)), "ssf_timer_stop")
        return ms.value

    def synchronize(self):
        self._check(self._lib.ssf_synchronize(self._h), "ssf_synchronize")

    def launchCount(self):
        n = C.c_uint64()
        self._check(self._lib.ssf_get_launch_count(self._h, C.byref(n)), "ssf_get_launch_count")
        return n.value
