"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")


def build(force=False):
    # make decides staleness (sources newer than the library); a box without the toolchain
    # but with a prebuilt library (the GPU box always has both) just uses it
    try:
        subprocess.check_call(["make", "-s", "-C", HERE, "-j8"] + (["-B"] if force else []))
    except (OSError, subprocess.CalledProcessError):
        if not os.path.exists(LIB_PATH):
            raise
    return LIB_PATH


class OrcCam(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("height", C.c_int), ("width", C.c_int)]


class OrcSurfels(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("colors", C.c_void_p), ("stamps", C.c_void_p),
                ("orientations", C.c_void_p), ("shapes", C.c_void_p), ("dims", C.c_void_p),
                ("confidences", C.c_void_p)]


class OrcConfig(C.Structure):
    _fields_ = [("cam", OrcCam), ("cell_size", C.c_int),
                ("lambda_pos", C.c_float), ("lambda_bound", C.c_float), ("lambda_size", C.c_float),
                ("lambda_disp", C.c_float), ("thresh_disp", C.c_float),
                ("seg_iter", C.c_int), ("seg_use_ransac", C.c_int), ("nb_samples", C.c_int),
                ("filter_iter", C.c_int),
                ("filter_alpha", C.c_float), ("filter_beta", C.c_float), ("filter_threshold", C.c_float),
                ("range_min", C.c_float), ("range_max", C.c_float), ("delta_t", C.c_int),
                ("conf_thresh", C.c_float), ("nb_supersurfels_max", C.c_int), ("icp_iter", C.c_int),
                ("icp_cov_thresh", C.c_double)]


class OrcIcpStats(C.Structure):
    _fields_ = [("valid", C.c_int), ("iters", C.c_int), ("inliers", C.c_float), ("error", C.c_double),
                ("last_system", C.c_float * 29)]


class OrcFuseCounts(C.Structure):
    _fields_ = [("nb_supersurfels", C.c_int), ("nb_visible", C.c_int), ("nb_removed", C.c_int),
                ("nb_matched", C.c_int), ("nb_inserted", C.c_int), ("nb_removed_stale", C.c_int),
                ("nb_removed_invalid", C.c_int), ("nb_removed_occluded", C.c_int)]


class OrcFrameStats(C.Structure):
    _fields_ = [("stamp", C.c_int), ("nb_supersurfels", C.c_int), ("nb_visible", C.c_int),
                ("nb_removed", C.c_int), ("icp_ran", C.c_int), ("icp_valid", C.c_int),
                ("icp_iters", C.c_int), ("icp_inliers", C.c_float), ("icp_error", C.c_double),
                ("nb_matched", C.c_int), ("nb_inserted", C.c_int), ("nb_removed_stale", C.c_int),
                ("nb_removed_invalid", C.c_int), ("nb_removed_occluded", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_tps_create.restype = C.c_void_p
        L.orc_engine_create.restype = C.c_void_p
        L.orc_engine_tps.restype = C.c_void_p
        L.orc_tps_nb_superpixels.restype = C.c_int
        L.orc_icp.restype = C.c_int
        L.orc_rng_create.restype = C.c_void_p
        L.orc_rng_u32.restype = C.c_uint
        L.orc_rng_uniform.restype = C.c_float
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def default_config(width=640, height=480, **kw):
    cfg = OrcConfig()
    lib().orc_config_default(C.byref(cfg))
    s = width / 640.0
    cfg.cam = OrcCam(525.0 * s, 525.0 * s, (319.5 + 0.5) * s - 0.5, (239.5 + 0.5) * height / 480.0 - 0.5,
                     height, width)
    for k, v in kw.items():
        if k == "cam":
            cfg.cam = OrcCam(*v)
        else:
            setattr(cfg, k, v)
    return cfg


def cam_of(cfg_or_cam):
    return cfg_or_cam.cam if hasattr(cfg_or_cam, "cam") else cfg_or_cam


class Surfels:
    """Host SoA in the reference's Supersurfels member layout."""
    FIELDS = (("positions", np.float32, 3), ("colors", np.float32, 3), ("stamps", np.int32, 2),
              ("orientations", np.float32, 9), ("shapes", np.float32, 6), ("dims", np.float32, 2),
              ("confidences", np.float32, 1))

    def __init__(self, n):
        self.n = n
        for name, dt, w in self.FIELDS:
            setattr(self, name, np.zeros((n, w) if w > 1 else (n,), dtype=dt))

    def struct(self):
        return OrcSurfels(*[_p(getattr(self, name)) for name, _, _ in self.FIELDS])

    def head(self, n):
        out = Surfels(0)
        out.n = n
        for name, _, _ in self.FIELDS:
            setattr(out, name, getattr(self, name)[:n].copy())
        return out

    def as_dict(self):
        return {name: getattr(self, name) for name, _, _ in self.FIELDS}


def icp_system(cam, src_pos, src_col, src_orient, tgt_col, tgt_orient, tgt_conf, R, t, labels, depth):
    out = np.zeros(29, np.float32)
    src_pos, src_col, src_orient = _f32(src_pos), _f32(src_col), _f32(src_orient)
    tgt_col, tgt_orient, tgt_conf = _f32(tgt_col), _f32(tgt_orient), _f32(tgt_conf)
    R, t = _f32(R).reshape(9), _f32(t).reshape(3)
    labels = np.ascontiguousarray(labels, np.int32)
    depth = _f32(depth)
    lib().orc_icp_system(C.c_int(len(src_pos)), _p(src_pos), _p(src_col), _p(src_orient), _p(tgt_col),
                         _p(tgt_orient), _p(tgt_conf), _p(R), _p(t), C.byref(cam), _p(labels), _p(depth),
                         _p(out))
    return out


def icp(cam, src_pos, src_col, src_orient, tgt_col, tgt_orient, tgt_conf, R_init, t_init, labels, depth,
        nb_iter=10, cov_thresh=0.04):
    src_pos, src_col, src_orient = _f32(src_pos), _f32(src_col), _f32(src_orient)
    tgt_col, tgt_orient, tgt_conf = _f32(tgt_col), _f32(tgt_orient), _f32(tgt_conf)
    R_init, t_init = _f32(R_init).reshape(9), _f32(t_init).reshape(3)
    labels = np.ascontiguousarray(labels, np.int32)
    depth = _f32(depth)
    R = np.zeros(9, np.float32)
    t = np.zeros(3, np.float32)
    st = OrcIcpStats()
    ok = lib().orc_icp(C.c_int(len(src_pos)), _p(src_pos), _p(src_col), _p(src_orient), _p(tgt_col),
                       _p(tgt_orient), _p(tgt_conf), _p(R_init), _p(t_init), C.byref(cam), _p(labels),
                       _p(depth), C.c_int(nb_iter), C.c_double(cov_thresh), _p(R), _p(t), C.byref(st))
    return bool(ok), R.reshape(3, 3), t, dict(valid=st.valid, iters=st.iters, inliers=st.inliers,
                                               error=st.error, system=np.array(st.last_system, np.float32))


def align(cam, src_pos, src_col, src_orient, src_conf, tgt_col, tgt_orient, tgt_conf, R_init, t_init, labels, depth,
          nb_iter=10, cov_thresh=0.04):
    """DenseRegistration::align (dense_registration.cu:52-243): keyframe -> current frame."""
    src_pos, src_col, src_orient, src_conf = _f32(src_pos), _f32(src_col), _f32(src_orient), _f32(src_conf)
    tgt_col, tgt_orient, tgt_conf = _f32(tgt_col), _f32(tgt_orient), _f32(tgt_conf)
    R_init, t_init = _f32(R_init).reshape(9), _f32(t_init).reshape(3)
    labels = np.ascontiguousarray(labels, np.int32)
    depth = _f32(depth)
    R = np.zeros(9, np.float32)
    t = np.zeros(3, np.float32)
    st = OrcIcpStats()
    lib().orc_align.restype = C.c_int
    ok = lib().orc_align(C.c_int(len(src_pos)), _p(src_pos), _p(src_col), _p(src_orient), _p(src_conf), _p(tgt_col),
                         _p(tgt_orient), _p(tgt_conf), _p(R_init), _p(t_init), C.byref(cam), _p(labels),
                         _p(depth), C.c_int(nb_iter), C.c_double(cov_thresh), _p(R), _p(t), C.byref(st))
    return bool(ok), R.reshape(3, 3), t, dict(valid=st.valid, iters=st.iters, pairs=st.inliers,
                                               error=st.error, system=np.array(st.last_system, np.float32))


def compose_pose(R, t, R_rel, t_rel):
    R = _f32(R).reshape(9).copy()
    t = _f32(t).reshape(3).copy()
    lib().orc_compose_pose(_p(R), _p(t), _p(_f32(R_rel).reshape(9)), _p(_f32(t_rel).reshape(3)))
    return R.reshape(3, 3), t


def generate_supersurfels(cam, n_superpixels, rgba, slanted, labels, inliers, bound, z_min, z_max, stamp):
    fr = Surfels(n_superpixels)
    st = fr.struct()
    rgba = np.ascontiguousarray(rgba, np.uint8)
    slanted = _f32(slanted)
    labels = np.ascontiguousarray(labels, np.int32)
    inliers = np.ascontiguousarray(inliers, np.uint8)
    bound = np.ascontiguousarray(bound, np.int32)
    lib().orc_generate_supersurfels(C.byref(cam), C.c_int(n_superpixels), _p(rgba), _p(slanted), _p(labels),
                                    _p(inliers), _p(bound), C.c_float(z_min), C.c_float(z_max),
                                    C.c_int(stamp), C.byref(st))
    return fr


def fuse(cam, frame, model, nb_max, R, t, labels, slanted, z_min, z_max, stamp, delta_t, conf_thresh,
         nb_supersurfels, nb_visible):
    """In-place on `model` (a Surfels with capacity nb_max). Returns the counts dict."""
    fc = OrcFuseCounts(nb_supersurfels, nb_visible, 0, 0, 0)
    fs, ms = frame.struct(), model.struct()
    R, t = _f32(R).reshape(9), _f32(t).reshape(3)
    labels = np.ascontiguousarray(labels, np.int32)
    slanted = _f32(slanted)
    lib().orc_fuse(C.byref(cam), C.c_int(frame.n), C.byref(fs), C.byref(ms), C.c_int(nb_max), _p(R), _p(t),
                   _p(labels), _p(slanted), C.c_float(z_min), C.c_float(z_max), C.c_int(stamp),
                   C.c_int(delta_t), C.c_float(conf_thresh), C.byref(fc))
    return dict(nb_supersurfels=fc.nb_supersurfels, nb_visible=fc.nb_visible, nb_removed=fc.nb_removed,
                nb_matched=fc.nb_matched, nb_inserted=fc.nb_inserted, nb_removed_stale=fc.nb_removed_stale,
                nb_removed_invalid=fc.nb_removed_invalid, nb_removed_occluded=fc.nb_removed_occluded)


def extract_local_point_cloud(surfels, conf_thresh, R_pose, t_pose, radius):
    """extractLocalPointCloud (supersurfel_fusion.cu:884-920) -> (positions, normals) in camera coordinates."""
    n = surfels.n
    pos = np.zeros((max(n, 1), 3), np.float32)
    nrm = np.zeros((max(n, 1), 3), np.float32)
    lib().orc_extract_local_point_cloud.restype = C.c_int
    cnt = lib().orc_extract_local_point_cloud(C.c_int(n), _p(surfels.positions), _p(surfels.orientations),
                                              _p(surfels.confidences), C.c_float(conf_thresh),
                                              _p(_f32(np.asarray(R_pose).reshape(9))), _p(_f32(np.asarray(t_pose).reshape(3))),
                                              C.c_float(radius), _p(pos), _p(nrm))
    return pos[:cnt], nrm[:cnt]


def transform_model(surfels, R, t):
    """applyTransformSuperSurfel (supersurfel_fusion_kernels.cu:467-488), in place."""
    lib().orc_transform_model(C.c_int(surfels.n), _p(surfels.positions), _p(surfels.orientations), _p(surfels.shapes),
                              _p(surfels.confidences), _p(_f32(np.asarray(R).reshape(9))), _p(_f32(np.asarray(t).reshape(3))))


class Tps:
    def __init__(self, cfg, handle=None):
        self.cfg = cfg
        self._own = handle is None
        self.h = C.c_void_p(lib().orc_tps_create(C.byref(cfg))) if handle is None else C.c_void_p(handle)
        self.S = lib().orc_tps_nb_superpixels(self.h)
        self.W, self.H = cfg.cam.width, cfg.cam.height

    def compute(self, rgb, depth):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = _f32(depth)
        assert rgb.shape == (self.H, self.W, 3) and depth.shape == (self.H, self.W)
        lib().orc_tps_compute(self.h, _p(rgb), _p(depth))
        return self.get()

    def get(self):
        H, W, S = self.H, self.W, self.S
        out = dict(labels=np.zeros((H, W), np.int32), bound=np.zeros((H, W), np.int32),
                   inliers=np.zeros((H, W), np.uint8), disp=np.zeros((H, W), np.float32),
                   superpixels=np.zeros((S, 12), np.float32), slanted=np.zeros((H, W), np.float32),
                   rgba=np.zeros((H, W, 4), np.uint8))
        lib().orc_tps_get(self.h, _p(out["labels"]), _p(out["bound"]), _p(out["inliers"]), _p(out["disp"]),
                          _p(out["superpixels"]), _p(out["slanted"]), _p(out["rgba"]))
        return out

    def samples(self):
        s = np.zeros((self.S, self.cfg.nb_samples, 4), np.float32)
        lib().orc_tps_get_samples(self.h, _p(s))
        return s

    def __del__(self):
        if getattr(self, "_own", False) and self.h:
            lib().orc_tps_destroy(self.h)
            self.h = None


class Engine:
    def __init__(self, cfg):
        self.cfg = cfg
        self.h = C.c_void_p(lib().orc_engine_create(C.byref(cfg)))
        self.tps = Tps(cfg, handle=lib().orc_engine_tps(self.h))
        self.S = self.tps.S
        self.last = None

    def process_frame(self, rgb, depth, prior=None, mask=None):
        """mask: optional uint8[S], the MOD hook (dynamic frame supersurfels -> confidence -1)."""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = _f32(depth)
        pr = None
        if prior is not None:
            pr = _f32(np.concatenate([np.asarray(prior[0]).reshape(9), np.asarray(prior[1]).reshape(3)]))
        mk = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        st = OrcFrameStats()
        lib().orc_engine_process_frame_masked(self.h, _p(rgb), _p(depth), _p(pr), _p(mk), C.byref(st))
        self.last = {k: getattr(st, k) for k, _ in OrcFrameStats._fields_}
        return self.last

    def pose(self):
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        lib().orc_engine_get_pose(self.h, _p(R), _p(t))
        return R.reshape(3, 3), t

    def set_pose(self, R, t):
        lib().orc_engine_set_pose(self.h, _p(_f32(np.asarray(R).reshape(9))), _p(_f32(np.asarray(t).reshape(3))))

    def transform_model(self, R, t):
        lib().orc_engine_transform_model(self.h, _p(_f32(np.asarray(R).reshape(9))), _p(_f32(np.asarray(t).reshape(3))))

    def local_cloud(self, radius):
        n = max(self.last["nb_supersurfels"] if self.last else 0, 1)
        pos = np.zeros((n, 3), np.float32)
        nrm = np.zeros((n, 3), np.float32)
        lib().orc_engine_local_cloud.restype = C.c_int
        cnt = lib().orc_engine_local_cloud(self.h, C.c_float(radius), _p(pos), _p(nrm))
        return pos[:cnt], nrm[:cnt]

    def model(self):
        n = self.last["nb_supersurfels"] if self.last else 0
        m = Surfels(n)
        st = m.struct()
        lib().orc_engine_get_model(self.h, C.byref(st))
        return m

    def frame(self):
        f = Surfels(self.S)
        st = f.struct()
        lib().orc_engine_get_frame(self.h, C.byref(st))
        return f

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_engine_destroy(self.h)
            self.h = None


def bilateral_filter(depth, kernel_size=-1, sigma_color=0.03, sigma_spatial=4.5):
    """cv::cuda::bilateralFilter(depth, depth, -1, 0.03, 4.5) restated out of place (oracle_ingest.cpp)."""
    depth = _f32(depth)
    out = np.empty_like(depth)
    h, w = depth.shape
    lib().orc_bilateral_filter(_p(depth), C.c_int(w), C.c_int(h), C.c_int(kernel_size), C.c_float(sigma_color),
                               C.c_float(sigma_spatial), _p(out))
    return out


def rgb_to_gray(rgb):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    out = np.empty(rgb.shape[:2], np.uint8)
    lib().orc_rgb_to_gray(_p(rgb), C.c_int(out.size), _p(out))
    return out


def depth16_to_metres(depth16, scale):
    depth16 = np.ascontiguousarray(depth16, np.uint16)
    out = np.empty(depth16.shape, np.float32)
    lib().orc_depth16_to_metres(_p(depth16), C.c_int(out.size), C.c_float(scale), _p(out))
    return out


def apply_deformation(surfels, node_pos, node_rot, node_trans, weights, nn):
    """applyDeformation (deformation_graph_kernels.cu:27-73), in place on a Surfels."""
    node_pos, node_rot, node_trans, weights = _f32(node_pos), _f32(node_rot), _f32(node_trans), _f32(weights)
    nn = np.ascontiguousarray(nn, np.int32)
    lib().orc_apply_deformation(_p(surfels.positions), _p(surfels.orientations), _p(surfels.shapes), _p(node_pos),
                                _p(node_rot), _p(node_trans), _p(weights), _p(nn), C.c_int(len(surfels.positions)))


def markers(surfels, conf_thresh):
    """publishModelMarker geometry (node/supersurfel_fusion_node.cpp:303-413): (points [n,6,3], colors [n,6,4])."""
    n = len(surfels.positions)
    pts = np.zeros((n, 6, 3), np.float32)
    col = np.zeros((n, 6, 4), np.float32)
    lib().orc_markers(_p(surfels.positions), _p(surfels.colors), _p(surfels.orientations), _p(surfels.dims),
                      _p(surfels.confidences), C.c_int(n), C.c_float(conf_thresh), _p(pts), _p(col))
    return pts, col


def format_tum_pose(R, t, timestamp):
    buf = C.create_string_buffer(256)
    lib().orc_format_tum_pose(_p(_f32(R).reshape(9)), _p(_f32(t).reshape(3)), timestamp.encode(), buf, C.c_int(256))
    return buf.value.decode()


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(n))
