"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the reference-kernel harness
(oracle/_ref/libssf_ref.so: the reference's own TPS_RGBD / DenseRegistration classes and
surfel kernels compiled unmodified for sm_100a, see oracle/ref_harness.cu).

Used by tests/ (to pin the CPU oracle and the CUDA product against the reference itself)
and by bench.py (to time "the reference's own kernels on one GPU of the same box").
The library is built in the development container (where /root/reference exists) and
travels to the GPU box as a prebuilt file; this module never reads /root/reference.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libssf_ref.so")


def available():
    return os.path.exists(LIB_PATH)


class RefParams(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("height", C.c_int), ("width", C.c_int), ("cell_size", C.c_int),
                ("lambda_pos", C.c_float), ("lambda_bound", C.c_float), ("lambda_size", C.c_float),
                ("lambda_disp", C.c_float), ("thresh_disp", C.c_float),
                ("seg_iter", C.c_int), ("seg_use_ransac", C.c_int), ("nb_samples", C.c_int), ("filter_iter", C.c_int),
                ("filter_alpha", C.c_float), ("filter_beta", C.c_float), ("filter_threshold", C.c_float),
                ("range_min", C.c_float), ("range_max", C.c_float), ("delta_t", C.c_int), ("conf_thresh", C.c_float),
                ("nb_supersurfels_max", C.c_int), ("icp_iter", C.c_int), ("icp_cov_thresh", C.c_double)]


class RefSurfelsHost(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("colors", C.c_void_p), ("stamps", C.c_void_p),
                ("orientations", C.c_void_p), ("shapes", C.c_void_p), ("dims", C.c_void_p),
                ("confidences", C.c_void_p)]


class RefStats(C.Structure):
    _fields_ = [("stamp", C.c_int), ("nb_supersurfels", C.c_int), ("nb_visible", C.c_int), ("nb_removed", C.c_int),
                ("icp_ran", C.c_int), ("icp_valid", C.c_int), ("ms_tps", C.c_float), ("ms_generate", C.c_float),
                ("ms_icp", C.c_float), ("ms_fuse", C.c_float), ("ms_total", C.c_float), ("wall_ms", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libssf_ref.so is not built (make -C oracle ref, needs /root/reference)")
        L = C.CDLL(LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_nb_superpixels.restype = C.c_int
        L.ref_icp.restype = C.c_int
        L.ref_align.restype = C.c_int
        L.ref_align.argtypes = None
        L.ref_icp_system_time.restype = C.c_float
        for name in ("ref_destroy", "ref_nb_superpixels", "ref_tps", "ref_get_segmentation", "ref_set_segmentation",
                     "ref_generate", "ref_get_frame", "ref_set_frame", "ref_get_model", "ref_set_model",
                     "ref_set_pose", "ref_get_pose", "ref_get_counts", "ref_icp_system", "ref_icp_system_time",
                     "ref_icp", "ref_fuse", "ref_process_frame"):
            getattr(L, name).argtypes = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _view(s):
    return RefSurfelsHost(_p(s.positions), _p(s.colors), _p(s.stamps), _p(s.orientations), _p(s.shapes), _p(s.dims),
                          _p(s.confidences))


class RefEngine:
    """The reference's hot path, driven through the harness."""

    def __init__(self, cam, surfels_cls, **params):
        self._cls = surfels_cls
        p = RefParams()
        p.fx, p.fy, p.cx, p.cy, p.height, p.width = cam
        for k, v in params.items():
            setattr(p, k, int(v) if isinstance(v, bool) else v)
        self.p = p
        self.h = C.c_void_p(lib().ref_create(C.byref(p)))
        self.S = lib().ref_nb_superpixels(self.h)
        self.H, self.W = p.height, p.width

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tps(self, rgb, depth):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        lib().ref_tps(self.h, _p(rgb), _p(depth))
        return self.segmentation()

    def segmentation(self):
        out = dict(labels=np.zeros((self.H, self.W), np.int32), bound=np.zeros((self.H, self.W), np.int32),
                   inliers=np.zeros((self.H, self.W), np.uint8), superpixels=np.zeros((self.S, 12), np.float32),
                   slanted=np.zeros((self.H, self.W), np.float32))
        lib().ref_get_segmentation(self.h, _p(out["labels"]), _p(out["bound"]), _p(out["inliers"]),
                                   _p(out["superpixels"]), _p(out["slanted"]))
        return out

    def set_segmentation(self, labels=None, bound=None, inliers=None, slanted=None, rgba=None):
        c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
        keep = [c(labels, np.int32), c(bound, np.int32), c(inliers, np.uint8), c(slanted, np.float32), c(rgba, np.uint8)]
        lib().ref_set_segmentation(self.h, *[_p(k) for k in keep])

    def generate(self, stamp):
        lib().ref_generate(self.h, C.c_int(stamp))
        return self.frame()

    def frame(self):
        f = self._cls(self.S)
        v = _view(f)
        lib().ref_get_frame(self.h, C.byref(v))
        return f

    def set_frame(self, f):
        v = _view(f)
        lib().ref_set_frame(self.h, C.byref(v))

    def model(self, n=None):
        n = self.counts()[0] if n is None else n
        m = self._cls(n)
        v = _view(m)
        lib().ref_get_model(self.h, C.byref(v), C.c_int(n))
        return m

    def set_model(self, m, n, n_visible):
        v = _view(m)
        lib().ref_set_model(self.h, C.byref(v), C.c_int(n), C.c_int(n_visible))

    def set_pose(self, R, t):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        t = np.ascontiguousarray(t, np.float32).reshape(3)
        lib().ref_set_pose(self.h, _p(R), _p(t))

    def pose(self):
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        lib().ref_get_pose(self.h, _p(R), _p(t))
        return R.reshape(3, 3), t

    def counts(self):
        c = np.zeros(4, np.int32)
        lib().ref_get_counts(self.h, _p(c))
        return tuple(int(x) for x in c)

    def icp_system(self, R, t, n):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        t = np.ascontiguousarray(t, np.float32).reshape(3)
        out = np.zeros(29, np.float32)
        lib().ref_icp_system(self.h, _p(R), _p(t), C.c_int(n), _p(out))
        return out

    def icp_system_time(self, R, t, n, launches):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        t = np.ascontiguousarray(t, np.float32).reshape(3)
        return float(lib().ref_icp_system_time(self.h, _p(R), _p(t), C.c_int(n), C.c_int(launches)))

    def icp(self, Rview, tview):
        Rv = np.ascontiguousarray(Rview, np.float32).reshape(9)
        tv = np.ascontiguousarray(tview, np.float32).reshape(3)
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        ok = lib().ref_icp(self.h, _p(Rv), _p(tv), _p(R), _p(t))
        return bool(ok), R.reshape(3, 3), t

    def align(self, source, R_init, t_init):
        """DenseRegistration::align: keyframe supersurfels (host) against the harness' current frame."""
        Ri = np.ascontiguousarray(R_init, np.float32).reshape(9)
        ti = np.ascontiguousarray(t_init, np.float32).reshape(3)
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        v = _view(source)
        ok = lib().ref_align(self.h, C.byref(v), C.c_int(len(source.positions)), _p(Ri), _p(ti), _p(R), _p(t))
        return bool(ok), R.reshape(3, 3), t

    def fuse(self, stamp):
        lib().ref_fuse(self.h, C.c_int(stamp))
        return self.counts()

    def process_frame(self, rgb, depth, prior=None):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        pr = None
        if prior is not None:
            pr = np.concatenate([np.asarray(prior[0], np.float32).reshape(9), np.asarray(prior[1], np.float32).reshape(3)])
        st = RefStats()
        lib().ref_process_frame(self.h, _p(rgb), _p(depth), _p(pr), C.byref(st))
        return {k: getattr(st, k) for k, _ in RefStats._fields_}
