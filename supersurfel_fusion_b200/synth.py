"""Deterministic synthetic RGB-D sequences shaped like TUM fr1/desk.

Scene: a desk top, three walls, two boxes and a tilted board, each a textured rectangle with
piecewise-constant Voronoi colour patches (about 30-60 px across at VGA) plus a low-amplitude
gradient.  Camera: pinhole 525/525/319.5/239.5 scaled with the resolution, moving on a smooth
Lissajous path (about 1 cm and 0.5 degrees per frame, never zero motion).  Depth: rendered
analytically in metres, Gaussian noise sigma(z) = 0.0012 + 0.0019 (z - 0.4)^2, quantised to
1/5000 m like TUM, 10-25 % dropout (random pixels, blobs, depth edges) written as 0.

Everything is a pure function of (seed, frame index, resolution), numpy only.
"""
import numpy as np


def _hash2(ix, iy, salt):
    """Integer lattice hash -> uint32 (vectorised)."""
    h = (ix.astype(np.int64) * 73856093) ^ (iy.astype(np.int64) * 19349663) ^ (int(salt) * 83492791)
    h = h & 0xFFFFFFFF
    h = (h ^ (h >> 16)) * 0x45D9F3B & 0xFFFFFFFF
    h = (h ^ (h >> 16)) * 0x45D9F3B & 0xFFFFFFFF
    h = h ^ (h >> 16)
    return h.astype(np.uint32)


def _rot(axis, ang):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    c, s = np.cos(ang), np.sin(ang)
    x, y, z = axis
    K = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return np.eye(3) * c + s * K + (1 - c) * np.outer(axis, axis)


class Rect:
    """Textured rectangle: origin p0, orthonormal in-plane axes e1, e2, extents l1, l2."""

    def __init__(self, p0, e1, e2, l1, l2, salt, patch=0.13):
        self.p0 = np.asarray(p0, np.float64)
        self.e1 = np.asarray(e1, np.float64) / np.linalg.norm(e1)
        self.e2 = np.asarray(e2, np.float64) / np.linalg.norm(e2)
        self.n = np.cross(self.e1, self.e2)
        self.l1, self.l2, self.salt, self.patch = l1, l2, salt, patch


def _box(x0, x1, y0, y1, z0, z1, salt):
    # faces that can face a camera near the origin: top (y = y0), front (z = z0), both sides
    return [
        Rect((x0, y0, z0), (1, 0, 0), (0, 0, 1), x1 - x0, z1 - z0, salt),
        Rect((x0, y0, z0), (1, 0, 0), (0, 1, 0), x1 - x0, y1 - y0, salt + 1),
        Rect((x0, y0, z0), (0, 0, 1), (0, 1, 0), z1 - z0, y1 - y0, salt + 2),
        Rect((x1, y0, z0), (0, 0, 1), (0, 1, 0), z1 - z0, y1 - y0, salt + 3),
    ]


def default_scene(seed):
    s = int(seed) * 101
    rects = [
        Rect((-3.0, -2.2, 2.6), (1, 0, 0), (0, 1, 0), 6.0, 2.9, s + 1),        # back wall
        Rect((-1.7, 0.55, 0.2), (1, 0, 0), (0, 0, 1), 3.6, 2.4, s + 2),        # desk top
        Rect((-1.7, -2.2, 0.0), (0, 0, 1), (0, 1, 0), 2.6, 2.9, s + 3),        # left wall
        Rect((1.9, -2.2, 0.0), (0, 0, 1), (0, 1, 0), 2.6, 2.9, s + 4),         # right wall
    ]
    rects += _box(-0.65, -0.2, 0.27, 0.55, 1.25, 1.65, s + 10)
    rects += _box(0.3, 0.85, 0.12, 0.55, 1.55, 1.95, s + 20)
    Rb = _rot((0, 1, 0), 0.45) @ _rot((1, 0, 0), -0.2)
    rects.append(Rect(np.array([-0.35, -0.55, 2.2]), Rb[:, 0], Rb[:, 1], 0.9, 0.6, s + 30))  # tilted board
    return rects


class SyntheticSequence:
    def __init__(self, width=640, height=480, seed=1234, n_frames=300, noise=True, dropout=True,
                 motion_scale=1.0):
        self.W, self.H, self.seed, self.n_frames = width, height, seed, n_frames
        sx = width / 640.0
        sy = height / 480.0
        self.fx, self.fy = 525.0 * sx, 525.0 * sy
        self.cx, self.cy = (319.5 + 0.5) * sx - 0.5, (239.5 + 0.5) * sy - 0.5
        self.noise, self.dropout = noise, dropout
        self.motion_scale = motion_scale
        self.rects = default_scene(seed)
        u = (np.arange(width, dtype=np.float64) - self.cx) / self.fx
        v = (np.arange(height, dtype=np.float64) - self.cy) / self.fy
        uu, vv = np.meshgrid(u, v)
        self.rays_cam = np.stack([uu, vv, np.ones_like(uu)], -1)   # z = 1 => ray parameter = depth
        ph = np.random.RandomState(seed).uniform(0, 2 * np.pi, 6)
        self._ph = ph

    def cam_param(self):
        return (self.fx, self.fy, self.cx, self.cy, self.H, self.W)

    def world_pose(self, k):
        """Camera-to-world (R, t) of frame k on the Lissajous path."""
        m, ph = self.motion_scale, self._ph
        w = 0.05
        t = m * np.array([0.16 * np.sin(w * k + ph[0]), 0.07 * np.sin(1.3 * w * k + ph[1]),
                          0.12 * np.sin(0.7 * w * k + ph[2])])
        yaw = m * np.deg2rad(7.0) * np.sin(0.9 * w * k + ph[3])
        pitch = m * np.deg2rad(4.0) * np.sin(1.1 * w * k + ph[4]) + np.deg2rad(6.0)
        roll = m * np.deg2rad(2.0) * np.sin(0.6 * w * k + ph[5])
        R = _rot((0, 1, 0), yaw) @ _rot((1, 0, 0), pitch) @ _rot((0, 0, 1), roll)
        return R, t

    def pose(self, k):
        """Ground-truth pose of frame k relative to frame 0 (the engine's world frame)."""
        R0, t0 = self.world_pose(0)
        Rk, tk = self.world_pose(k)
        R = R0.T @ Rk
        t = R0.T @ (tk - t0)
        return R.astype(np.float32), t.astype(np.float32)

    def _texture(self, rect, a, b):
        c = rect.patch
        ga, gb = a / c, b / c
        ia, ib = np.floor(ga).astype(np.int64), np.floor(gb).astype(np.int64)
        best = np.full(a.shape, 1e30)
        best_h = np.zeros(a.shape, np.uint32)
        for da in (-1, 0, 1):
            for db in (-1, 0, 1):
                ja, jb = ia + da, ib + db
                h = _hash2(ja, jb, rect.salt)
                sa = ja + ((h & 0xFFFF).astype(np.float64) / 65535.0)
                sb = jb + (((h >> 16) & 0xFFFF).astype(np.float64) / 65535.0)
                d2 = (ga - sa) ** 2 + (gb - sb) ** 2
                upd = d2 < best
                best = np.where(upd, d2, best)
                best_h = np.where(upd, h, best_h)
        h2 = _hash2(best_h.astype(np.int64), (best_h >> 7).astype(np.int64), rect.salt + 7)
        base = np.stack([20 + (h2 & 0xFF) % 216, 20 + ((h2 >> 8) & 0xFF) % 216, 20 + ((h2 >> 16) & 0xFF) % 216], -1)
        grad = 6.0 * np.sin(3.0 * a + 2.0 * b)
        return np.clip(base.astype(np.float64) + grad[..., None], 0, 255)

    def frame(self, k):
        """-> rgb uint8 (H,W,3) in R,G,B order, depth float32 (H,W) metres (0 = missing)."""
        R, t = self.world_pose(k)
        rays = self.rays_cam @ R.T
        depth = np.full((self.H, self.W), np.inf)
        rgb = np.zeros((self.H, self.W, 3))
        for rect in self.rects:
            denom = rays @ rect.n
            num = float((rect.p0 - t) @ rect.n)
            with np.errstate(divide="ignore", invalid="ignore"):
                s = num / denom
            hit = np.isfinite(s) & (s > 0.05) & (s < depth)
            if not hit.any():
                continue
            P = t + rays * s[..., None]
            rel = P - rect.p0
            a, b = rel @ rect.e1, rel @ rect.e2
            hit &= (a >= 0) & (a <= rect.l1) & (b >= 0) & (b <= rect.l2)
            if not hit.any():
                continue
            col = self._texture(rect, a[hit], b[hit])
            depth[hit] = s[hit]
            rgb[hit] = col
        valid = np.isfinite(depth)
        z = np.where(valid, depth, 0.0)
        rs = np.random.RandomState((self.seed * 7919 + k * 104729) % (2 ** 31 - 1))
        if self.noise:
            sigma = 0.0012 + 0.0019 * (z - 0.4) ** 2
            z = z + rs.standard_normal(z.shape) * sigma * valid
        z = np.round(z * 5000.0) / 5000.0
        if self.dropout:
            drop = rs.uniform(size=z.shape) < 0.06
            # blobs
            yy, xx = np.mgrid[0:self.H, 0:self.W]
            for _ in range(10):
                cx, cy = rs.uniform(0, self.W), rs.uniform(0, self.H)
                r = rs.uniform(8, 34) * self.W / 640.0
                drop |= (xx - cx) ** 2 + (yy - cy) ** 2 < r * r
            # depth edges
            gx = np.abs(np.diff(depth, axis=1, prepend=depth[:, :1]))
            gy = np.abs(np.diff(depth, axis=0, prepend=depth[:1, :]))
            with np.errstate(invalid="ignore"):
                edge = (gx > 0.08) | (gy > 0.08)
            edge = edge | np.roll(edge, 1, 0) | np.roll(edge, -1, 0) | np.roll(edge, 1, 1) | np.roll(edge, -1, 1)
            drop |= edge
            z = np.where(drop, 0.0, z)
        z = np.where(valid, z, 0.0)
        return np.ascontiguousarray(rgb.round().astype(np.uint8)), np.ascontiguousarray(z.astype(np.float32))


def synthetic_icp_problem(n_src, width=2560, height=1920, cell=16, seed=1234, inlier_frac=0.6):
    """Roofline-sizing input for the ICP system kernel: n_src visible model
    supersurfels uniform in the frustum, about `inlier_frac` of them passing every gate,
    against a synthetic planar-patch frame of the given size.  Returns a dict of numpy
    arrays in the reference member layout plus the frame-side maps."""
    rs = np.random.RandomState(seed)
    s = width / 640.0
    fx = fy = 525.0 * s
    cx, cy = (319.5 + 0.5) * s - 0.5, (239.5 + 0.5) * height / 480.0 - 0.5
    gx, gy = (width + cell - 1) // cell, (height + cell - 1) // cell
    S = gx * gy
    # frame: one fronto-parallel-ish planar patch per grid cell
    yy, xx = np.mgrid[0:height, 0:width]
    labels = ((yy // cell) * gx + (xx // cell)).astype(np.int32)
    cell_depth = rs.uniform(0.8, 3.5, S).astype(np.float32)
    depth = cell_depth[labels]
    tgt_col = rs.uniform(20, 235, (S, 3)).astype(np.float32)
    nrm = np.stack([rs.uniform(-0.15, 0.15, S), rs.uniform(-0.15, 0.15, S), -np.ones(S)], -1)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    tgt_ori = np.zeros((S, 9), np.float32)
    tgt_ori[:, 6:9] = nrm
    tgt_conf = np.where(rs.uniform(size=S) < 0.95, 200.0, -1.0).astype(np.float32)
    # sources: sample a pixel, back-project at the frame depth, perturb
    u = rs.randint(0, width, n_src)
    v = rs.randint(0, height, n_src)
    lab = labels[v, u]
    z = depth[v, u]
    # ~20 % of the good ones land in a neighbouring (different-depth) cell once the view
    # transform shifts the projection, and 5 % of the frame supersurfels are invalid
    good = rs.uniform(size=n_src) < inlier_frac / (0.95 * 0.8)
    dz = np.where(good, rs.uniform(-0.01, 0.01, n_src), rs.uniform(0.15, 0.6, n_src)).astype(np.float32)
    zz = z + dz
    pos = np.stack([(u - cx) / fx * zz, (v - cy) / fy * zz, zz], -1).astype(np.float32)
    src_col = (tgt_col[lab] + rs.uniform(-3, 3, (n_src, 3))).astype(np.float32)
    src_ori = np.zeros((n_src, 9), np.float32)
    sn = nrm[lab] + rs.uniform(-0.05, 0.05, (n_src, 3))
    sn /= np.linalg.norm(sn, axis=1, keepdims=True)
    src_ori[:, 6:9] = sn
    return dict(cam=(fx, fy, cx, cy, height, width), S=S, labels=labels, depth=depth.astype(np.float32),
                tgt_col=tgt_col, tgt_ori=tgt_ori, tgt_conf=tgt_conf, src_pos=pos, src_col=src_col,
                src_ori=src_ori)
