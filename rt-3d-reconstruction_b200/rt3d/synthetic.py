"""Seeded synthetic depth/mask frames for tests and bench (the ZED SDK and YOLO are out of scope).

A small analytic scene in the robot base frame -- a gently wavy table at z ~ 0 and a few
ellipsoid "bottles/cups" standing on it -- is ray-cast through a pin-hole camera placed with
the reference's own calibration constants (2cam/2cams.py:100-114, data not code), so that
two cameras see the *same* objects and the centroid fusion / workspace subtraction stages
have something meaningful to do.  Depth is Z-depth in metres (what ZED ``MEASURE.DEPTH``
returns), with NaN / +inf / -inf / 0 holes as the ZED produces them.

Intrinsics are not in the reference repo (read from the camera at run time,
2cam/2cams.py:90-96); SURVEY.md 8(d) fixes fx=527.8*s, fy=527.6*s, cx=643.2*s, cy=361.7*s,
s = W/1280.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# --- calibration constants of the reference rig (2cam/2cams.py:100-114) ------------------
T_CHESS_CAM = {
    1: np.array([[0.8401, 0.3007, -0.4515, 0.5914],
                 [-0.5424, 0.4647, -0.6999, 1.1329],
                 [-0.0006, 0.8329, 0.5535, -0.7193],
                 [0.0, 0.0, 0.0, 1.0]]),
    2: np.array([[0.4608, -0.4105, 0.7869, -0.8716],
                 [0.8874, 0.2010, -0.4148, 0.8318],
                 [0.0121, 0.8895, 0.4569, -0.7200],
                 [0.0, 0.0, 0.0, 1.0]]),
}
T_ROBOT_CHESS = np.array([[-1.0, 0.0, 0.0, 0.3580],
                          [0.0, 1.0, 0.0, 0.0300],
                          [0.0, 0.0, -1.0, 0.0060],
                          [0.0, 0.0, 0.0, 1.0]])

# crop box of process_point_clouds (2cam/vision_pipeline_utils.py:241-243)
WORKSPACE_BOX = ((-0.25, 0.75), (-0.5, 1.75), (-0.05, 2.0))
WORKSPACE_CENTRE = np.array([0.36, 0.25, 0.0])

RESOLUTIONS = {"HD720": (720, 1280), "HD1080": (1080, 1920), "HD2K": (1242, 2208), "VGA360": (360, 640),
               "QVGA": (180, 320)}


def extrinsics(cam: int, rig: int = 2):
    """(R (3,3) f64, t (3,) f64) camera -> robot base.  cam 1/2 are the reference's two ZEDs.

    For ``rig=8`` (config 4) camera k in 0..7 is the reference cam-1 pose rotated by k*45 deg
    about the vertical axis through the workspace centre.
    """
    if rig == 2:
        T = T_ROBOT_CHESS @ T_CHESS_CAM[cam]
        return T[:3, :3].copy(), T[:3, 3].copy()
    T = T_ROBOT_CHESS @ T_CHESS_CAM[1]
    a = np.deg2rad(45.0 * cam)
    Rz = np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
    R = Rz @ T[:3, :3]
    t = Rz @ (T[:3, 3] - WORKSPACE_CENTRE) + WORKSPACE_CENTRE
    return R, t


def intrinsics(W: int):
    """(cx, cy, fx, fy) as Python floats, scaled from the HD720 values."""
    s = W / 1280.0
    return 643.2 * s, 361.7 * s, 527.8 * s, 527.6 * s


@dataclass
class Ellipsoid:
    centre: np.ndarray  # robot frame, metres
    radii: np.ndarray
    class_id: int


@dataclass
class Scene:
    objects: list = field(default_factory=list)
    table_amp: float = 0.008

    def table_height(self, x, y):
        return self.table_amp * np.sin(7.0 * x) * np.cos(5.0 * y)


def make_scene(seed: int, n_objects: int = 1, jitter: float = 0.0, overlap: int = 0) -> Scene:
    """n ellipsoids on a jittered grid over the table; class ids alternate 39 (bottle) / 41 (cup).

    The last ``overlap`` of them are not given a grid position of their own: object n - overlap + k intersects
    object k (same class, centre 4-6 cm away), so that two clouds of one class have nearly the same centroid and
    the greedy matcher of fuse_point_clouds_centroid has a real choice to make (BASELINE.json configs[2]:
    "10 non-overlapping + 2 deliberately overlapping ellipses")."""
    rng = np.random.default_rng(seed)
    sc = Scene()
    base_n = n_objects - overlap
    cols = int(np.ceil(np.sqrt(max(base_n, 1))))
    for i in range(base_n):
        gx, gy = i % cols, i // cols
        base = np.array([0.16 + 0.40 * (gx + 0.5) / cols, -0.05 + 0.60 * (gy + 0.5) / cols])
        if base_n == 1:
            base = np.array([0.36, 0.25])
        xy = base + rng.uniform(-0.02, 0.02, 2) + jitter * rng.uniform(-1, 1, 2)
        rad = np.array([rng.uniform(0.05, 0.07), rng.uniform(0.05, 0.07), rng.uniform(0.10, 0.15)])
        if cols > 3:     # keep a dozen objects apart on the 0.4 x 0.6 m table
            rad[:2] *= 3.0 / cols
        sc.objects.append(Ellipsoid(np.array([xy[0], xy[1], rad[2]]), rad, 39 if i % 2 == 0 else 41))
    for k in range(overlap):
        host = sc.objects[k % max(base_n, 1)]
        off = np.array([rng.uniform(0.04, 0.06), rng.uniform(-0.02, 0.02), 0.0])
        rad = host.radii * rng.uniform(0.8, 1.0, 3)
        sc.objects.append(Ellipsoid(np.array([host.centre[0] + off[0], host.centre[1] + off[1], rad[2]]), rad, host.class_id))
    return sc


def render(scene: Scene, R, t, H: int, W: int, seed: int, holes: bool = True, noise_mm: float = 1.0,
           mask_dilate: int = 2, flying: float = 0.004, max_depth: float = 6.0):
    """Ray-cast ``scene`` -> (depth (H,W) f32, masks (n,H,W) u8, class_ids (n,) int).

    ``mask_dilate`` grows each instance mask by that many pixels (segmentation masks are
    never tight -- the reference erodes them for that reason, 2cam/2cams_mask_cpu.py:54-57);
    ``flying`` is the fraction of mask pixels whose depth is pushed 5-40 cm off the surface
    (mixed-pixel outliers for the statistical outlier filter to find).
    """
    rng = np.random.default_rng(seed)
    cx, cy, fx, fy = intrinsics(W)
    u = (np.arange(W, dtype=np.float64) - cx) / fx
    v = (np.arange(H, dtype=np.float64) - cy) / fy
    dc = np.stack(np.broadcast_arrays(u[None, :], v[:, None], np.ones((1, 1))), axis=-1)  # (H,W,3), z = 1
    dr = dc @ np.asarray(R, np.float64).T                                                   # robot-frame ray per unit depth
    o = np.asarray(t, np.float64)

    # table: z = h(x, y); two fixed-point refinements from the flat plane
    with np.errstate(divide="ignore", invalid="ignore"):
        d = -o[2] / dr[..., 2]
        for _ in range(2):
            hx = o[0] + d * dr[..., 0]
            hy = o[1] + d * dr[..., 1]
            d = (scene.table_height(hx, hy) - o[2]) / dr[..., 2]
    d = np.where((d > 0.2) & (d < max_depth), d, np.inf)

    masks = np.zeros((len(scene.objects), H, W), np.uint8)
    for k, ob in enumerate(scene.objects):
        # |(o + d*dr - c) / r|^2 = 1
        q = (o - ob.centre) / ob.radii
        w = dr / ob.radii
        a = (w * w).sum(-1); b = 2.0 * (w * q).sum(-1); c = (q * q).sum() - 1.0
        disc = b * b - 4 * a * c
        with np.errstate(invalid="ignore"):
            root = (-b - np.sqrt(disc)) / (2 * a)
        hit = (disc > 0) & (root > 0.2) & (root < d)
        d = np.where(hit, root, d)
        m = hit
        for _ in range(mask_dilate):
            g = m.copy()
            g[1:, :] |= m[:-1, :]; g[:-1, :] |= m[1:, :]; g[:, 1:] |= m[:, :-1]; g[:, :-1] |= m[:, 1:]
            m = g
        masks[k] = m

    depth = d + rng.normal(0.0, noise_mm * 1e-3, d.shape)
    if flying > 0 and len(scene.objects):
        anym = masks.any(0)
        fl = anym & (rng.random(d.shape) < flying)
        depth = np.where(fl, depth + rng.uniform(0.05, 0.4, d.shape) * rng.choice([-1.0, 1.0], d.shape), depth)
    depth = depth.astype(np.float32)
    if holes:
        r = rng.random(d.shape)
        depth[r < 0.010] = np.nan
        depth[(r >= 0.010) & (r < 0.015)] = np.inf
        depth[(r >= 0.015) & (r < 0.020)] = -np.inf
        depth[(r >= 0.020) & (r < 0.022)] = 0.0
    class_ids = np.array([ob.class_id for ob in scene.objects], dtype=np.float32)
    return depth, masks, class_ids


def xyz_image(depth: np.ndarray):
    """Camera-frame XYZ image (H,W,4) fp32 like ZED ``MEASURE.XYZ`` (4th channel is padding)."""
    H, W = depth.shape
    cx, cy, fx, fy = intrinsics(W)
    u = np.arange(W, dtype=np.float32)[None, :]
    v = np.arange(H, dtype=np.float32)[:, None]
    out = np.full((H, W, 4), np.nan, np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        bad = ~(np.isfinite(depth) & (depth > 0))
        out[..., 0] = (u - np.float32(cx)) * depth / np.float32(fx)
        out[..., 1] = (v - np.float32(cy)) * depth / np.float32(fy)
        out[..., 2] = depth
    out[bad, :3] = np.nan
    return out


def frame(config: str = "HD720", seed: int = 10, n_objects: int = 1, cams=(1, 2), rig: int = 2, jitter: float = 0.0,
          scene_seed=None, overlap: int = 0, shuffle_cams: bool = False, **kw):
    """One multi-camera frame: dict(depth [C,H,W] f32, masks [C,n,H,W] u8, class_ids [C,n], R [C,3,3] f32,
    t [C,3] f32, intr (cx,cy,fx,fy)).  Camera c uses render seed ``seed + index``."""
    H, W = RESOLUTIONS[config] if isinstance(config, str) else config
    sc = make_scene(seed if scene_seed is None else scene_seed, n_objects, jitter, overlap)
    depth, masks, cls, Rs, ts = [], [], [], [], []
    for i, cam in enumerate(cams):
        R, t = extrinsics(cam, rig)
        d, m, c = render(sc, R, t, H, W, seed + i, **kw)
        if shuffle_cams and len(c) > 1:
            perm = np.random.default_rng(7919 * seed + i).permutation(len(c))
            m, c = m[perm], c[perm]
        depth.append(d); masks.append(m); cls.append(c)
        Rs.append(R.astype(np.float32)); ts.append(t.astype(np.float32))
    return dict(depth=np.stack(depth), masks=np.stack(masks), class_ids=np.stack(cls), R=np.stack(Rs), t=np.stack(ts),
                intr=intrinsics(W), scene=sc, H=H, W=W)
