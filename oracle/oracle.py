"""numpy front-end of the C oracle (oracle/rt3d_oracle.c).  TEST INFRASTRUCTURE ONLY.

Every function names the reference lines it restates (paths relative to
/root/reference).  See oracle/__init__.py for who may import this and for the
parity status of each stage.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librt3d_oracle.so")
_lib = None

__all__ = [
    "build", "lib", "threads", "mask_nonzero", "backproject_indices", "transform", "transform_f64",
    "crop", "xyz_ingest", "voxel_snap", "voxel_snap_unique", "voxel_index", "voxel_centroid", "centroid",
    "knn_avg", "knn_avg_kdtree", "sor_mask", "filter_outliers_sor", "subtract_mask", "subtract_mask_grid", "subtract_point_clouds_gpu",
    "subtract_radius_mask", "fuse_point_clouds_centroid", "process_masks_cloud",
]

_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    """Compile oracle/rt3d_oracle.c -> librt3d_oracle.so (gcc, OpenMP, -ffp-contract=off)."""
    src = os.path.join(_HERE, "rt3d_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.rt3d_oracle_mask_nonzero.restype = C.c_int64
        L.rt3d_oracle_backproject_indices.restype = C.c_int64
        L.rt3d_oracle_crop.restype = C.c_int64
        L.rt3d_oracle_xyz_ingest.restype = C.c_int64
        L.rt3d_oracle_voxel_snap_unique.restype = C.c_int64
        L.rt3d_oracle_voxel_centroid.restype = C.c_int64
        L.rt3d_oracle_sor.restype = C.c_int64
        L.rt3d_oracle_sor_threshold.restype = C.c_int64
        L.rt3d_oracle_subtract.restype = C.c_int64
        L.rt3d_oracle_subtract_grid.restype = C.c_int64
        L.rt3d_oracle_subtract_radius.restype = C.c_int64
        _lib = L
    return _lib


def threads() -> int:
    return int(lib().rt3d_oracle_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


# ---------------------------------------------------------------- B0 / B1 ----
def mask_nonzero(mask: np.ndarray) -> np.ndarray:
    """torch.nonzero(mask, as_tuple=False) -- 2cam/vision_pipeline_utils.py:296."""
    mask = np.ascontiguousarray(mask)
    if mask.dtype == np.bool_:
        mask = mask.view(np.uint8)
    if mask.dtype not in (np.uint8, np.float32):
        raise TypeError("mask must be uint8/bool or float32")
    H, W = mask.shape
    dt = 0 if mask.dtype == np.uint8 else 1
    n = lib().rt3d_oracle_mask_nonzero(C.c_void_p(mask.ctypes.data), dt, H, W, None)
    out = np.empty((n, 2), np.int64)
    lib().rt3d_oracle_mask_nonzero(C.c_void_p(mask.ctypes.data), dt, H, W, _p(out, _i64p))
    return out


def backproject_indices(idx, depth, cx, cy, fx, fy, recip=False) -> np.ndarray:
    """convert_mask_to_3d_points -- 2cam/vision_pipeline_utils.py:19-36."""
    idx = np.ascontiguousarray(idx, dtype=np.int64).reshape(-1, 2)
    depth = _f32(depth)
    H, W = depth.shape
    out = np.empty((idx.shape[0], 3), np.float32)
    n = lib().rt3d_oracle_backproject_indices(
        _p(idx, _i64p), C.c_int64(idx.shape[0]), _p(depth, _f32p), H, W,
        C.c_double(cx), C.c_double(cy), C.c_double(fx), C.c_double(fy), int(bool(recip)), _p(out, _f32p))
    return out[:n].copy()


# ---------------------------------------------------------------- T1 / C1 / W0
def transform(P, R, t) -> np.ndarray:
    """torch.mm(P, R.T) + t -- 2cam/vision_pipeline_utils.py:301 (== :239-240)."""
    P = _f32(P).reshape(-1, 3); R = _f32(R); t = _f32(t)
    out = np.empty_like(P)
    lib().rt3d_oracle_transform(_p(P, _f32p), C.c_int64(P.shape[0]), _p(R, _f32p), _p(t, _f32p), _p(out, _f32p))
    return out


def transform_f64(P, R, t) -> np.ndarray:
    """np.dot(R64, P.T).T + t64 -- 2cam/2cams_mask_cpu.py:595,618."""
    P = _f32(P).reshape(-1, 3); R = _f64(R); t = _f64(t)
    out = np.empty(P.shape, np.float64)
    lib().rt3d_oracle_transform_f64(_p(P, _f32p), C.c_int64(P.shape[0]), _p(R, _f64p), _p(t, _f64p), _p(out, _f64p))
    return out


def crop(P, x_bounds, y_bounds, z_bounds) -> np.ndarray:
    """crop_point_cloud_gpu -- 2cam/vision_pipeline_utils.py:78-84."""
    P = _f32(P).reshape(-1, 3)
    box = np.array([x_bounds[0], x_bounds[1], y_bounds[0], y_bounds[1], z_bounds[0], z_bounds[1]], np.float32)
    out = np.empty_like(P)
    n = lib().rt3d_oracle_crop(_p(P, _f32p), C.c_int64(P.shape[0]), _p(box, _f32p), _p(out, _f32p))
    return out[:n].copy()


def xyz_ingest(xyz) -> np.ndarray:
    """[:, :, :3].reshape(-1,3) + isfinite().all(1) -- 2cam/vision_pipeline_utils.py:233-238."""
    xyz = _f32(xyz)
    stride = xyz.shape[-1]
    flat = xyz.reshape(-1, stride)
    out = np.empty((flat.shape[0], 3), np.float32)
    n = lib().rt3d_oracle_xyz_ingest(_p(flat, _f32p), C.c_int64(flat.shape[0]), stride, _p(out, _f32p))
    return out[:n].copy()


# --------------------------------------------------------------------- V1 / V2
def voxel_snap(P, v, recip=False) -> np.ndarray:
    P = _f32(P).reshape(-1, 3)
    out = np.empty_like(P)
    lib().rt3d_oracle_voxel_snap(_p(P, _f32p), C.c_int64(P.shape[0]), C.c_double(v), int(bool(recip)), _p(out, _f32p))
    return out


def voxel_snap_unique(P, v, recip=False) -> np.ndarray:
    """downsample_point_cloud_gpu -- 2cam/vision_pipeline_utils.py:40-44."""
    P = _f32(P).reshape(-1, 3)
    out = np.empty_like(P)
    n = lib().rt3d_oracle_voxel_snap_unique(_p(P, _f32p), C.c_int64(P.shape[0]), C.c_double(v), int(bool(recip)), _p(out, _f32p))
    return out[:n].copy()


def voxel_index(P, v, recip=False) -> np.ndarray:
    """Integer voxel indices rint(p / v) (what the packed device keys hold)."""
    P = _f32(P).reshape(-1, 3)
    out = np.empty(P.shape, np.int32)
    lib().rt3d_oracle_voxel_index(_p(P, _f32p), C.c_int64(P.shape[0]), C.c_double(v), int(bool(recip)), _p(out, _i32p))
    return out


def voxel_centroid(P, voxel_size=0.01):
    """downsample_point_cloud (Open3D tensor voxel_down_sample, mean) -- 2cam/2cams_mask_cpu.py:60-65.

    Returns (keys int64 (U,3) sorted lexicographically, centroids (U,3) in the input
    dtype, counts int32 (U,)).  Open3D's own row order is arbitrary (hash order).
    """
    P = np.ascontiguousarray(P)
    if P.dtype not in (np.float32, np.float64):
        P = P.astype(np.float64)
    P = P.reshape(-1, 3)
    n = P.shape[0]
    keys = np.empty((n, 3), np.int64); xyz = np.empty((n, 3), np.float64); cnt = np.empty(n, np.int32)
    m = lib().rt3d_oracle_voxel_centroid(C.c_void_p(P.ctypes.data), 0 if P.dtype == np.float32 else 1, C.c_int64(n),
                                         C.c_double(voxel_size), _p(keys, _i64p), _p(xyz, _f64p), _p(cnt, _i32p))
    return keys[:m].copy(), xyz[:m].astype(P.dtype), cnt[:m].copy()


def centroid(P) -> np.ndarray:
    """calculate_centroid = np.mean(axis=0) -- 2cam/vision_pipeline_utils.py:74-75."""
    P = np.ascontiguousarray(P).reshape(-1, 3)
    if P.dtype == np.float32:
        out = np.empty(3, np.float32)
        lib().rt3d_oracle_centroid_f32(_p(P, _f32p), C.c_int64(P.shape[0]), _p(out, _f32p))
    else:
        P = _f64(P); out = np.empty(3, np.float64)
        lib().rt3d_oracle_centroid_f64(_p(P, _f64p), C.c_int64(P.shape[0]), _p(out, _f64p))
    return out


# -------------------------------------------------------------------------- O1
def knn_avg(P, k=20, brute=False) -> np.ndarray:
    P = _f64(P).reshape(-1, 3)
    avg = np.empty(P.shape[0], np.float64)
    lib().rt3d_oracle_knn_avg(_p(P, _f64p), C.c_int64(P.shape[0]), int(k), int(bool(brute)), _p(avg, _f64p))
    return avg


def knn_avg_kdtree(P, k=20) -> np.ndarray:
    """mean k-NN distance through the nanoflann-faithful KD-tree (leaf size 15, the pruning rule of searchLevel)"""
    P = _f64(P).reshape(-1, 3)
    avg = np.empty(P.shape[0], np.float64)
    lib().rt3d_oracle_knn_avg_kdtree(_p(P, _f64p), C.c_int64(P.shape[0]), int(k), _p(avg, _f64p))
    return avg


def sor_mask(P, nb_neighbors=20, std_ratio=1.5, brute=False):
    """Open3D remove_statistical_outlier restated; returns (keep uint8 (N,), avg (N,), stats[mu,sigma,thr,valid])."""
    if nb_neighbors < 1 or std_ratio <= 0:
        raise RuntimeError("[RemoveStatisticalOutliers] Illegal input parameters, the number of neighbors "
                           "and standard deviation ratio must be positive.")
    P = _f64(P).reshape(-1, 3)
    n = P.shape[0]
    keep = np.zeros(n, np.uint8); avg = np.empty(n, np.float64); stats = np.zeros(4, np.float64)
    if n:
        lib().rt3d_oracle_sor(_p(P, _f64p), C.c_int64(n), int(nb_neighbors), C.c_double(std_ratio), int(bool(brute)),
                              _p(keep, _u8p), _p(avg, _f64p), _p(stats, _f64p))
    return keep, avg, stats


def filter_outliers_sor(point_cloud, nb_neighbors=20, std_ratio=1.5):
    """filter_outliers_sor -- 2cam/vision_pipeline_utils.py:48-58 (float64 out, order kept)."""
    P = _f64(point_cloud).reshape(-1, 3)
    keep, _, _ = sor_mask(P, nb_neighbors, std_ratio)
    return P[keep.astype(bool)]


# -------------------------------------------------------------------------- S1
def subtract_mask(W, O, thr, band=None):
    """keep mask of subtract_point_clouds_gpu -- 2cam/vision_pipeline_utils.py:175-182, exact predicate.  Large inputs
    without a `band` request go through the grid-accelerated evaluation of the same predicate."""
    W = _f32(W).reshape(-1, 3); O = _f32(O).reshape(-1, 3)
    if band is None and W.shape[0] * O.shape[0] > 2_000_000_000:
        return subtract_mask_grid(W, O, thr)
    keep = np.empty(W.shape[0], np.uint8)
    near = np.empty(W.shape[0], np.uint8) if band is not None else None
    lib().rt3d_oracle_subtract(_p(W, _f32p), C.c_int64(W.shape[0]), _p(O, _f32p), C.c_int64(O.shape[0]), C.c_float(thr),
                               _p(keep, _u8p), C.c_double(band or 0.0), _p(near, _u8p) if near is not None else None)
    return (keep, near) if band is not None else keep


def subtract_mask_grid(W, O, thr):
    W = _f32(W).reshape(-1, 3); O = _f32(O).reshape(-1, 3)
    keep = np.empty(W.shape[0], np.uint8)
    rc = lib().rt3d_oracle_subtract_grid(_p(W, _f32p), C.c_int64(W.shape[0]), _p(O, _f32p), C.c_int64(O.shape[0]), C.c_float(thr),
                                         _p(keep, _u8p))
    if rc < 0:
        raise MemoryError("rt3d_oracle_subtract_grid")
    return keep


def subtract_point_clouds_gpu(workspace_pc, objects_pc, distance_threshold=0.005):
    """subtract_point_clouds_gpu -- 2cam/vision_pipeline_utils.py:168-187 (exact-predicate contract)."""
    O = _f32(objects_pc).reshape(-1, 3)
    if O.shape[0] == 0:
        # the reference's distances.min(dim=1) over a zero-size dim raises IndexError [measured]
        raise IndexError("min(): Expected reduction dim 1 to have non-zero size.")
    W = _f32(workspace_pc).reshape(-1, 3)
    return W[subtract_mask(W, O, distance_threshold).astype(bool)]


def subtract_radius_mask(W, O, r):
    """subtract_point_clouds (KD-tree radius) -- 2cam/2cams_mask_cpu.py:246-270: strict d2 < r*r removes."""
    W = _f64(W).reshape(-1, 3); O = _f64(O).reshape(-1, 3)
    keep = np.empty(W.shape[0], np.uint8)
    lib().rt3d_oracle_subtract_radius(_p(W, _f64p), C.c_int64(W.shape[0]), _p(O, _f64p), C.c_int64(O.shape[0]),
                                      C.c_double(r), _p(keep, _u8p))
    return keep


# -------------------------------------------------------------------------- F1
def fuse_point_clouds_centroid(point_clouds_camera1, point_clouds_camera2, distance_threshold=0.1,
                               sor=filter_outliers_sor, sor_single_pair=True, sor_candidates=True):
    """Control flow of fuse_point_clouds_centroid -- 2cam/vision_pipeline_utils.py:88-165.

    ``sor_single_pair=False, sor_candidates=False`` gives the script-local variant F2
    (2cam/2cams_mask_cpu.py:163-243).  Quirks kept on purpose: class ids iterate in set
    order (:117); a matched cam-2 cloud is removed by array equality with its *filtered*
    copy (:155) and so reappears when SOR dropped a point (:162-163); the returned
    pcs1/pcs2 are those of the last class visited (:165).
    """
    by_class = [{}, {}]
    for cam, clouds in enumerate((point_clouds_camera1, point_clouds_camera2)):
        for pc, cid in clouds:
            by_class[cam].setdefault(cid, []).append(pc)
    pcs1, pcs2, fused = [], [], []
    for cid in set(by_class[0].keys()).union(by_class[1].keys()):
        pcs1 = by_class[0].get(cid, [])
        pcs2 = by_class[1].get(cid, [])
        if len(pcs1) == 1 and len(pcs2) == 1:
            pair = np.vstack((pcs1[0], pcs2[0]))
            fused.append((sor(pair) if sor_single_pair else pair, cid))
            continue
        for pc1 in pcs1:
            pc1 = sor(pc1)
            c1 = np.mean(pc1, axis=0)
            best, match = float("inf"), None
            for pc2 in pcs2:
                d = np.linalg.norm(c1 - np.mean(pc2, axis=0))
                if d < best and d < distance_threshold:
                    best, match = d, (sor(pc2) if sor_candidates else pc2)
            if match is None:
                fused.append((pc1, cid))
            else:
                fused.append((np.vstack((pc1, match)), cid))
                pcs2 = [pc for pc in pcs2 if not np.array_equal(pc, match)]
        fused.extend((pc2, cid) for pc2 in pcs2)
    return pcs1, pcs2, fused


# -------------------------------------------------------------------------- P1
def process_masks_cloud(mask, depth, cx, cy, fx, fy, R, t, voxel=0.005, recip=False):
    """One object of process_masks -- 2cam/vision_pipeline_utils.py:296-303: B0->B1->T1->V1."""
    pts = backproject_indices(mask_nonzero(mask), depth, cx, cy, fx, fy, recip)
    if pts.shape[0] == 0:
        return None
    return voxel_snap_unique(transform(pts, R, t), voxel, recip)
