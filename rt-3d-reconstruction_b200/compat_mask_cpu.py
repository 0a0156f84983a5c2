"""The helper functions that ``2cam/2cams_mask_cpu.py`` defines privately (lines 47-309), on the rt3d kernels.

That script does not import ``vision_pipeline_utils``; it carries its own copies of the geometry
helpers, several of them different from the shared module.  A
maintainer replaces the block of local definitions with::

    from compat_mask_cpu import *          # instead of 2cam/2cams_mask_cpu.py:47-309

Names, positional order and defaults are the script's.  Everything that is identical to the shared
module is re-exported from :mod:`vision_pipeline_utils`; what differs is defined here:

====================================  =========================  ===========================================
function (2cams_mask_cpu.py)          lines                      difference from vision_pipeline_utils
====================================  =========================  ===========================================
``erode_mask_gpu(mask, 12)``          47-50                      default kernel 12 (3 there)
``erode_mask(mask, iterations=1)``    54-57                      cv2.erode, 12 x 12 ones kernel
``downsample_point_cloud(pc, 0.01)``  60-65                      Open3D tensor voxel_down_sample: per-voxel MEAN
``convert_mask_to_3d_points_gpu``     69-86                      same arithmetic, different name
``convert_mask_to_3d_points_cpu``     89-108                     takes a NumPy index array
``calculate_centroid_gpu(pc)``        149-150                    torch.mean(dim=0) of a device tensor
``fuse_point_clouds_centroid``        163-243                    1-vs-1 pairs are stacked WITHOUT the outlier
                                                                 filter; candidates are matched unfiltered
``subtract_point_clouds(ws, obj, r)`` 246-270                    KD-tree radius search: removes d^2 < r^2
                                                                 (strict -- ties stay, the opposite of the
                                                                 cdist variant)
``voxel_grid_subtract``               290-309                    set difference of legacy-voxel centroids
====================================  =========================  ===========================================

There is no CPU fallback: every function needs a CUDA device and the built ``librt3d.so``.
"""
from __future__ import annotations

import numpy as np
import torch

from rt3d import ops
from vision_pipeline_utils import (_Fuser, _to_device, calculate_centroid, crop_point_cloud_gpu,  # noqa: F401
                                   downsample_point_cloud_gpu, filter_outliers_sor, point_clouds_equal,
                                   subtract_point_clouds_gpu, visualize_point_cloud)

__all__ = [
    "erode_mask_gpu", "erode_mask", "downsample_point_cloud", "convert_mask_to_3d_points_gpu",
    "convert_mask_to_3d_points_cpu", "downsample_point_cloud_gpu", "filter_outliers_sor", "point_clouds_equal",
    "visualize_point_cloud", "calculate_centroid", "calculate_centroid_gpu", "crop_point_cloud_gpu",
    "fuse_point_clouds_centroid", "subtract_point_clouds", "subtract_point_clouds_gpu", "voxel_grid_subtract",
]


def erode_mask_gpu(mask, kernel_size=12):
    """2cams_mask_cpu.py:47-50 -- conv2d(ones(k, k), padding=k // 2) > 0 (a box dilation, sic; output is
    (H + 1, W + 1) for an even kernel, like conv2d's)."""
    return ops.box_dilate(mask, kernel_size)


def erode_mask(mask, iterations=1):
    """2cams_mask_cpu.py:54-57 -- cv2.erode with a 12 x 12 ones kernel (anchor (6, 6), pixels beyond the border
    ignored).  NumPy uint8 in, NumPy uint8 out, like the script (which calls it on ``mask.cpu().numpy()``)."""
    host = isinstance(mask, np.ndarray)
    m = _to_device(mask, torch.uint8) if host else mask
    out = ops.box_erode(m, 12, iterations)
    return out.cpu().numpy() if host else out


def downsample_point_cloud(point_cloud, voxel_size=0.01):
    """2cams_mask_cpu.py:60-65 -- Open3D tensor ``voxel_down_sample``: key = floor(p / v), one point per voxel =
    the mean of its points (accumulated in float32, returned in the input dtype).  Open3D's row order is its
    hash-buffer order; here voxels come out sorted by key."""
    pts = np.asarray(point_cloud)
    if pts.shape[0] == 0:
        return pts.reshape(0, 3).copy()
    dt = torch.float32 if pts.dtype == np.float32 else torch.float64
    return ops.voxel_centroid(_to_device(pts, dt), voxel_size).cpu().numpy()


def convert_mask_to_3d_points_gpu(mask_indices, depth_map, cx, cy, fx, fy):
    """2cams_mask_cpu.py:69-86 -- same arithmetic as vision_pipeline_utils.convert_mask_to_3d_points."""
    return ops.backproject_indices(mask_indices, depth_map, cx, cy, fx, fy)


def convert_mask_to_3d_points_cpu(mask_indices, depth_map, cx, cy, fx, fy, device="cuda"):
    """2cams_mask_cpu.py:89-108 -- the index array arrives as NumPy (``np.argwhere``) and the depth map as a host
    or device tensor; both are moved to ``device`` first."""
    idx = torch.as_tensor(np.ascontiguousarray(mask_indices), dtype=torch.int64).to(device)
    depth = depth_map if isinstance(depth_map, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(depth_map))
    return ops.backproject_indices(idx, depth.to(device=device, dtype=torch.float32), cx, cy, fx, fy)


def calculate_centroid_gpu(point_cloud):
    """2cams_mask_cpu.py:149-150 -- the mean row of a device tensor, in the tensor's dtype (sequential row sum,
    like NumPy; torch.mean's own reduction order on CUDA is unspecified and agrees within rounding)."""
    return ops.centroid(point_cloud)


def fuse_point_clouds_centroid(point_clouds_camera1, point_clouds_camera2, distance_threshold=0.1):
    """2cams_mask_cpu.py:163-243 -- the script-local fusion: a class seen once by each camera is simply stacked
    (:198-202, no outlier filter), camera-2 candidates are matched and fused unfiltered (:218-232, so the
    equality removal at :232 works), camera-1 clouds are filtered first (:207)."""
    return _Fuser(distance_threshold, sor_single_pair=False, sor_candidates=False).run(point_clouds_camera1, point_clouds_camera2)


def subtract_point_clouds(workspace_pc, objects_pc, distance_threshold=0.01):
    """2cams_mask_cpu.py:246-270 -- a workspace point is removed iff ``search_radius_vector_3d(point, r)`` finds an
    object point, i.e. d^2 < r^2 strictly (nanoflann's radius search).  The kernel evaluates the predicate on the
    coordinates rounded to float32 (exact for the float32-valued clouds the pipeline produces; for genuinely float64
    inputs a point within ~1e-7 relative of the radius may be decided differently from the reference's float64
    KD-tree -- documented in INTEGRATION.md).  The rows returned are the CALLER's rows (float64, unrounded), selected
    with the kernel's keep mask, as ``select_by_index`` does (:265-268)."""
    ws = np.asarray(workspace_pc, dtype=np.float64).reshape(-1, 3)
    ob = np.asarray(objects_pc)
    if ob.shape[0] == 0 or ws.shape[0] == 0:
        return ws.copy()
    _, keep = ops.subtract(_to_device(ws, torch.float32), _to_device(ob, torch.float32), float(distance_threshold), return_mask=True,
                           strict_radius=True)
    return ws[keep.cpu().numpy().astype(bool)]     # np.asarray(pcd.points) is float64


def voxel_grid_subtract(workspace_pc, objects_pc, voxel_size=0.02):
    """2cams_mask_cpu.py:290-309 (never called by the script) -- legacy ``voxel_down_sample`` of both clouds (grid
    origin = the cloud's min bound - v / 2, per-voxel float64 mean), then the set difference of the centroid
    tuples.  Parity unpinned (Open3D is not installable here); rows come back sorted."""
    def centroids(pc):
        p = np.asarray(pc, dtype=np.float64).reshape(-1, 3)
        if p.shape[0] == 0:
            return p
        origin = p.min(axis=0) - voxel_size * 0.5
        c = ops.voxel_centroid(_to_device(p - origin, torch.float64), voxel_size).cpu().numpy()
        return c + origin
    ws, ob = centroids(workspace_pc), centroids(objects_pc)
    gone = set(map(tuple, ob))
    rows = [r for r in map(tuple, ws) if r not in gone]
    return np.array(sorted(rows), dtype=np.float64).reshape(-1, 3)
