"""Drop-in for the reference's ``2cam/vision_pipeline_utils.py``, backed by sm_100a CUDA kernels.

Put this directory on ``sys.path`` ahead of the reference's ``2cam/`` and ``2cams.py``
(``from vision_pipeline_utils import *``, 2cam/2cams.py:1) and ``2cams_mask_gpu.py`` (named
imports, 2cam/2cams_mask_gpu.py:8) pick it up unchanged: every public name, positional
signature and default of the reference module is here, including the module globals the scripts
use unqualified through the star import (np, cv2, time, torch, F, csv, sl, o3d, YOLO).
``compat_mask_cpu.py`` adds the extra names that ``2cams_mask_cpu.py`` defines privately.

All geometry runs in librt3d.so (see include/rt3d.h); host code here is only the reference's
orchestration: list/dict bookkeeping, timing, host<->device staging and the greedy
centroid matcher.  There is no CPU fallback -- without a CUDA device and the built library the
compute functions raise.
"""
from __future__ import annotations

import csv
import time

import numpy as np
import torch
import torch.nn.functional as F

try:  # display only (process_iteration_end); the geometric path never needs it
    import cv2
except ImportError:  # pragma: no cover
    cv2 = None
try:  # camera SDK: absent in this image; a shim with the same enum names keeps the wrappers callable
    import pyzed.sl as sl
except ImportError:
    from rt3d.fakes import sl_shim as sl
try:
    import open3d as o3d  # only visualize_point_cloud needs it
except ImportError:
    o3d = None
try:
    from ultralytics import YOLO
except ImportError:
    YOLO = None

from rt3d import ops

_VOXEL = 0.005                                            # literal at :246-247, :302
_WORKSPACE_BOX = ((-0.25, 0.75), (-0.5, 1.75), (-0.05, 2))  # literals at :241-243


def _dev(device):
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError(f"rt3d runs on CUDA devices only (got {device!r}); there is no CPU fallback")
    return d


def _to_device(array, dtype, device="cuda"):
    return torch.as_tensor(np.ascontiguousarray(array), dtype=dtype).to(_dev(device), non_blocking=False)


# ----------------------------------------------------------------------------------------------------
# geometry primitives (reference: 2cam/vision_pipeline_utils.py:12-187)
# ----------------------------------------------------------------------------------------------------
def erode_mask_gpu(mask, kernel_size=3):
    """:12-15 -- despite its name the reference's conv2d(ones) > 0 test *grows* the mask by
    kernel_size // 2 pixels; kept as is (unused by the reference's main loops)."""
    return ops.box_dilate(mask, kernel_size)


def convert_mask_to_3d_points(mask_indices, depth_map, cx, cy, fx, fy):
    """:19-36 -- (K,2) int64 [row, col] indices + depth map -> (N,3) fp32 camera-frame points."""
    return ops.backproject_indices(mask_indices, depth_map, cx, cy, fx, fy)


def downsample_point_cloud_gpu(point_cloud, voxel_size):
    """:40-44 -- round(p / v) * v, distinct rows in lexicographic order."""
    return ops.voxel_snap_unique(point_cloud, voxel_size)


def filter_outliers_sor(point_cloud, nb_neighbors=20, std_ratio=1.5):
    """:48-58 -- Open3D remove_statistical_outlier; numpy in, float64 numpy out, order kept."""
    pts = np.asarray(point_cloud)
    if nb_neighbors < 1 or std_ratio <= 0:
        raise RuntimeError("[RemoveStatisticalOutliers] Illegal input parameters, the number of neighbors and "
                           "standard deviation ratio must be positive.")
    if pts.shape[0] == 0:
        return np.empty((0, 3), np.float64)
    dt = torch.float32 if pts.dtype == np.float32 else torch.float64
    return ops.sor(_to_device(pts, dt), nb_neighbors, std_ratio).cpu().numpy()


def point_clouds_equal(pc1, pc2):
    """:62-63"""
    return np.array_equal(pc1, pc2)


def visualize_point_cloud(point_cloud, title="Point Cloud"):
    """:67-71 -- GUI helper; needs Open3D, which is outside the geometric path."""
    if o3d is None:
        raise RuntimeError("visualize_point_cloud needs open3d, which is not installed")
    cloud = o3d.geometry.PointCloud()
    cloud.points = o3d.utility.Vector3dVector(point_cloud)
    o3d.visualization.draw_geometries([cloud], window_name=title)


def calculate_centroid(point_cloud):
    """:74-75 -- np.mean(axis=0), same summation order and dtype, computed on the device."""
    pts = np.asarray(point_cloud)
    dt = torch.float32 if pts.dtype == np.float32 else torch.float64
    return ops.centroid(_to_device(pts, dt)).cpu().numpy()


def crop_point_cloud_gpu(point_cloud, x_bounds, y_bounds, z_bounds):
    """:78-84 -- closed-interval box test, order preserved."""
    return ops.crop(point_cloud, x_bounds, y_bounds, z_bounds)


class _Fuser:
    """Greedy per-class centroid matching of :88-165, with the reference's observable quirks:
    classes are visited in ``set`` order (:117); a matched camera-2 cloud is dropped by array
    equality with its *filtered* copy (:155), so it is emitted again when the filter removed a
    point (:162-163); the lists handed back are those of the last class visited (:165)."""

    def __init__(self, threshold, sor_single_pair=True, sor_candidates=True):
        self.threshold = threshold
        self.sor_single_pair = sor_single_pair
        self.sor_candidates = sor_candidates

    @staticmethod
    def _group(clouds):
        groups = {}
        for cloud, class_id in clouds:
            groups.setdefault(class_id, []).append(cloud)
        return groups

    def _match(self, anchor, candidates):
        """Nearest camera-2 centroid under the threshold; returns the (filtered) partner or None."""
        c1 = calculate_centroid(anchor)
        best, partner = float("inf"), None
        for cand in candidates:
            gap = np.linalg.norm(c1 - calculate_centroid(cand))
            if gap < best and gap < self.threshold:
                best = gap
                partner = filter_outliers_sor(cand) if self.sor_candidates else cand
        return partner

    def run(self, cam1, cam2):
        g1, g2 = self._group(cam1), self._group(cam2)
        fused, left, right = [], [], []
        for class_id in set(g1.keys()).union(g2.keys()):
            left, right = g1.get(class_id, []), g2.get(class_id, [])
            if len(left) == 1 and len(right) == 1:
                pair = np.vstack((left[0], right[0]))
                fused.append((filter_outliers_sor(pair) if self.sor_single_pair else pair, class_id))
                continue
            for cloud in left:
                cloud = filter_outliers_sor(cloud)
                partner = self._match(cloud, right)
                if partner is None:
                    fused.append((cloud, class_id))
                    continue
                fused.append((np.vstack((cloud, partner)), class_id))
                right = [other for other in right if not point_clouds_equal(other, partner)]
            fused.extend((other, class_id) for other in right)
        return left, right, fused


def fuse_point_clouds_centroid(point_clouds_camera1, point_clouds_camera2, distance_threshold=0.1):
    """:88-165 -> (pcs1, pcs2, [(cloud, class_id), ...])."""
    return _Fuser(distance_threshold).run(point_clouds_camera1, point_clouds_camera2)


def subtract_point_clouds_gpu(workspace_pc, objects_pc, distance_threshold=0.005):
    """:168-187 -- workspace rows farther than the threshold from every object point (numpy in/out, fp32).

    The reference builds the dense M x N cdist matrix; the kernel evaluates the same predicate exactly
    without it.  An empty object cloud raises IndexError like the reference's ``min(dim=1)``."""
    objects = np.asarray(objects_pc)
    if objects.shape[0] == 0:
        raise IndexError("min(): Expected reduction dim 1 to have non-zero size.")
    ws = _to_device(workspace_pc, torch.float32)
    ob = _to_device(objects, torch.float32)
    return ops.subtract(ws, ob, distance_threshold).cpu().numpy()


# ----------------------------------------------------------------------------------------------------
# stage wrappers (reference: :190-381): same signatures, timing keys and side effects
# ----------------------------------------------------------------------------------------------------
def _stamp(timings, key, t0, label=None):
    dt = time.time() - t0
    timings[key].append(dt)
    if label:
        print(f"{label}: {dt:.4f} seconds")
    return dt


def retrieve_frames(zed1, zed2, image1, image2, timings):
    """:190-216 -- left images of both cameras, BGRA -> BGR."""
    t0 = time.time()
    frames = []
    for zed, image in ((zed1, image1), (zed2, image2)):
        zed.retrieve_image(image, view=sl.VIEW.LEFT)
        bgra = image.get_data()
        frames.append(cv2.cvtColor(bgra, cv2.COLOR_BGRA2BGR) if cv2 is not None else np.ascontiguousarray(bgra[..., :3]))
    _stamp(timings, "Frame Retrieval", t0, "Frame retrieval time")
    return frames[0], frames[1]


def retrieve_depth_maps(zed1, zed2, depth1, depth2, timings):
    """:218-227"""
    t0 = time.time()
    status1 = zed1.retrieve_measure(depth1, measure=sl.MEASURE.DEPTH)
    status2 = zed2.retrieve_measure(depth2, measure=sl.MEASURE.DEPTH)
    maps = depth1.get_data(), depth2.get_data()
    _stamp(timings, "Depth Retrieval", t0, "Depth retrieval time")
    return status1, status2, maps[0], maps[1]


def _workspace_cloud(zed, mat, resolution, rotation, origin, device):
    """One camera of :231-247: XYZ image -> finite rows -> robot frame -> crop -> 5 mm voxel set."""
    zed.retrieve_measure(mat, measure=sl.MEASURE.XYZ, resolution=resolution)
    xyz = _to_device(mat.get_data(), torch.float32, device)
    cropped = ops.xyz_select(xyz, finite_filter=True, R=rotation, t=origin, box=_WORKSPACE_BOX)
    return ops.voxel_snap_unique(cropped, _VOXEL)


def process_point_clouds(zed1, zed2, point_cloud1_ws, point_cloud2_ws, resolution, rotation1_torch, origin1_torch,
                         rotation2_torch, origin2_torch, device, timings):
    """:229-254 -- both cameras' workspace voxel sets, concatenated (not re-uniqued, :248), as numpy."""
    t0 = time.time()
    ws1 = _workspace_cloud(zed1, point_cloud1_ws, resolution, rotation1_torch, origin1_torch, device)
    ws2 = _workspace_cloud(zed2, point_cloud2_ws, resolution, rotation2_torch, origin2_torch, device)
    fused_point_cloud_ws = torch.cat((ws1, ws2), dim=0).cpu().numpy()
    print(f"Fused Point Cloud Workspace shape: {fused_point_cloud_ws.shape}")
    _stamp(timings, "Point Cloud Processing", t0, "Point cloud processing time")
    return fused_point_cloud_ws


def perform_yolo_inference(model, frame1, frame2, device, timings):
    """:256-290 -- detector/tracker glue (out of scope for the kernels; kept so the star import is complete)."""
    t0 = time.time()
    common = dict(classes=[39, 41], persist=True, retina_masks=True, conf=0.1, device=device,
                  tracker="ultralytics/cfg/trackers/bytetrack.yaml")
    res1 = model.track(source=frame1, **common)
    res2 = model.track(source=frame2, imgsz=640, **common)
    out = (res1[0].plot(line_width=2, font_size=18), res2[0].plot(line_width=2, font_size=18), res1[0].masks, res2[0].masks,
           res1[0].boxes.cls.cpu().numpy(), res2[0].boxes.cls.cpu().numpy())
    _stamp(timings, "YOLO11 Inference", t0, "YOLO11 processing time")
    return out


def process_masks(masks, class_ids, zed_depth_np, cx, cy, fx, fy, rotation_robot_cam_torch, origin_cam_torch, device,
                  point_clouds_camera):
    """:292-304 -- per instance mask: nonzero -> back-project -> robot frame -> 5 mm voxel set -> host.
    One fused kernel does the first three steps straight from the mask plane."""
    if masks is None:
        return
    depth = _to_device(zed_depth_np, torch.float32, device)
    intr = (cx, cy, fx, fy)
    for i, mask in enumerate(masks.data):
        pts = ops.backproject_mask(depth, mask, intr, R=rotation_robot_cam_torch, t=origin_cam_torch)
        if pts.shape[0] > 0:
            cloud = ops.voxel_snap_unique(pts, _VOXEL)
            point_clouds_camera.append((cloud.cpu().numpy(), int(class_ids[i])))


def fuse_point_clouds(point_clouds_camera1, point_clouds_camera2, distance_threshold, timings):
    """:307-320"""
    t0 = time.time()
    pcs1, pcs2, fused = fuse_point_clouds_centroid(point_clouds_camera1, point_clouds_camera2, distance_threshold)
    _stamp(timings, "Point Cloud Fusion", t0)
    clouds = [cloud for cloud, _ in fused]
    merged = np.vstack(clouds) if clouds else np.empty((0, 3))
    print(f"Fused Point Cloud Objects Concatenated shape: {merged.shape}")
    return pcs1, pcs2, merged


def subtract_point_clouds(fused_point_cloud_ws, fused_pc_objects_concatenated, distance_threshold, timings):
    """:322-327"""
    t0 = time.time()
    remaining = subtract_point_clouds_gpu(fused_point_cloud_ws, fused_pc_objects_concatenated, distance_threshold)
    _stamp(timings, "Subtraction", t0)
    return remaining


def process_iteration_end(start_time, timings, frame_count, fps_values, fps_log_file, annotated_frame1, annotated_frame2,
                          point_clouds_camera1, point_clouds_camera2, pcs1, pcs2, timings_csv="../timings.csv",
                          window="YOLO11 Segmentation+Tracking"):
    """:329-381 -- loop tail: FPS moving average (10 frames), fps_log.csv row, timings.csv rewrite,
    optional display, and clearing of the per-frame lists."""
    elapsed = time.time() - start_time
    timings["Total Time per Iteration"].append(elapsed)
    print("Total time per loop iteration: ", elapsed)
    frame_count += 1
    fps_values.append(1.0 / elapsed)
    del fps_values[:-10]
    mean_fps = sum(fps_values) / len(fps_values)

    with open(fps_log_file, mode="a", newline="") as fh:
        csv.writer(fh).writerow([time.time(), fps_values[-1]])
    with open(timings_csv, "w", newline="") as fh:
        out = csv.writer(fh)
        out.writerow(["Step", "Timings"])
        for step, values in timings.items():
            out.writerow([step, ",".join(str(v) for v in values)])

    shown = [f for f in (annotated_frame1, annotated_frame2) if f is not None]
    if cv2 is not None and shown:
        for frame in shown:
            cv2.putText(frame, f"FPS: {mean_fps:.2f}", (10, 30), cv2.FONT_HERSHEY_SIMPLEX, 1, (0, 255, 0), 2)
        if len(shown) == 2:
            both = cv2.hconcat(shown)
            both = cv2.resize(both, (both.shape[1] // 2, both.shape[0] // 2))
            try:
                cv2.imshow(window, both)
            except cv2.error:  # headless build
                pass

    for per_frame in (point_clouds_camera1, point_clouds_camera2, pcs1, pcs2):
        per_frame.clear()
    return frame_count, fps_values
