"""The reference's op sequence in torch eager -- MEASUREMENT / TEST INFRASTRUCTURE ONLY (bench.py --impl torch_eager).

SURVEY.md 8d(iii) asks for "the reference as deployed": its own torch functions on the same GPU.  The reference
module cannot travel to the GPU box (/root/reference exists in the build container only), so this file restates,
call for call, which torch ops one iteration of 2cam/2cams.py:165-208 issues for the geometric path -- the same
library kernels (nonzero, advanced indexing, mm, round, unique(dim=0), cdist, min), the same host round trips
(.cpu().numpy() after every stage, torch.tensor(numpy, device=) before the next), frame by frame and object by
object -- so that its launch + sync cost can be timed next to rt3d_frame_batch_run.  tests/test_oracle_vs_reference.py
checks (in the build container, on CPU tensors) that every stage returns exactly what the imported reference returns.
The outlier filter goes through the CPU port (oracle.filter_outliers_sor): Open3D's remove_statistical_outlier is a
host routine in the reference as well (2cam/vision_pipeline_utils.py:48-58).

Stage names follow the reference's `timings` keys (2cam/2cams.py:24-33).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import oracle as _o

BOX = ((-0.25, 0.75), (-0.5, 1.75), (-0.05, 2))          # 2cam/vision_pipeline_utils.py:241-243


def back_project(idx, depth, cx, cy, fx, fy):
    """2cam/vision_pipeline_utils.py:19-36"""
    u, v = idx[:, 1], idx[:, 0]
    d = depth[v, u]
    ok = (d > 0) & ~torch.isnan(d) & ~torch.isinf(d)
    u, v, d = u[ok], v[ok], d[ok]
    return torch.stack(((u - cx) * d / fx, (v - cy) * d / fy, d), dim=-1)


def voxel_unique(pc, voxel):
    """2cam/vision_pipeline_utils.py:40-44"""
    return torch.unique(torch.round(pc / voxel) * voxel, dim=0)


def crop(pc, xb, yb, zb):
    """2cam/vision_pipeline_utils.py:78-84"""
    keep = ((pc[:, 0] >= xb[0]) & (pc[:, 0] <= xb[1]) & (pc[:, 1] >= yb[0]) & (pc[:, 1] <= yb[1]) &
            (pc[:, 2] >= zb[0]) & (pc[:, 2] <= zb[1]))
    return pc[keep]


def workspace_cloud(depth_np, intr, R, t, voxel, dev):
    """process_point_clouds for one camera (:233-247).  The ZED's XYZ measure is replaced by a full-frame
    back-projection of the synthetic depth map, then the reference's own steps."""
    depth = torch.tensor(depth_np, dtype=torch.float32, device=dev)
    idx = torch.nonzero(torch.ones_like(depth), as_tuple=False)
    pts = back_project(idx, depth, *intr)
    pts = pts[torch.isfinite(pts).all(dim=1)]
    pts = torch.mm(R, pts.T).T + t
    return voxel_unique(crop(pts, *BOX), voxel)


def object_clouds(masks, class_ids, depth_np, intr, R, t, voxel, dev, out):
    """process_masks (:292-304): per instance mask nonzero -> back-project -> transform -> voxels -> host"""
    depth = torch.tensor(depth_np, dtype=torch.float32, device=dev)
    for i, mask in enumerate(masks):
        idx = torch.nonzero(mask, as_tuple=False)
        pts = back_project(idx, depth, *intr)
        if pts.size(0) > 0:
            pc = voxel_unique(torch.mm(pts, R.T) + t, voxel)
            out.append((pc.cpu().numpy(), int(class_ids[i])))


def subtract(ws_np, obj_np, thr, dev):
    """subtract_point_clouds_gpu (:168-187): dense cdist, row-min, boolean index, host copy.  The M x N matrix is
    formed in row blocks when it would not fit (the reference itself would raise an out-of-memory error there)."""
    ws = torch.tensor(ws_np, dtype=torch.float32, device=dev)
    ob = torch.tensor(obj_np, dtype=torch.float32, device=dev)
    rows = max(1, int(4e9 // max(1, 4 * ob.shape[0])))
    if rows >= ws.shape[0]:
        keep = torch.cdist(ws, ob).min(dim=1)[0] > thr
    else:
        keep = torch.cat([torch.cdist(ws[a:a + rows], ob).min(dim=1)[0] > thr for a in range(0, ws.shape[0], rows)])
    return ws[keep].cpu().numpy()


def frame(fr, w, dev):
    """One loop iteration of 2cam/2cams.py:165-208, geometry only.  Returns seconds per stage (wall clock between the
    host round trips that end every stage, like the reference's own time.time() brackets)."""
    intr = fr["intr"]
    R = [torch.tensor(fr["R"][c], dtype=torch.float32, device=dev) for c in range(2)]       # 2cams.py:127-130
    t = [torch.tensor(fr["t"][c], dtype=torch.float32, device=dev) for c in range(2)]
    voxel, out = w["voxel"], {}
    t0 = time.perf_counter()
    ws = torch.cat([workspace_cloud(fr["depth"][c], intr, R[c], t[c], voxel, dev) for c in range(2)], dim=0).cpu().numpy()
    t1 = time.perf_counter()
    lists = [[], []]
    for c in range(2):
        masks = torch.tensor(fr["masks"][c], dtype=torch.float32, device=dev)               # masks.data: (n, H, W) fp32 on the device
        object_clouds(masks, fr["class_ids"][c], fr["depth"][c], intr, R[c], t[c], voxel, dev, lists[c])
    torch.cuda.synchronize(dev) if dev.type == "cuda" else None
    t2 = time.perf_counter()
    fused = [pc for pc, _ in _o.fuse_point_clouds_centroid(lists[0], lists[1], w["fuse_thr"])[2]]
    objs = np.vstack(fused) if fused else np.empty((0, 3))
    t3 = time.perf_counter()
    kept = subtract(ws, objs, w["thr"], dev) if objs.shape[0] else ws
    t4 = time.perf_counter()
    out["Point Cloud Processing"], out["Mask Processing"], out["Point Cloud Fusion"], out["Subtraction"] = t1 - t0, t2 - t1, t3 - t2, t4 - t3
    out["_kept_rows"] = float(kept.shape[0])
    return out
