"""Per-stage CUDA-event timings of the stand-alone ops on one synthetic HD720 frame (development aid)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rt-3d-reconstruction_b200"))
from rt3d import ops, synthetic as syn  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main(config="HD720", voxel=0.005):
    fr = syn.frame(config, seed=10, n_objects=1)
    dev = "cuda:0"
    depth = torch.tensor(fr["depth"][0], device=dev)
    mask = torch.tensor(fr["masks"][0, 0], device=dev)
    R, t = torch.tensor(fr["R"][0], device=dev), torch.tensor(fr["t"][0], device=dev)
    intr = fr["intr"]
    H, W = depth.shape
    ws_pts = ops.backproject_mask(depth, None, intr, R, t, syn.WORKSPACE_BOX)
    ob_pts = ops.backproject_mask(depth, mask, intr, R, t)
    ws_vox = ops.voxel_snap_unique(ws_pts, voxel)
    ob_vox = ops.voxel_snap_unique(ob_pts, voxel)
    pair = torch.cat([ob_vox, ob_vox + 0.0025])
    fused = ops.sor(pair, 20, 1.5)
    ws2 = torch.cat([ws_vox, ws_vox])
    print(f"{config}: ws pts {ws_pts.shape[0]}, ws voxels {ws_vox.shape[0]}, obj pts {ob_pts.shape[0]}, obj voxels {ob_vox.shape[0]}, pair {pair.shape[0]} -> {fused.shape[0]}")
    rows = [
        ("backproject ws (crop)", lambda: ops.backproject_mask(depth, None, intr, R, t, syn.WORKSPACE_BOX), 4 * H * W + 12 * ws_pts.shape[0]),
        ("backproject obj (u8 mask)", lambda: ops.backproject_mask(depth, mask, intr, R, t), 5 * H * W + 12 * ob_pts.shape[0]),
        ("voxel ws", lambda: ops.voxel_snap_unique(ws_pts, voxel), 12 * ws_pts.shape[0] + 12 * ws_vox.shape[0]),
        ("voxel obj", lambda: ops.voxel_snap_unique(ob_pts, voxel), 12 * ob_pts.shape[0] + 12 * ob_vox.shape[0]),
        ("sor pair", lambda: ops.sor(pair, 20, 1.5), 12 * pair.shape[0]),
        ("subtract 0.3", lambda: ops.subtract(ws2, fused.float(), 0.3), 12 * ws2.shape[0]),
        ("subtract 0.06", lambda: ops.subtract(ws2, fused.float(), 0.06), 12 * ws2.shape[0]),
        ("subtract 0.005", lambda: ops.subtract(ws2, fused.float(), 0.005), 12 * ws2.shape[0]),
        ("centroid obj", lambda: ops.centroid(ob_vox), 12 * ob_vox.shape[0]),
        ("sor ws (10mm-scale)", lambda: ops.sor(ws2, 20, 1.5), 12 * ws2.shape[0]),
    ]
    for name, fn, nbytes in rows:
        ms = timeit(fn, reps=10 if "sor ws" in name else 20)
        print(f"  {name:28s} {ms*1e3:9.1f} us   ({nbytes/ms/1e6:8.1f} GB/s algorithmic, includes host call + count read-back)")


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["HD720"]))
