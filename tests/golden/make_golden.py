"""Generate tests/golden/*.npz by running the UNMODIFIED reference on torch-CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The reference module is imported in place with stub modules for its missing dependencies
(oracle/ref_import.py); nothing of it is copied.  Open3D is not installable, so the two fixtures
that involve filter_outliers_sor (fusion) run the reference's control flow on top of the C
oracle's restatement of remove_statistical_outlier -- those pin the control flow, not Open3D.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rt-3d-reconstruction_b200"))

from oracle import oracle, ref_import  # noqa: E402
from rt3d import synthetic as syn  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def geometry_fixture(name, config, seed, n_objects, light=False):
    ref = ref_import.load_reference(cpu_torch=True)
    fr = syn.frame(config, seed=seed, n_objects=n_objects)
    cx, cy, fx, fy = fr["intr"]
    out = dict(depth=fr["depth"], masks=fr["masks"], class_ids=fr["class_ids"], R=fr["R"], t=fr["t"],
               intr=np.array(fr["intr"], np.float64), box=np.array(syn.WORKSPACE_BOX, np.float64))
    ws_sets, cams = [], []
    for c in range(2):
        depth_t = torch.tensor(fr["depth"][c])
        R, t = torch.tensor(fr["R"][c]), torch.tensor(fr["t"][c])
        # workspace flavour: every pixel -> B1 -> T1 -> C1 -> V1
        idx = torch.nonzero(torch.ones_like(depth_t), as_tuple=False)
        pts = ref.convert_mask_to_3d_points(idx, depth_t, cx, cy, fx, fy)
        tr = torch.mm(pts, R.T) + t
        cr = ref.crop_point_cloud_gpu(tr, *syn.WORKSPACE_BOX)
        vx = ref.downsample_point_cloud_gpu(cr, 0.005)
        out[f"ws{c}_voxels"] = vx.numpy()
        if not light:
            out[f"ws{c}_points"] = pts.numpy(); out[f"ws{c}_transformed"] = tr.numpy()
            out[f"ws{c}_cropped"] = cr.numpy()
            out[f"ws{c}_voxels_10mm"] = ref.downsample_point_cloud_gpu(cr, 0.01).numpy()
            out[f"ws{c}_voxels_2mm"] = ref.downsample_point_cloud_gpu(cr, 0.002).numpy()
        ws_sets.append(vx)
        # object flavour: the reference's own process_masks, unmodified (fp32 0/1 planes as ultralytics gives them)
        masks = types.SimpleNamespace(data=torch.tensor(fr["masks"][c]).float())
        clouds = []
        ref.process_masks(masks, fr["class_ids"][c], fr["depth"][c], cx, cy, fx, fy, R, t, "cpu", clouds)
        cams.append(clouds)
        out[f"obj{c}_n"] = np.array(len(clouds))
        for k, (pc, cid) in enumerate(clouds):
            out[f"obj{c}_{k}_cloud"] = pc; out[f"obj{c}_{k}_class"] = np.array(cid)
        for k in range(fr["masks"].shape[1]):
            mi = torch.nonzero(masks.data[k], as_tuple=False)
            out[f"obj{c}_{k}_nonzero"] = mi.numpy().astype(np.int32)
            out[f"obj{c}_{k}_points"] = ref.convert_mask_to_3d_points(mi, depth_t, cx, cy, fx, fy).numpy()
    ws = torch.cat(ws_sets).numpy()
    out["ws_fused"] = ws
    # fusion: reference control flow (:88-165) over the oracle's SOR restatement
    for thr, tag in ((0.3, "thr030"), (0.02, "thr002")):
        pcs1, pcs2, fused = ref.fuse_point_clouds_centroid(cams[0], cams[1], thr)
        out[f"fuse_{tag}_n"] = np.array(len(fused))
        for k, (pc, cid) in enumerate(fused):
            out[f"fuse_{tag}_{k}_cloud"] = np.asarray(pc); out[f"fuse_{tag}_{k}_class"] = np.array(cid)
        out[f"fuse_{tag}_pcs1_n"] = np.array(len(pcs1)); out[f"fuse_{tag}_pcs2_n"] = np.array(len(pcs2))
        if tag == "thr030":
            objs = np.vstack([np.asarray(pc) for pc, _ in fused])
            out["objects_concat"] = objs
    # subtraction: the reference's cdist formulation on torch-CPU + the near-tie rows where its
    # fp32 rounding noise (abs err ~1.7e-5) may legitimately disagree with the exact predicate
    for thr, tag in ((0.3, "thr030"), (0.06, "thr006"), (0.005, "thr0005")):
        kept = ref.subtract_point_clouds_gpu(ws, objs, thr)
        keep_exact, near = oracle.subtract_mask(ws, objs, thr, band=4e-5)
        out[f"sub_{tag}_kept"] = kept; out[f"sub_{tag}_near"] = near
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if k.endswith(("voxels", "fused", "kept", "concat"))})


def edge_fixture():
    """Tiny hand-checkable cases run through the reference's functions."""
    ref = ref_import.load_reference(cpu_torch=True)
    out = {}
    # 3x4 depth patch with every kind of invalid value
    depth = np.array([[1.0, np.nan, np.inf, -np.inf], [0.0, -1.5, 2.0, 1e-3], [3.25, 0.5, 19.75, 1.0]], np.float32)
    mask = np.array([[1, 1, 1, 1], [1, 1, 1, 0], [0, 1, 1, 1]], np.float32)
    idx = torch.nonzero(torch.tensor(mask), as_tuple=False)
    out["patch_depth"] = depth; out["patch_mask"] = mask; out["patch_idx"] = idx.numpy()
    out["patch_points"] = ref.convert_mask_to_3d_points(idx, torch.tensor(depth), 1.7, 1.2, 3.3, 3.1).numpy()
    # half-to-even voxel boundaries and signed zeros
    pts = np.array([[0.0025, 0.0075, -0.0025], [0.0125, -0.0075, 0.0], [-0.0, 0.0049, -0.0049], [0.0025, 0.0075, -0.0025],
                    [1.0, 2.0, 3.0], [-1.23456, 0.7, 1e-9]], np.float32)
    out["snap_in"] = pts
    for v, tag in ((0.005, "5mm"), (0.01, "10mm"), (0.002, "2mm")):
        out[f"snap_{tag}"] = ref.downsample_point_cloud_gpu(torch.tensor(pts), v).numpy()
    # crop: closed intervals, bounds rounded to fp32
    cp = np.array([[-0.25, -0.5, -0.05], [0.75, 1.75, 2.0], [-0.05, 0.0, -0.05], [0.0, 0.0, np.float32(-0.05)],
                   [0.7500001, 0, 0], [0.5, 1.7500001, 1], [np.nan, 0, 0], [0.1, 0.2, 0.3]], np.float32)
    out["crop_in"] = cp
    out["crop_out"] = ref.crop_point_cloud_gpu(torch.tensor(cp), (-0.25, 0.75), (-0.5, 1.75), (-0.05, 2)).numpy()
    np.savez_compressed(os.path.join(OUT, "edge_cases.npz"), **out)
    print("edge_cases", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if not ref_import.available():
        sys.exit("reference not present; golden fixtures can only be regenerated in the build container")
    geometry_fixture("qvga_1obj", "QVGA", seed=10, n_objects=1)
    geometry_fixture("qvga_4obj", "QVGA", seed=20, n_objects=4, light=True)
    edge_fixture()
