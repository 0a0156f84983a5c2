"""GPU: fuse_point_clouds_centroid (2cam/vision_pipeline_utils.py:88-165) on the device, inside the batched frame
pipeline (fuse="centroid"), against the oracle's restatement of the reference's control flow
(oracle.fuse_point_clouds_centroid, itself pinned by tests/golden against the reference source): every fused cloud
bit-identical and in the reference's list order, the subtracted workspace bit-identical.  BASELINE.json configs[2]
(2 x HD1080, 10 + 2 overlapping objects, class ids alternating 39 / 41, 2 mm) runs at full size."""
import numpy as np
import pytest
import torch

from test_gpu_parity import assert_same

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def oracle_frame(oracle_mod, fr, voxel, fuse_thr, sub_thr, class_ids=None):
    from rt3d import synthetic as syn
    cx, cy, fx, fy = fr["intr"]
    n_obj = fr["masks"].shape[1]
    cls = fr["class_ids"] if class_ids is None else class_ids
    ws, lists = [], [[], []]
    ones = np.ones(fr["depth"][0].shape, np.uint8)
    for c in range(2):
        p = oracle_mod.backproject_indices(oracle_mod.mask_nonzero(ones), fr["depth"][c], cx, cy, fx, fy)
        ws.append(oracle_mod.voxel_snap_unique(oracle_mod.crop(oracle_mod.transform(p, fr["R"][c], fr["t"][c]), *syn.WORKSPACE_BOX), voxel))
        for o in range(n_obj):
            pc = oracle_mod.process_masks_cloud(fr["masks"][c, o], fr["depth"][c], cx, cy, fx, fy, fr["R"][c], fr["t"][c], voxel)
            if pc is not None:                                   # process_masks skips empty objects (:300)
                lists[c].append((pc, int(cls[c][o])))            # (:304)
    _, _, fused = oracle_mod.fuse_point_clouds_centroid(lists[0], lists[1], fuse_thr)
    ws_cat = np.vstack(ws)
    rows = [pc for pc, _ in fused if pc.shape[0]]
    if rows:
        all_obj = np.vstack(rows).astype(np.float32)             # subtract_point_clouds_gpu casts to fp32 (:171)
        kept = ws_cat[oracle_mod.subtract_mask(ws_cat, all_obj, sub_thr).astype(bool)]
    else:
        kept = ws_cat
    return dict(ws=ws, fused=fused, kept=kept)


def run_centroid(frames, voxel, fuse_thr, sub_thr, class_ids=None, mask_dtype=torch.uint8, **kw):
    from rt3d import synthetic as syn
    from rt3d.pipeline import FramePipeline
    F = len(frames)
    C, H, W = frames[0]["depth"].shape
    n_obj = frames[0]["masks"].shape[1]
    depth = torch.tensor(np.stack([f["depth"] for f in frames]), device=DEV)
    masks = torch.tensor(np.stack([f["masks"] for f in frames]), device=DEV).to(mask_dtype)
    pipe = FramePipeline(F, C, H, W, n_obj, frames[0]["intr"], frames[0]["R"], frames[0]["t"], syn.WORKSPACE_BOX, voxel=voxel,
                         distance_threshold=sub_thr, device=DEV, fuse="centroid", fuse_threshold=fuse_thr, mask_dtype=mask_dtype, **kw)
    pipe.set_class_ids(np.stack([f["class_ids"] for f in frames]) if class_ids is None else class_ids)
    pipe.run(depth, masks)
    res = pipe.results()
    pipe.run(depth, masks)      # second run on the self-cleaned arena
    for a, b in zip(res, pipe.results()):
        assert torch.equal(a["kept"], b["kept"]) and len(a["fused"]) == len(b["fused"])
        for x, y in zip(a["fused"], b["fused"]):
            assert torch.equal(x["points"], y["points"])
    return pipe, res


def check(oracle_mod, frames, res, voxel, fuse_thr, sub_thr, class_ids=None):
    kinds = {}
    for f, fr in enumerate(frames):
        want = oracle_frame(oracle_mod, fr, voxel, fuse_thr, sub_thr, None if class_ids is None else class_ids[f])
        for c, w in enumerate(want["ws"]):
            assert_same(f"frame {f} workspace cam {c}", res[f]["workspace_per_cam"][c].cpu().numpy(), w)
        got = res[f]["fused"]
        assert [g["class_id"] for g in got] == [cid for _, cid in want["fused"]], f"frame {f}: class order of the fused list"
        for i, (g, (w, cid)) in enumerate(zip(got, want["fused"])):
            # the reference returns float64 wherever the filter ran and the untouched float32 array otherwise
            assert (g["kind"] == "leftover") == (w.dtype == np.float32), f"frame {f} cloud {i}: kind {g['kind']} vs dtype {w.dtype}"
            assert_same(f"frame {f} fused cloud {i} ({g['kind']}, class {cid})", g["points"].cpu().numpy(), w.astype(np.float64))
            kinds[g["kind"]] = kinds.get(g["kind"], 0) + 1
        assert_same(f"frame {f} kept", res[f]["kept"].cpu().numpy(), want["kept"])
    return kinds


@pytest.mark.parametrize("voxel,fuse_thr,sub_thr", [(0.005, 0.3, 0.3), (0.005, 0.06, 0.06), (0.002, 0.3, 0.06), (0.01, 0.1, 0.005)])
def test_centroid_fusion_twelve_objects_qvga(oracle_mod, voxel, fuse_thr, sub_thr):
    """configs[2] in miniature: 10 + 2 overlapping objects, two classes, every camera listing its masks in its own order"""
    from rt3d import synthetic as syn
    frames = [syn.frame("QVGA", seed=20 + 2 * i, n_objects=12, overlap=2, shuffle_cams=True, jitter=0.01) for i in range(3)]
    _, res = run_centroid(frames, voxel, fuse_thr, sub_thr)
    kinds = check(oracle_mod, frames, res, voxel, fuse_thr, sub_thr)
    assert kinds.get("matched", 0) > 0 and kinds.get("leftover", 0) > 0      # the multi-match branch and the :155 quirk both ran
    if fuse_thr < 0.1:
        assert kinds.get("unmatched", 0) > 0


def test_centroid_fusion_with_bit_plane_neighbour_search(oracle_mod):
    """single-cloud groups (no duplicates) and pairs: bit-plane search (default) and grid search alone give the same bits"""
    from rt3d import synthetic as syn
    frames = [syn.frame("VGA360", seed=24 + 2 * i, n_objects=6, overlap=1, shuffle_cams=True, jitter=0.01) for i in range(2)]
    pipe, res = run_centroid(frames, 0.005, 0.3, 0.06, sor_bitplanes=True)
    check(oracle_mod, frames, res, 0.005, 0.3, 0.06)
    assert pipe.counts.tolist()[6] > 0
    pipe0, res0 = run_centroid(frames, 0.005, 0.3, 0.06, sor_bitplanes=False)
    check(oracle_mod, frames, res0, 0.005, 0.3, 0.06)
    assert pipe0.counts.tolist()[6] == 0


def test_centroid_fusion_mixed_classes(oracle_mod):
    """a class seen once per camera (1-vs-1, :123-126), one seen twice, one seen by camera 2 only (:162-163), and one
    seen by camera 1 only (:160) in the same frame"""
    from rt3d import synthetic as syn
    frames = [syn.frame("QVGA", seed=51 + i, n_objects=5, jitter=0.02) for i in range(2)]
    cls = np.array([[[39, 41, 41, 67, 67], [39, 41, 41, 0, 0]]] * 2)
    for fr in frames:
        fr["masks"][1, 3] = 0     # class 67: camera 1 sees two objects, camera 2 sees ...
        fr["masks"][1, 4] = 0     # ... none (its masks are empty; their class id 0 never reaches a dictionary)
    cls[:, 1, 3:] = 8
    frames[1]["masks"][0, 0] = 0  # frame 1: class 39 seen by camera 2 only
    _, res = run_centroid(frames, 0.005, 0.3, 0.3, class_ids=cls)
    kinds = check(oracle_mod, frames, res, 0.005, 0.3, 0.3, class_ids=cls)
    assert kinds.get("pair", 0) >= 1 and kinds.get("unmatched", 0) >= 2 and kinds.get("leftover", 0) >= 1


def test_centroid_fusion_duplicate_and_tiny_clouds(oracle_mod):
    """the same mask listed twice by camera 2 (two equal clouds: the :155 equality removes BOTH only when the filter
    drops nothing) and a one-pixel mask (its filtered cloud is empty: centroid NaN, no match, an empty fused entry)"""
    from rt3d import synthetic as syn
    fr = syn.frame("QVGA", seed=57, n_objects=4, jitter=0.02, flying=0.0)
    fr["masks"][1, 1] = fr["masks"][1, 0]                    # camera 2 reports object 0 twice
    one = np.zeros_like(fr["masks"][0, 3])
    ys, xs = np.nonzero(fr["masks"][0, 3] & np.isfinite(fr["depth"][0]) & (fr["depth"][0] > 0))
    one[ys[0], xs[0]] = 1
    fr["masks"][0, 3] = one                                  # camera 1 sees one pixel of object 3
    cls = np.array([[[39, 39, 39, 39], [39, 39, 39, 39]]])
    _, res = run_centroid([fr], 0.005, 0.3, 0.06, class_ids=cls)
    check(oracle_mod, [fr], res, 0.005, 0.3, 0.06, class_ids=cls)
    assert any(g["kind"] == "unmatched" and g["points"].shape[0] == 0 for g in res[0]["fused"])


def test_centroid_fusion_fp32_mask_planes(oracle_mod):
    """masks handed over as the ultralytics fp32 planes (masks.data, :295), non-zero = set"""
    from rt3d import synthetic as syn
    frames = [syn.frame("QVGA", seed=71, n_objects=3, jitter=0.02, shuffle_cams=True)]
    frames[0]["class_ids"][:] = 39
    pipe, res = run_centroid(frames, 0.005, 0.3, 0.3, mask_dtype=torch.float32)
    check(oracle_mod, frames, res, 0.005, 0.3, 0.3)
    assert pipe.count_launches() > 30


def test_centroid_fusion_hd1080_ten_objects_2mm(oracle_mod):
    """BASELINE.json configs[2] at full size: 2 x 1920x1080, 10 + 2 overlapping objects, ids alternating 39 / 41, 2 mm
    voxels, SOR(20, 1.5), fusion threshold 0.3 (2cams.py:197), subtraction 0.06 (2cams_mask_gpu.py:397)"""
    from rt3d import synthetic as syn
    frames = [syn.frame("HD1080", seed=20, n_objects=12, overlap=2, shuffle_cams=True)]
    _, res = run_centroid(frames, 0.002, 0.3, 0.06)
    kinds = check(oracle_mod, frames, res, 0.002, 0.3, 0.06)
    assert kinds.get("matched", 0) >= 5


def test_centroid_fusion_hd1080_thr03(oracle_mod):
    """same scene at the 2cams.py subtraction threshold 0.3 and 5 mm (the reference's own literals at HD1080)"""
    from rt3d import synthetic as syn
    frames = [syn.frame("HD1080", seed=21, n_objects=12, overlap=2, shuffle_cams=True)]
    _, res = run_centroid(frames, 0.005, 0.3, 0.3, project_kernel="tma")    # 12 masks through the tensor-map kernel: 4 staged, 8 from global memory
    check(oracle_mod, frames, res, 0.005, 0.3, 0.3)


def test_centroid_fusion_replays_from_a_cuda_graph(oracle_mod):
    from rt3d import synthetic as syn
    from rt3d.pipeline import FramePipeline
    fa = [syn.frame("QVGA", seed=81, n_objects=4, shuffle_cams=True)]
    fb = [syn.frame("QVGA", seed=83, n_objects=4, shuffle_cams=True)]
    depth = torch.tensor(np.stack([f["depth"] for f in fa]), device=DEV)
    masks = torch.tensor(np.stack([f["masks"] for f in fa]), device=DEV)
    pipe = FramePipeline(1, 2, 180, 320, 4, fa[0]["intr"], fa[0]["R"], fa[0]["t"], syn.WORKSPACE_BOX, distance_threshold=0.3,
                         device=DEV, fuse="centroid", fuse_threshold=0.3)
    pipe.set_class_ids(np.stack([f["class_ids"] for f in fa]))
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        pipe.run(depth, masks)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        pipe.run(depth, masks)
    for frames in (fb, fa):
        depth.copy_(torch.tensor(np.stack([f["depth"] for f in frames])))
        masks.copy_(torch.tensor(np.stack([f["masks"] for f in frames])))
        pipe.set_class_ids(np.stack([f["class_ids"] for f in frames]))
        graph.replay()
        torch.cuda.synchronize()
        check(oracle_mod, frames, pipe.results(), 0.005, 0.3, 0.3)


def test_invalid_class_rank_is_reported():
    from rt3d import synthetic as syn
    from rt3d._lib import Rt3dError
    from rt3d.pipeline import FramePipeline
    fr = syn.frame("QVGA", seed=8, n_objects=2)
    pipe = FramePipeline(1, 2, 180, 320, 2, fr["intr"], fr["R"], fr["t"], syn.WORKSPACE_BOX, device=DEV, fuse="centroid")
    pipe.set_class_ids(fr["class_ids"][None])
    pipe.class_rank.fill_(99)
    pipe.run(torch.tensor(fr["depth"][None], device=DEV), torch.tensor(fr["masks"][None], device=DEV))
    with pytest.raises(Rt3dError, match="out of range"):
        pipe.results()
