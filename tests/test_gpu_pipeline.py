"""GPU: the batched frame pipeline (rt3d_frame_batch_run) against the oracle composed stage by stage."""
import numpy as np
import pytest
import torch

from test_gpu_parity import assert_same

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def oracle_frame(oracle_mod, fr, voxel, thr, k=20, ratio=1.5, recip=False):
    from rt3d import synthetic as syn
    cx, cy, fx, fy = fr["intr"]
    C, n_obj = fr["depth"].shape[0], fr["masks"].shape[1]
    ws, objs = [], [[None] * C for _ in range(n_obj)]
    ones = np.ones(fr["depth"][0].shape, np.uint8)
    for c in range(C):
        p = oracle_mod.backproject_indices(oracle_mod.mask_nonzero(ones), fr["depth"][c], cx, cy, fx, fy, recip=recip)
        ws.append(oracle_mod.voxel_snap_unique(oracle_mod.crop(oracle_mod.transform(p, fr["R"][c], fr["t"][c]), *syn.WORKSPACE_BOX), voxel, recip=recip))
        for o in range(n_obj):
            objs[o][c] = oracle_mod.process_masks_cloud(fr["masks"][c, o], fr["depth"][c], cx, cy, fx, fy, fr["R"][c], fr["t"][c], voxel, recip=recip)
    fused = []
    for o in range(n_obj):
        parts = [x for x in objs[o] if x is not None]
        if not parts:
            fused.append(np.empty((0, 3), np.float64))
        elif objs[o][0] is None:      # camera-0 cloud missing: the reference passes camera 2's cloud through (:162-163)
            fused.append(np.vstack(parts).astype(np.float64))
        else:
            fused.append(oracle_mod.filter_outliers_sor(np.vstack(parts), k, ratio))
    ws_cat = np.vstack(ws)
    all_obj = np.vstack(fused) if fused else np.empty((0, 3))
    kept = ws_cat if all_obj.shape[0] == 0 else ws_cat[oracle_mod.subtract_mask(ws_cat, all_obj.astype(np.float32), thr).astype(bool)]
    return dict(ws=ws, fused=fused, kept=kept)


def run_batch(frames, voxel, thr, flavour=None, **kw):
    from rt3d import synthetic as syn
    from rt3d.pipeline import FramePipeline
    F = len(frames)
    C, H, W = frames[0]["depth"].shape
    n_obj = frames[0]["masks"].shape[1]
    depth = torch.tensor(np.stack([f["depth"] for f in frames]), device=DEV)
    masks = torch.tensor(np.stack([f["masks"] for f in frames]), device=DEV) if n_obj else None
    pipe = FramePipeline(F, C, H, W, n_obj, frames[0]["intr"], frames[0]["R"], frames[0]["t"], syn.WORKSPACE_BOX, voxel=voxel,
                         flavour=flavour, distance_threshold=thr, device=DEV, **kw)
    pipe.run(depth, masks)
    res = pipe.results()
    # a second run on the same (self-cleaned) arena must give the same answer
    pipe.run(depth, masks)
    res2 = pipe.results()
    for a, b in zip(res, res2):
        assert torch.equal(a["kept"], b["kept"]) and torch.equal(a["workspace"], b["workspace"])
    return pipe, res


def check(oracle_mod, frames, res, voxel, thr, recip=False):
    for f, fr in enumerate(frames):
        want = oracle_frame(oracle_mod, fr, voxel, thr, recip=recip)
        for c, w in enumerate(want["ws"]):
            assert_same(f"frame {f} workspace cam {c}", res[f]["workspace_per_cam"][c].cpu().numpy(), w)
        for o, w in enumerate(want["fused"]):
            got = res[f]["objects"][o].cpu().numpy()
            assert got.dtype == np.float64
            assert_same(f"frame {f} fused object {o}", got, w)
        assert_same(f"frame {f} kept", res[f]["kept"].cpu().numpy(), want["kept"])


@pytest.mark.parametrize("voxel,thr", [(0.005, 0.3), (0.01, 0.06), (0.002, 0.005)])
def test_batch_qvga_two_objects(oracle_mod, voxel, thr):
    from rt3d import synthetic as syn
    frames = [syn.frame("QVGA", seed=40 + 3 * i, n_objects=2, jitter=0.03) for i in range(3)]
    _, res = run_batch(frames, voxel, thr)
    check(oracle_mod, frames, res, voxel, thr)


@pytest.mark.parametrize("project_kernel", ["tma", "cp_async"])
def test_batch_hd720_headline_config(oracle_mod, project_kernel):
    """config 2: 2 cameras HD720, 1 object, 5 mm, fusion threshold irrelevant (1-vs-1), subtraction 0.3; through both
    project kernels (the headline batch of 256 frame pairs takes the tensor-map one)"""
    from rt3d import synthetic as syn
    frames = [syn.frame("HD720", seed=10 + 2 * i, n_objects=1, jitter=0.05) for i in range(2)]
    pipe, res = run_batch(frames, 0.005, 0.3, project_kernel=project_kernel)
    check(oracle_mod, frames, res, 0.005, 0.3)
    assert pipe.count_launches() > 20


def test_batch_torch_cuda_flavour(oracle_mod):
    from rt3d import synthetic as syn
    frames = [syn.frame("QVGA", seed=77, n_objects=1)]
    _, res = run_batch(frames, 0.005, 0.3, flavour="torch_cuda")
    check(oracle_mod, frames, res, 0.005, 0.3, recip=True)


def test_batch_missing_objects_and_no_objects(oracle_mod):
    from rt3d import synthetic as syn
    fr = syn.frame("QVGA", seed=5, n_objects=2)
    fr["masks"][0, 0] = 0          # object 0 unseen by camera 0 -> camera 1's cloud passes through unfiltered
    fr["masks"][:, 1] = 0          # object 1 unseen by both cameras
    _, res = run_batch([fr], 0.005, 0.06)
    check(oracle_mod, [fr], res, 0.005, 0.06)
    # no instance masks at all: workspace only, nothing subtracted
    fr0 = syn.frame("QVGA", seed=6, n_objects=0)
    _, res0 = run_batch([fr0], 0.005, 0.3)
    want = oracle_frame(oracle_mod, fr0, 0.005, 0.3)
    assert_same("kept == workspace", res0[0]["kept"].cpu().numpy(), np.vstack(want["ws"]))


def test_batch_capacity_overflow_is_reported():
    from rt3d import synthetic as syn
    from rt3d._lib import Rt3dError
    fr = syn.frame("QVGA", seed=8, n_objects=1)
    with pytest.raises(Rt3dError, match="capacity"):
        run_batch([fr], 0.005, 0.3, obj_capacity=64)


def test_batch_host_fed_path(oracle_mod):
    from rt3d import synthetic as syn
    from rt3d.pipeline import FramePipeline
    frames = [syn.frame("QVGA", seed=90 + i, n_objects=1) for i in range(2)]
    depth = torch.tensor(np.stack([f["depth"] for f in frames])).pin_memory()
    masks = torch.tensor(np.stack([f["masks"] for f in frames])).pin_memory()
    pipe = FramePipeline(2, 2, 180, 320, 1, frames[0]["intr"], frames[0]["R"], frames[0]["t"], syn.WORKSPACE_BOX, device=DEV)
    h2d, d2h = pipe.run_host(depth, masks)
    assert h2d == depth.numel() * 4 + masks.numel() and d2h > 0
    want = [oracle_frame(oracle_mod, fr, 0.005, 0.3) for fr in frames]
    for f, hr in enumerate(pipe.host_results()):
        assert_same(f"host-fed kept {f}", hr["kept"], want[f]["kept"])
        assert_same(f"host-fed fused {f}", hr["objects"][0], want[f]["fused"][0])
    # a sequence of batches: copy-in, kernels and copy-out of neighbouring batches overlap; every batch's rows are
    # copied with their exact size after its counts were read and its status word checked
    seen = []
    pipe.run_host_pipelined([(depth, masks)] * 5, on_result=lambda i, h: seen.append((i, int(h["counts"][5]))))
    assert [i for i, _ in seen] == list(range(5)) and len({n for _, n in seen}) == 1
    for f, hr in enumerate(pipe.host_results()):
        assert_same(f"pipelined kept {f}", hr["kept"], want[f]["kept"])
    # instance masks already resident on the device (as YOLO leaves them in the reference): only depth crosses PCIe
    h2d, _ = pipe.run_host_pipelined([(depth, None)] * 3, device_masks=masks.to(DEV))
    assert h2d == depth.numel() * 4
    for f, hr in enumerate(pipe.host_results()):
        assert_same(f"device-mask kept {f}", hr["kept"], want[f]["kept"])


def test_batch_host_fed_overflow_is_loud():
    """the host-fed path must not hand back truncated rows when a capacity is exceeded"""
    from rt3d import synthetic as syn
    from rt3d._lib import Rt3dError
    from rt3d.pipeline import FramePipeline
    fr = syn.frame("QVGA", seed=90, n_objects=1)
    depth = torch.tensor(fr["depth"][None]).pin_memory()
    masks = torch.tensor(fr["masks"][None]).pin_memory()
    pipe = FramePipeline(1, 2, 180, 320, 1, fr["intr"], fr["R"], fr["t"], syn.WORKSPACE_BOX, device=DEV, ws_capacity=2000)
    with pytest.raises(Rt3dError, match="capacity"):
        pipe.run_host(depth, masks)


def test_batch_many_objects_fine_voxels(oracle_mod):
    """more instance masks than the project kernel stages in shared memory (4): the rest are read straight from
    global memory; 2 mm voxels as in configs[2]"""
    from rt3d import synthetic as syn
    frames = [syn.frame("QVGA", seed=61 + i, n_objects=6, jitter=0.02) for i in range(2)]
    for kernel in ("tma", "cp_async"):
        _, res = run_batch(frames, 0.002, 0.06, project_kernel=kernel)
        check(oracle_mod, frames, res, 0.002, 0.06)


def test_batch_single_camera_and_odd_width(oracle_mod):
    """configs[0]-style single camera; an image width that is a multiple of 4 but not of 16 (per-lane staging only)"""
    from rt3d import synthetic as syn
    frames = [syn.frame((90, 164), seed=17 + i, n_objects=1, cams=(1,)) for i in range(3)]   # H % 4 != 0: a cull cell straddles the last row
    _, res = run_batch(frames, 0.005, 0.3, project_kernel="tma")      # W % 16 != 0: the request quietly takes cp.async
    check(oracle_mod, frames, res, 0.005, 0.3)


def test_batch_eight_camera_rig(oracle_mod):
    """configs[3]-style rig: 8 cameras around the workspace, one object"""
    from rt3d import synthetic as syn
    fr = syn.frame("QVGA", seed=33, n_objects=1, cams=tuple(range(8)), rig=8)
    _, res = run_batch([fr], 0.005, 0.06, project_kernel="tma")
    check(oracle_mod, [fr], res, 0.005, 0.06)


def test_batch_replays_from_a_cuda_graph(oracle_mod):
    """the whole batch (40 launches, no host round trip) captured once and replayed on new inputs"""
    from rt3d import synthetic as syn
    from rt3d.pipeline import FramePipeline
    frames_a = [syn.frame("QVGA", seed=70 + i, n_objects=1) for i in range(2)]
    frames_b = [syn.frame("QVGA", seed=80 + i, n_objects=1) for i in range(2)]
    depth = torch.tensor(np.stack([f["depth"] for f in frames_a]), device=DEV)
    masks = torch.tensor(np.stack([f["masks"] for f in frames_a]), device=DEV)
    pipe = FramePipeline(2, 2, 180, 320, 1, frames_a[0]["intr"], frames_a[0]["R"], frames_a[0]["t"], syn.WORKSPACE_BOX,
                         distance_threshold=0.3, device=DEV)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        pipe.run(depth, masks)                     # warm-up outside the capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        pipe.run(depth, masks)
    for frames in (frames_b, frames_a):
        depth.copy_(torch.tensor(np.stack([f["depth"] for f in frames])))
        masks.copy_(torch.tensor(np.stack([f["masks"] for f in frames])))
        graph.replay()
        torch.cuda.synchronize()
        check(oracle_mod, frames, pipe.results(), 0.005, 0.3)


@pytest.mark.parametrize("seed", range(32))
def test_batch_random_sweep(oracle_mod, seed):
    """seeded sweep over image sizes (heights that are no multiple of the 4 x 4 cull cells, widths that are no multiple
    of the 64-pixel warp tile; the batch API wants W % 4 == 0, which every ZED resolution satisfies), object counts,
    voxel sizes and subtraction thresholds: every stage bit-identical to the oracle"""
    from rt3d import synthetic as syn
    rng = np.random.default_rng(9000 + seed)
    H = int(rng.integers(97, 260))
    W = 4 * int(rng.integers(33, 100)) if seed % 2 else 16 * int(rng.integers(9, 25))
    n_obj = int(rng.integers(1, 4))
    F = int(rng.integers(1, 5))
    voxel = float(rng.choice([0.003, 0.004, 0.005, 0.0075, 0.01, 0.02]))
    thr = float(rng.choice([0.004, 0.02, 0.06, 0.15, 0.3, 0.45]))
    frames = [syn.frame((H, W), seed=500 + 17 * seed + i, n_objects=n_obj, jitter=0.04) for i in range(F)]
    # both project kernels: the tensor-map (TMA) pipeline needs W % 16 == 0 and falls back to cp.async otherwise
    _, res = run_batch(frames, voxel, thr, project_kernel=("tma", "cp_async", "auto")[seed % 3])
    check(oracle_mod, frames, res, voxel, thr)


@pytest.mark.parametrize("config,voxel,n_obj", [("HD720", 0.005, 2), ("QVGA", 0.005, 2), ("VGA360", 0.01, 3), ("QVGA", 0.002, 1)])
def test_batch_bit_plane_neighbour_search(oracle_mod, config, voxel, n_obj):
    """sor_bitplanes=True (the default): the outlier filter settles the dense part of every cloud by probing lattice offsets in two
    occupancy bit planes per group and hands the rest (sparse regions, outliers outside the 3-sigma core box) to the
    grid search: the same bits as the oracle either way, and most queries do take the new path"""
    from rt3d import synthetic as syn
    frames = [syn.frame(config, seed=310 + 5 * i, n_objects=n_obj, jitter=0.03) for i in range(2)]
    pipe, res = run_batch(frames, voxel, 0.3, sor_bitplanes=True)
    check(oracle_mod, frames, res, voxel, 0.3)
    counts = pipe.counts.tolist()
    if config == "HD720":      # dense clouds (pixel footprint below the voxel size); QVGA clouds are too sparse for 3-voxel shells
        assert counts[6] > 0.3 * counts[2], "bit planes settled too few queries"
    pipe0, res0 = run_batch(frames, voxel, 0.3, sor_bitplanes=False)     # grid search only: the same bits
    check(oracle_mod, frames, res0, voxel, 0.3)
    assert pipe0.counts.tolist()[6] == 0


@pytest.mark.parametrize("config,voxel,thr,n_obj", [("VGA360", 0.01, 0.06, 1), ("QVGA", 0.005, 0.3, 2), ("QVGA", 0.01, 0.06, 0)])
def test_batch_workspace_outlier_filter(oracle_mod, config, voxel, thr, n_obj):
    """workspace_sor=True: every frame's fused workspace (cam1 || cam2 voxel sets, duplicates kept) goes through
    remove_statistical_outlier(20, 1.5) before the subtraction, as 2cam/2cams_mask_cpu.py:523-530 does (voxel 0.01,
    subtraction 0.06 there); the reported per-camera voxel sets stay unfiltered"""
    from rt3d import synthetic as syn
    frames = [syn.frame(config, seed=120 + 3 * i, n_objects=n_obj, jitter=0.02) for i in range(3)]
    pipe, res = run_batch(frames, voxel, thr, workspace_sor=True)
    total = 0
    for f, fr in enumerate(frames):
        want = oracle_frame(oracle_mod, fr, voxel, thr)
        for c, w in enumerate(want["ws"]):
            assert_same(f"frame {f} workspace cam {c} (unfiltered)", res[f]["workspace_per_cam"][c].cpu().numpy(), w)
        ws_cat = np.vstack(want["ws"])
        ws_f = oracle_mod.filter_outliers_sor(ws_cat).astype(np.float32)        # float64 out of Open3D, fp32 again at :170
        total += ws_f.shape[0]
        assert 0 < ws_f.shape[0] < ws_cat.shape[0]
        objs = [o for o in want["fused"] if o.shape[0]]
        kept = ws_f[oracle_mod.subtract_mask(ws_f, np.vstack(objs).astype(np.float32), thr).astype(bool)] if objs else ws_f
        assert_same(f"frame {f} kept (filtered workspace)", res[f]["kept"].cpu().numpy(), kept)
    assert pipe.counts.tolist()[7] == total
