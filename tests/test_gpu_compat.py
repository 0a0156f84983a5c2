"""GPU: compat_mask_cpu (the helpers 2cam/2cams_mask_cpu.py defines privately) against the oracle, and the
opt-in bulk-copy staging of the batched project kernel against the default one."""
import os

import numpy as np
import pytest
import torch

from test_gpu_parity import assert_same, cu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def compat():
    import compat_mask_cpu
    return compat_mask_cpu


def lattice_scene(rng, n_ws=6000, n_obj=400, v=0.005):
    """workspace / object clouds on the same voxel lattice, so that exact ties d == r are common"""
    ws = (rng.integers(-60, 60, size=(n_ws, 3)) * v).astype(np.float32)
    ob = (rng.integers(-8, 8, size=(n_obj, 3)) * v).astype(np.float32)
    return np.unique(ws, axis=0), np.unique(ob, axis=0)


@pytest.mark.parametrize("radius", [0.01, 0.005, 0.025, 0.3, 0.06])
def test_kdtree_subtract_strict_radius(compat, oracle_mod, radius):
    """2cams_mask_cpu.py:246-270: removed iff d^2 < r^2 -- ties are KEPT (the cdist variant removes them)."""
    from rt3d import ops
    rng = np.random.default_rng(int(radius * 1e4))
    ws, ob = lattice_scene(rng)
    want = oracle_mod.subtract_radius_mask(ws.astype(np.float64), ob.astype(np.float64), radius)
    got = compat.subtract_point_clouds(ws, ob, radius)
    assert got.dtype == np.float64
    assert_same(f"strict radius {radius}", got, ws[want.astype(bool)].astype(np.float64))
    # an exact tie: a point at distance exactly r (representable in fp32) stays under "<", goes under "<="
    ws2 = np.array([[0.25, 0.0, 0.0], [0.125, 0.0, 0.0]], np.float32)
    ob2 = np.zeros((1, 3), np.float32)
    assert compat.subtract_point_clouds(ws2, ob2, 0.25).shape[0] == 1 and compat.subtract_point_clouds(ws2, ob2, 0.125).shape[0] == 2
    assert ops.subtract(cu(ws2), cu(ob2), 0.25).shape[0] == 0 and ops.subtract(cu(ws2), cu(ob2), 0.125).shape[0] == 1


def test_kdtree_subtract_returns_the_callers_float64_rows(compat):
    """2cams_mask_cpu.py feeds float64 clouds; the rows that come back are the caller's, not fp32-rounded copies"""
    rng = np.random.default_rng(4)
    ws = rng.uniform(-0.3, 0.3, size=(3000, 3))                       # float64, not representable in fp32
    ob = rng.uniform(-0.05, 0.05, size=(200, 3))
    got = compat.subtract_point_clouds(ws, ob, 0.02)
    assert got.dtype == np.float64 and 0 < got.shape[0] < ws.shape[0]
    rows = {r.tobytes() for r in ws}
    assert all(r.tobytes() in rows for r in got)
    d = np.sqrt(((ws[:, None, :] - ob[None, :, :]) ** 2).sum(-1)).min(1)
    far = ws[d > 0.02 * (1 + 1e-6)]
    assert {r.tobytes() for r in far} <= {r.tobytes() for r in got}


def test_kdtree_subtract_empty_inputs(compat):
    ws = np.zeros((5, 3), np.float32)
    assert compat.subtract_point_clouds(ws, np.empty((0, 3), np.float32)).shape == (5, 3)
    assert compat.subtract_point_clouds(np.empty((0, 3), np.float32), ws).shape == (0, 3)


def test_mean_voxel_downsample(compat, oracle_mod):
    rng = np.random.default_rng(3)
    for dtype in (np.float32, np.float64):
        p = (rng.normal(size=(20000, 3)) * 0.2).astype(dtype)
        _, want, _ = oracle_mod.voxel_centroid(p, 0.01)
        got = compat.downsample_point_cloud(p, 0.01)
        assert got.dtype == dtype and got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-5      # north_star tolerance for centroids (fp32 accumulation order)
    assert compat.downsample_point_cloud(np.empty((0, 3), np.float32)).shape == (0, 3)


def test_script_local_fusion_skips_the_filter_for_single_pairs(compat, oracle_mod):
    from rt3d import synthetic as syn
    fr = syn.frame("QVGA", seed=21, n_objects=3)
    cx, cy, fx, fy = fr["intr"]
    clouds = [[], []]
    for c in range(2):
        for o in range(3):
            pc = oracle_mod.process_masks_cloud(fr["masks"][c, o], fr["depth"][c], cx, cy, fx, fy, fr["R"][c], fr["t"][c], 0.005)
            if pc is not None:
                clouds[c].append((pc, 39 if o < 2 else 41))      # class 39 twice per camera -> matching branch; 41 once -> plain stack
    p1, p2, fused = compat.fuse_point_clouds_centroid(clouds[0], clouds[1], 0.3)
    want = oracle_mod.fuse_point_clouds_centroid(clouds[0], clouds[1], 0.3, sor_single_pair=False, sor_candidates=False)
    assert len(fused) == len(want[2])
    for (a, ca), (b, cb) in zip(fused, want[2]):
        assert ca == cb
        assert_same("fused cloud", a, b)
    single = [pc for pc, cid in fused if cid == 41]
    assert len(single) == 1 and single[0].dtype == np.float32      # stacked, never filtered: still float32


def test_erode_and_misc(compat):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    m = (rng.random((180, 320)) > 0.02).astype(np.uint8)
    m[60:120, 100:200] = 1
    want = cv2.erode(m, np.ones((12, 12), np.uint8), iterations=2)
    assert_same("erode_mask", compat.erode_mask(m, iterations=2), want)
    pc = torch.tensor(rng.normal(size=(1000, 3)), dtype=torch.float32, device=DEV)
    got = compat.calculate_centroid_gpu(pc).cpu().numpy()
    assert np.abs(got - pc.mean(dim=0).cpu().numpy()).max() < 1e-5
    idx = np.argwhere(m[:20, :30] > 0)
    depth = torch.rand(180, 320)
    a = compat.convert_mask_to_3d_points_cpu(idx, depth, 160.1, 90.2, 131.9, 131.7)
    b = compat.convert_mask_to_3d_points_gpu(torch.tensor(idx, device=DEV), depth.to(DEV), 160.1, 90.2, 131.9, 131.7)
    assert torch.equal(a, b)


def test_voxel_grid_subtract_is_a_set_difference(compat):
    rng = np.random.default_rng(9)
    ws = rng.uniform(-0.2, 0.2, size=(4000, 3))
    out = compat.voxel_grid_subtract(ws, ws[:10], 0.02)
    assert out.ndim == 2 and out.shape[1] == 3 and out.shape[0] <= 4000
    assert compat.voxel_grid_subtract(ws, ws, 0.02).shape[0] == 0       # identical clouds voxelise identically
