"""GPU: the voxel-key set operations behind the optional box-wide fusion (BASELINE.json configs[3]):
rt3d_keys_rebase, rt3d_keys_decode, rt3d_keys_sort_unique, ops.keys_canonical and multi.box_wide_voxels -- the
latter over a real NCCL process group (world size 1 is enough for the driver's one-GPU pytest; the N > 1 run is
bench.py --config 4 under torchrun).  Expected values come from the oracle's voxel indices and np.unique."""
import os

import numpy as np
import pytest
import torch

from test_gpu_parity import assert_same, cu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BIAS = 1 << 20


def canonical_np(idx):
    """(N,3) int voxel indices -> canonical packed keys: 21 bits per axis, biased by 2^20, x major"""
    b = idx.astype(np.int64) + BIAS
    return ((b[:, 0] << 42) | (b[:, 1] << 21) | b[:, 2]).astype(np.int64)


def camera_points(oracle_mod, fr, c):
    from rt3d import synthetic as syn
    cx, cy, fx, fy = fr["intr"]
    ones = np.ones(fr["depth"][c].shape, np.uint8)
    p = oracle_mod.backproject_indices(oracle_mod.mask_nonzero(ones), fr["depth"][c], cx, cy, fx, fy)
    return oracle_mod.crop(oracle_mod.transform(p, fr["R"][c], fr["t"][c]), *syn.WORKSPACE_BOX)


@pytest.fixture(scope="module")
def rig_frame():
    from rt3d import synthetic as syn
    return syn.frame("QVGA", seed=30, n_objects=2, cams=tuple(range(8)), rig=8)


@pytest.mark.parametrize("voxel", [0.002, 0.005, 0.01])
def test_keys_canonical_and_decode(oracle_mod, rig_frame, voxel):
    from rt3d import ops
    pts = camera_points(oracle_mod, rig_frame, 0)
    lattice, keys, layout = ops.voxel_snap_unique(cu(pts), voxel, return_keys=True)
    want_idx = np.unique(oracle_mod.voxel_index(pts, voxel), axis=0)          # lexicographic = key order
    # raw per-call keys decode back to the rows they were returned with
    assert_same("decode(raw keys)", ops.keys_decode(keys, layout, voxel).cpu().numpy(), lattice.cpu().numpy())
    canon, canon_layout = ops.keys_canonical(keys, layout)
    assert_same("canonical keys", canon.cpu().numpy(), canonical_np(want_idx))
    assert canon_layout.cpu().tolist()[:6] == [-BIAS, -BIAS, -BIAS, 21, 21, 21]
    assert_same("decode(canonical keys)", ops.keys_decode(canon, canon_layout, voxel).cpu().numpy(),
                oracle_mod.voxel_snap_unique(pts, voxel))
    # canonical keys of different calls are comparable: a shifted cloud shares no layout but the same lattice
    shifted = pts + np.float32(40 * voxel)
    _, k2, l2 = ops.voxel_snap_unique(cu(shifted), voxel, return_keys=True)
    c2, _ = ops.keys_canonical(k2, l2)
    assert_same("canonical keys (shifted cloud)", c2.cpu().numpy(),
                canonical_np(np.unique(oracle_mod.voxel_index(shifted, voxel), axis=0)))


def test_keys_canonical_empty_and_range():
    from rt3d import ops
    from rt3d._lib import Rt3dError
    empty = torch.empty((0, 3), dtype=torch.float32, device=DEV)
    lattice, keys, layout = ops.voxel_snap_unique(empty, 0.005, return_keys=True)
    canon, canon_layout = ops.keys_canonical(keys, layout)
    assert canon.numel() == 0 and ops.keys_decode(canon, canon_layout, 0.005).shape == (0, 3)
    # 2^20 voxels of 1 mm = 1048.576 m: a point beyond that does not fit the 21-bit canonical axis
    far = torch.tensor([[0.0, 0.0, 0.0], [1100.0, 0.0, 0.0]], dtype=torch.float32, device=DEV)
    _, k, l = ops.voxel_snap_unique(far, 0.001, return_keys=True)
    with pytest.raises(Rt3dError, match="out of range"):
        ops.keys_canonical(k, l)


def _nccl_world1():
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29631")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
    return dist


@pytest.mark.parametrize("voxel", [0.005, 0.002])
def test_box_wide_voxels_eight_cameras(oracle_mod, rig_frame, voxel):
    """configs[3] in miniature: 8 camera streams, each voxelised on its own, canonical keys gathered and merged."""
    from rt3d import multi, ops
    dist = _nccl_world1()
    per_cam = [camera_points(oracle_mod, rig_frame, c) for c in range(8)]
    want_idx = np.unique(np.vstack([oracle_mod.voxel_index(p, voxel) for p in per_cam]), axis=0)
    want = (want_idx.astype(np.float32) * np.float32(voxel))
    # (a) 8 logical shards on one device: per-stream canonical keys, one sort-unique over their concatenation
    canon = []
    for p in per_cam:
        _, k, l = ops.voxel_snap_unique(cu(p), voxel, return_keys=True)
        canon.append(ops.keys_canonical(k, l))
    merged = ops.keys_sort_unique(torch.cat([c for c, _ in canon]), 63)
    assert_same("merged canonical keys", merged.cpu().numpy(), canonical_np(want_idx))
    assert_same("merged lattice", ops.keys_decode(merged, canon[0][1], voxel).cpu().numpy(), want)
    # (b) the same through the NCCL all-gather (this rank owns all 8 streams)
    got = multi.box_wide_voxels(cu(np.vstack(per_cam)), voxel, group=dist.group.WORLD)
    assert_same("box_wide_voxels", got.cpu().numpy(), want)
    parts = multi.all_gather_varlen(merged)
    assert len(parts) == 1 and torch.equal(parts[0], merged)


def test_box_wide_voxels_hd2k_camera(oracle_mod):
    """one full-size HD2K (2208 x 1242) stream of the rig at 5 mm"""
    from rt3d import multi, synthetic as syn
    dist = _nccl_world1()
    fr = syn.frame("HD2K", seed=31, n_objects=1, cams=(3,), rig=8)
    pts = camera_points(oracle_mod, fr, 0)
    want = oracle_mod.voxel_snap_unique(pts, 0.005)
    got = multi.box_wide_voxels(cu(pts), 0.005, group=dist.group.WORLD)
    assert_same("HD2K box_wide_voxels", got.cpu().numpy(), want)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_ops_on_a_non_current_device(oracle_mod, rig_frame):
    """a tensor on cuda:1 while cuda:0 is current (sort with payload needs > 48 KB of shared memory there)"""
    from rt3d import ops
    pts = camera_points(oracle_mod, rig_frame, 1)
    torch.cuda.set_device(0)
    t = torch.as_tensor(pts).to("cuda:1")
    assert_same("V1 on cuda:1", ops.voxel_snap_unique(t, 0.005).cpu().numpy(), oracle_mod.voxel_snap_unique(pts, 0.005))
    keys, want, cnt = oracle_mod.voxel_centroid(pts.astype(np.float64), 0.01)
    assert_same("V2 on cuda:1", ops.voxel_centroid(t.double(), 0.01).cpu().numpy(), want)
    assert_same("O1 on cuda:1", ops.sor(t[:5000]).cpu().numpy(), oracle_mod.filter_outliers_sor(pts[:5000]))
