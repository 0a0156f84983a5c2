"""CPU, build container only: the oracle against the live reference module at full HD720 size.

/root/reference does not travel to the GPU box -- these tests skip there; the committed golden
vectors (tests/golden, same generator) cover the same ground at QVGA size everywhere."""
import numpy as np
import pytest

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    return ref_import.load_reference(cpu_torch=True)


def test_hd720_workspace_and_object_chain(oracle_mod, ref):
    import torch
    from rt3d import synthetic as syn
    fr = syn.frame("HD720", seed=10, n_objects=1)
    cx, cy, fx, fy = fr["intr"]
    for c in range(2):
        d = fr["depth"][c]
        dt, R, t = torch.tensor(d), torch.tensor(fr["R"][c]), torch.tensor(fr["t"][c])
        idx = torch.nonzero(torch.ones_like(dt), as_tuple=False)
        pts = ref.convert_mask_to_3d_points(idx, dt, cx, cy, fx, fy)
        tr = torch.mm(R, pts.T).T + t          # the :239-240 form
        cr = ref.crop_point_cloud_gpu(tr, *syn.WORKSPACE_BOX)
        vx = ref.downsample_point_cloud_gpu(cr, 0.005)
        o_pts = oracle_mod.backproject_indices(oracle_mod.mask_nonzero(np.ones(d.shape, np.uint8)), d, cx, cy, fx, fy)
        o_tr = oracle_mod.transform(o_pts, fr["R"][c], fr["t"][c])
        o_cr = oracle_mod.crop(o_tr, *syn.WORKSPACE_BOX)
        assert np.array_equal(o_pts, pts.numpy()) and np.array_equal(o_tr, tr.numpy()) and np.array_equal(o_cr, cr.numpy())
        assert np.array_equal(oracle_mod.voxel_snap_unique(o_cr, 0.005), vx.numpy())
        # xyz ingest path of process_point_clouds (:233-238)
        xyz = syn.xyz_image(d)
        flat = torch.tensor(xyz[:, :, :3]).reshape(-1, 3)
        assert np.array_equal(oracle_mod.xyz_ingest(xyz), flat[torch.isfinite(flat).all(dim=1)].numpy())


def test_centroid_is_numpy_mean(oracle_mod, ref):
    rng = np.random.default_rng(0)
    a = rng.normal(size=(4000, 3)).astype(np.float32)
    assert np.array_equal(oracle_mod.centroid(a), ref.calculate_centroid(a))


def test_torch_eager_restatement_matches_the_reference_call_for_call(ref):
    """oracle/torch_eager.py (bench.py --impl torch_eager: the reference's op sequence, timed on the GPU box where the
    reference itself is absent) returns exactly what the imported reference functions return, stage by stage"""
    import types

    import torch
    from oracle import torch_eager as te
    from rt3d import synthetic as syn
    fr = syn.frame("QVGA", seed=40, n_objects=3, jitter=0.03)
    intr = fr["intr"]
    dev = torch.device("cpu")
    for c in range(2):
        depth, R, t = torch.tensor(fr["depth"][c]), torch.tensor(fr["R"][c]), torch.tensor(fr["t"][c])
        idx = torch.nonzero(torch.tensor(fr["masks"][c, 0]), as_tuple=False)
        assert torch.equal(te.back_project(idx, depth, *intr), ref.convert_mask_to_3d_points(idx, depth, *intr))
        pts = torch.mm(ref.convert_mask_to_3d_points(idx, depth, *intr), R.T) + t
        assert torch.equal(te.voxel_unique(pts, 0.005), ref.downsample_point_cloud_gpu(pts, 0.005))
        assert torch.equal(te.crop(pts, *te.BOX), ref.crop_point_cloud_gpu(pts, *te.BOX))
        got, want = [], []
        masks = torch.tensor(fr["masks"][c], dtype=torch.float32)
        te.object_clouds(masks, fr["class_ids"][c], fr["depth"][c], intr, R, t, 0.005, dev, got)
        ref.process_masks(types.SimpleNamespace(data=masks), fr["class_ids"][c], fr["depth"][c], *intr, R, t, dev, want)
        assert len(got) == len(want) and all(np.array_equal(a, b) and ca == cb for (a, ca), (b, cb) in zip(got, want))
    ws = np.vstack([te.workspace_cloud(fr["depth"][c], intr, torch.tensor(fr["R"][c]), torch.tensor(fr["t"][c]), 0.005, dev).numpy() for c in range(2)])
    assert np.array_equal(te.subtract(ws, got[0][0], 0.06, dev), ref.subtract_point_clouds_gpu(ws, got[0][0], 0.06))
    assert te.BOX == ((-0.25, 0.75), (-0.5, 1.75), (-0.05, 2))
