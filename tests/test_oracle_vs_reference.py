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
