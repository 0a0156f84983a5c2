"""GPU: every CUDA stage, called through the C ABI, against the CPU oracle and the golden vectors.

Bars (BASELINE.json north_star): point counts, voxel sets and outlier masks bit-exact; coordinates
within 1e-5 m -- in practice everything below except the fp64 transform is compared bit for bit.
"""
import numpy as np
import pytest
import torch

from conftest import lex_sorted

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from rt3d import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def hd720():
    from rt3d import synthetic as syn
    return syn.frame("HD720", seed=10, n_objects=1)


def cu(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def report_diff(name, got, want):
    got, want = np.asarray(got), np.asarray(want)
    if got.shape != want.shape:
        return f"{name}: shape {got.shape} != {want.shape}"
    bad = np.nonzero(~((got == want) | (np.isnan(got) & np.isnan(want))))
    if len(bad[0]) == 0:
        return None
    i = tuple(b[0] for b in bad)
    return f"{name}: {len(bad[0])} of {got.size} elements differ; first at {i}: got {got[i]!r} want {want[i]!r}"


def assert_same(name, got, want):
    msg = report_diff(name, got, want)
    assert msg is None, msg


# ------------------------------------------------------------------------------------------- B0
@pytest.mark.parametrize("shape", [(3, 4), (7, 13), (64, 100), (180, 320), (720, 1280), (1, 1), (5, 1023)])
@pytest.mark.parametrize("dtype", ["u8", "f32"])
def test_mask_nonzero(ops, oracle_mod, shape, dtype):
    rng = np.random.default_rng(hash(shape) % 1000)
    for density in (0.0, 0.03, 0.5, 1.0):
        m = (rng.random(shape) < density)
        m = m.astype(np.uint8) if dtype == "u8" else m.astype(np.float32) * rng.choice([1.0, -2.5, np.nan], shape).astype(np.float32)
        if dtype == "f32":
            m = np.where(np.isnan(m) & (rng.random(shape) < 0.5), 0.0, m).astype(np.float32)
        got = ops.mask_nonzero(cu(m)).cpu().numpy()
        assert_same(f"nonzero{shape}{dtype}{density}", got, oracle_mod.mask_nonzero(m))


# ------------------------------------------------------------------------------------------- B1
@pytest.mark.parametrize("flavour,recip", [("ieee", False), ("torch_cuda", True)])
def test_backproject_indices_hd720(ops, oracle_mod, hd720, flavour, recip):
    cx, cy, fx, fy = hd720["intr"]
    for c in range(2):
        depth = hd720["depth"][c]
        for mask in (hd720["masks"][c, 0], np.ones(depth.shape, np.uint8)):
            idx = oracle_mod.mask_nonzero(mask)
            want = oracle_mod.backproject_indices(idx, depth, cx, cy, fx, fy, recip=recip)
            got = ops.backproject_indices(cu(idx), cu(depth), cx, cy, fx, fy, flavour=flavour).cpu().numpy()
            assert_same(f"b1 cam{c}", got, want)


def test_backproject_indices_edge(ops, oracle_mod, golden):
    g = golden("edge_cases")
    got = ops.backproject_indices(cu(g["patch_idx"]), cu(g["patch_depth"]), 1.7, 1.2, 3.3, 3.1).cpu().numpy()
    assert_same("patch", got, g["patch_points"])
    empty = ops.backproject_indices(torch.empty((0, 2), dtype=torch.int64, device=DEV), cu(g["patch_depth"]), 1.7, 1.2, 3.3, 3.1)
    assert tuple(empty.shape) == (0, 3)
    # unordered + repeated indices keep their order
    idx = np.array([[2, 3], [0, 0], [2, 3], [1, 2], [2, 0]], np.int64)
    want = oracle_mod.backproject_indices(idx, g["patch_depth"], 1.7, 1.2, 3.3, 3.1)
    assert_same("repeat", ops.backproject_indices(cu(idx), cu(g["patch_depth"]), 1.7, 1.2, 3.3, 3.1).cpu().numpy(), want)


# ------------------------------------------------------------------------------- fused B0+B1+T1+C1
@pytest.mark.parametrize("mask_kind", ["none", "u8", "f32", "bool"])
@pytest.mark.parametrize("with_rt,with_box", [(False, False), (True, False), (True, True)])
def test_backproject_mask_fused_hd720(ops, oracle_mod, hd720, mask_kind, with_rt, with_box):
    from rt3d import synthetic as syn
    cx, cy, fx, fy = hd720["intr"]
    for c in range(2):
        depth = hd720["depth"][c]
        m = None if mask_kind == "none" else hd720["masks"][c, 0]
        want = oracle_mod.backproject_indices(oracle_mod.mask_nonzero(np.ones(depth.shape, np.uint8) if m is None else m), depth, cx, cy, fx, fy)
        if with_rt:
            want = oracle_mod.transform(want, hd720["R"][c], hd720["t"][c])
        if with_box:
            want = oracle_mod.crop(want, *syn.WORKSPACE_BOX)
        mt = None if m is None else cu(m, {"u8": torch.uint8, "f32": torch.float32, "bool": torch.bool}[mask_kind])
        got = ops.backproject_mask(cu(depth), mt, hd720["intr"], R=cu(hd720["R"][c]) if with_rt else None,
                                   t=cu(hd720["t"][c]) if with_rt else None, box=syn.WORKSPACE_BOX if with_box else None)
        assert_same(f"fused cam{c}", got.cpu().numpy(), want)


@pytest.mark.parametrize("shape", [(5, 7), (33, 130), (180, 320), (241, 322)])
def test_backproject_mask_ragged_shapes(ops, oracle_mod, shape):
    """widths that are not multiples of 4 take the scalar path; sliced (unaligned) planes too"""
    rng = np.random.default_rng(7)
    H, W = shape
    depth = rng.uniform(-0.2, 3.0, shape).astype(np.float32)
    depth[rng.random(shape) < 0.1] = np.nan
    depth[rng.random(shape) < 0.05] = np.inf
    mask = (rng.random(shape) < 0.4).astype(np.uint8)
    R = np.linalg.qr(rng.normal(size=(3, 3)))[0].astype(np.float32)
    t = rng.normal(size=3).astype(np.float32)
    intr = (W / 2 + 0.3, H / 2 - 0.2, 200.7, 201.3)
    want = oracle_mod.transform(oracle_mod.backproject_indices(oracle_mod.mask_nonzero(mask), depth, *intr), R, t)
    got = ops.backproject_mask(cu(depth), cu(mask), intr, R=cu(R), t=cu(t))
    assert_same(f"ragged{shape}", got.cpu().numpy(), want)
    # a plane sliced out of a stack starting at an odd offset
    stack = torch.zeros((H * W + 1,), dtype=torch.float32, device=DEV)
    stack[1:] = cu(depth).reshape(-1)
    got2 = ops.backproject_mask(stack[1:].reshape(H, W), cu(mask), intr, R=cu(R), t=cu(t))
    assert_same(f"unaligned{shape}", got2.cpu().numpy(), want)


# ------------------------------------------------------------------------------------- W0 / C1 / T1
def test_xyz_select_crop_transform(ops, oracle_mod, hd720):
    from rt3d import synthetic as syn
    for c in range(2):
        xyz = syn.xyz_image(hd720["depth"][c])
        pts = oracle_mod.xyz_ingest(xyz)
        assert_same("ingest", ops.xyz_select(cu(xyz)).cpu().numpy(), pts)
        tr = oracle_mod.transform(pts, hd720["R"][c], hd720["t"][c])
        assert_same("transform", ops.transform(cu(pts), cu(hd720["R"][c]), cu(hd720["t"][c])).cpu().numpy(), tr)
        cr = oracle_mod.crop(tr, *syn.WORKSPACE_BOX)
        assert_same("crop", ops.crop(cu(tr), *syn.WORKSPACE_BOX).cpu().numpy(), cr)
        fused = ops.xyz_select(cu(xyz), True, cu(hd720["R"][c]), cu(hd720["t"][c]), syn.WORKSPACE_BOX)
        assert_same("ingest+T+C", fused.cpu().numpy(), cr)
        assert_same("stride3", ops.xyz_select(cu(np.ascontiguousarray(xyz[..., :3]))).cpu().numpy(), pts)
        t64 = oracle_mod.transform_f64(pts, hd720["R"][c].astype(np.float64), hd720["t"][c].astype(np.float64))
        g64 = ops.transform_f64(cu(pts), hd720["R"][c].astype(np.float64), hd720["t"][c].astype(np.float64)).cpu().numpy()
        assert g64.dtype == np.float64 and np.abs(g64 - t64).max() <= 1e-12


def test_crop_edge_cases(ops, golden):
    g = golden("edge_cases")
    got = ops.crop(cu(g["crop_in"]), (-0.25, 0.75), (-0.5, 1.75), (-0.05, 2)).cpu().numpy()
    assert_same("crop edge", got, g["crop_out"])
    assert tuple(ops.crop(torch.empty((0, 3), device=DEV), (0, 1), (0, 1), (0, 1)).shape) == (0, 3)


# ------------------------------------------------------------------------------------------- V1
@pytest.mark.parametrize("v", [0.005, 0.01, 0.002])
@pytest.mark.parametrize("flavour,recip", [("ieee", False), ("torch_cuda", True)])
def test_voxel_snap_unique_workspace(ops, oracle_mod, hd720, v, flavour, recip):
    from rt3d import synthetic as syn
    cx, cy, fx, fy = hd720["intr"]
    depth = hd720["depth"][0]
    pts = oracle_mod.crop(oracle_mod.transform(oracle_mod.backproject_indices(
        oracle_mod.mask_nonzero(np.ones(depth.shape, np.uint8)), depth, cx, cy, fx, fy, recip=recip),
        hd720["R"][0], hd720["t"][0]), *syn.WORKSPACE_BOX)
    want = oracle_mod.voxel_snap_unique(pts, v, recip=recip)
    got = ops.voxel_snap_unique(cu(pts), v, flavour=flavour).cpu().numpy()
    assert_same(f"voxel {v} {flavour}", got, want)


def test_voxel_snap_unique_edges_and_sizes(ops, oracle_mod, golden):
    g = golden("edge_cases")
    for v, tag in ((0.005, "5mm"), (0.01, "10mm"), (0.002, "2mm")):
        assert_same(f"snap {tag}", ops.voxel_snap_unique(cu(g["snap_in"]), v).cpu().numpy(), g[f"snap_{tag}"])
    assert tuple(ops.voxel_snap_unique(torch.empty((0, 3), device=DEV), 0.005).shape) == (0, 3)
    one = np.array([[0.1234, -5.0, 19.9]], np.float32)
    assert_same("single", ops.voxel_snap_unique(cu(one), 0.005).cpu().numpy(), oracle_mod.voxel_snap_unique(one, 0.005))
    rng = np.random.default_rng(11)
    for n in (2, 31, 4095, 4096, 4097, 8191, 100_003, 1_000_000):
        # wide, uncropped range (object clouds are not cropped): +-20 m
        pts = (rng.normal(size=(n, 3)) * np.array([4.0, 1.0, 6.0])).astype(np.float32)
        pts[rng.random(n) < 0.3] = pts[0]  # heavy duplication
        want = oracle_mod.voxel_snap_unique(pts, 0.005)
        got, keys, layout = ops.voxel_snap_unique(cu(pts), 0.005, return_keys=True)
        assert_same(f"random n={n}", got.cpu().numpy(), want)
        k = keys.cpu().numpy()
        assert (np.diff(k) > 0).all()      # packed keys ascend with the rows
    # idempotence: snapping lattice points again returns them unchanged
    again = ops.voxel_snap_unique(got, 0.005)
    assert_same("idempotent", again.cpu().numpy(), got.cpu().numpy())


def test_voxel_snap_unique_rejects_non_finite(ops):
    from rt3d._lib import Rt3dError
    bad = torch.tensor([[0.0, 1.0, float("nan")], [1.0, 2.0, 3.0]], device=DEV)
    with pytest.raises(Rt3dError, match="out of range"):
        ops.voxel_snap_unique(bad, 0.005)


def test_keys_sort_unique_and_decode(ops):
    rng = np.random.default_rng(5)
    for n, bits in ((1, 64), (17, 12), (4096, 64), (4097, 33), (250_000, 26), (1_000_003, 40)):
        hi = (1 << bits) - 1
        k = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True)
        k[rng.random(n) < 0.2] = k[0]
        got = ops.keys_sort_unique(cu(k.view(np.int64)), bits).cpu().numpy().view(np.uint64)
        assert_same(f"keys n={n} bits={bits}", got, np.unique(k))


# ------------------------------------------------------------------------------------------- V2
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_voxel_centroid(ops, oracle_mod, hd720, dtype):
    cx, cy, fx, fy = hd720["intr"]
    pts32 = oracle_mod.backproject_indices(oracle_mod.mask_nonzero(hd720["masks"][0, 0]), hd720["depth"][0], cx, cy, fx, fy)
    pts = oracle_mod.transform_f64(pts32, hd720["R"][0].astype(np.float64), hd720["t"][0].astype(np.float64)).astype(dtype)
    for v in (0.01, 0.005):
        keys, want, cnt = oracle_mod.voxel_centroid(pts, v)
        got, gcnt = ops.voxel_centroid(cu(pts), v, return_counts=True)
        got = got.cpu().numpy()
        assert got.dtype == dtype and got.shape == want.shape
        assert np.array_equal(gcnt.cpu().numpy(), cnt)
        assert np.abs(got - want).max() <= 1e-5          # north_star tolerance for centroids
        assert_same("centroid (deterministic input-order accumulation)", got, want)
        # every centroid lies in its voxel: floor(c / v) == key (up to fp32 mean rounding at the faces)
    assert tuple(ops.voxel_centroid(torch.empty((0, 3), dtype=torch.float64, device=DEV), 0.01).shape) == (0, 3)


# ------------------------------------------------------------------------------------------- O1
def _object_pair(oracle_mod, fr):
    cx, cy, fx, fy = fr["intr"]
    clouds = [oracle_mod.process_masks_cloud(fr["masks"][c, 0], fr["depth"][c], cx, cy, fx, fy, fr["R"][c], fr["t"][c]) for c in range(2)]
    return np.vstack(clouds)


def test_sor_object_pair_bit_exact(ops, oracle_mod, hd720):
    P = _object_pair(oracle_mod, hd720)
    for dtype in (np.float32, np.float64):
        keep, avg, stats = oracle_mod.sor_mask(P.astype(dtype), 20, 1.5)
        out, gkeep, gavg, gstats = ops.sor(cu(P.astype(dtype)), 20, 1.5, return_details=True)
        assert_same("avg", gavg.cpu().numpy(), avg)
        assert_same("stats", gstats.cpu().numpy(), stats)
        assert_same("mask", gkeep.cpu().numpy(), keep)
        assert_same("rows", out.cpu().numpy(), P.astype(np.float64)[keep.astype(bool)])
        assert out.dtype == torch.float64 and 0 < keep.sum() < len(keep)


@pytest.mark.parametrize("k,ratio", [(1, 1.0), (2, 1.0), (20, 1.5), (32, 0.5), (33, 2.0), (64, 1.5)])
def test_sor_random_clouds(ops, oracle_mod, k, ratio):
    rng = np.random.default_rng(k)
    for n in (1, 2, 5, 19, 20, 21, 500, 5000):
        P = rng.normal(size=(n, 3))
        P[rng.random(n) < 0.1] *= 25.0                      # isolated far points -> exhaustive-scan path
        if n > 50:
            P[10:20] = P[10]                                # exact duplicates
        keep, avg, stats = oracle_mod.sor_mask(P, k, ratio, brute=(n <= 500))
        out, gkeep, gavg, gstats = ops.sor(cu(P), k, ratio, return_details=True)
        assert_same(f"avg n={n} k={k}", gavg.cpu().numpy(), avg)
        assert_same(f"mask n={n} k={k}", gkeep.cpu().numpy(), keep)
        if n > 1:
            assert_same(f"stats n={n} k={k}", gstats.cpu().numpy(), stats)


def test_sor_known_answers_and_errors(ops):
    P = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0], [100, 0, 0]], np.float64)
    out, keep, avg, stats = ops.sor(cu(P), 2, 1.0, return_details=True)
    assert np.allclose(avg.cpu().numpy(), [0.5, 0.5, 0.5, 0.5, 48.5]) and keep.cpu().tolist() == [1, 1, 1, 1, 0]
    assert tuple(ops.sor(torch.empty((0, 3), dtype=torch.float64, device=DEV)).shape) == (0, 3)
    with pytest.raises(ValueError, match="Illegal input parameters"):
        ops.sor(cu(P), 0, 1.5)
    with pytest.raises(ValueError, match="Illegal input parameters"):
        ops.sor(cu(P), 20, 0.0)
    D = np.vstack([np.zeros((5, 3)), [[1.0, 0, 0], [1.1, 0, 0], [1.2, 0, 0]]])
    _, keep, avg, _ = ops.sor(cu(D), 3, 5.0, return_details=True)
    assert (avg.cpu().numpy()[:5] == 0).all() and keep.cpu().numpy()[:5].sum() == 0


def test_sor_workspace_scale(ops, oracle_mod, hd720):
    """the mask_cpu flavour runs SOR on the fused workspace (2cam/2cams_mask_cpu.py:530): ~160 k points"""
    from rt3d import synthetic as syn
    cx, cy, fx, fy = hd720["intr"]
    ws = []
    for c in range(2):
        d = hd720["depth"][c]
        p = oracle_mod.backproject_indices(oracle_mod.mask_nonzero(np.ones(d.shape, np.uint8)), d, cx, cy, fx, fy)
        ws.append(oracle_mod.voxel_snap_unique(oracle_mod.crop(oracle_mod.transform(p, hd720["R"][c], hd720["t"][c]), *syn.WORKSPACE_BOX), 0.01))
    W = np.vstack(ws)
    keep, avg, stats = oracle_mod.sor_mask(W, 20, 1.5)
    out, gkeep, gavg, gstats = ops.sor(cu(W), 20, 1.5, return_details=True)
    assert_same("avg", gavg.cpu().numpy(), avg)
    assert_same("stats", gstats.cpu().numpy(), stats)
    assert_same("mask", gkeep.cpu().numpy(), keep)


# ------------------------------------------------------------------------------------- centroid
def test_centroid_numpy_order(ops):
    rng = np.random.default_rng(2)
    for n in (1, 7, 8, 9, 1000, 4001):
        a = (rng.normal(size=(n, 3)) + 1).astype(np.float32)
        assert_same(f"c32 n={n}", ops.centroid(cu(a)).cpu().numpy(), np.mean(a, axis=0))
        a64 = a.astype(np.float64) * 1.000000123
        assert_same(f"c64 n={n}", ops.centroid(cu(a64)).cpu().numpy(), np.mean(a64, axis=0))


# ------------------------------------------------------------------------------------------- S1
@pytest.mark.parametrize("name", ["qvga_1obj", "qvga_4obj"])
@pytest.mark.parametrize("thr,tag", [(0.3, "thr030"), (0.06, "thr006"), (0.005, "thr0005"), (0.0, None), (10.0, None)])
def test_subtract_vs_oracle_and_reference(ops, oracle_mod, golden, name, thr, tag):
    g = golden(name)
    ws, objs = g["ws_fused"], g["objects_concat"]
    want = oracle_mod.subtract_mask(ws, objs, thr)
    out, keep = ops.subtract(cu(ws), cu(objs, torch.float32), thr, return_mask=True)
    assert_same(f"keep {thr}", keep.cpu().numpy(), want)
    assert_same(f"rows {thr}", out.cpu().numpy(), ws[want.astype(bool)])
    if tag:  # the reference's own (cdist) result may differ only inside the near-tie band
        ref_rows = {r.tobytes() for r in g[f"sub_{tag}_kept"]}
        ref_keep = np.array([r.tobytes() in ref_rows for r in ws])
        near = g[f"sub_{tag}_near"].astype(bool)
        assert not ((keep.cpu().numpy().astype(bool) != ref_keep) & ~near).any()


def test_subtract_hd720_and_empty(ops, oracle_mod, hd720):
    from rt3d import synthetic as syn
    cx, cy, fx, fy = hd720["intr"]
    d = hd720["depth"][0]
    p = oracle_mod.backproject_indices(oracle_mod.mask_nonzero(np.ones(d.shape, np.uint8)), d, cx, cy, fx, fy)
    ws = oracle_mod.voxel_snap_unique(oracle_mod.crop(oracle_mod.transform(p, hd720["R"][0], hd720["t"][0]), *syn.WORKSPACE_BOX), 0.005)
    objs = _object_pair(oracle_mod, hd720)
    for thr in (0.3, 0.06, 0.005):
        want = oracle_mod.subtract_mask(ws, objs, thr)
        out, keep = ops.subtract(cu(ws), cu(objs), thr, return_mask=True)
        assert_same(f"hd720 keep {thr}", keep.cpu().numpy(), want)
    with pytest.raises(IndexError):
        ops.subtract(cu(ws), torch.empty((0, 3), device=DEV), 0.3)
    assert tuple(ops.subtract(torch.empty((0, 3), device=DEV), cu(objs), 0.3).shape) == (0, 3)
    # a single object point, and objects spread over a huge range
    for objs2 in (objs[:1], np.vstack([objs, objs + 50.0, objs - 30.0])):
        want = oracle_mod.subtract_mask(ws, objs2, 0.06)
        assert_same("odd objects", ops.subtract(cu(ws), cu(objs2), 0.06, return_mask=True)[1].cpu().numpy(), want)


# ------------------------------------------------------------------------------ mask pre-processing
def test_box_dilate_and_erode(ops):
    import cv2
    import torch.nn.functional as F
    rng = np.random.default_rng(9)
    m = (rng.random((97, 130)) < 0.02).astype(np.float32)
    for k in (1, 3, 5, 12):
        t = torch.tensor(m)
        want = (F.conv2d(t[None, None], torch.ones(1, 1, k, k), padding=k // 2).squeeze() > 0).float().numpy()
        assert_same(f"dilate k={k}", ops.box_dilate(cu(m), k).cpu().numpy(), want)
    blob = (rng.random((97, 130)) < 0.97).astype(np.uint8)
    for k, it in ((12, 1), (3, 1), (3, 2), (12, 3)):
        want = cv2.erode(blob, np.ones((k, k), np.uint8), iterations=it)
        assert_same(f"erode k={k} it={it}", ops.box_erode(cu(blob), k, it).cpu().numpy(), want)
