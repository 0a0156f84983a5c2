"""CPU: the property a lattice-offset neighbour search for the batched outlier filter would rest on.

The filter's inputs in the batched path are the voxel-lattice points of V1 (vision_pipeline_utils.py:40-44), both
cameras stacked (:123-126).  For such clouds the neighbours of a point can be enumerated by integer lattice offset
in shell order (dx^2 + dy^2 + dz^2 = 0, 1, 2, ...).  This test checks, against the oracle's k-nearest-neighbour
means (Open3D's arithmetic, O1), that
  * the fp64 distances of different shells never interleave (shells are separated by far more than the fp32
    rounding of the lattice coordinates), so the shell holding the k-th neighbour follows from counts alone, and
  * the mean over the k smallest exact distances of the shells up to that one is bit-identical to the oracle's,
and reports how many points the boundary set holds (the kernel would rank them with one value per lane).
"""
import numpy as np
import pytest


def lattice_offsets(max_m):
    r = int(np.floor(np.sqrt(max_m)))
    g = np.arange(-r, r + 1)
    d = np.array(np.meshgrid(g, g, g, indexing="ij")).reshape(3, -1).T
    m = (d * d).sum(1)
    keep = m <= max_m
    d, m = d[keep], m[keep]
    order = np.argsort(m, kind="stable")
    return d[order], m[order]


@pytest.mark.parametrize("voxel", [0.005, 0.01])
def test_shell_order_gives_the_oracle_knn_mean(oracle_mod, voxel):
    from rt3d import synthetic as syn
    fr = syn.frame("HD720", seed=77, n_objects=1, jitter=0.03)     # dense: 2 mm pixel footprint at 1.4 m
    cx, cy, fx, fy = fr["intr"]
    parts = [oracle_mod.process_masks_cloud(fr["masks"][c, 0], fr["depth"][c], cx, cy, fx, fy, fr["R"][c], fr["t"][c], voxel)
             for c in range(2)]
    P32 = np.vstack([p for p in parts if p is not None]).astype(np.float32)
    P = P32.astype(np.float64)
    n, k = P.shape[0], 20
    assert n > 200
    want = oracle_mod.knn_avg(P, k)
    ijk = np.rint(P / np.float64(np.float32(voxel))).astype(np.int64)
    assert np.array_equal((ijk.astype(np.float32) * np.float32(voxel)), P32)      # the points ARE lattice points
    table = {}
    for i, key in enumerate(map(tuple, ijk)):
        table.setdefault(key, []).append(i)
    assert max(len(v) for v in table.values()) <= 2                              # one point per voxel and camera
    offs, ms = lattice_offsets(9)
    assert len(offs) == 123
    rng = np.random.default_rng(5)
    resolved = 0
    boundary_sizes = []
    for i in rng.choice(n, size=min(n, 400), replace=False):
        found_m, found_d2 = [], []
        for d, m in zip(offs, ms):
            for j in table.get(tuple(ijk[i] + d), ()):
                dx, dy, dz = P[i] - P[j]
                found_m.append(m)
                found_d2.append((dx * dx + dy * dy) + dz * dz)                   # nanoflann's L2 accumulation order
        found_m, found_d2 = np.array(found_m), np.array(found_d2)
        shells = np.unique(found_m)
        for a, b in zip(shells[:-1], shells[1:]):                                # shells never interleave
            assert found_d2[found_m == a].max() < found_d2[found_m == b].min()
        if len(found_m) < k:
            continue                                                             # sparse spot: wider search / grid path
        cum = np.cumsum([np.sum(found_m == s) for s in shells])
        m_b = shells[np.searchsorted(cum, k)]
        # every point outside the enumerated shells is at least sqrt(10) voxels away: the k-th neighbour is inside
        sel = np.sort(found_d2[found_m <= m_b])
        boundary_sizes.append(len(sel))
        acc = 0.0
        for v in np.sqrt(sel[:k]):
            acc += v
        assert acc / k == want[i]
        resolved += 1
    assert resolved > 100
    print(f"\nvoxel {voxel}: {resolved} of 400 sampled queries resolved within 123 offsets; boundary set size "
          f"median {int(np.median(boundary_sizes))}, max {max(boundary_sizes)}, <= 32 for "
          f"{100 * np.mean(np.array(boundary_sizes) <= 32):.0f} %")


def test_wide_shells_stay_ordered_up_to_ten_voxels(oracle_mod):
    """A wider bit-plane search for the sparse queries (measured in round 2 and rejected on speed) would take
    every occupied voxel up to the k-th one's squared lattice distance within 10 voxels: the fp64 distances of different
    shells do not interleave up to r2 = 100 either, and the mean over the k smallest is the oracle's."""
    from rt3d import synthetic as syn
    voxel = 0.005
    fr = syn.frame("HD720", seed=78, n_objects=1, jitter=0.03)
    cx, cy, fx, fy = fr["intr"]
    parts = [oracle_mod.process_masks_cloud(fr["masks"][c, 0], fr["depth"][c], cx, cy, fx, fy, fr["R"][c], fr["t"][c], voxel)
             for c in range(2)]
    P32 = np.vstack([p for p in parts if p is not None]).astype(np.float32)
    P = P32.astype(np.float64)
    n, k = P.shape[0], 20
    want = oracle_mod.knn_avg(P, k)
    ijk = np.rint(P / np.float64(np.float32(voxel))).astype(np.int64)
    table = {}
    for i, key in enumerate(map(tuple, ijk)):
        table.setdefault(key, []).append(i)
    offs3, _ = lattice_offsets(9)
    offs, ms = lattice_offsets(100)
    assert len(offs) == 4169
    rng = np.random.default_rng(6)
    wide = 0
    for i in rng.permutation(n):
        if sum(len(table.get(tuple(ijk[i] + d), ())) for d in offs3) >= k:
            continue                                                             # settled by the 123-offset search
        found_m, found_d2 = [], []
        for d, m in zip(offs, ms):
            for j in table.get(tuple(ijk[i] + d), ()):
                dx, dy, dz = P[i] - P[j]
                found_m.append(m)
                found_d2.append((dx * dx + dy * dy) + dz * dz)
        found_m, found_d2 = np.array(found_m), np.array(found_d2)
        shells = np.unique(found_m)
        for a, b in zip(shells[:-1], shells[1:]):
            assert found_d2[found_m == a].max() < found_d2[found_m == b].min()
        if len(found_m) < k:
            continue                                                             # farther than 10 voxels: grid search
        r2k = np.sort(found_m)[k - 1]
        sel = np.sort(found_d2[found_m <= r2k])
        acc = 0.0
        for v in np.sqrt(sel[:k]):
            acc += v
        assert acc / k == want[i]
        wide += 1
        if wide >= 60:
            break
    assert wide >= 20
