"""CPU: librt3d.so loads without a GPU and exports exactly what include/rt3d.h declares."""
import ctypes
import os
import re

import pytest

from conftest import PKG, ROOT


@pytest.fixture(scope="module")
def built_lib():
    from rt3d import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "rt3d.h")).read()
    return sorted(set(re.findall(r"RT3D_API[^;(]*?\b(rt3d_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    names = _declared()
    assert len(names) >= 17
    L = ctypes.CDLL(built_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in rt3d.h but not exported: {missing}"


def test_binding_table_matches_header(built_lib):
    assert sorted(built_lib.SIGNATURES) == _declared()


def test_host_only_calls_work_without_gpu(built_lib):
    L = built_lib.lib()
    assert L.rt3d_abi_version() == 2
    sizes = [L.rt3d_scratch_bytes(op, 921600, 8000) for op in range(6)]
    assert all(s > 0 for s in sizes)
    assert L.rt3d_scratch_bytes(99, 1, 1) == 0 and b"unknown op" in L.rt3d_last_error()


def test_frame_batch_host_queries_without_gpu(built_lib):
    """slot count / scratch size / argument checks of the batched front-end are host arithmetic"""
    import ctypes as C

    from rt3d.pipeline import FrameConfig
    L = built_lib.lib()
    cfg = FrameConfig()
    cfg.F, cfg.C, cfg.H, cfg.W, cfg.n_obj = 4, 2, 180, 320, 3
    for c in range(2):
        for k, v in enumerate((160.8, 90.4, 131.9, 131.9)):
            cfg.intr[c][k] = v
        for k, v in enumerate((1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0)):
            cfg.Rt[c][k] = v
    for k, v in enumerate((-0.25, 0.75, -0.5, 1.75, -0.05, 2.0)):
        cfg.box[k] = v
    cfg.voxel, cfg.max_range, cfg.sor_k, cfg.sor_std_ratio, cfg.sub_thr = 0.005, 20.0, 20, 1.5, 0.3
    cfg.ws_capacity, cfg.obj_capacity = 100000, 50000
    assert L.rt3d_frame_batch_slots(C.byref(cfg)) == 3
    by_index = L.rt3d_frame_batch_scratch_bytes(C.byref(cfg))
    cfg.fuse_mode, cfg.fuse_thr, cfg.fused_capacity = 1, 0.3, 100000
    assert L.rt3d_frame_batch_slots(C.byref(cfg)) == 6
    assert L.rt3d_frame_batch_scratch_bytes(C.byref(cfg)) > by_index > 0
    cfg.C = 3      # fuse_point_clouds_centroid is a two-camera function
    assert L.rt3d_frame_batch_scratch_bytes(C.byref(cfg)) == 0 and b"exactly 2 cameras" in L.rt3d_last_error()
    cfg.C = 2
    # run() before init() is refused (cfg.plan == 0), without touching the device
    rc = L.rt3d_frame_batch_run(C.byref(cfg), *([None] * 12), 0, None)
    assert rc == built_lib.RT3D_EINVAL and b"rt3d_frame_batch_init" in L.rt3d_last_error()


def test_class_ranks_follow_the_reference_set_order():
    from rt3d.pipeline import class_ranks
    ids = [[[39, 41, 39, 41], [41, 41, 39, 39]], [[0, 67, 39, 39], [39, 47, 0, 8]]]
    ranks = class_ranks(ids)
    for f, (a, b) in enumerate(ids):
        d1, d2 = dict.fromkeys(a), dict.fromkeys(b)
        order = list(set(d1.keys()).union(d2.keys()))        # 2cam/vision_pipeline_utils.py:117
        assert [order[r] for r in ranks[f, 0]] == a and [order[r] for r in ranks[f, 1]] == b
        assert ranks[f].max() < 2 * len(a)


def test_argument_validation_is_loud_without_gpu(built_lib):
    L = built_lib.lib()
    # null pointers / bad sizes are rejected before anything touches the device
    assert L.rt3d_transform(None, 5, None, None, None) == built_lib.RT3D_EINVAL
    assert L.rt3d_sor(None, 1, 0, 0, 1.5, None, None, None, None, None, None, 0, None) == built_lib.RT3D_EINVAL
    assert b"bad arguments" in L.rt3d_last_error() or b"Illegal" in L.rt3d_last_error()


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle or fall back to torch/NumPy compute."""
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{f} touches the oracle"


def test_ops_refuse_cpu_tensors(built_lib):
    import torch
    from rt3d import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.voxel_snap_unique(torch.zeros(4, 3), 0.005)
