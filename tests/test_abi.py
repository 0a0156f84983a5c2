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
    assert L.rt3d_abi_version() == 1
    sizes = [L.rt3d_scratch_bytes(op, 921600, 8000) for op in range(6)]
    assert all(s > 0 for s in sizes)
    assert L.rt3d_scratch_bytes(99, 1, 1) == 0 and b"unknown op" in L.rt3d_last_error()


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
