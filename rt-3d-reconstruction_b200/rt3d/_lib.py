"""ctypes binding of librt3d.so (the C ABI declared in include/rt3d.h).

There is deliberately no fallback: if the CUDA library is missing or cannot be loaded the
import of any compute function raises -- the product never routes through NumPy, torch eager
ops or the test oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("RT3D_LIB_PATH") or os.path.join(_PKG_DIR, "librt3d.so")   # override: kernel experiments only
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")

# enums of include/rt3d.h
RT3D_OK, RT3D_EINVAL, RT3D_ECUDA, RT3D_ESCRATCH, RT3D_ERANGE = 0, -1, -2, -3, -4
RT3D_U8, RT3D_F32, RT3D_F64 = 0, 1, 2
RT3D_DIV_IEEE, RT3D_DIV_RECIP = 0, 1
RT3D_DEV_KEY_RANGE, RT3D_DEV_LOOKBACK, RT3D_DEV_CAPACITY = 1, 2, 4
OP_SELECT, OP_VOXEL_UNIQUE, OP_VOXEL_CENTROID, OP_SOR, OP_SUBTRACT, OP_KEYS_UNIQUE = range(6)

_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); kept in step with include/rt3d.h (tests/test_abi.py checks both ways)
SIGNATURES = {
    "rt3d_abi_version": (_i, []),
    "rt3d_last_error": (C.c_char_p, []),
    "rt3d_scratch_bytes": (_sz, [_i, _i64, _i64]),
    "rt3d_box_dilate": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "rt3d_box_erode": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "rt3d_mask_nonzero": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "rt3d_backproject_indices": (_i, [_vp, _i64, _vp, _i, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "rt3d_backproject_mask": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "rt3d_xyz_select": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rt3d_transform": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "rt3d_transform_f64": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "rt3d_voxel_snap_unique": (_i, [_vp, _i64, _vp, _d, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rt3d_voxel_centroid": (_i, [_vp, _i, _i64, _d, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rt3d_sor": (_i, [_vp, _i, _i64, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rt3d_centroid": (_i, [_vp, _i, _i64, _vp, _vp]),
    "rt3d_subtract": (_i, [_vp, _i64, _vp, _i64, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rt3d_subtract_radius": (_i, [_vp, _i64, _vp, _i64, _d, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rt3d_keys_sort_unique": (_i, [_vp, _i64, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "rt3d_keys_rebase": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rt3d_keys_decode": (_i, [_vp, _i64, _vp, _vp, _f, _vp, _vp]),
    # batched frame front-end; the first argument is a `const rt3d_frame_config *` (rt3d.pipeline.FrameConfig)
    "rt3d_frame_batch_slots": (_i, [_vp]),
    "rt3d_frame_batch_scratch_bytes": (_sz, [_vp]),
    "rt3d_frame_batch_init": (_i, [_vp, _vp, _sz, _vp]),
    "rt3d_frame_batch_run": (_i, [_vp] * 13 + [_sz, _vp]),
    "rt3d_frame_batch_run_timed": (_i, [_vp] * 13 + [_sz, _vp, _vp]),
    "rt3d_stage_timer_create": (_vp, []),
    "rt3d_stage_timer_destroy": (None, [_vp]),
    "rt3d_stage_timer_read": (_i, [_vp, _vp]),
}

_lib = None


class Rt3dError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into librt3d.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise Rt3dError("building librt3d.so failed (see output above)")
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Rt3dError(f"{LIB_PATH} is missing: build it with `make -C {CSRC_DIR}` (or __graft_entry__.build()); "
                            "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.rt3d_abi_version() != 2:
            raise Rt3dError("librt3d.so ABI version mismatch; rebuild it")
        _lib = L
    return _lib


def last_error() -> str:
    return lib().rt3d_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc == RT3D_OK:
        return
    msg = last_error()
    if rc == RT3D_EINVAL:
        raise ValueError(f"{what}: {msg}")
    raise Rt3dError(f"{what}: error {rc}: {msg}")


def decode_status(bits: int) -> str:
    names = []
    if bits & RT3D_DEV_KEY_RANGE:
        names.append("voxel/grid index out of range (non-finite or absurdly large coordinates, or index out of bounds)")
    if bits & RT3D_DEV_LOOKBACK:
        names.append("decoupled look-back timed out")
    if bits & RT3D_DEV_CAPACITY:
        names.append("buffer capacity exceeded")
    return "; ".join(names) or "ok"
