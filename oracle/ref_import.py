"""Import the UNMODIFIED reference module for pinning the oracle.  TEST INFRASTRUCTURE ONLY.

/root/reference exists only in the build container (never on the GPU box), so
everything here is used by (a) tests/golden/make_golden.py, which writes the
committed fixtures, and (b) ``-m "not gpu"`` tests that skip when the reference
is absent.  Nothing is copied from the reference: its source file is executed in
place with stub modules standing in for the dependencies this image lacks
(``pyzed``, ``open3d``, ``ultralytics``).

The ``open3d`` stub implements just the two legacy calls the reference makes in
``filter_outliers_sor`` (2cam/vision_pipeline_utils.py:50-57) on top of the C
oracle's restatement of ``remove_statistical_outlier``; that lets the reference's
own ``fuse_point_clouds_centroid`` control flow (:88-165) run here unmodified.
The reference's ``subtract_point_clouds_gpu`` hard-codes ``device='cuda'``
(:170-171); ``cpu_torch=True`` swaps the module's ``torch`` global for a proxy
that maps that device to the CPU so the function can run in this container.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("RT3D_REFERENCE_ROOT", "/root/reference")
_UTILS = os.path.join(REFERENCE_ROOT, "2cam", "vision_pipeline_utils.py")


def available() -> bool:
    return os.path.exists(_UTILS)


class _Vector3dVector:
    def __init__(self, arr):
        self.arr = np.asarray(arr, dtype=np.float64).reshape(-1, 3)

    def __array__(self, dtype=None, copy=None):
        return self.arr if dtype is None else self.arr.astype(dtype)

    def __len__(self):
        return self.arr.shape[0]


class _LegacyPointCloud:
    """Stand-in for o3d.geometry.PointCloud: holds float64 points, SOR via the C oracle."""

    def __init__(self):
        self._points = _Vector3dVector(np.empty((0, 3)))

    @property
    def points(self):
        return self._points

    @points.setter
    def points(self, v):
        self._points = v if isinstance(v, _Vector3dVector) else _Vector3dVector(v)

    def remove_statistical_outlier(self, nb_neighbors, std_ratio):
        from . import oracle
        keep, _, _ = oracle.sor_mask(self._points.arr, nb_neighbors, std_ratio)
        idx = np.nonzero(keep)[0]
        out = _LegacyPointCloud()
        out.points = _Vector3dVector(self._points.arr[idx])
        return out, idx.tolist()


def _stub_modules():
    stubs = {}
    for name in ("pyzed", "pyzed.sl", "open3d", "open3d.core", "ultralytics"):
        stubs[name] = types.ModuleType(name)
    stubs["pyzed"].sl = stubs["pyzed.sl"]
    stubs["ultralytics"].YOLO = type("YOLO", (), {})
    o3d = stubs["open3d"]
    o3d.core = stubs["open3d.core"]
    o3d.geometry = types.SimpleNamespace(PointCloud=_LegacyPointCloud)
    o3d.utility = types.SimpleNamespace(Vector3dVector=_Vector3dVector)
    return stubs


class _CpuTorch:
    """Proxy for the reference module's ``torch`` global: device='cuda' -> 'cpu' when there is no GPU."""

    def __init__(self, torch):
        self._torch = torch

    def __getattr__(self, name):
        return getattr(self._torch, name)

    def tensor(self, data, *a, **kw):
        if kw.get("device") == "cuda" and not self._torch.cuda.is_available():
            kw["device"] = "cpu"
        return self._torch.tensor(data, *a, **kw)


def load_reference(cpu_torch: bool = False, path: str = _UTILS, name: str = "_rt3d_reference_vpu"):
    """Execute the reference's vision_pipeline_utils.py in place and return the module."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"reference not present at {path} (it only exists in the build container)")
    saved = {}
    stubs = _stub_modules()
    for k, m in stubs.items():
        saved[k] = sys.modules.get(k)
        if saved[k] is None:
            sys.modules[k] = m
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, m in saved.items():
            if m is None:
                sys.modules.pop(k, None)
    if cpu_torch:
        import torch
        mod.torch = _CpuTorch(torch)
    return mod
