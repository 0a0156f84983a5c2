"""Stand-ins for the hardware / model SDKs that sit either side of the geometric path.

The ZED SDK (`pyzed.sl`, 2cam/2cams.py:53-96,140-150) and ultralytics YOLO are out of scope and
absent from this image; the drop-in functions only touch a handful of their attributes, so tests
and the bench feed them synthetic frames through these duck-typed fakes.
"""
from __future__ import annotations

import enum
import types

import numpy as np

from . import synthetic


class _View(enum.Enum):
    LEFT = 0


class _Measure(enum.Enum):
    DEPTH = 1
    XYZ = 2


class _ErrorCode(enum.Enum):
    SUCCESS = 0
    FAILURE = 1


class Resolution:
    def __init__(self, width=0, height=0):
        self.width, self.height = int(width), int(height)


class FakeMat:
    """sl.Mat: a host buffer filled by retrieve_*; get_data() returns the numpy array."""

    def __init__(self, *_, **__):
        self._data = None

    def get_data(self):
        return self._data


# what `import pyzed.sl as sl` provides, as far as vision_pipeline_utils uses it
sl_shim = types.SimpleNamespace(VIEW=_View, MEASURE=_Measure, ERROR_CODE=_ErrorCode, Mat=FakeMat, Resolution=Resolution)


class FakeZed:
    """One synthetic camera: serves a fixed depth map, its XYZ image and a grey BGRA frame."""

    def __init__(self, depth: np.ndarray):
        self.depth = np.ascontiguousarray(depth, dtype=np.float32)

    def grab(self):
        return _ErrorCode.SUCCESS

    def retrieve_image(self, mat, view=None):
        h, w = self.depth.shape
        mat._data = np.full((h, w, 4), 127, np.uint8)
        return _ErrorCode.SUCCESS

    def retrieve_measure(self, mat, measure=None, resolution=None):
        name = getattr(measure, "name", str(measure))
        if name == "DEPTH":
            mat._data = self.depth
        elif name == "XYZ":
            d = self.depth
            if resolution is not None and (resolution.height, resolution.width) != d.shape and resolution.width > 0:
                ys = (np.arange(resolution.height) * d.shape[0]) // resolution.height
                xs = (np.arange(resolution.width) * d.shape[1]) // resolution.width
                full = synthetic.xyz_image(d)
                mat._data = np.ascontiguousarray(full[ys][:, xs])
            else:
                mat._data = synthetic.xyz_image(d)
        else:
            raise ValueError(f"FakeZed cannot serve measure {measure!r}")
        return _ErrorCode.SUCCESS


def yolo_masks(masks_device):
    """ultralytics `Masks` stand-in: an object with a ``.data`` (n,H,W) tensor (vision_pipeline_utils.py:295)."""
    return types.SimpleNamespace(data=masks_device)
