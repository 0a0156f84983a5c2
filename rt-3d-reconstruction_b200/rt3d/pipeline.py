"""Batched, device-resident frame front-end over rt3d_frame_batch_run.

One `FramePipeline.run()` = the geometric part of F iterations of the reference's main loop
(2cam/2cams.py:165-208: process_point_clouds -> process_masks -> fuse_point_clouds ->
subtract_point_clouds) for C cameras each: ~30 kernel launches per batch, no host round trip,
CUDA-graph capturable.  `run_host()` is the same call fed from pinned host buffers with the copies
on the stream -- the end-to-end path a capture thread would use.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .ops import arith_flags


class FrameConfig(C.Structure):
    """Mirror of rt3d_frame_config (include/rt3d.h)."""
    _fields_ = [
        ("F", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("n_obj", C.c_int32),
        ("intr", (C.c_double * 4) * 8), ("Rt", (C.c_float * 12) * 8), ("box", C.c_float * 6),
        ("voxel", C.c_double), ("flags", C.c_int32), ("max_range", C.c_double),
        ("sor_k", C.c_int32), ("sor_std_ratio", C.c_double), ("sub_thr", C.c_float),
        ("ws_capacity", C.c_int64), ("obj_capacity", C.c_int64),
    ]


def _bind():
    return _lib.lib()


class FramePipeline:
    """F frames x C cameras per call.  All buffers are allocated once; outputs are device tensors
    (compacted rows + per-frame offsets) valid until the next run().

    intr: (cx, cy, fx, fy) shared by all cameras, or one tuple per camera.
    R, t: per-camera (C,3,3) / (C,3) camera -> robot transforms.
    box:  ((xlo,xhi),(ylo,yhi),(zlo,zhi)) workspace crop (2cam/vision_pipeline_utils.py:241-243).
    """

    def __init__(self, F, C_, H, W, n_obj, intr, R, t, box, voxel=0.005, flavour=None, nb_neighbors=20, std_ratio=1.5,
                 distance_threshold=0.3, max_range=20.0, ws_capacity=None, obj_capacity=None, device="cuda:0"):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("FramePipeline needs a CUDA device; rt3d has no CPU path")
        self.F, self.C, self.H, self.W, self.n_obj = int(F), int(C_), int(H), int(W), int(n_obj)
        cfg = FrameConfig()
        cfg.F, cfg.C, cfg.H, cfg.W, cfg.n_obj = self.F, self.C, self.H, self.W, self.n_obj
        intr = np.asarray(intr, dtype=np.float64)
        if intr.ndim == 1:
            intr = np.tile(intr, (self.C, 1))
        R = np.asarray(R.detach().cpu() if isinstance(R, torch.Tensor) else R, dtype=np.float32).reshape(self.C, 9)
        t = np.asarray(t.detach().cpu() if isinstance(t, torch.Tensor) else t, dtype=np.float32).reshape(self.C, 3)
        for c in range(self.C):
            for k in range(4):
                cfg.intr[c][k] = float(intr[c, k])
            for k in range(9):
                cfg.Rt[c][k] = float(R[c, k])
            for k in range(3):
                cfg.Rt[c][9 + k] = float(t[c, k])
        flat_box = [float(b) for axis in box for b in axis]
        for k in range(6):
            cfg.box[k] = flat_box[k]
        cfg.voxel, cfg.flags, cfg.max_range = float(voxel), arith_flags(flavour), float(max_range)
        cfg.sor_k, cfg.sor_std_ratio, cfg.sub_thr = int(nb_neighbors), float(std_ratio), float(np.float32(distance_threshold))
        npx = self.H * self.W
        cfg.ws_capacity = int(ws_capacity or self.F * self.C * npx // 2)
        cfg.obj_capacity = int(obj_capacity or max(4096, self.F * self.C * max(self.n_obj, 1) * npx // 16))
        self.cfg = cfg
        L = _bind()
        nbytes = L.rt3d_frame_batch_scratch_bytes(C.byref(cfg))
        if nbytes == 0:
            raise ValueError(f"rt3d_frame_batch: {_lib.last_error()}")
        self.stream = torch.cuda.current_stream(self.dev)
        self.scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        _lib.check(L.rt3d_frame_batch_init(C.byref(cfg), self.scratch.data_ptr(), nbytes, self.stream.cuda_stream),
                   "rt3d_frame_batch_init")
        nimg, npairs = self.F * self.C, self.F * self.n_obj
        d = self.dev
        self.ws_xyz = torch.empty((cfg.ws_capacity, 3), dtype=torch.float32, device=d)
        self.ws_off = torch.zeros(nimg + 1, dtype=torch.int32, device=d)
        self.obj_xyz = torch.empty((cfg.obj_capacity, 3), dtype=torch.float64, device=d)
        self.obj_off = torch.zeros(npairs + 1, dtype=torch.int32, device=d)
        self.kept_xyz = torch.empty((cfg.ws_capacity, 3), dtype=torch.float32, device=d)
        self.kept_off = torch.zeros(self.F + 1, dtype=torch.int32, device=d)
        self.counts = torch.zeros(6, dtype=torch.int32, device=d)
        # staging for run_host(): device copies of the inputs + pinned result mirrors
        self._d_depth = self._d_masks = None
        self._h = None
        self.launches_per_run = None

    # kernel launches of one run (for bench.py's gpu_launches claim): fixed by the configuration
    def count_launches(self) -> int:
        def passes(bits):
            return (bits + 7) // 8
        import math
        n = 2 + 5 + 4  # project, clamp; extract = count, slab scan, image scan, emit, bitmap clear; keep scan, write, per-frame counts, offsets scan
        if self.n_obj == 0:
            n += 1     # no flag kernel to count the kept rows per tile on the way: a counting sweep
        if self.n_obj > 0:
            kmax = math.ceil(self.cfg.max_range / self.cfg.voxel) + 2
            obits = max(1, math.ceil(math.log2(2 * kmax)))
            seg_bits = max(1, math.ceil(math.log2(max(2, self.F * self.n_obj * self.C))))
            pair_bits = max(1, math.ceil(math.log2(max(2, self.F * self.n_obj))))
            frame_bits = max(1, math.ceil(math.log2(max(2, self.F))))
            n += 1 + passes(3 * obits + seg_bits) + 2            # hist + passes, unique select, segment offsets
            n += 2 + passes(21 + pair_bits) + 6                  # prepare, cell keys, passes, gather, column table, knn, stats, select, offsets
            n += 2 + passes(12 + frame_bits) + 3                 # prepare, cell keys, passes, cells, super cells, flags
        return n

    STAGES = ("project", "extract", "objects", "fuse", "subtract")

    def enable_stage_timing(self, slots: int = 1):
        """Record CUDA events between the stages of every run(); `slots` independent event sets let a
        timed loop record every step without a host synchronise in between (read them afterwards)."""
        self._timers = getattr(self, "_timers", [])
        while len(self._timers) < slots:
            t = _bind().rt3d_stage_timer_create()
            if not t:
                raise _lib.Rt3dError(_lib.last_error())
            self._timers.append(t)
        return self

    def stage_ms(self, slot: int = 0):
        buf = (C.c_float * 5)()
        _lib.check(_bind().rt3d_stage_timer_read(self._timers[slot], buf), "rt3d_stage_timer_read")
        return dict(zip(self.STAGES, (float(x) for x in buf)))

    def __del__(self):
        for t in getattr(self, "_timers", []):
            try:
                _bind().rt3d_stage_timer_destroy(t)
            except Exception:
                pass

    def run(self, depth: torch.Tensor, masks: torch.Tensor | None, slot: int | None = None):
        """depth (F,C,H,W) fp32 and masks (F,C,n_obj,H,W) uint8, both on the device.  Asynchronous."""
        if not depth.is_cuda or depth.dtype != torch.float32 or not depth.is_contiguous():
            raise ValueError("depth must be a contiguous fp32 CUDA tensor")
        if depth.numel() != self.F * self.C * self.H * self.W:
            raise ValueError("depth has the wrong number of elements for this pipeline")
        mptr = None
        if self.n_obj > 0:
            if masks is None or not masks.is_cuda or masks.dtype not in (torch.uint8, torch.bool) or not masks.is_contiguous():
                raise ValueError("masks must be a contiguous uint8 CUDA tensor")
            if masks.numel() != self.F * self.C * self.n_obj * self.H * self.W:
                raise ValueError("masks has the wrong number of elements for this pipeline")
            mptr = masks.data_ptr()
        L = _bind()
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        _lib.check(L.rt3d_frame_batch_run_timed(C.byref(self.cfg), depth.data_ptr(), mptr, self.ws_xyz.data_ptr(),
                                                self.ws_off.data_ptr(), self.obj_xyz.data_ptr(), self.obj_off.data_ptr(),
                                                self.kept_xyz.data_ptr(), self.kept_off.data_ptr(), self.counts.data_ptr(),
                                                self.scratch.data_ptr(), self.scratch.numel(), stream,
                                                self._timers[slot] if slot is not None else None), "rt3d_frame_batch_run")
        return self

    def _check(self, counts):
        if counts[0]:
            raise _lib.Rt3dError(f"frame batch: device status {counts[0]}: {_lib.decode_status(counts[0])}")

    def results(self):
        """Synchronise and slice the outputs per frame: list of dicts(workspace, objects[list], kept)."""
        counts = self.counts.tolist()
        self._check(counts)
        ws_off, obj_off, kept_off = self.ws_off.tolist(), self.obj_off.tolist(), self.kept_off.tolist()
        out = []
        for f in range(self.F):
            out.append(dict(
                workspace=self.ws_xyz[ws_off[f * self.C]:ws_off[(f + 1) * self.C]],
                workspace_per_cam=[self.ws_xyz[ws_off[f * self.C + c]:ws_off[f * self.C + c + 1]] for c in range(self.C)],
                objects=[self.obj_xyz[obj_off[f * self.n_obj + o]:obj_off[f * self.n_obj + o + 1]] for o in range(self.n_obj)],
                kept=self.kept_xyz[kept_off[f]:kept_off[f + 1]],
            ))
        return out

    # ---------------------------------------------------------------- host-fed path (end-to-end)
    def _ensure_staging(self):
        if self._d_depth is None:
            self._d_depth = torch.empty((self.F, self.C, self.H, self.W), dtype=torch.float32, device=self.dev)
            self._d_masks = (torch.empty((self.F, self.C, self.n_obj, self.H, self.W), dtype=torch.uint8, device=self.dev)
                             if self.n_obj else None)
            self._h = dict(
                counts=torch.empty(6, dtype=torch.int32).pin_memory(),
                kept_off=torch.empty(self.F + 1, dtype=torch.int32).pin_memory(),
                obj_off=torch.empty(self.F * self.n_obj + 1, dtype=torch.int32).pin_memory(),
                kept=torch.empty((self.cfg.ws_capacity, 3), dtype=torch.float32).pin_memory(),
                objs=torch.empty((self.cfg.obj_capacity, 3), dtype=torch.float64).pin_memory(),
            )

    def run_host_pipelined(self, batches, expect_kept=None, expect_objs=None):
        """End-to-end over a SEQUENCE of host batches [(depth_host, masks_host), ...] (pinned): the H2D copy of
        batch i+1 runs on a copy stream into the second input buffer while batch i computes, so a capture loop
        is bound by the slower of PCIe and the kernels instead of their sum.  Every batch's results are copied
        back to the pinned mirrors (overwritten batch after batch); returns (bytes_h2d, bytes_d2h) per batch.
        Asynchronous: synchronise the device (or the current stream) before reading the mirrors."""
        self._ensure_staging()
        if getattr(self, "_in2", None) is None:
            self._in2 = [(self._d_depth, self._d_masks),
                         (torch.empty_like(self._d_depth), torch.empty_like(self._d_masks) if self.n_obj else None)]
            self._copy_stream = torch.cuda.Stream(self.dev)
            self._ev_in = [torch.cuda.Event(), torch.cuda.Event()]
            self._ev_free = [torch.cuda.Event(), torch.cuda.Event()]
        main = torch.cuda.current_stream(self.dev)
        cs = self._copy_stream
        cs.wait_stream(main)     # whatever used the input buffers before is done
        h = self._h
        nk = int(expect_kept if expect_kept is not None else self.cfg.ws_capacity)
        no = int(expect_objs if expect_objs is not None else self.cfg.obj_capacity)
        h2d = d2h = 0
        for i, (depth_host, masks_host) in enumerate(batches):
            b = i & 1
            d_depth, d_masks = self._in2[b]
            with torch.cuda.stream(cs):
                if i >= 2:
                    cs.wait_event(self._ev_free[b])
                d_depth.copy_(depth_host.view_as(d_depth), non_blocking=True)
                h2d = depth_host.numel() * 4
                if self.n_obj:
                    d_masks.copy_(masks_host.view_as(d_masks), non_blocking=True)
                    h2d += masks_host.numel()
                self._ev_in[b].record(cs)
            main.wait_event(self._ev_in[b])
            self.run(d_depth, d_masks)
            self._ev_free[b].record(main)
            h["counts"].copy_(self.counts, non_blocking=True)
            h["kept_off"].copy_(self.kept_off, non_blocking=True)
            h["obj_off"].copy_(self.obj_off, non_blocking=True)
            h["kept"][:nk].copy_(self.kept_xyz[:nk], non_blocking=True)
            h["objs"][:no].copy_(self.obj_xyz[:no], non_blocking=True)
            d2h = nk * 12 + no * 24 + (6 + self.F + 1 + self.F * self.n_obj + 1) * 4
        return h2d, d2h

    def run_host(self, depth_host: torch.Tensor, masks_host: torch.Tensor | None, expect_kept=None, expect_objs=None):
        """Pinned host inputs -> H2D -> run -> D2H of the per-frame results (subtracted workspace, fused
        objects, offsets, counts).  The result rows are copied with sizes known from the previous batch
        (`expect_*`, rows) so that the whole call is one stream of work with a single final synchronise;
        returns (bytes_h2d, bytes_d2h)."""
        self._ensure_staging()
        self._d_depth.copy_(depth_host.view_as(self._d_depth), non_blocking=True)
        h2d = depth_host.numel() * 4
        if self.n_obj:
            self._d_masks.copy_(masks_host.view_as(self._d_masks), non_blocking=True)
            h2d += masks_host.numel()
        self.run(self._d_depth, self._d_masks)
        h = self._h
        h["counts"].copy_(self.counts, non_blocking=True)
        h["kept_off"].copy_(self.kept_off, non_blocking=True)
        h["obj_off"].copy_(self.obj_off, non_blocking=True)
        nk = int(expect_kept if expect_kept is not None else self.cfg.ws_capacity)
        no = int(expect_objs if expect_objs is not None else self.cfg.obj_capacity)
        h["kept"][:nk].copy_(self.kept_xyz[:nk], non_blocking=True)
        h["objs"][:no].copy_(self.obj_xyz[:no], non_blocking=True)
        d2h = nk * 12 + no * 24 + (6 + self.F + 1 + self.F * self.n_obj + 1) * 4
        return h2d, d2h
