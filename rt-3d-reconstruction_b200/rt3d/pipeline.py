"""Batched, device-resident frame front-end over rt3d_frame_batch_run.

One `FramePipeline.run()` = the geometric part of F iterations of the reference's main loop
(2cam/2cams.py:165-208: process_point_clouds -> process_masks -> fuse_point_clouds ->
subtract_point_clouds) for C cameras each: a few dozen kernel launches per batch, no host round trip,
CUDA-graph capturable.  `run_host_pipelined()` is the same call fed from pinned host buffers with the
copies on their own streams -- the end-to-end path a capture thread would use.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .ops import arith_flags

FUSE_BY_INDEX, FUSE_CENTROID = 0, 1
CLOUD_EMPTY, CLOUD_PAIR, CLOUD_MATCHED, CLOUD_UNMATCHED, CLOUD_LEFTOVER = range(5)
CLOUD_KINDS = ("empty", "pair", "matched", "unmatched", "leftover")


class FrameConfig(C.Structure):
    """Mirror of rt3d_frame_config (include/rt3d.h)."""
    _fields_ = [
        ("F", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("n_obj", C.c_int32),
        ("intr", (C.c_double * 4) * 8), ("Rt", (C.c_float * 12) * 8), ("box", C.c_float * 6),
        ("voxel", C.c_double), ("flags", C.c_int32), ("max_range", C.c_double),
        ("sor_k", C.c_int32), ("sor_std_ratio", C.c_double), ("sub_thr", C.c_float),
        ("ws_capacity", C.c_int64), ("obj_capacity", C.c_int64),
        ("fuse_mode", C.c_int32), ("fuse_thr", C.c_double), ("mask_dtype", C.c_int32), ("fused_capacity", C.c_int64),
        ("ws_sor", C.c_int32), ("plan", C.c_int32),
    ]


def _bind():
    return _lib.lib()


def class_ranks(class_ids) -> np.ndarray:
    """(F, 2, n) class ids -> (F, 2, n) int32 ranks: the position of each class in the iteration order of
    `set(class_dict1.keys()).union(class_dict2.keys())` (2cam/vision_pipeline_utils.py:117), the dictionaries keyed
    by `int(class_id)` in first-appearance order (:98-107, :304).  The reference builds them from the NON-EMPTY
    clouds only; which clouds are empty is known on the device alone, so the ranks are those of the full id lists
    (identical whenever every class has at least one non-empty cloud, i.e. always in practice)."""
    ids = np.asarray(class_ids)
    if ids.ndim != 3 or ids.shape[1] != 2:
        raise ValueError("class_ids must have shape (F, 2, n_obj)")
    out = np.zeros(ids.shape, np.int32)
    cache = {}
    for f in range(ids.shape[0]):
        a, b = tuple(int(c) for c in ids[f, 0]), tuple(int(c) for c in ids[f, 1])
        rank = cache.get((a, b))
        if rank is None:
            order = list(set(dict.fromkeys(a).keys()).union(dict.fromkeys(b).keys()))
            rank = cache[(a, b)] = {cid: i for i, cid in enumerate(order)}
        out[f, 0] = [rank[c] for c in a]
        out[f, 1] = [rank[c] for c in b]
    return out


class _Outputs:
    """One set of device-resident result buffers of a run."""

    def __init__(self, p):
        d, cfg = p.dev, p.cfg
        nimg, nslots = p.F * p.C, p.F * p.slots
        self.ws_xyz = torch.empty((cfg.ws_capacity, 3), dtype=torch.float32, device=d)
        self.ws_off = torch.zeros(nimg + 1, dtype=torch.int32, device=d)
        self.obj_xyz = torch.empty((p.fused_capacity, 3), dtype=torch.float64, device=d)
        self.obj_off = torch.zeros(nslots + 1, dtype=torch.int32, device=d)
        self.obj_meta = torch.zeros((max(nslots, 1), 4), dtype=torch.int32, device=d)
        self.kept_xyz = torch.empty((cfg.ws_capacity, 3), dtype=torch.float32, device=d)
        self.kept_off = torch.zeros(p.F + 1, dtype=torch.int32, device=d)
        self.counts = torch.zeros(8, dtype=torch.int32, device=d)


class FramePipeline:
    """F frames x C cameras per call.  All buffers are allocated once; outputs are device tensors
    (compacted rows + per-frame offsets) valid until the next run().

    intr: (cx, cy, fx, fy) shared by all cameras, or one tuple per camera.
    R, t: per-camera (C,3,3) / (C,3) camera -> robot transforms.
    box:  ((xlo,xhi),(ylo,yhi),(zlo,zhi)) workspace crop (2cam/vision_pipeline_utils.py:241-243).
    fuse: "index" -- mask o of every camera is the same object (1-vs-1 branch, :123-126), or
          "centroid" -- the whole fuse_point_clouds_centroid (:88-165) per frame, C == 2; needs class ids
          (set_class_ids) and `fuse_threshold` (2cams.py:197 passes 0.3).
    mask_dtype: torch.uint8 (0/1 bytes) or torch.float32 (the ultralytics masks.data planes, :295).
    project_kernel: "auto" (by batch size / mask count), "tma" or "cp_async" -- same results, see include/rt3d.h.
    sor_bitplanes: occupancy bit-plane neighbour search in front of the outlier filter's grid search (same results): None =
          automatic (large batches), True / False = forced on / off.
    workspace_sor: filter every frame's fused workspace with remove_statistical_outlier(nb_neighbors, std_ratio) before the
          subtraction (2cam/2cams_mask_cpu.py:530); `workspace` in results() stays the unfiltered voxel sets.
    """

    def __init__(self, F, C_, H, W, n_obj, intr, R, t, box, voxel=0.005, flavour=None, nb_neighbors=20, std_ratio=1.5,
                 distance_threshold=0.3, max_range=20.0, ws_capacity=None, obj_capacity=None, device="cuda:0",
                 fuse="index", fuse_threshold=0.3, mask_dtype=torch.uint8, fused_capacity=None, sor_bitplanes=None,
                 project_kernel="auto", workspace_sor=False):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("FramePipeline needs a CUDA device; rt3d has no CPU path")
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        if fuse not in ("index", "centroid"):
            raise ValueError("fuse must be 'index' or 'centroid'")
        if mask_dtype not in (torch.uint8, torch.bool, torch.float32):
            raise ValueError("mask_dtype must be torch.uint8 or torch.float32")
        self.F, self.C, self.H, self.W, self.n_obj = int(F), int(C_), int(H), int(W), int(n_obj)
        self.mask_dtype = torch.float32 if mask_dtype == torch.float32 else torch.uint8
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
        self.sor_bitplanes = sor_bitplanes            # occupancy bit-plane neighbour search in front of the grid search: None = automatic
        if project_kernel not in ("auto", "tma", "cp_async"):
            raise ValueError("project_kernel must be 'auto', 'tma' or 'cp_async'")
        flags = arith_flags(flavour) | (0 if sor_bitplanes is None else 16 if sor_bitplanes else 2) | {"auto": 0, "tma": 4, "cp_async": 8}[project_kernel]
        cfg.voxel, cfg.flags, cfg.max_range = float(voxel), flags, float(max_range)
        cfg.sor_k, cfg.sor_std_ratio, cfg.sub_thr = int(nb_neighbors), float(std_ratio), float(np.float32(distance_threshold))
        npx = self.H * self.W
        # defaults: half the pixels of the batch become workspace voxels; instance masks cover 1/16 of an image each.
        # A scene that exceeds them is reported (device status "capacity"), never silently truncated.
        cfg.ws_capacity = int(ws_capacity or self.F * self.C * npx // 2)
        cfg.obj_capacity = int(obj_capacity or max(4096, self.F * self.C * max(self.n_obj, 1) * npx // 16))
        cfg.fuse_mode = FUSE_CENTROID if fuse == "centroid" else FUSE_BY_INDEX
        cfg.fuse_thr = float(fuse_threshold)
        cfg.mask_dtype = _lib.RT3D_F32 if self.mask_dtype == torch.float32 else _lib.RT3D_U8
        # "centroid" can emit a camera-2 cloud once per camera-1 cloud that matches it and once more as a leftover
        self.fused_capacity = int(fused_capacity or (2 * cfg.obj_capacity if fuse == "centroid" else cfg.obj_capacity))
        cfg.fused_capacity = self.fused_capacity
        cfg.ws_sor = 1 if workspace_sor else 0
        self.workspace_sor = bool(workspace_sor)
        cfg.plan = 0
        self.cfg = cfg
        self.slots = 2 * self.n_obj if fuse == "centroid" else self.n_obj
        self.fuse = fuse
        L = _bind()
        nbytes = L.rt3d_frame_batch_scratch_bytes(C.byref(cfg))
        if nbytes == 0:
            raise ValueError(f"rt3d_frame_batch: {_lib.last_error()}")
        assert L.rt3d_frame_batch_slots(C.byref(cfg)) == self.slots
        with torch.cuda.device(self.dev):
            self.scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            _lib.check(L.rt3d_frame_batch_init(C.byref(cfg), self.scratch.data_ptr(), nbytes,
                                               torch.cuda.current_stream(self.dev).cuda_stream), "rt3d_frame_batch_init")
            self._outs = [_Outputs(self)]
        self._bind_outputs(0)
        self.class_rank = None        # device int32 (F, C, n_obj), fuse="centroid"
        self.class_names = None       # per frame: rank -> class id
        self._timers = []
        self._in2 = None
        self._h = None

    def _bind_outputs(self, b):
        o = self._outs[b]
        self.ws_xyz, self.ws_off, self.obj_xyz, self.obj_off, self.obj_meta = o.ws_xyz, o.ws_off, o.obj_xyz, o.obj_off, o.obj_meta
        self.kept_xyz, self.kept_off, self.counts = o.kept_xyz, o.kept_off, o.counts

    def set_class_ids(self, class_ids):
        """Class id of every instance mask, (F, 2, n_obj) (the `cls` arrays the reference gets from YOLO, :277-282),
        for fuse="centroid".  Host data; ranked here in the reference's set order and uploaded (into the same device
        buffer on every call, so a captured CUDA graph sees the new ranks)."""
        ids = np.asarray(class_ids).reshape(self.F, self.C, self.n_obj)
        ranks = class_ranks(ids)
        self.class_names = [{int(r): int(c) for r, c in zip(ranks[f].ravel(), ids[f].ravel())} for f in range(self.F)]
        if self.class_rank is None:
            self.class_rank = torch.from_numpy(ranks).to(self.dev)
        else:
            self.class_rank.copy_(torch.from_numpy(ranks))      # in place: a captured CUDA graph holds this address
        return self

    # kernel launches of one run (for bench.py's gpu_launches claim): fixed by the configuration
    def count_launches(self) -> int:
        def passes(bits):
            return (bits + 7) // 8
        n = 2 + 5 + 4  # project, clamp; extract = count, slab scan, image scan, emit, bitmap clear; keep scan, write, per-frame counts, offsets scan
        if self.n_obj == 0:
            n += 1     # no flag kernel to count the kept rows per tile on the way: a counting sweep
        if self.n_obj > 0:
            if self.mask_dtype == torch.float32:
                n += 1                                           # fp32 mask planes -> bytes
            kmax = math.ceil(self.cfg.max_range / self.cfg.voxel) + 2
            obits = max(1, math.ceil(math.log2(2 * kmax)))
            seg_bits = max(1, math.ceil(math.log2(max(2, self.F * self.n_obj * self.C))))
            group_bits = max(1, math.ceil(math.log2(max(2, self.F * self.slots))))
            frame_bits = max(1, math.ceil(math.log2(max(2, self.F))))
            n += 1 + passes(3 * obits + seg_bits) + 2            # hist + passes, unique select, segment offsets
            n += 4 + passes(21 + group_bits) + 6                 # plan, regroup, prepare, cell keys, passes, gather, column table, knn, stats, select, offsets
            planes = self.sor_bitplanes if self.sor_bitplanes is not None else (self.cfg.obj_capacity >= 1000000 and self.cfg.voxel >= 0.003)
            if planes and self.cfg.sor_k <= 32 and (self.fuse == "centroid" or self.C <= 2):
                n += 4                                           # bit planes: plan, zero, fill, lattice neighbour search
            n += 4 if self.fuse == "centroid" else 1             # centroids, matching, scan, assembly | meta rows
            n += 2 + passes(12 + frame_bits) + 3                 # subtraction: prepare, cell keys, passes, cells, super cells, flags
        if self.workspace_sor:
            frame_bits = max(1, math.ceil(math.log2(max(2, self.F))))
            n += 2 + 2 + (21 + frame_bits + 7) // 8 + 2 + 2 + 2  # groups, rows; prepare, cell keys; passes; gather, columns; knn, stats; select, offsets
        return n

    @property
    def project_kernel_used(self) -> str:
        """which project kernel rt3d_frame_batch_run launches for this configuration (mirror of the rule in frame_batch.cu)"""
        choice = (self.cfg.flags >> 2) & 3
        wide = self.W % 16 == 0
        tiles = self.F * self.C * ((self.H + 7) // 8) * (((self.W + 255) // 256) * 4)
        tma = wide and self.n_obj <= 2 and tiles >= 2048 * 148
        if choice == 1:
            tma = wide
        elif choice == 2:
            tma = False
        return "fb_project_tma_kernel" if tma else "fb_project_kernel"

    STAGES = ("project", "extract", "objects", "fuse", "subtract")

    def enable_stage_timing(self, slots: int = 1):
        """Record CUDA events between the stages of every run(); `slots` independent event sets let a
        timed loop record every step without a host synchronise in between (read them afterwards)."""
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
        """depth (F,C,H,W) fp32 and masks (F,C,n_obj,H,W) of the pipeline's mask_dtype, both on the device.
        Asynchronous on the current stream of the pipeline's device."""
        if not depth.is_cuda or depth.device != self.dev or depth.dtype != torch.float32 or not depth.is_contiguous():
            raise ValueError(f"depth must be a contiguous fp32 tensor on {self.dev}")
        if depth.numel() != self.F * self.C * self.H * self.W:
            raise ValueError("depth has the wrong number of elements for this pipeline")
        mptr = rptr = None
        if self.n_obj > 0:
            ok = (torch.uint8, torch.bool) if self.mask_dtype == torch.uint8 else (torch.float32,)
            if masks is None or not masks.is_cuda or masks.device != self.dev or masks.dtype not in ok or not masks.is_contiguous():
                raise ValueError(f"masks must be a contiguous {self.mask_dtype} tensor on {self.dev}")
            if masks.numel() != self.F * self.C * self.n_obj * self.H * self.W:
                raise ValueError("masks has the wrong number of elements for this pipeline")
            mptr = masks.data_ptr()
            if self.fuse == "centroid":
                if self.class_rank is None:
                    raise ValueError("fuse='centroid' needs the class ids of the instance masks: call set_class_ids() first")
                rptr = self.class_rank.data_ptr()
        L = _bind()
        with torch.cuda.device(self.dev):
            stream = torch.cuda.current_stream(self.dev).cuda_stream
            _lib.check(L.rt3d_frame_batch_run_timed(C.byref(self.cfg), depth.data_ptr(), mptr, rptr, self.ws_xyz.data_ptr(),
                                                    self.ws_off.data_ptr(), self.obj_xyz.data_ptr(), self.obj_off.data_ptr(),
                                                    self.obj_meta.data_ptr(), self.kept_xyz.data_ptr(), self.kept_off.data_ptr(),
                                                    self.counts.data_ptr(), self.scratch.data_ptr(), self.scratch.numel(), stream,
                                                    self._timers[slot] if slot is not None else None), "rt3d_frame_batch_run")
        return self

    @staticmethod
    def _check(counts):
        if counts[0]:
            raise _lib.Rt3dError(f"frame batch: device status {counts[0]}: {_lib.decode_status(counts[0])}")

    def results(self):
        """Synchronise and slice the outputs per frame: list of dicts
        workspace / workspace_per_cam / kept: fp32 rows;
        objects: one fp64 tensor per output slot (fuse="index": slot o = object o);
        fused: the non-empty slots in the reference's list order as dicts(points, kind, class_id) -- kind "leftover"
               clouds are float32 in the reference (:162-163), all others float64."""
        counts = self.counts.tolist()
        self._check(counts)
        ws_off, obj_off, kept_off = self.ws_off.tolist(), self.obj_off.tolist(), self.kept_off.tolist()
        meta = self.obj_meta.tolist() if self.n_obj else []
        out = []
        for f in range(self.F):
            objects = [self.obj_xyz[obj_off[f * self.slots + o]:obj_off[f * self.slots + o + 1]] for o in range(self.slots)]
            fused = []
            for o in range(self.slots):
                kind, rank, na, nb = meta[f * self.slots + o]
                if kind == CLOUD_EMPTY:
                    continue
                cid = self.class_names[f].get(rank, rank) if (self.fuse == "centroid" and self.class_names) else rank
                fused.append(dict(points=objects[o], kind=CLOUD_KINDS[kind], class_id=cid, rows=(na, nb)))
            out.append(dict(
                workspace=self.ws_xyz[ws_off[f * self.C]:ws_off[(f + 1) * self.C]],
                workspace_per_cam=[self.ws_xyz[ws_off[f * self.C + c]:ws_off[f * self.C + c + 1]] for c in range(self.C)],
                objects=objects, fused=fused,
                kept=self.kept_xyz[kept_off[f]:kept_off[f + 1]],
            ))
        return out

    # ---------------------------------------------------------------- host-fed path (end-to-end)
    def _ensure_staging(self, masks_on_host=True):
        if self._in2 is None:
            with torch.cuda.device(self.dev):
                def inputs():
                    d = torch.empty((self.F, self.C, self.H, self.W), dtype=torch.float32, device=self.dev)
                    m = (torch.empty((self.F, self.C, self.n_obj, self.H, self.W), dtype=self.mask_dtype, device=self.dev)
                         if (self.n_obj and masks_on_host) else None)
                    return [d, m]
                self._in2 = [inputs(), inputs()]
                while len(self._outs) < 2:
                    self._outs.append(_Outputs(self))
                self._copy_in = torch.cuda.Stream(self.dev)
                self._copy_out = torch.cuda.Stream(self.dev)
                self._ev_in = [torch.cuda.Event(), torch.cuda.Event()]
                self._ev_in_free = [torch.cuda.Event(), torch.cuda.Event()]
                self._ev_run = [torch.cuda.Event(), torch.cuda.Event()]
                self._ev_small = [torch.cuda.Event(), torch.cuda.Event()]
                self._ev_out_free = [torch.cuda.Event(), torch.cuda.Event()]
            nslots = self.F * self.slots

            def mirrors():
                return dict(counts=torch.empty(8, dtype=torch.int32).pin_memory(),
                            kept_off=torch.empty(self.F + 1, dtype=torch.int32).pin_memory(),
                            obj_off=torch.empty(nslots + 1, dtype=torch.int32).pin_memory(),
                            obj_meta=torch.empty((max(nslots, 1), 4), dtype=torch.int32).pin_memory(),
                            kept=torch.empty((self.cfg.ws_capacity, 3), dtype=torch.float32).pin_memory(),
                            objs=torch.empty((self.fused_capacity, 3), dtype=torch.float64).pin_memory())
            self._h2 = [mirrors(), mirrors()]
            self._h = self._h2[0]
        elif masks_on_host and self.n_obj and self._in2[0][1] is None:
            for pair in self._in2:
                pair[1] = torch.empty((self.F, self.C, self.n_obj, self.H, self.W), dtype=self.mask_dtype, device=self.dev)

    def run_host_pipelined(self, batches, device_masks=None, on_result=None):
        """End-to-end over a SEQUENCE of host batches [(depth_host, masks_host), ...] (pinned memory).

        Three streams: the H2D copy of batch i+1 (copy-in stream, second input buffer) and the D2H copy of batch
        i-1's results (copy-out stream, second output set) overlap batch i's kernels, so a capture loop is bound by
        the slowest of {PCIe in, kernels, PCIe out} instead of their sum.  Result rows are copied with their exact
        size: the host reads the batch's counts / offsets first (a 100-byte copy it waits for while the NEXT batch is
        already running), checks the device status word and the capacities, then issues the row copies.

        `device_masks`: instance masks already resident on the device (as they are in the reference, where YOLO
        leaves `masks.data` on the GPU, :295) -- then every batch is `(depth_host, None)` and only depth crosses PCIe.
        `on_result(i, mirrors)` is called once batch i's rows have been queued for copy (they are complete after
        the copy-out stream's next event; the final batch is complete when this function returns).
        Returns (h2d_bytes, d2h_bytes) of the last batch.  Synchronises the copy-out stream before returning."""
        self._ensure_staging(masks_on_host=device_masks is None)
        main = torch.cuda.current_stream(self.dev)
        cin, cout = self._copy_in, self._copy_out
        cin.wait_stream(main)
        cout.wait_stream(main)
        h2d = d2h = 0
        pending = None   # (index, buffer) of the batch whose rows have not been copied yet

        def finish(i, b):
            """exact-size D2H of batch i's rows (outputs set b)"""
            nonlocal d2h
            h, o = self._h2[b], self._outs[b]
            self._ev_small[b].synchronize()
            counts = h["counts"].tolist()
            self._check(counts)
            nk, no = int(h["kept_off"][self.F]), int(h["obj_off"][self.F * self.slots])
            if nk != counts[5] or (self.n_obj and no != counts[3]) or nk > self.cfg.ws_capacity or no > self.fused_capacity:
                raise _lib.Rt3dError(f"frame batch {i}: inconsistent result sizes (kept {nk}/{counts[5]}, fused {no}/{counts[3]})")
            with torch.cuda.stream(cout):
                h["kept"][:nk].copy_(o.kept_xyz[:nk], non_blocking=True)
                if self.n_obj:
                    h["objs"][:no].copy_(o.obj_xyz[:no], non_blocking=True)
                self._ev_out_free[b].record(cout)
            d2h = nk * 12 + no * 24 + (8 + self.F + 1 + (self.F * self.slots + 1) + 4 * self.F * self.slots) * 4
            self._h = h
            if on_result is not None:
                on_result(i, h)

        for i, (depth_host, masks_host) in enumerate(batches):
            b = i & 1
            d_depth, d_masks = self._in2[b]
            with torch.cuda.stream(cin):
                if i >= 2:
                    cin.wait_event(self._ev_in_free[b])
                d_depth.copy_(depth_host.view_as(d_depth), non_blocking=True)
                h2d = depth_host.numel() * 4
                if self.n_obj and device_masks is None:
                    d_masks.copy_(masks_host.view_as(d_masks), non_blocking=True)
                    h2d += masks_host.numel() * masks_host.element_size()
                self._ev_in[b].record(cin)
            main.wait_event(self._ev_in[b])
            if i >= 2:
                main.wait_event(self._ev_out_free[b])     # output set b has been copied out
            self._bind_outputs(b)
            self.run(d_depth, device_masks if device_masks is not None else d_masks)
            self._ev_in_free[b].record(main)
            self._ev_run[b].record(main)
            # the previous batch's rows go out while this batch computes (queued on the copy-out stream BEFORE this
            # batch's small copies, which wait for its kernels)
            if pending is not None:
                finish(*pending)
            h, o = self._h2[b], self._outs[b]
            with torch.cuda.stream(cout):
                cout.wait_event(self._ev_run[b])
                h["counts"].copy_(o.counts, non_blocking=True)
                h["kept_off"].copy_(o.kept_off, non_blocking=True)
                h["obj_off"].copy_(o.obj_off, non_blocking=True)
                h["obj_meta"].copy_(o.obj_meta, non_blocking=True)
                self._ev_small[b].record(cout)
            pending = (i, b)
        if pending is not None:
            finish(*pending)
        cout.synchronize()
        return h2d, d2h

    def run_host(self, depth_host: torch.Tensor, masks_host: torch.Tensor | None):
        """One batch from pinned host buffers: H2D -> run -> exact-size D2H of the results (see run_host_pipelined).
        Returns (bytes_h2d, bytes_d2h); the pinned mirrors are in `host_results()`."""
        return self.run_host_pipelined([(depth_host, masks_host)])

    def host_results(self):
        """The last batch copied out by run_host*(): per frame dict(kept, objects) of NumPy views into the pinned
        mirrors (status and sizes were checked when the rows were copied)."""
        h = self._h
        ko, oo = h["kept_off"].tolist(), h["obj_off"].tolist()
        kept, objs = h["kept"].numpy(), h["objs"].numpy()
        return [dict(kept=kept[ko[f]:ko[f + 1]],
                     objects=[objs[oo[f * self.slots + o]:oo[f * self.slots + o + 1]] for o in range(self.slots)])
                for f in range(self.F)]
