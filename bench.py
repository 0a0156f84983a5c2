#!/usr/bin/env python
"""bench.py -- fused 2-camera HD720 frames/s @ 5 mm voxels (BASELINE.json metric, configs[1]) and its siblings.

One "step" = one batch of F synthetic multi-camera frames through the whole geometric path
(back-project -> transform -> crop -> voxels -> SOR fusion -> workspace subtraction) of
rt3d_frame_batch_run.  `value` is device-resident throughput (inputs already in HBM, CUDA-event
timed, max over ranks); `e2e` feeds the same call from pinned host buffers with the H2D / D2H copies
inside the timed region.  `--impl reference` times the CPU restatement of the reference path
(oracle/, the "port" -- Open3D cannot be installed, the reference is pure Python) on the host cores;
`--impl torch_eager` times the reference's own sequence of torch ops, restated in oracle/torch_eager.py,
on the same GPU ("the reference as deployed", SURVEY.md 8d iii).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--frames F] [--res HD720] [--voxel 0.005]
                  [--objects n] [--thr 0.3] [--distinct D] [--impl rt3d|reference|torch_eager]
  torchrun --nproc-per-node N bench.py --gpus N ...   (one rank per GPU, frames sharded, no collective)

configs (BASELINE.json `configs`, 0-based there):
  2  (default, the headline)  2 cameras HD720, 1 object, 5 mm, fusion by index (1-vs-1), subtraction 0.3, F = 256
  3  2 cameras HD1080, 10 + 2 overlapping objects, 2 mm, the whole fuse_point_clouds_centroid (threshold 0.3), F = 8
  4  8-camera HD2K rig, stream c -> GPU c mod G, canonical voxel keys all-gathered over NCCL, merged and decoded
  5  = config 2 swept over voxel 2 / 5 / 10 mm (one JSON line per voxel size)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rt-3d-reconstruction_b200"))

import numpy as np  # noqa: E402

UNIT = "frames/s"
PRESETS = {
    2: dict(res="HD720", objects=1, overlap=0, voxel=0.005, thr=0.3, fuse="index", fuse_thr=0.3, frames=256, cams=2),
    3: dict(res="HD1080", objects=12, overlap=2, voxel=0.002, thr=0.3, fuse="centroid", fuse_thr=0.3, frames=8, cams=2),
    4: dict(res="HD2K", objects=0, overlap=0, voxel=0.005, thr=0.3, fuse="index", fuse_thr=0.3, frames=4, cams=8),
    5: dict(res="HD720", objects=1, overlap=0, voxel=0.005, thr=0.3, fuse="index", fuse_thr=0.3, frames=256, cams=2),
}


def resolve(args):
    """the workload of this invocation: preset of --config overridden by the explicit flags"""
    w = dict(PRESETS[args.config])
    for k, a in (("res", args.res), ("objects", args.objects), ("voxel", args.voxel), ("thr", args.thr), ("frames", args.frames),
                 ("fuse", args.fuse), ("fuse_thr", args.fuse_thr)):
        if a is not None:
            w[k] = a
    if args.objects is not None and args.objects < w["overlap"] + 1:
        w["overlap"] = 0
    w["distinct"] = min(args.distinct or 8, w["frames"])
    w["ws_sor"] = bool(getattr(args, "ws_sor", False))
    w["sor_bitplanes"] = False if getattr(args, "sor_grid_only", False) else None      # None: automatic
    return w


def describe(w):
    return _describe(w) + (" + SOR(20,1.5) of the fused workspace before the subtraction (2cams_mask_cpu.py:530)" if w.get("ws_sor") else "")


def _describe(w):
    mm = f"{w['voxel'] * 1e3:g}"
    if w["cams"] == 8:
        return (f"configs[3]: synthetic 8-camera {w['res']} rig, camera stream c -> GPU c mod G, voxel {mm} mm, crop + transform + voxel set per "
                "stream, canonical voxel keys all-gathered (NCCL over NVLink), merge-unique + decode = box-wide fused workspace")
    if w["fuse"] == "centroid":
        return (f"configs[2]: 2 cameras {w['res']} (fp32 depth + {w['objects']} uint8 instance masks each: {w['objects'] - w['overlap']} + "
                f"{w['overlap']} overlapping objects, class ids alternating 39/41, per-camera mask order), voxel {mm} mm, crop + transform + "
                f"fuse_point_clouds_centroid (SOR(20,1.5), threshold {w['fuse_thr']:g}) + workspace subtraction thr {w['thr']:g}")
    return (f"configs[1]: 2 cameras {w['res']} ({'x'.join(str(v) for v in reversed(RES(w['res'])))} fp32 depth + {w['objects']} uint8 instance mask"
            f"{'s' if w['objects'] != 1 else ''} each), {w['objects']} object{'s' if w['objects'] != 1 else ''}, voxel {mm} mm, "
            f"crop + transform + SOR(20,1.5) fusion + workspace subtraction thr {w['thr']:g} (2cams.py defaults)")


def RES(name):
    from rt3d import synthetic as syn
    return syn.RESOLUTIONS[name]


def metric_name(w):
    mm = f"{w['voxel'] * 1e3:g}".replace(".", "p")
    if w["cams"] == 8:
        return f"fused_8cam_{w['res'].lower()}_rig_frames_per_sec_{mm}mm"
    return f"fused_2cam_{w['res'].lower()}_frames_per_sec_{mm}mm"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def project_traffic(w):
    """dram__bytes_read.sum + dram__bytes_write.sum of one fb_project_kernel launch, from the committed
    `ncu --set full` capture of this very configuration (profiles/); None when no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "project_traffic.json")) as fh:
            for row in json.load(fh)["captures"]:
                if (int(row["frames_per_step"]) == int(w["frames"]) and row.get("res", "HD720") == w["res"]
                        and int(row.get("objects", 1)) == int(w["objects"])
                        and abs(float(row.get("voxel", w["voxel"])) - float(w["voxel"])) < 1e-9):
                    return float(row["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def make_frames(w, n_unique, seed0=1000):
    """the config's scene with the object positions jittered per frame (SURVEY.md 8d config 5 seeds 1000..)"""
    from rt3d import synthetic as syn
    if w["cams"] == 8:
        return [syn.frame(w["res"], seed=30 + 8 * i, n_objects=2, cams=tuple(range(8)), rig=8) for i in range(n_unique)]
    return [syn.frame(w["res"], seed=seed0 + 2 * i, n_objects=w["objects"], overlap=w["overlap"], jitter=0.05 if w["objects"] == 1 else 0.01,
                      shuffle_cams=w["fuse"] == "centroid") for i in range(n_unique)]


_FULL_AFFINITY = None


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank to the CPU cores next to its GPU (NVML's affinity mask) and prefer that NUMA node for its
    memory: the pinned host buffers of the end-to-end path then sit on the GPU's own socket, and 8 ranks do not all
    pull from one.  Best effort (a cpuset may forbid either); what happened goes to stderr."""
    global _FULL_AFFINITY
    note = []
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        _FULL_AFFINITY = os.sched_getaffinity(0)
        near = cpus & _FULL_AFFINITY
        if near:
            os.sched_setaffinity(0, near)
        note.append(f"cpus {len(near)} of {len(cpus)} near (allowed {len(_FULL_AFFINITY)})")
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        node = -1
        for cand in (bus, bus[4:] if len(bus) > 12 else bus):       # NVML prints an 8-digit domain, sysfs a 4-digit one
            path = f"/sys/bus/pci/devices/{cand}/numa_node"
            if os.path.exists(path):
                node = int(open(path).read().strip())
                break
        if node >= 0:
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            MPOL_PREFERRED = 1
            rc = libc.syscall(238, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))   # set_mempolicy (x86-64)
            note.append(f"numa node {node} mempolicy rc={rc}")
        else:
            note.append("numa node unknown")
    except Exception as e:   # noqa: BLE001
        note.append(f"not bound: {e!r}")
    print(f"bench.py: gpu {gpu_index}: " + "; ".join(note), file=sys.stderr)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML from a
    thread every ~5 ms (a timed region lasts tens of milliseconds), nvidia-smi as the fallback."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.idx, self.sm, self.bits, self.mx, self.run, self.thread, self.src = gpu_index, [], 0, None, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while self.run:
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.run, self.src = True, "nvml"
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.src = None
        return self

    def stop(self):
        self.run = False
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if not self.sm:      # NVML unavailable: one nvidia-smi reading taken right after the timed region
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     stdout=subprocess.PIPE, text=True, timeout=10).stdout.strip().split(",")
                self.sm, self.mx, self.src = [float(out[0])], float(out[1]), "nvidia-smi (after the timed region)"
            except Exception:
                pass
        reasons = [name for name, bit in self.REASONS if self.bits & bit]
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(self.sm), "source": self.src}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_port_fps(w, frames, cores, budget_s=20.0):
    """The oracle (CPU restatement of the reference path) on the host cores: frames are independent, so
    `cores` worker threads each run whole frames (ctypes releases the GIL; OpenMP inside is set to 1)."""
    from concurrent.futures import ThreadPoolExecutor

    import oracle
    from rt3d import synthetic as syn
    oracle.build()
    oracle.lib().rt3d_oracle_set_threads(1)
    voxel, thr = w["voxel"], w["thr"]

    def one(fr):
        cx, cy, fx, fy = fr["intr"]
        ones = np.ones(fr["depth"][0].shape, np.uint8)
        C, n_obj = fr["depth"].shape[0], (fr["masks"].shape[1] if w["objects"] else 0)
        ws, lists = [], [[] for _ in range(C)]
        for c in range(C):
            p = oracle.backproject_indices(oracle.mask_nonzero(ones), fr["depth"][c], cx, cy, fx, fy)
            ws.append(oracle.voxel_snap_unique(oracle.crop(oracle.transform(p, fr["R"][c], fr["t"][c]), *syn.WORKSPACE_BOX), voxel))
            for o in range(n_obj):
                pc = oracle.process_masks_cloud(fr["masks"][c, o], fr["depth"][c], cx, cy, fx, fy, fr["R"][c], fr["t"][c], voxel)
                if pc is not None:
                    lists[c].append((pc, int(fr["class_ids"][c][o])))
        if C == 8:     # the rig: box-wide union of the streams' voxel sets
            return np.unique(np.vstack(ws), axis=0).shape[0]
        if w["fuse"] == "centroid":
            fused = [pc for pc, _ in oracle.fuse_point_clouds_centroid(lists[0], lists[1], w["fuse_thr"])[2] if pc.shape[0]]
        else:
            fused = [oracle.filter_outliers_sor(np.vstack([pc for lst in lists for pc, _ in lst]))]
        if w.get("ws_sor"):
            ws = [oracle.filter_outliers_sor(np.vstack(ws)).astype(np.float32)]
        kept = oracle.subtract_point_clouds_gpu(np.vstack(ws), np.vstack(fused), thr)
        return kept.shape[0]

    one(frames[0])  # warm (library load, page faults)
    t0 = time.perf_counter()
    one(frames[0])
    t_one = time.perf_counter() - t0
    rounds = max(1, int(budget_s / max(t_one, 1e-3)))       # rounds of `cores` frames that fit the budget
    rounds = min(rounds, 8)
    work = [frames[i % len(frames)] for i in range(cores * rounds)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(one, work))
    dt = time.perf_counter() - t0
    return len(work) / dt, len(work), dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    w = resolve(args)
    cores = os.cpu_count() or 1
    frames = make_frames(w, min(w["distinct"], 4))
    per_step = []
    sample = None
    for i in range(args.warmup + args.steps):
        fps, n, dt = cpu_port_fps(w, frames, cores, budget_s=max(2.0, 60.0 / max(1, args.warmup + args.steps)))
        sample = f"{n} fused {w['res']} frames per step on {cores} threads ({dt:.1f} s)"
        if i >= args.warmup:
            per_step.append((fps, dt))
    value = statistics.median(f for f, _ in per_step)
    line = {
        "impl": "reference", "metric": metric_name(w), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(d for _, d in per_step), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": describe(w), "note": "CPU restatement of the reference path (oracle port; parity vs the in-repo restatement "
                   "for the Open3D stages); the reference's Open3D/pyzed/ultralytics dependencies are not installable here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arms
def init_dist(world, dev):
    import torch.distributed as dist
    if world <= 1:
        return
    # NCCL prints its version banner to stdout when the communicator is created (NCCL_DEBUG=VERSION or higher,
    # as set on some boxes): route fd 1 to stderr until that has happened, so stdout carries the one JSON line
    import torch
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def host_copy_ceiling(dev, world, nbytes=1 << 30, reps=4):
    """Platform ceiling of the end-to-end path: all ranks run a pinned H2D copy loop of `nbytes` at the same time
    (barrier before, max time over ranks after); returns aggregate GB/s.  What the box's PCIe / host memory gives N
    concurrent GPUs, independent of any kernel of ours."""
    import torch
    import torch.distributed as dist
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del src, dst
    return world * reps * nbytes / (t.item() * 1e-3) / 1e9


def run_rt3d_arm(args, rank, local_rank, world, w=None, quiet_cpu=False):
    import torch
    import torch.distributed as dist

    from rt3d import synthetic as syn
    from rt3d.pipeline import FramePipeline

    w = w or resolve(args)
    dev = torch.device("cuda", local_rank)
    F, C, n_obj = w["frames"], w["cams"], w["objects"]
    D = w["distinct"]
    frames = make_frames(w, D, seed0=1000 + 100 * rank)
    H, W = frames[0]["H"], frames[0]["W"]
    reps = (F + D - 1) // D
    depth_h = torch.from_numpy(np.stack([f["depth"] for f in frames] * reps)[:F]).contiguous().pin_memory()
    masks_h = torch.from_numpy(np.stack([f["masks"] for f in frames] * reps)[:F]).contiguous().pin_memory() if n_obj else None
    cls = np.stack([f["class_ids"] for f in frames] * reps)[:F] if n_obj else None
    depth_d = depth_h.to(dev)
    masks_d = masks_h.to(dev) if n_obj else None
    obj_px_guess = int(masks_h.sum()) if n_obj else 0
    pipe = FramePipeline(F, C, H, W, n_obj, frames[0]["intr"], frames[0]["R"], frames[0]["t"], syn.WORKSPACE_BOX, voxel=w["voxel"],
                         distance_threshold=w["thr"], device=dev, fuse=w["fuse"], fuse_threshold=w["fuse_thr"],
                         obj_capacity=max(4096, int(obj_px_guess * 1.1)) if n_obj else None,
                         ws_capacity=(F * C * H * W // 8 if w["ws_sor"] else None), workspace_sor=w["ws_sor"],
                         sor_bitplanes=w["sor_bitplanes"]).enable_stage_timing(args.steps)
    if w["fuse"] == "centroid":
        pipe.set_class_ids(cls)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput
    for _ in range(max(args.warmup, 3)):
        pipe.run(depth_d, masks_d)
    counts = pipe.counts.tolist()
    if counts[0]:
        raise SystemExit(f"bench.py: device status {counts[0]} after warm-up")
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    stage_ms = {s: [] for s in pipe.STAGES}
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        pipe.run(depth_d, masks_d, slot=i)        # per-step stage events, no host sync inside the timed region
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    for i in range(args.steps):
        for k, v in pipe.stage_ms(i).items():
            stage_ms[k].append(v)
    clocks = sampler.stop() if sampler else None
    counts = pipe.counts.tolist()

    # ---- single-frame latency of the same configuration (F = 1, CUDA-graph replay), rank 0 only
    latency = None
    if rank == 0 and args.latency:
        latency = frame_latency(w, frames, dev, args)

    # ---- end to end: pinned host buffers in, results out, copies inside the timed region
    def e2e(device_masks=None):
        batches = [(depth_h, masks_h if device_masks is None else None)] * args.steps
        pipe.run_host_pipelined(batches[:2], device_masks=device_masks)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h2d, d2h = pipe.run_host_pipelined(batches, device_masks=device_masks)   # every step: H2D of its inputs + D2H of its results, status checked
        e1.record()
        barrier()
        return e0.elapsed_time(e1), h2d, d2h

    e2e_ms, h2d, d2h = e2e()
    e2e_dm_ms, h2d_dm, d2h_dm = e2e(device_masks=masks_d) if n_obj else (e2e_ms, h2d, d2h)
    ceiling = host_copy_ceiling(dev, world) if args.ceiling else None

    t = torch.tensor([elapsed_ms, e2e_ms, e2e_dm_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, e2e_dm_ms = t.tolist()
    if rank != 0:
        return None

    total_frames = F * args.steps * world
    value = total_frames / (elapsed_ms * 1e-3)
    p50 = {k: statistics.median(v) for k, v in stage_ms.items()}
    peak, peak_src = measured_peaks()
    obj_px = counts[1]
    proj_bytes = F * C * H * W * (4 + n_obj) + 8 * obj_px          # depth + uint8 masks in, packed object keys out
    achieved = proj_bytes / (p50["project"] * 1e-3) / 1e9
    e2e_fps = total_frames / (e2e_ms * 1e-3)
    line = {
        "metric": metric_name(w), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": f"synthetic ({D} distinct ray-cast frames per rank, tiled to {F})",
        "config": {"workload": describe(w), "frames_per_step_per_gpu": F, "l2_policy": "inputs larger than L2 "
                   f"({(depth_d.numel() * 4 + (masks_d.numel() if n_obj else 0)) / 1e6:.0f} MB per step vs 126 MB L2)",
                   "arith": os.environ.get("RT3D_ARITH", "ieee"), "parallelism": f"frames sharded x{world}, no collective"},
        "stage_p50_ms": p50,
        "points": {"object_px": obj_px, "object_voxels": counts[2], "fused": counts[3], "workspace_voxels": counts[4], "kept": counts[5],
                   "sor_queries_settled_on_bit_planes": counts[6]},
        "roofline": {"bound": "hbm", "kernel": pipe.project_kernel_used + " (stage A: back-project + transform + crop + voxel bit / key)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": project_traffic(w),
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": proj_bytes},
        "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "path": "FramePipeline.run_host_pipelined: per step pinned host depth+masks -> H2D (copy-in stream) -> rt3d_frame_batch_run -> "
                        "counts/offsets D2H, status check, exact-size D2H of kept and fused rows (copy-out stream); neighbouring steps overlap",
                "h2d_gbs": world * h2d * args.steps / (e2e_ms * 1e-3) / 1e9,
                "device_masks": {"value": total_frames / (e2e_dm_ms * 1e-3), "h2d_bytes_per_step": h2d_dm, "d2h_bytes_per_step": d2h_dm,
                                 "note": "instance masks already on the device (YOLO leaves masks.data there, vision_pipeline_utils.py:295): only depth crosses PCIe"}},
        "gpu_launches": pipe.count_launches() * args.steps,
        "clocks": clocks,
    }
    if ceiling is not None:
        line["e2e"]["platform_h2d_ceiling_gbs"] = ceiling
        line["e2e"]["frac_of_ceiling"] = line["e2e"]["h2d_gbs"] / ceiling
    if latency is not None:
        line["latency"] = latency
    if world == 1 and not args.no_cpu_baseline and not quiet_cpu:
        if _FULL_AFFINITY:
            os.sched_setaffinity(0, _FULL_AFFINITY)   # the CPU arm gets every host core again
        cores = os.cpu_count() or 1
        fps, n, dt = cpu_port_fps(w, frames[:4], cores, budget_s=15.0)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n} fused {w['res']} frames on {cores} threads ({dt:.1f} s)"}
    return line


def frame_latency(w, frames, dev, args, reps=50):
    """One frame at a time, as the reference's real-time loop runs (2cam/2cams.py:165-208): F = 1 and F = 8 through a
    captured CUDA graph, device-resident inputs; p50 of `reps` replays, per-stage p50 from un-captured runs."""
    import torch

    from rt3d import synthetic as syn
    from rt3d.pipeline import FramePipeline
    out = {}
    for F in (1, 8):
        fr = frames[:F] if len(frames) >= F else (frames * F)[:F]
        n_obj = w["objects"]
        depth = torch.from_numpy(np.stack([f["depth"] for f in fr])).to(dev)
        masks = torch.from_numpy(np.stack([f["masks"] for f in fr])).to(dev) if n_obj else None
        pipe = FramePipeline(F, w["cams"], fr[0]["H"], fr[0]["W"], n_obj, fr[0]["intr"], fr[0]["R"], fr[0]["t"], syn.WORKSPACE_BOX,
                             voxel=w["voxel"], distance_threshold=w["thr"], device=dev, fuse=w["fuse"], fuse_threshold=w["fuse_thr"]).enable_stage_timing(reps)
        if w["fuse"] == "centroid":
            pipe.set_class_ids(np.stack([f["class_ids"] for f in fr]))
        side = torch.cuda.Stream(dev)
        with torch.cuda.stream(side):
            for _ in range(3):
                pipe.run(depth, masks)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            pipe.run(depth, masks)
        ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            ms.append(e0.elapsed_time(e1))
        stage = {s: [] for s in pipe.STAGES}
        for i in range(reps):
            pipe.run(depth, masks, slot=i)
        torch.cuda.synchronize(dev)
        for i in range(reps):
            for k, v in pipe.stage_ms(i).items():
                stage[k].append(v)
        out[f"F{F}"] = {"graph_replay_p50_ms": statistics.median(ms), "graph_replay_min_ms": min(ms),
                        "stream_launch_stage_p50_ms": {k: statistics.median(v) for k, v in stage.items()},
                        "launches": pipe.count_launches()}
        del pipe, graph
    return out


def run_rig_arm(args, rank, local_rank, world):
    """configs[3]: 8 camera streams, stream c -> GPU c mod G.  Per step and rank: its streams' depth maps -> voxel sets
    (rt3d_frame_batch_run with n_obj = 0), canonical keys of the rank's union, NCCL all-gather of the variable-length key
    arrays, one merge-unique + decode on every rank.  The three phases are timed separately (CUDA events, max over ranks)."""
    import torch
    import torch.distributed as dist

    from rt3d import multi, ops, synthetic as syn
    from rt3d.pipeline import FramePipeline
    w = resolve(args)
    dev = torch.device("cuda", local_rank)
    if 8 % world:
        raise SystemExit("bench.py --config 4: the rig has 8 streams; use 1, 2, 4 or 8 GPUs")
    mine = [c for c in range(8) if multi.stream_owner(c, world) == rank]
    F = w["frames"]
    frames = make_frames(w, min(w["distinct"], F))
    frames = (frames * F)[:F]
    H, W = frames[0]["H"], frames[0]["W"]
    depth = torch.from_numpy(np.stack([f["depth"][mine] for f in frames])).contiguous().to(dev)
    R, t = frames[0]["R"][mine], frames[0]["t"][mine]
    pipe = FramePipeline(F, len(mine), H, W, 0, frames[0]["intr"], R, t, syn.WORKSPACE_BOX, voxel=w["voxel"], device=dev)

    def step(ev=None):
        pipe.run(depth, None)
        if ev:
            ev[0].record()
        # frame f of the rig: union over this rank's streams (rows ws_off[f*C] .. ws_off[(f+1)*C])
        off = pipe.ws_off.tolist()        # one small read per step: the per-frame key arrays have data-dependent lengths
        outs = []
        for f in range(F):
            pts = pipe.ws_xyz[off[f * len(mine)]:off[(f + 1) * len(mine)]]
            _, keys, layout = ops.voxel_snap_unique(pts, w["voxel"], return_keys=True)
            canon, canon_layout = ops.keys_canonical(keys, layout)
            outs.append((canon, canon_layout))
        if ev:
            ev[1].record()
        if world > 1:   # every rig frame of the step in one exchange: 2 collectives per step, not per frame
            per_rank = multi.all_gather_ragged(torch.cat([c for c, _ in outs]), [int(c.numel()) for c, _ in outs])
            gathered = [torch.cat([segs[f] for segs in per_rank]) for f in range(F)]
        else:
            gathered = [c for c, _ in outs]
        if ev:
            ev[2].record()
        fused = [ops.keys_decode(ops.keys_sort_unique(g, 63), outs[0][1], w["voxel"]) for g in gathered]
        if ev:
            ev[3].record()
        return outs, fused

    for _ in range(max(args.warmup, 3)):
        outs, fused = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    phases = []
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        start = torch.cuda.Event(enable_timing=True)
        start.record()
        outs, fused = step(ev)
        phases.append((start, ev))
    t1 = torch.cuda.Event(enable_timing=True)
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    total_ms = t0.elapsed_time(t1)
    ph = np.array([[s.elapsed_time(e[0]), e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])] for s, e in phases])
    med = np.median(ph, axis=0)
    tt = torch.tensor([total_ms, *med], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    key_bytes = sum(int(c.numel()) * 8 for c, _ in outs)
    kb = torch.tensor([key_bytes], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(kb, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    total_ms, voxel_ms, keys_ms, gather_ms, merge_ms = tt.tolist()
    return {
        "metric": metric_name(w), "value": F * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": f"synthetic ({min(w['distinct'], F)} distinct 8-camera rig frames, tiled to {F})",
        "config": {"workload": describe(w), "rig_frames_per_step": F, "streams_per_gpu": len(mine),
                   "parallelism": f"camera streams sharded x{world}; one ragged NCCL all-gather of voxel keys per step (lengths + padded keys)"},
        "phase_p50_ms": {"voxel_sets": voxel_ms, "rank_union_canonical_keys": keys_ms, "all_gather": gather_ms, "merge_unique_decode": merge_ms},
        "all_gather": {"bytes_per_step_all_ranks": int(kb.item()), "latency_us_per_rig_frame": 1e3 * gather_ms / F,
                       "collectives_per_step": 2 if world > 1 else 0},
        "points": {"box_wide_voxels_frame0": int(fused[0].shape[0])},
        "gpu_launches": None,
    }


def run_torch_eager_arm(args, rank, local_rank, world):
    """the reference's own op sequence (torch eager on the GPU + the CPU outlier filter, as deployed)
    on the same B200, frame by frame like its main loop.  oracle/torch_eager.py restates vision_pipeline_utils.py call
    for call (the reference itself cannot travel to the GPU box)."""
    if rank != 0:
        return None
    import torch

    from oracle import torch_eager
    w = resolve(args)
    dev = torch.device("cuda", local_rank)
    frames = make_frames(w, min(w["distinct"], 4))
    n = max(1, min(args.steps, 20))
    for fr in frames[:2]:
        torch_eager.frame(fr, w, dev)
    torch.cuda.synchronize(dev)
    stage = {}
    t0 = time.perf_counter()
    for i in range(n):
        for k, v in torch_eager.frame(frames[i % len(frames)], w, dev).items():
            stage.setdefault(k, []).append(v)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    fps = n / dt
    return {"impl": "torch_eager", "metric": metric_name(w), "value": fps, "unit": UNIT, "n_gpus": 1, "steps": n, "warmup": 2,
            "ms_per_step": 1e3 * dt / n, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": describe(w), "note": "the reference's op sequence in torch eager on this GPU, one frame per step, host "
                       "round trips as in the reference; SOR through the CPU port (Open3D's filter is a CPU routine too)"},
            "stage_p50_ms": {k: 1e3 * statistics.median(v) for k, v in stage.items() if not k.startswith("_")},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(PRESETS), help="BASELINE.json config (1-based as in SURVEY.md 8d); default 2 = the headline")
    ap.add_argument("--frames", type=int, default=None, help="fused frames per step per GPU (config 2/5: 256)")
    ap.add_argument("--res", default=None, choices=["QVGA", "VGA360", "HD720", "HD1080", "HD2K"])
    ap.add_argument("--voxel", type=float, default=None)
    ap.add_argument("--objects", type=int, default=None)
    ap.add_argument("--thr", type=float, default=None, help="workspace subtraction threshold")
    ap.add_argument("--fuse", default=None, choices=["index", "centroid"])
    ap.add_argument("--fuse-thr", dest="fuse_thr", type=float, default=None)
    ap.add_argument("--distinct", type=int, default=None, help="distinct synthetic frames per rank (default 8, tiled to --frames)")
    ap.add_argument("--impl", default="rt3d", choices=["rt3d", "reference", "torch_eager"])
    ap.add_argument("--latency", action="store_true", help="also report F = 1 / F = 8 CUDA-graph latency of the configuration")
    ap.add_argument("--ceiling", action="store_true", help="also measure the platform's concurrent pinned-H2D ceiling")
    ap.add_argument("--sor-grid-only", dest="sor_grid_only", action="store_true", help="outlier filter without the occupancy bit-plane neighbour search")
    ap.add_argument("--ws-sor", dest="ws_sor", action="store_true", help="also filter every frame's fused workspace (2cams_mask_cpu.py:530)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- rt3d has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(local_rank)   # before the pinned buffers are allocated (first touch decides their node)
    init_dist(world, dev)
    try:
        if args.impl == "torch_eager":
            lines = [run_torch_eager_arm(args, rank, local_rank, world)]
        elif args.config == 4:
            lines = [run_rig_arm(args, rank, local_rank, world)]
        elif args.config == 5:
            lines = []
            for v in (0.002, 0.005, 0.01):
                w = resolve(args)
                w["voxel"] = v
                lines.append(run_rt3d_arm(args, rank, local_rank, world, w=w, quiet_cpu=True))
        else:
            lines = [run_rt3d_arm(args, rank, local_rank, world)]
        for line in lines:
            if line is not None:
                print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
