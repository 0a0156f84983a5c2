#!/usr/bin/env python
"""bench.py -- fused 2-camera HD720 frames/s @ 5 mm voxels (BASELINE.json metric, configs[1]).

One "step" = one batch of F synthetic 2-camera frames through the whole geometric path
(back-project -> transform -> crop -> voxels -> SOR fusion -> workspace subtraction) of
rt3d_frame_batch_run.  `value` is device-resident throughput (inputs already in HBM, CUDA-event
timed, max over ranks); `e2e` feeds the same call from pinned host buffers with the H2D / D2H copies
inside the timed region.  `--impl reference` times the CPU restatement of the reference path
(oracle/, the "port" -- Open3D cannot be installed, the reference is pure Python) on the host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl rt3d|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...   (one rank per GPU, frames sharded, no collective)
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

METRIC = "fused_2cam_hd720_frames_per_sec_5mm"
UNIT = "frames/s"
WORKLOAD = "configs[1]: 2 cameras HD720 (1280x720 fp32 depth + 1 uint8 instance mask each), 1 object, voxel 5 mm, " \
           "crop + transform + SOR(20,1.5) fusion + workspace subtraction thr 0.3 (2cams.py defaults)"
VOXEL, THR, UNIQUE = 0.005, 0.3, 8


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def project_traffic(frames):
    """dram__bytes_read.sum + dram__bytes_write.sum of one fb_project_kernel launch, from the committed
    `ncu --set full` capture of this very configuration (profiles/); None when no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "project_traffic.json")) as fh:
            for row in json.load(fh)["captures"]:
                if int(row["frames_per_step"]) == int(frames):
                    return float(row["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def make_frames(n_unique, seed0=1000):
    """config-2 scene with the object position jittered per frame (SURVEY.md 8d config 5 seeds)"""
    from rt3d import synthetic as syn
    return [syn.frame("HD720", seed=seed0 + 2 * i, n_objects=1, jitter=0.05) for i in range(n_unique)]


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
def cpu_port_fps(frames, cores, budget_s=20.0):
    """The oracle (CPU restatement of the reference path) on the host cores: frames are independent, so
    `cores` worker threads each run whole frames (ctypes releases the GIL; OpenMP inside is set to 1)."""
    from concurrent.futures import ThreadPoolExecutor

    import oracle
    from rt3d import synthetic as syn
    oracle.build()
    oracle.lib().rt3d_oracle_set_threads(1)

    def one(fr):
        cx, cy, fx, fy = fr["intr"]
        ones = np.ones(fr["depth"][0].shape, np.uint8)
        ws, objs = [], []
        for c in range(2):
            p = oracle.backproject_indices(oracle.mask_nonzero(ones), fr["depth"][c], cx, cy, fx, fy)
            ws.append(oracle.voxel_snap_unique(oracle.crop(oracle.transform(p, fr["R"][c], fr["t"][c]), *syn.WORKSPACE_BOX), VOXEL))
            objs.append(oracle.process_masks_cloud(fr["masks"][c, 0], fr["depth"][c], cx, cy, fx, fy, fr["R"][c], fr["t"][c], VOXEL))
        fused = oracle.filter_outliers_sor(np.vstack([o for o in objs if o is not None]))
        kept = oracle.subtract_point_clouds_gpu(np.vstack(ws), fused, THR)
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
    cores = os.cpu_count() or 1
    frames = make_frames(min(UNIQUE, 4))
    per_step = []
    sample = None
    for i in range(args.warmup + args.steps):
        fps, n, dt = cpu_port_fps(frames, cores, budget_s=max(2.0, 60.0 / max(1, args.warmup + args.steps)))
        sample = f"{n} fused HD720 frames per step on {cores} threads ({dt:.1f} s)"
        if i >= args.warmup:
            per_step.append((fps, dt))
    value = statistics.median(f for f, _ in per_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(d for _, d in per_step), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference path (oracle port); the reference's "
                   "Open3D/pyzed/ultralytics dependencies are not installable here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_rt3d_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from rt3d import synthetic as syn
    from rt3d.pipeline import FramePipeline

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- rt3d has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(local_rank)   # before the pinned buffers are allocated (first touch decides their node)
    if world > 1:
        # NCCL prints its version banner to stdout when the communicator is created (NCCL_DEBUG=VERSION or higher,
        # as set on some boxes): route fd 1 to stderr until that has happened, so stdout carries the one JSON line
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

    F = args.frames
    frames = make_frames(UNIQUE, seed0=1000 + 100 * rank)
    H, W = frames[0]["H"], frames[0]["W"]
    reps = (F + UNIQUE - 1) // UNIQUE
    depth_h = torch.from_numpy(np.stack([f["depth"] for f in frames] * reps)[:F]).contiguous().pin_memory()
    masks_h = torch.from_numpy(np.stack([f["masks"] for f in frames] * reps)[:F]).contiguous().pin_memory()
    depth_d, masks_d = depth_h.to(dev), masks_h.to(dev)
    pipe = FramePipeline(F, 2, H, W, 1, frames[0]["intr"], frames[0]["R"], frames[0]["t"], syn.WORKSPACE_BOX, voxel=VOXEL,
                         distance_threshold=THR, device=dev).enable_stage_timing(args.steps)

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

    # ---- end to end: pinned host buffers in, results out, copies inside the timed region
    counts = pipe.counts.tolist()
    exp_kept, exp_objs = counts[5], counts[3]
    host_batches = [(depth_h, masks_h)] * args.steps
    h2d, d2h = pipe.run_host_pipelined(host_batches[:2], exp_kept, exp_objs)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h2d, d2h = pipe.run_host_pipelined(host_batches, exp_kept, exp_objs)   # every step: H2D of its inputs + D2H of its results
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    t = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_frames = F * args.steps * world
    value = total_frames / (elapsed_ms * 1e-3)
    p50 = {k: statistics.median(v) for k, v in stage_ms.items()}
    peak, peak_src = measured_peaks()
    obj_px = counts[1]
    proj_bytes = F * 2 * H * W * (4 + 1) + 8 * obj_px          # depth + uint8 mask in, packed object keys out
    achieved = proj_bytes / (p50["project"] * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": f"synthetic ({UNIQUE} distinct ray-cast frames per rank, tiled to {F})",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": F, "l2_policy": "inputs larger than L2 "
                   f"({(depth_d.numel() * 4 + masks_d.numel()) / 1e6:.0f} MB per step vs 126 MB L2)",
                   "arith": os.environ.get("RT3D_ARITH", "ieee"), "parallelism": f"frames sharded x{world}, no collective"},
        "stage_p50_ms": p50,
        "points": {"object_px": obj_px, "object_voxels": counts[2], "fused": counts[3], "workspace_voxels": counts[4], "kept": counts[5]},
        "roofline": {"bound": "hbm", "kernel": "fb_project_kernel (stage A: back-project + transform + crop + voxel bit / key)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": project_traffic(F),
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": proj_bytes},
        "e2e": {"value": total_frames / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "path": "FramePipeline.run_host_pipelined: per step pinned host depth+masks -> H2D (copy stream, overlapping the "
                        "previous step's kernels) -> rt3d_frame_batch_run -> D2H of counts, offsets, kept and fused rows"},
        "gpu_launches": pipe.count_launches() * args.steps,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        if _FULL_AFFINITY:
            os.sched_setaffinity(0, _FULL_AFFINITY)   # the CPU arm gets every host core again
        cores = os.cpu_count() or 1
        fps, n, dt = cpu_port_fps(frames[:4], cores, budget_s=15.0)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n} fused HD720 frames on {cores} threads ({dt:.1f} s)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=256, help="fused frames per step per GPU (config 5 batches 256 HD720 frames)")
    ap.add_argument("--impl", default="rt3d", choices=["rt3d", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_rt3d_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
