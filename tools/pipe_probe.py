"""Per-stage CUDA-event timings of the batched frame pipeline at a given batch size (development aid).

    python tools/pipe_probe.py [F=64] [RES=HD720] [VOXEL=0.005] [N_OBJ=1] [STEPS=10]

Run it under `timeout`: a kernel that never finishes must not hold the GPU box."""
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rt-3d-reconstruction_b200"))
from rt3d import synthetic as syn  # noqa: E402
from rt3d.pipeline import FramePipeline  # noqa: E402


def main(F=64, res="HD720", voxel=0.005, n_obj=1, steps=10):
    dev = torch.device("cuda:0")
    D = min(F, 8)
    frames = [syn.frame(res, seed=1000 + 2 * i, n_objects=n_obj, jitter=0.05) for i in range(D)]
    reps = (F + D - 1) // D
    depth = torch.from_numpy(np.stack([f["depth"] for f in frames] * reps)[:F]).to(dev)
    masks = torch.from_numpy(np.stack([f["masks"] for f in frames] * reps)[:F]).to(dev) if n_obj else None
    H, W = frames[0]["H"], frames[0]["W"]
    pipe = FramePipeline(F, 2, H, W, n_obj, frames[0]["intr"], frames[0]["R"], frames[0]["t"], syn.WORKSPACE_BOX, voxel=voxel,
                         device=dev, sor_bitplanes=(False if os.environ.get("PROBE_GRID_ONLY") else True if os.environ.get("PROBE_BITPLANES") else None)).enable_stage_timing(steps)
    print(f"probe: F={F} {res} voxel={voxel} n_obj={n_obj}: first run ...", flush=True)
    pipe.run(depth, masks)
    torch.cuda.synchronize()
    print("probe: first run done, counts", pipe.counts.tolist(), flush=True)
    pipe.run(depth, masks)
    torch.cuda.synchronize()
    print("probe: second run (synchronised) done, counts", pipe.counts.tolist(), flush=True)
    for _ in range(2):
        pipe.run(depth, masks)
    torch.cuda.synchronize()
    print("probe: two runs back to back done, counts", pipe.counts.tolist(), flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    status = torch.zeros(steps, dtype=torch.int32, device=dev)
    e0.record()
    marks[0].record()
    for i in range(steps):
        pipe.run(depth, masks, slot=i)
        status[i:i + 1].copy_(pipe.counts[:1])
        marks[i + 1].record()
    e1.record()
    torch.cuda.synchronize()
    print("probe: per-step ms", [round(marks[i].elapsed_time(marks[i + 1]), 2) for i in range(steps)], "status", status.tolist(),
          "project ms", [round(pipe.stage_ms(i)["project"], 2) for i in range(steps)], flush=True)
    ms = e0.elapsed_time(e1) / steps
    st = {k: statistics.median(pipe.stage_ms(i)[k] for i in range(steps)) for k in pipe.STAGES}
    pb = F * 2 * H * W * (4 + n_obj) + 8 * pipe.counts.tolist()[1]
    print("probe: final counts", pipe.counts.tolist(), flush=True)
    print(f"probe: {ms:.3f} ms/step, {F / ms * 1e3:.0f} frames/s, stages {', '.join(f'{k} {v:.3f}' for k, v in st.items())}; "
          f"project {pb / st['project'] / 1e6:.0f} GB/s = {pb / st['project'] / 1e6 / 6545.3:.3f} of 6545.3", flush=True)


if __name__ == "__main__":
    a = sys.argv[1:]
    main(int(a[0]) if a else 64, a[1] if len(a) > 1 else "HD720", float(a[2]) if len(a) > 2 else 0.005,
         int(a[3]) if len(a) > 3 else 1, int(a[4]) if len(a) > 4 else 10)
