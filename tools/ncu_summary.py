"""Summaries of an Nsight Compute report, as committed under profiles/ (run here: ncu reads .ncu-rep without a GPU).

    python tools/ncu_summary.py raw     gpurun_out/x.ncu-rep [extra-metric-regex]   # the per-kernel metric table
    python tools/ncu_summary.py regions gpurun_out/x.ncu-rep UNITS [min-share]      # SASS regions by executed count

`regions` groups consecutive SASS lines that executed (nearly) the same number of times, which is how the loops of a
kernel show up without source correlation: executed / UNITS = iterations per work item (UNITS = queries, row groups,
slabs, ... that the launch processed)."""
import csv
import re
import subprocess
import sys

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "lts__t_requests_srcunit_tex_op_atom_dot_alu.sum", "lts__t_requests_srcunit_tex_op_red.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "sm__cycles_elapsed.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
]


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(out.splitlines()))


def raw(rep, extra=None):
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("-----", r[idx["Kernel Name"]][:100])
        for h in hdr:
            stall = h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h
            if h in RAW_METRICS or stall or (extra and re.search(extra, h)):
                if r[idx[h]] not in ("", "0"):
                    print(f"{h:92s} {r[idx[h]]:>18s} {units[idx[h]]}")


def regions(rep, units, min_share=0.015):
    rows = page(rep, "source")
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx["Instructions Executed"]].isdigit()]
    total = sum(int(r[idx["Instructions Executed"]]) for r in data)
    print(f"{rows[0][1] if len(rows[0]) > 1 else ''}: {total} warp instructions, {total / units:.1f} per unit")
    runs = []
    for i, r in enumerate(data):
        e = int(r[idx["Instructions Executed"]])
        key = e / units
        if runs and abs(runs[-1][0] - key) <= 0.03 * max(key, 0.05):
            runs[-1][2] = i
            runs[-1][3] += e
        else:
            runs.append([key, i, i, e])
    for key, a, b, e in runs:
        if e / total > min_share:
            print(f"  lines {a:5d}-{b:5d} ({b - a + 1:4d})  executed/unit {key:8.2f}  {e / 1e6:8.2f} M ({100 * e / total:4.1f} %)  "
                  f"first: {data[a][idx['Source']].strip()[:60]}")


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("raw", "regions"):
        sys.exit(__doc__)
    if sys.argv[1] == "raw":
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        regions(sys.argv[2], float(sys.argv[3]), float(sys.argv[4]) if len(sys.argv) > 4 else 0.015)
