#!/bin/bash
# Development aid: the evidence set committed under profiles/ (run on the GPU box through gpurun; every ncu capture only
# after the same command has exited 0 without ncu).  usage: bash tools/gpu_evidence_run.sh <tag>
t=${1:-r05}; o=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $o/${t}_pytest_gpu.log 2>&1; tail -3 $o/${t}_pytest_gpu.log
timeout 600 python bench.py --latency --ceiling > $o/${t}_bench_c2.json 2> $o/${t}_bench_c2.err
timeout 600 python bench.py --config 3 --latency > $o/${t}_bench_c3.json 2>/dev/null
timeout 400 python bench.py --config 5 --steps 10 --no-cpu-baseline > $o/${t}_bench_c5.json 2>/dev/null
timeout 300 python bench.py --res HD1080 --frames 64 --no-cpu-baseline --steps 10 > $o/${t}_bench_hd1080_5mm.json 2>/dev/null
timeout 300 python bench.py --config 4 --steps 5 > $o/${t}_bench_c4.json 2>/dev/null
timeout 300 python bench.py --frames 64 --voxel 0.01 --thr 0.06 --ws-sor --no-cpu-baseline --steps 10 > $o/${t}_bench_ws_sor.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${t}_bench_reference_arm.json 2>/dev/null
B="python bench.py --no-cpu-baseline --steps 2 --warmup 3"
timeout 300 $B > $o/${t}_plain.json 2>/dev/null || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 700 --csv --log-file $o/${t}_launches_f256.csv $B > $o/${t}_ncu_list.log 2>&1
for k in fb_project_tma:3 sub_flag:3 fb_sor_knn:3 fb_sor_lattice:3 fb_ws_emit:3; do
  n=${k%%:*}; s=${k##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$n -s $s -c 1 -f -o $o/${t}_$n $B > $o/${t}_ncu_$n.log 2>&1
done
