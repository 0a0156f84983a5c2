timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r03h_pytest_gpu.log 2>&1; tail -5 gpurun_out/r03h_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r03h_bench_c2.json 2> gpurun_out/r03h_bench_c2.err
timeout 300 python bench.py --config 3 --no-cpu-baseline --steps 10 > gpurun_out/r03h_bench_c3.json 2>/dev/null
timeout 300 python bench.py --thr 0.06 --no-cpu-baseline --steps 10 > gpurun_out/r03h_bench_thr006.json 2>/dev/null
