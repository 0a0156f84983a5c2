timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r03f_pytest_gpu.log 2>&1; tail -3 gpurun_out/r03f_pytest_gpu.log
timeout 600 python bench.py --latency --ceiling > gpurun_out/r03f_bench_c2.json 2> gpurun_out/r03f_bench_c2.err
timeout 400 python bench.py --config 5 --steps 10 --no-cpu-baseline > gpurun_out/r03f_bench_c5.json 2>/dev/null
timeout 400 python bench.py --config 5 --steps 10 --no-cpu-baseline --sor-grid-only > gpurun_out/r03f_bench_c5_grid.json 2>/dev/null
timeout 600 python bench.py --config 3 --latency --no-cpu-baseline > gpurun_out/r03f_bench_c3.json 2>/dev/null
timeout 300 python bench.py --res HD1080 --frames 64 --no-cpu-baseline --steps 10 > gpurun_out/r03f_bench_hd1080_5mm.json 2>/dev/null
