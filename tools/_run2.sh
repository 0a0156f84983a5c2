t=r04b; o=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $o/${t}_pytest_gpu.log 2>&1; tail -5 $o/${t}_pytest_gpu.log
B="python bench.py --no-cpu-baseline --steps 10 --warmup 3"
for v in "" _r8 _yfast; do
  RT3D_LIB_PATH=$PWD/rt-3d-reconstruction_b200/librt3d$v.so timeout 300 $B > $o/${t}_bench$v.json 2>$o/${t}_bench$v.err
  python - <<PY
import json
d=json.loads(open("$o/${t}_bench$v.json").read().strip().splitlines()[-1])
print("$v", d["ms_per_step"], d["stage_p50_ms"], d["roofline"]["frac"])
PY
done
