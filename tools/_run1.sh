t=r04a; o=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > $o/${t}_pytest_gpu.log 2>&1; tail -3 $o/${t}_pytest_gpu.log
B="python bench.py --no-cpu-baseline --steps 10 --warmup 3"
timeout 300 $B > $o/${t}_plain.json 2>$o/${t}_plain.err || exit 1
cat $o/${t}_plain.json
B="python bench.py --no-cpu-baseline --steps 2 --warmup 3"
for k in fb_project_tma:3 sub_flag:3 fb_sor_knn:3 fb_sor_lattice:3; do
  n=${k%%:*}; s=${k##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$n -s $s -c 1 -f -o $o/${t}_$n $B > $o/${t}_ncu_$n.log 2>&1
done
ls -la $o
