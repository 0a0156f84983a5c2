timeout 600 python -m pytest tests -x -q -m gpu -k "subtract or batch or pipeline or fusion" > gpurun_out/r03j_pytest_gpu.log 2>&1; tail -3 gpurun_out/r03j_pytest_gpu.log
for v in "" _sub5 _sub4; do
  echo "== variant '$v'"
  RT3D_LIB_PATH=$PWD/rt-3d-reconstruction_b200/librt3d$v.so timeout 200 python tools/pipe_probe.py 256 HD720 0.005 1 10 2>&1 | grep "ms/step"
done
