# chunkyseal bring-up on the GPU: the regular suite first, then the chunkyseal-width tests in their own process (a hung or
# trapped kernel must not take the rest down), then the layer walk if they failed
mkdir -p gpurun_out
( time timeout 240 python -m pytest tests -m gpu -q -k "not chunkyseal" ) > gpurun_out/pytest_main.log 2>&1; echo "main rc=$?" ; tail -4 gpurun_out/pytest_main.log
( time timeout 200 python -m pytest tests/test_e2e_gpu.py -m gpu -q -k chunkyseal ) > gpurun_out/pytest_chunky.log 2>&1; rc=$?; echo "chunky rc=$rc"; tail -30 gpurun_out/pytest_chunky.log | cut -c1-400
if [ $rc -ne 0 ]; then
  timeout -s KILL 150 python tests/gpu_diag.py --net chunkyseal:tiny > gpurun_out/diag_chunky.log 2>&1; rc2=$?; echo "diag rc=$rc2"; cut -c1-300 gpurun_out/diag_chunky.log | tail -60
  if [ $rc2 -eq 137 ]; then
    VSB_NO_DIRECT1=1 timeout -s KILL 150 python tests/gpu_diag.py --net chunkyseal:tiny > gpurun_out/diag_chunky_nodirect1.log 2>&1; echo "diag(no direct 1x1) rc=$?"; cut -c1-300 gpurun_out/diag_chunky_nodirect1.log | tail -60
  fi
fi
