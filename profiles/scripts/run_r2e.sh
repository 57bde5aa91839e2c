# round 2, call e: pointwise kernels v3 (predicate-free resize, vertical-first JND), facade GPU tests; ncu of ring dwconv + pwconv1
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2e_pytest.log 2>&1; tail -4 gpurun_out/r2e_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-clip-leg --video --size 768 > gpurun_out/r2e_bench_video768.json 2>gpurun_out/r2e_bench_video768_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2e_bench_video768.json')); print(b['value'], b['ms_per_step']); print(json.dumps(b['roofline_hbm'])[:1500])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dwconv7_ln_ring|conv_gemm_kernel" --launch-skip 1 --launch-count 2 -o gpurun_out/r2e_cnx python tests/prof_detect.py > gpurun_out/r2e_ncu_cnx.log 2>&1; tail -2 gpurun_out/r2e_ncu_cnx.log
