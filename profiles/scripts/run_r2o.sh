# round 2, call o: residual read straight from global for the long-K convs (one more pipeline stage); ncu source view of the ring kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_conv_gemm_gpu.py tests/test_config_size_gpu.py -m gpu -q -x -k "not chunky" 2>&1 | tail -3
timeout 300 python tests/prof_cases.py --time p_conv3_bott p_conv1_bott 2>&1 | tail -3
VSB_RESID_RING=1 timeout 300 python tests/prof_cases.py --time p_conv3_bott 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-cpu-baseline --no-hbm-leg --no-e2e --profile-out gpurun_out/r2o_step_profile.json > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2o_bench.json')); print(b['value'], b['ms_per_step'])
for r in b['top_kernels'][:4]: print(r['name'], r['avg_us'], r['launches_per_step'], r['ms_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dwconv7_ln_ring" --launch-skip 3 --launch-count 1 -o gpurun_out/r2o_ring python tests/prof_detect.py > gpurun_out/r2o_ncu_ring.log 2>&1; tail -1 gpurun_out/r2o_ncu_ring.log
