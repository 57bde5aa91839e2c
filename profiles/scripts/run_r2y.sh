# round 2, call y: CTA-pair kernel without the cluster-scope release fences on the remote arrivals
mkdir -p gpurun_out
timeout 300 python tests/prof_cases.py --time p_conv3_bott 2>&1 | tail -1
VSB_PAIR_STAGES=4 timeout 300 python tests/prof_cases.py --time p_conv3_bott 2>&1 | tail -1
VSB_NO_PAIR=1 timeout 300 python tests/prof_cases.py --time p_conv3_bott 2>&1 | tail -1
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_e2e_gpu.py tests/test_config_size_gpu.py -m gpu -q -x -k "not chunky" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-cpu-baseline --no-e2e --no-hbm-leg --profile-out gpurun_out/r2y_step_profile.json > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2y_bench.json')); print(b['value'], b['ms_per_step'])
for r in b['top_kernels'][:3]: print(r['name'], r['avg_us'], r['launches_per_step'], r.get('tflops'))"
