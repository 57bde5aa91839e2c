# round 2, call n: A/B of the ring depthwise kernel with its weights in shared memory (3 blocks per SM)
mkdir -p gpurun_out
VSB_DW_WS=1 timeout 600 python -m pytest tests/test_e2e_gpu.py tests/test_config_size_gpu.py -m gpu -q -x -k "v1_image or golden or config1 or repeated" 2>&1 | tail -3
VSB_DW_WS=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-cpu-baseline --no-hbm-leg --no-e2e --profile-out gpurun_out/r2n_step_profile_ws.json > gpurun_out/r2n_bench_ws.json 2> gpurun_out/r2n_bench_ws_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2n_bench_ws.json')); print('WS', b['value'], b['ms_per_step'])
d=json.load(open('gpurun_out/r2n_step_profile_ws.json'))
for r in d['table']:
    if 'dwconv' in r['name']: print(r['name'], round(r['avg_us'],1), r['launches_per_step'])"
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-cpu-baseline --no-hbm-leg --no-e2e --profile-out gpurun_out/r2n_step_profile.json > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2n_bench.json')); print('regs', b['value'], b['ms_per_step'])
d=json.load(open('gpurun_out/r2n_step_profile.json'))
for r in d['table']:
    if 'dwconv' in r['name']: print(r['name'], round(r['avg_us'],1), r['launches_per_step'])"
