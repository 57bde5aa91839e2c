# round 2, call r: TMA-fed blend kernel (jnd_blend3_kernel), packed fp32 heat-map arithmetic
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_config_size_gpu.py -m gpu -q -x -k "not chunky" 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-cpu-baseline --no-e2e --profile-out gpurun_out/r2r_step_profile.json > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2r_bench.json')); print(b['value'], b['ms_per_step']); print(json.dumps(b['roofline_hbm']['kernels']))"
VSB_BLEND_OLD=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-clip-leg --no-cpu-baseline --no-e2e > gpurun_out/r2r_bench_old.json 2> gpurun_out/r2r_bench_old_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2r_bench_old.json')); print(b['value'], b['ms_per_step']); print(json.dumps(b['roofline_hbm']['kernels']))"
timeout 600 python bench.py --video --size 768 --steps 10 --warmup 3 --no-clip-leg --no-cpu-baseline --no-e2e --no-hbm-leg > gpurun_out/r2r_bench_video768.json 2> gpurun_out/r2r_bench_video768_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2r_bench_video768.json')); print(b['value'], b['ms_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jnd_blend3" --launch-skip 1 --launch-count 1 -f -o gpurun_out/r2r_blend3 python tests/prof_pointwise.py > gpurun_out/r2r_ncu_blend.log 2>&1; tail -1 gpurun_out/r2r_ncu_blend.log
