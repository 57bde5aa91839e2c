# round 2, call v: resize staging rounds as single bulk copies (mbarrier ring, 2 rounds in flight)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_config_size_gpu.py -m gpu -q -x -k "not chunky" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-cpu-baseline --no-e2e --profile-out gpurun_out/r2v_step_profile.json > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2v_bench.json')); print(b['value'], b['ms_per_step']); print(json.dumps(b['roofline_hbm']['kernels']))"
timeout 600 python bench.py --video --size 768 --steps 10 --warmup 3 --no-clip-leg --no-cpu-baseline --no-e2e --no-hbm-leg > gpurun_out/r2v_bench_video768.json 2> gpurun_out/r2v_bench_video768_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2v_bench_video768.json')); print(b['value'], b['ms_per_step'])"
