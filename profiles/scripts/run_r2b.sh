# round 2, call b: new pointwise kernels (separable resize, tiled JND blend): whole GPU suite, default bench (+768 HBM leg), video 768 bench
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/r2b_pytest.log 2>&1; tail -5 gpurun_out/r2b_pytest.log; grep "\[parity\]\|^FAILED\|^ERROR" gpurun_out/r2b_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/r2b_step_profile.json > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2b_bench.json')); print(b['value'], b['ms_per_step'], b['e2e']['value']); print(json.dumps(b['roofline_hbm'])[:3000])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --video --size 768 --profile-out gpurun_out/r2b_step_profile_video768.json > gpurun_out/r2b_bench_video768.json 2>gpurun_out/r2b_bench_video768_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2b_bench_video768.json')); print(b['value'], b['ms_per_step']); print(json.dumps(b['roofline_hbm'])[:3000])"
