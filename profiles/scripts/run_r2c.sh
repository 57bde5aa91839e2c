# round 2, call c: second-generation pointwise kernels + hygiene + pipelined host entry + clip leg at N=1; ncu of the HBM kernels
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2c_pytest.log 2>&1; tail -4 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/r2c_step_profile.json > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2c_bench.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['u8_frames']['value']); print(json.dumps(b['roofline_hbm'])[:1500]); print(json.dumps(b['clip']))" || tail -5 gpurun_out/r2c_bench_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-clip-leg --video --size 768 > gpurun_out/r2c_bench_video768.json 2>gpurun_out/r2c_bench_video768_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2c_bench_video768.json')); print(b['value'], b['ms_per_step']); print(json.dumps(b['roofline_hbm'])[:1500])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jnd_blend2|resize_sep" --launch-skip 3 --launch-count 3 -o gpurun_out/r2c_pointwise python tests/prof_pointwise.py > gpurun_out/r2c_ncu_pw.log 2>&1; tail -2 gpurun_out/r2c_ncu_pw.log
