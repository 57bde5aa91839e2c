# round 2, call d: ring dwconv (TMA-fed register rolling), single-MUFU packed GELU epilogue, PDL default
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2d_pytest.log 2>&1; tail -4 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-hbm-leg --profile-out gpurun_out/r2d_step_profile.json > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2d_bench.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['u8_frames']['value'])
for r in b['top_kernels']: print(r['name'], r['avg_us'], r['launches_per_step'], r['ms_per_step'])" || tail -5 gpurun_out/r2d_bench_err.log
VSB_DW_NO_RING=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-hbm-leg --no-e2e --no-cpu-baseline --profile-out gpurun_out/r2d_step_profile_noring.json > gpurun_out/r2d_bench_noring.json 2> /dev/null; python -c "
import json; b=json.load(open('gpurun_out/r2d_bench_noring.json')); print('noring', b['value'], b['ms_per_step'])"
