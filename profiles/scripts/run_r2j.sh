# round 2, call j: JND channel variants, all-warp LN epilogue, border-fix loop restored
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2j_pytest.log 2>&1; grep -n "passed\|failed" gpurun_out/r2j_pytest.log | tail -2; grep "^FAILED" gpurun_out/r2j_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-cpu-baseline --profile-out gpurun_out/r2j_step_profile.json > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2j_bench.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['u8_frames']['value']); h=b['roofline_hbm']; print([(k['name'], round(k['avg_us'],1), round(k['frac'],3)) for k in h['kernels']])
d=json.load(open('gpurun_out/r2j_step_profile.json'))
for r in d['table']:
    if any(k in r['name'] for k in ('upborder','stem','unet.first')): print(r['name'], round(r['avg_us'],1), r['launches_per_step'])" || tail -5 gpurun_out/r2j_bench_err.log
