# round 2, call i: border-fix loads batched, unet.first 4 px x 8 ch per thread, smem-tiled up-conv gather, 128-bit fixed-order GRN sums
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2i_pytest.log 2>&1; grep -n "passed\|failed" gpurun_out/r2i_pytest.log | tail -2; grep "^FAILED" gpurun_out/r2i_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-cpu-baseline --no-hbm-leg --profile-out gpurun_out/r2i_step_profile.json > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2i_bench.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['u8_frames']['value'])
d=json.load(open('gpurun_out/r2i_step_profile.json'))
for r in d['table']:
    if any(k in r['name'] for k in ('upborder','upgather','unet.first','grn_','head_tail')): print(r['name'], round(r['avg_us'],1), r['launches_per_step'])" || tail -5 gpurun_out/r2i_bench_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-clip-leg --no-cpu-baseline --no-hbm-leg --no-e2e --card pixelseal --batch 32 --size 768 > gpurun_out/r2i_bench_pixelseal768.json 2>/dev/null; python -c "
import json; b=json.load(open('gpurun_out/r2i_bench_pixelseal768.json')); print('pixelseal768', b['value'], b['ms_per_step'])"
