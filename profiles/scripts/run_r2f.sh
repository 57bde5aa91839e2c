# round 2, call f: ring dwconv v2 (k-outer, shared emission, R=32), resident weight slabs for multi-N-tile GEMMs, bias from global (no
# epilogue barrier), deterministic GRN partial sums, facade fix; ncu of blend v3
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2f_pytest.log 2>&1; tail -4 gpurun_out/r2f_pytest.log | head -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --profile-out gpurun_out/r2f_step_profile.json > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2f_bench.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['u8_frames']['value'])
for r in b['top_kernels']: print(r['name'], r['avg_us'], r['launches_per_step'], r['ms_per_step'])" || tail -5 gpurun_out/r2f_bench_err.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jnd_blend2" --launch-skip 1 --launch-count 1 -o gpurun_out/r2f_blend python tests/prof_pointwise.py 32 > gpurun_out/r2f_ncu_blend.log 2>&1; tail -1 gpurun_out/r2f_ncu_blend.log
