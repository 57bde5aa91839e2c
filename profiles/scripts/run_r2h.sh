# round 2, call h: 16 epilogue warps for all TMA GEMMs, per-stage deterministic GRN, blend/resize v3
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2h_pytest.log 2>&1; grep -n "passed\|failed" gpurun_out/r2h_pytest.log | tail -2; grep "^FAILED" gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-cpu-baseline --profile-out gpurun_out/r2h_step_profile.json > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2h_bench.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['u8_frames']['value']); h=b['roofline_hbm']; print([(k['name'], round(k['avg_us'],1), round(k['frac'],3)) for k in h['kernels']])
for r in b['top_kernels']: print(r['name'], r['avg_us'], r['launches_per_step'], r['ms_per_step'])" || tail -5 gpurun_out/r2h_bench_err.log
timeout 300 python tests/prof_cases.py --time p_conv1_bott p_pw1_384 p_pw2_384 p_pw1_96 p_conv3_bott 2>&1 | tail -6
