# round 2, call g: blend CD template, batched GRN partial sums, deterministic head pool; ncu of the small-K GEMM epilogues
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2g_pytest.log 2>&1; grep -n "passed\|failed" gpurun_out/r2g_pytest.log | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-cpu-baseline --profile-out gpurun_out/r2g_step_profile.json > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2g_bench.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['u8_frames']['value']); h=b['roofline_hbm']; print([(k['name'], round(k['avg_us'],1), round(k['frac'],3)) for k in h['kernels']])
for r in b['top_kernels']: print(r['name'], r['avg_us'], r['launches_per_step'], r['ms_per_step'])" || tail -5 gpurun_out/r2g_bench_err.log
timeout 300 python tests/prof_cases.py --time p_conv1_bott p_pw1_384 p_pw2_384 p_pw1_96 p_conv3_bott 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel" --launch-skip 0 --launch-count 2 -o gpurun_out/r2g_gemm python tests/prof_cases.py p_conv1_bott p_pw2_384 > gpurun_out/r2g_ncu_gemm.log 2>&1; tail -1 gpurun_out/r2g_ncu_gemm.log
