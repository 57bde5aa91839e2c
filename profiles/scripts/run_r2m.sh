# round 2, call m: swizzled staging boxes for the TMA-store epilogue
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2m_pytest.log 2>&1; grep -n "passed\|failed" gpurun_out/r2m_pytest.log | tail -2; grep "^FAILED" gpurun_out/r2m_pytest.log
timeout 300 python tests/prof_cases.py --time p_conv1_bott p_pw1_384 p_pw1_96 p_pw2_96t p_pw2_192t p_uptap_128 2>&1 | tail -7
VSB_TMA_STORE_NOSWZ=1 timeout 300 python tests/prof_cases.py --time p_conv1_bott p_pw1_384 p_pw1_96 p_pw2_96t p_pw2_192t p_uptap_128 2>&1 | tail -7
timeout 600 python bench.py --steps 20 --warmup 3 --no-clip-leg --no-cpu-baseline --no-hbm-leg --profile-out gpurun_out/r2m_step_profile.json > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench_err.log; python -c "
import json; b=json.load(open('gpurun_out/r2m_bench.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['u8_frames']['value'])
for r in b['top_kernels']: print(r['name'], r['avg_us'], r['launches_per_step'], r['ms_per_step'])" || tail -5 gpurun_out/r2m_bench_err.log
