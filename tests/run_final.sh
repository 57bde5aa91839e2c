# round 2, final refresh of profiles/ for the committed state: GPU suite, bench (+ event profile, HBM leg, clip leg, CPU baseline),
# the other BASELINE configs, ncu launch list, ncu --set full of the dominant tensor kernel and of the two full-resolution kernels
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/pytest.log 2>&1; grep -n "passed\|failed" gpurun_out/pytest.log | tail -2; grep "^FAILED" gpurun_out/pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/step_profile.json > gpurun_out/bench.json 2> gpurun_out/bench_err.log; tail -c 400 gpurun_out/bench.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-clip-leg --card pixelseal --batch 32 --size 768 > gpurun_out/bench_pixelseal768.json 2>/dev/null
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-clip-leg --video --size 768 > gpurun_out/bench_v1_video768.json 2>/dev/null
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-clip-leg --no-hbm-leg --card chunkyseal --batch 16 --size 512 > gpurun_out/bench_chunky.json 2>/dev/null
for f in pixelseal768 v1_video768 chunky; do python -c "
import json; b=json.load(open('gpurun_out/bench_$f.json')); print('$f', round(b['value']), round(b['ms_per_step'],2))"; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-clip-leg --no-hbm-leg > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_pair -c 1 -f -o gpurun_out/final_dominant python tests/prof_cases.py p_conv3_bott > gpurun_out/ncu_dom.log 2>&1; tail -1 gpurun_out/ncu_dom.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"jnd_blend3|resize_sep" --launch-skip 3 --launch-count 3 -f -o gpurun_out/final_pointwise python tests/prof_pointwise.py > gpurun_out/ncu_pw.log 2>&1; tail -1 gpurun_out/ncu_pw.log
