# refresh of profiles/ for the committed state: tests, bench (+per-kernel event profile), ncu launch list
mkdir -p gpurun_out
date +%s > gpurun_out/t0
( time timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
date +%s > gpurun_out/t1
timeout 300 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/step_profile.json > gpurun_out/bench.json 2> gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench.json
date +%s > gpurun_out/t2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
date +%s > gpurun_out/t3
wc -l gpurun_out/launches.csv
