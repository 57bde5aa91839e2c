timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/prof_dwc.json 2>gpurun_out/bench_err.log | tail -1 | cut -c1-300
