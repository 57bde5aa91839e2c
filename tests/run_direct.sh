CASES=direct1_c16,direct1_c32,direct1_c64,direct1_wrap,direct_c16_w256,direct_outc
timeout 300 python tests/gpu_diag.py --cases $CASES --skip-net 2>&1 | tail -8
timeout 120 python tests/prof_cases.py --time p_direct1_c16 p_direct_c16 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/prof_direct1.json 2>gpurun_out/bench_err.log | tail -1 | cut -c1-300
