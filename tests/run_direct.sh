CASES=direct_c16_w64,direct_c16_w256,direct_c32,direct_c64,direct_c32_n16,direct_outc,direct_outc3,direct_ringwrap,direct_wrap_c64
timeout 300 python tests/gpu_diag.py --cases $CASES --skip-net 2>&1 | tail -10
timeout 120 python tests/prof_cases.py --time p_direct_c16 p_direct_c32 p_direct_c64 2>&1 | tail -4
timeout 200 ncu --set full --import-source on --clock-control none -k regex:conv3_direct -c 1 -o gpurun_out/prof_direct_c16 -f python tests/prof_cases.py p_direct_c16 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/prof_direct.json 2>gpurun_out/bench_err.log | tail -1 | cut -c1-300
