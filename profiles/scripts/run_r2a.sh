# round 2, first call: the whole GPU suite incl. the config-size parity tests, the default bench, and the A/B of the opt-in build
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/r2a_pytest.log 2>&1; tail -5 gpurun_out/r2a_pytest.log; grep "\[parity\]" gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/r2a_step_profile.json > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench_err.log; tail -c 1500 gpurun_out/r2a_bench.json
VSB200_LIB=$PWD/videoseal_b200/libvsb200_pdl.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/r2a_step_profile_pdl.json > gpurun_out/r2a_bench_pdl.json 2> gpurun_out/r2a_bench_pdl_err.log; head -c 400 gpurun_out/r2a_bench_pdl.json; echo
VSB200_LIB=$PWD/videoseal_b200/libvsb200_pdl.so VSB_NO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_nopdl.json 2> /dev/null; head -c 400 gpurun_out/r2a_bench_nopdl.json; echo
VSB200_LIB=$PWD/videoseal_b200/libvsb200_pdl.so VSB_DW2=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/r2a_step_profile_dw2.json > gpurun_out/r2a_bench_dw2.json 2> gpurun_out/r2a_bench_dw2_err.log; head -c 400 gpurun_out/r2a_bench_dw2.json; echo
VSB200_LIB=$PWD/videoseal_b200/libvsb200_pdl.so VSB_DW2=1 timeout 300 python -m pytest tests/test_e2e_gpu.py -m gpu -x -q -k "v1_image or golden" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --video --size 768 --profile-out gpurun_out/r2a_step_profile_video768.json > gpurun_out/r2a_bench_video768.json 2>/dev/null; head -c 400 gpurun_out/r2a_bench_video768.json; echo
