timeout 300 python tests/gpu_diag.py --net videoseal_1.0 2>&1 | grep RESULT | grep -E '"down0|"down1|up2_conv|delta|imgs_w|logits' | cut -c1-200
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/prof_dwc.json 2>gpurun_out/bench_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']['value'], d['e2e']['u8_frames']['value'])"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -c 1 -o gpurun_out/prof_conv1_bott -f python tests/prof_cases.py p_conv1_bott 2>&1 | tail -1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -c 1 -o gpurun_out/prof_pw1_96 -f python tests/prof_cases.py p_pw1_96 2>&1 | tail -1
