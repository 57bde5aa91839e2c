timeout 300 python tests/gpu_diag.py --net videoseal_1.0 2>&1 | grep RESULT | grep -E '"up2_conv|delta|imgs_w|logits' | cut -c1-200
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/prof_dwc.json 2>gpurun_out/bench_err.log | tail -1 | cut -c1-300
