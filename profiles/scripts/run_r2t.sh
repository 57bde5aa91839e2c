# round 2, call t: source-level ncu of the separable resize kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"resize_sep" --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2t_resize python tests/prof_pointwise.py 32 > gpurun_out/r2t_ncu.log 2>&1; tail -1 gpurun_out/r2t_ncu.log
