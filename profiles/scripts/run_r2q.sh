# round 2, call q: source-level ncu of the blend kernel (instruction counts per line)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jnd_blend2" --launch-skip 1 --launch-count 1 -o gpurun_out/r2q_blend python tests/prof_pointwise.py > gpurun_out/r2q_ncu_blend.log 2>&1; tail -1 gpurun_out/r2q_ncu_blend.log
