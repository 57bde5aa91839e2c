# round 2, call z: ncu --set full of the full-resolution kernels on exactly the roofline_hbm workload (image mode, 32 x 3x768x768)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"jnd_blend3|resize_sep" --launch-skip 3 --launch-count 3 -f -o gpurun_out/final_pointwise_img32 python tests/prof_pointwise.py 32 image > gpurun_out/ncu_pw32.log 2>&1; tail -1 gpurun_out/ncu_pw32.log
