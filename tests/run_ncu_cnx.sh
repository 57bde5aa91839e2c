# one ncu --set full capture of the two extractor kernels furthest from their roofline (profiles/r1_layer_roofline.md):
# the first dwconv7+LN (96 channels @ 64x64) and the first pwconv1 (GELU + GRN statistics epilogue) of a 64-frame detect()
mkdir -p gpurun_out
timeout 170 ncu --set full --import-source on --clock-control none -k regex:"dwconv7_ln_c_kernel|conv_gemm_kernel" --launch-skip 1 --launch-count 2 -o gpurun_out/ncu_cnx -f python tests/prof_detect.py > gpurun_out/ncu_cnx.log 2>&1
tail -3 gpurun_out/ncu_cnx.log
ncu -i gpurun_out/ncu_cnx.ncu-rep --page raw --csv > gpurun_out/ncu_cnx_raw.csv 2> gpurun_out/ncu_cnx_raw.err; wc -c gpurun_out/ncu_cnx_raw.csv; ls -la gpurun_out/ncu_cnx.ncu-rep
