# round 2, call x: why is the CTA-pair kernel at half rate?  A/B on tile width and pipeline depth, ncu of the kernel
mkdir -p gpurun_out
timeout 300 python tests/prof_cases.py --time p_conv3_bott 2>&1 | tail -1
VSB_PAIR_BN=128 timeout 300 python tests/prof_cases.py --time p_conv3_bott 2>&1 | tail -1
VSB_PAIR_STAGES=3 timeout 300 python tests/prof_cases.py --time p_conv3_bott 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_pair -c 1 -f -o gpurun_out/r2x_pair python tests/prof_cases.py p_conv3_bott > gpurun_out/r2x_ncu.log 2>&1; tail -1 gpurun_out/r2x_ncu.log
