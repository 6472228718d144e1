#!/bin/bash
# Round 2, last GPU minutes: one `ncu --set full` capture each of the conv forward kernel and of the out_proj GEMM (the two kernels
# DESIGN section 6 still lists as below their roofline), final build, config-2 layer shapes.
mkdir -p gpurun_out
timeout 75 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_tok4 -s 3 -c 1 -f -o gpurun_out/r02e_conv python scripts/conv_sweep.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 75 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 26 -c 1 -f -o gpurun_out/r02e_outproj python scripts/gemm_bench.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
