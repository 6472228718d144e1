mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_bwd_tok|block_tail_bwd" -c 2 -o gpurun_out/conv_bwd python scripts/profile_bwd.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_conv.log
