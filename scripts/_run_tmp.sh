mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bwd.py tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider --timeout=600 > gpurun_out/pytest_bwd.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_bwd.log
for d in bf16 fp32; do BS=16 DTYPE=$d PROFILE=1 timeout 600 python scripts/train_bench.py 2>&1 | grep -v "^$" | cut -c1-220; done > gpurun_out/train_bench.log 2>&1
grep "^{" gpurun_out/train_bench.log | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"scan_bwd_q4|scan_fwd_tpc2|conv_bwd_tok|add_norm_bwd" -c 4 -o gpurun_out/bwd_kernels python scripts/profile_bwd.py > gpurun_out/ncu_bwd.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_bwd.log
