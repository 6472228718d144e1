mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bwd.py tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider --timeout=600 > gpurun_out/pytest_bwd.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_bwd.log
for bs in 16 64; do BS=$bs timeout 300 python scripts/bwd_bench.py; done 2>&1 | tee gpurun_out/bwd_bench.log
ZG_SCAN_BWD_Q4=0 BS=16 timeout 300 python scripts/bwd_bench.py 2>&1 | head -3 | tee -a gpurun_out/bwd_bench.log
