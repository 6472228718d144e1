mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bwd.py tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider --timeout=600 > gpurun_out/pytest_bwd.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_bwd.log
for bs in 16 64; do BS=$bs timeout 300 python scripts/bwd_bench.py 2>&1 | grep -v "ours vs reference"; done | tee gpurun_out/bwd_bench.log
ZG_SCAN_BWD_STAGED=0 BS=16 timeout 300 python scripts/bwd_bench.py 2>&1 | head -2 | tee -a gpurun_out/bwd_bench.log
BS=16 DTYPE=bf16 timeout 600 python scripts/train_bench.py 2>&1 | grep "^{" | cut -c1-300
