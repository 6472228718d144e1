mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train.py tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider --timeout=600 > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_train.log
for bs in 16 64; do BS=$bs DTYPE=bf16 timeout 600 python scripts/train_bench.py 2>&1 | grep "^{" | cut -c1-420; done | tee gpurun_out/train_bench_bs.log
BS=16 DTYPE=amp timeout 600 python scripts/train_bench.py 2>&1 | grep "^{" | cut -c1-700 | tee -a gpurun_out/train_bench_bs.log
