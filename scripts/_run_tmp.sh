mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bwd.py tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider --timeout=600 -k "conv or token_major" > gpurun_out/pytest_bwd.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_bwd.log | cut -c1-300
for f in 1 0; do echo "FAST=$f: $(ZG_CONV_BWD_FAST=$f BS=16 timeout 300 python scripts/bwd_bench.py 2>&1 | grep 'token-major conv')  bs64: $(ZG_CONV_BWD_FAST=$f BS=64 timeout 300 python scripts/bwd_bench.py 2>&1 | grep 'token-major conv')"; done
BS=16 DTYPE=bf16 PROFILE=1 timeout 600 python scripts/train_bench.py 2>&1 | grep -E "^\{|block_tail|conv_bwd|Self CUDA time" | cut -c1-72,150-222
