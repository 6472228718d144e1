mkdir -p gpurun_out
ZG_SCAN_TPC2_WARP=1 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_bwd.py -m gpu -q -x -p no:cacheprovider --timeout=600 -k "scan or model or zigma or sampler or token_major or staged" > gpurun_out/pytest_w.log 2>&1; echo "pytest(warp) rc=$?"; tail -3 gpurun_out/pytest_w.log | cut -c1-200
for w in 0 1; do echo "WARP=$w: $(ZG_SCAN_TPC2_WARP=$w timeout 200 python scripts/scan_sweep.py | tail -1)"; done
for w in 0 1; do echo "WARP=$w bench: $(ZG_SCAN_TPC2_WARP=$w timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-train 2>/dev/null | tail -1 | cut -c1-260)"; done
for w in 0 1; do echo "WARP=$w train-fwd: $(ZG_SCAN_TPC2_WARP=$w BS=16 timeout 300 python scripts/bwd_bench.py 2>&1 | grep 'token-major scan')"; done
