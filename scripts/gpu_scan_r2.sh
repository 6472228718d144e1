#!/bin/bash
# Round-2 GPU visit for the scan kernel: parity tests, timings of the kernel variants, sanitizer, ncu captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
echo "== pytest (scan)"; timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "scan" -p no:cacheprovider --timeout=600 > gpurun_out/pytest_scan.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_scan.log
echo "== pytest (scan, ZG_SCAN_PROD=1)"; ZG_SCAN_PROD=1 timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "scan" -p no:cacheprovider --timeout=600 > gpurun_out/pytest_scan_prod1.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_scan_prod1.log
echo "== timings"
( ZG_SCAN_TMA=0 FUSED=0 timeout 200 python scripts/scan_sweep.py | tail -1
  for pr in ${PROD_SET:-2 1}; do for n in ${NPOLY_SET:-0 1}; do ZG_SCAN_PROD=$pr ZG_SCAN_TMA_NPOLY=$n timeout 200 python scripts/scan_sweep.py | tail -1; done; done
  for cfg in ${SHAPES:-"16 4096 1536" "1024 256 1536" "4096 16 1536"}; do set -- $cfg; BS=$1 SEQ=$2 EDIM=$3 timeout 200 python scripts/scan_sweep.py | tail -1; BS=$1 SEQ=$2 EDIM=$3 ZG_SCAN_TMA=0 FUSED=0 timeout 200 python scripts/scan_sweep.py | tail -1; done
) 2>&1 | tee gpurun_out/scan_sweep.log
if [ -n "$DO_SANITIZER" ]; then
echo "== compute-sanitizer"; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "tma_pipeline or fused_dt" -p no:cacheprovider > gpurun_out/sanitizer.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitizer.log | tail -5
fi
echo "== pytest -m gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
if [ -n "$DO_BENCH" ]; then
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.log | cut -c1-1500; tail -5 gpurun_out/bench.err
fi
if [ -n "$DO_NCU" ]; then
echo "== ncu full (unfused, fused)"
FUSED=0 ZG_SCAN_TMA_NPOLY=${NCU_NPOLY:-0} timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_tma -s 3 -c 1 -o gpurun_out/r02_scan_tma python scripts/scan_sweep.py > gpurun_out/ncu1.log 2>&1; echo "ncu1 rc=$?"
ZG_SCAN_TMA_NPOLY=${NCU_NPOLY:-0} timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_tma -s 25 -c 1 -o gpurun_out/r02_scan_tma_fused python scripts/scan_sweep.py > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
fi
echo done
