#!/bin/bash
# GPU visit: two-channels-per-lane warp-private scan (scan_fwd_wp2.cuh, ZG_SCAN_WP=3|4): bit-identity, timings, ncu, sanitizer, suite, bench.
mkdir -p gpurun_out
sw() { FUSED=0 timeout 200 python scripts/scan_sweep.py 2>&1 | tail -1; }
echo "== pytest (scan: warp-private pipelines)"
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "warp_private" -p no:cacheprovider --timeout=800 > gpurun_out/pytest_wp2.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_wp2.log
echo "== timings (config 2 layer shape)"
( ZG_SCAN_WP=0 sw
  ZG_SCAN_WP=1 ZG_SCAN_WP_WARPS=18 sw
  ZG_SCAN_WP=3 sw
  ZG_SCAN_WP=4 sw
  for w in 2 3 5 6; do ZG_SCAN_WP=3 ZG_SCAN_WP_WARPS=$w sw; done
  ZG_SCAN_WP=3 sw
  ZG_SCAN_WP=0 sw
) | tee gpurun_out/scan_wp2_sweep.log
echo "== other shapes"
( for cfg in "16 1024 1280" "32 4096 1536" "256 256 1536" "4096 16 1536"; do set -- $cfg
    for wp in 0 3; do BS=$1 SEQ=$2 EDIM=$3 ZG_SCAN_WP=$wp sw; done
  done ) | tee gpurun_out/scan_wp2_shapes.log
echo "== ncu full (ZG_SCAN_WP=3)"
ZG_SCAN_WP=3 FUSED=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_ -s 3 -c 1 -f -o gpurun_out/r02d_scan_wp3 python scripts/scan_sweep.py > gpurun_out/ncu_wp3.log 2>&1; echo "ncu rc=$?"
echo "== compute-sanitizer memcheck (mode 3, bf16)"
ZG_SCAN_WP=3 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "tma_pipeline or out_reverse or temporal_layout" -p no:cacheprovider > gpurun_out/sanitizer_wp3.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitizer_wp3.log | tail -5
echo "== pytest -m gpu (all) with ZG_SCAN_WP=3"
ZG_SCAN_WP=3 timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 > gpurun_out/pytest_gpu_wp3.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_wp3.log
echo "== quick bench lines (no side measurements)"
for cfg in zigzag8_b1 faceshq1024; do for wp in 0 3; do
  ZG_SCAN_WP=$wp timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-train --no-configs > gpurun_out/bench_${cfg}_wp$wp.json 2> gpurun_out/bench_${cfg}_wp$wp.err; echo "bench $cfg wp=$wp rc=$?"
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/bench_${cfg}_wp$wp.json").read().strip().splitlines()[-1])
    print("$cfg wp=$wp", d["ms_per_step"], "ms/step", d["value"], d["unit"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"])
except Exception as ex:
    print("bench parse failed", ex)
P
done; done
echo done
