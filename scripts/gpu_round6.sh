#!/bin/bash
# GPU visit: mixed-warp scan (scan_fwd_wph.cuh, ZG_SCAN_WP=5): bit-identity, timings, ncu, sanitizer, suite, quick bench.
mkdir -p gpurun_out
sw() { FUSED=0 timeout 200 python scripts/scan_sweep.py 2>&1 | tail -1; }
echo "== pytest (scan: warp-private pipelines)"
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "warp_private" -p no:cacheprovider --timeout=800 > gpurun_out/pytest_wph.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_wph.log
echo "== timings (config 2 layer shape)"
( ZG_SCAN_WP=0 sw
  ZG_SCAN_WP=3 sw
  ZG_SCAN_WP=5 sw
  ZG_SCAN_WP=5 ZG_SCAN_WPH_ND=9 ZG_SCAN_WPH_NS=0 sw
  ZG_SCAN_WP=5 ZG_SCAN_WPH_ND=7 ZG_SCAN_WPH_NS=2 sw
  ZG_SCAN_WP=5 ZG_SCAN_WPH_ND=6 ZG_SCAN_WPH_NS=4 sw
  ZG_SCAN_WP=5 sw
  ZG_SCAN_WP=0 sw
) | tee gpurun_out/scan_wph_sweep.log
echo "== ncu full (ZG_SCAN_WP=5)"
ZG_SCAN_WP=5 FUSED=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_ -s 3 -c 1 -f -o gpurun_out/r02e_scan_wp5 python scripts/scan_sweep.py > gpurun_out/ncu_wp5.log 2>&1; echo "ncu rc=$?"
echo "== compute-sanitizer memcheck (mode 5)"
ZG_SCAN_WP=5 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "tma_pipeline or out_reverse or temporal_layout" -p no:cacheprovider > gpurun_out/sanitizer_wp5.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitizer_wp5.log | tail -5
echo "== pytest -m gpu (all) with ZG_SCAN_WP=5"
ZG_SCAN_WP=5 timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 > gpurun_out/pytest_gpu_wp5.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_wp5.log
echo "== quick bench lines (no side measurements), alternating"
for wp in 0 5 0 5; do
  ZG_SCAN_WP=$wp timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-train --no-configs > gpurun_out/bench_wp$wp.json 2> gpurun_out/bench_wp$wp.err; echo "bench wp=$wp rc=$?"
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/bench_wp$wp.json").read().strip().splitlines()[-1])
    print("wp=$wp", d["ms_per_step"], "ms/step", d["value"], d["unit"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"])
except Exception as ex:
    print("bench parse failed", ex)
P
done
echo done
