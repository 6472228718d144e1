#!/bin/bash
# Round-2 (second half) GPU visit: the warp-private scan pipeline (scan_fwd_wp.cuh) -- bit-identity against the CTA-wide kernel,
# timings of its variants at the BASELINE layer shapes, ncu captures, then the whole GPU suite and quick bench lines with it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
echo "== pytest (scan: warp-private pipeline)"
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "warp_private" -p no:cacheprovider --timeout=500 > gpurun_out/pytest_wp.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_wp.log
echo "== timings (config 2 layer shape)"
sw() { FUSED=0 timeout 200 python scripts/scan_sweep.py 2>&1 | tail -1; }
( ZG_SCAN_WP=0 sw
  ZG_SCAN_WP=1 sw
  ZG_SCAN_WP=2 sw
  ZG_SCAN_WP=1 ZG_SCAN_WP_WARPS=4 sw
  ZG_SCAN_WP=1 ZG_SCAN_WP_WARPS=2 sw
  ZG_SCAN_WP=1 ZG_SCAN_WP_WARPS=7 sw
  ZG_SCAN_WP=1 ZG_SCAN_WP_NPOLY=1 sw
  ZG_SCAN_WP=0 ZG_SCAN_TMA_NPOLY=1 sw
  for lib in $(ls zigma_b200/lib/libzigma_exp*.so 2>/dev/null); do ZIGMA_B200_LIB=$PWD/$lib ZG_SCAN_WP=1 sw; done
  for cfg in "16 1024 1280" "32 4096 1536" "256 256 1536" "4096 16 1536"; do set -- $cfg
    for wp in 0 1; do BS=$1 SEQ=$2 EDIM=$3 ZG_SCAN_WP=$wp sw; done; done
) | tee gpurun_out/scan_wp_sweep.log
echo "== ncu full"
for v in "1 0" "0 0" "1 1"; do set -- $v
  ZG_SCAN_WP=$1 ZG_SCAN_WP_NPOLY=$2 FUSED=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_ -s 3 -c 1 -f -o gpurun_out/r02b_scan_wp$1_np$2 python scripts/scan_sweep.py > gpurun_out/ncu_wp$1_np$2.log 2>&1; echo "ncu wp=$1 npoly=$2 rc=$?"
done
echo "== pytest -m gpu (all) with the warp-private scan"
ZG_SCAN_WP=1 timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 > gpurun_out/pytest_gpu_wp1.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_wp1.log
echo "== quick bench lines (no side measurements)"
for wp in 0 1; do
  ZG_SCAN_WP=$wp timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-train --no-configs > gpurun_out/bench_wp$wp.json 2> gpurun_out/bench_wp$wp.err; echo "bench wp=$wp rc=$?"
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/bench_wp$wp.json").read().strip().splitlines()[-1])
    print("wp=$wp", d["ms_per_step"], "ms/step", d["value"], d["unit"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"])
except Exception as ex:
    print("bench wp=$wp parse failed", ex)
P
done
echo done
