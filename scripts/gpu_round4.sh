#!/bin/bash
# GPU visit: CTA shape / fairness-barrier sweep of the warp-private scan pipeline, ncu of the best setting, tests, quick bench.
mkdir -p gpurun_out
sw() { FUSED=0 timeout 200 python scripts/scan_sweep.py 2>&1 | tail -1; }
echo "== timings (config 2 layer shape)"
( ZG_SCAN_WP=0 sw
  ZG_SCAN_WP=1 ZG_SCAN_WP_WARPS=4 ZG_SCAN_WP_SYNC=0 sw
  for w in 18 12; do for k in 0 1 2 4 8 16 32; do ZG_SCAN_WP=1 ZG_SCAN_WP_WARPS=$w ZG_SCAN_WP_SYNC=$k sw; done; done
  for w in 9 7 6; do for k in 0 4; do ZG_SCAN_WP=1 ZG_SCAN_WP_WARPS=$w ZG_SCAN_WP_SYNC=$k sw; done; done
  ZG_SCAN_WP=2 ZG_SCAN_WP_WARPS=18 ZG_SCAN_WP_SYNC=4 sw
  ZG_SCAN_WP=1 ZG_SCAN_WP_WARPS=18 ZG_SCAN_WP_SYNC=4 ZG_SCAN_WP_NPOLY=1 sw
) | tee gpurun_out/scan_wp_sweep2.log
best=$(python - <<'P'
import re
best = None
for line in open("gpurun_out/scan_wp_sweep2.log"):
    m = re.search(r"\[(.*?)\].*scan ([0-9.]+) ms", line)
    if not m or "ZG_SCAN_WP=1" not in m.group(1) or "NPOLY" in m.group(1): continue
    t = float(m.group(2))
    if best is None or t < best[0]: best = (t, m.group(1))
print(best[1] if best else "ZG_SCAN_WP=1")
P
)
echo "best: $best"
echo "$best" > gpurun_out/scan_wp_best.txt
echo "== other shapes with the default shape rule vs the CTA-wide kernel"
( for cfg in "16 1024 1280" "32 4096 1536" "256 256 1536" "4096 16 1536"; do set -- $cfg
    BS=$1 SEQ=$2 EDIM=$3 ZG_SCAN_WP=0 sw
    for k in 4 8; do BS=$1 SEQ=$2 EDIM=$3 ZG_SCAN_WP=1 ZG_SCAN_WP_SYNC=$k sw; done
  done ) | tee gpurun_out/scan_wp_shapes.log
echo "== ncu full of the best setting"
env $best FUSED=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_ -s 3 -c 1 -f -o gpurun_out/r02c_scan_wp_best python scripts/scan_sweep.py > gpurun_out/ncu_wp_best.log 2>&1; echo "ncu rc=$?"
echo "== pytest -m gpu (all) with the best setting"
env $best timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 > gpurun_out/pytest_gpu_wpbest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_wpbest.log
echo "== quick bench lines (no side measurements)"
for v in "ZG_SCAN_WP=0" "$best"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-train --no-configs > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench [$v] rc=$?"
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("[$v]", d["ms_per_step"], "ms/step", d["value"], d["unit"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"])
except Exception as ex:
    print("bench parse failed", ex)
P
done
echo done
