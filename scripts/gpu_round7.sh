#!/bin/bash
# GPU visit: the per-call kernel choice (scan_auto_choice) + L2 hint, and the FMA-pipe reciprocal experiment (libzigma_exprcpfma.so).
mkdir -p gpurun_out
EXP=$PWD/zigma_b200/lib/libzigma_exprcpfma.so
sw() { FUSED=0 timeout 200 python scripts/scan_sweep.py 2>&1 | tail -1; }
echo "== timings"
( sw
  ZIGMA_B200_LIB=$EXP sw
  ZG_SCAN_WP=0 sw
  ZG_SCAN_WP=0 ZIGMA_B200_LIB=$EXP sw
  ZG_SCAN_WP=3 ZIGMA_B200_LIB=$EXP sw
  for k in 1 2 4 8 16; do ZG_SCAN_WP_SYNC=$k sw; done          # staggered fairness barrier every k stages (mixed CTAs)
  ZG_SCAN_WP=1 ZG_SCAN_WP_WARPS=18 ZG_SCAN_WP_SYNC=4 sw
  sw
  ZIGMA_B200_LIB=$EXP sw
  for cfg in "16 1024 1280" "32 1024 1280" "32 4096 1536" "256 256 1536" "4096 16 1536"; do set -- $cfg
    BS=$1 SEQ=$2 EDIM=$3 ZG_SCAN_WP=0 sw; BS=$1 SEQ=$2 EDIM=$3 sw; BS=$1 SEQ=$2 EDIM=$3 ZIGMA_B200_LIB=$EXP sw
  done ) | tee gpurun_out/scan_auto_sweep.log
echo "== conv row-prefetch experiment (libzigma_expconvpf*.so)"
( timeout 120 python scripts/conv_sweep.py 2>&1 | tail -1
  for v in convpf6 convpf5 convpf4r; do echo -n "$v: "; ZIGMA_B200_LIB=$PWD/zigma_b200/lib/libzigma_exp$v.so timeout 120 python scripts/conv_sweep.py 2>&1 | tail -1; done
  timeout 120 python scripts/conv_sweep.py 2>&1 | tail -1 ) | tee gpurun_out/conv_prefetch_sweep.log
for v in convpf6 convpf5; do ZIGMA_B200_LIB=$PWD/zigma_b200/lib/libzigma_exp$v.so timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv" -p no:cacheprovider > gpurun_out/pytest_$v.log 2>&1; echo "pytest conv ($v) rc=$?"; tail -2 gpurun_out/pytest_$v.log; done
echo "== ncu full (default library, kernel chosen per call)"
FUSED=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_ -s 3 -c 1 -f -o gpurun_out/r02f_scan_auto python scripts/scan_sweep.py > gpurun_out/ncu_auto.log 2>&1; echo "ncu rc=$?"
echo "== pytest -m gpu (all), default library"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 > gpurun_out/pytest_gpu_auto.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_auto.log
echo "== pytest -m gpu (all), FMA-pipe reciprocal library"
ZIGMA_B200_LIB=$EXP timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 > gpurun_out/pytest_gpu_rcpfma.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_rcpfma.log
echo "== quick bench lines (no side measurements), alternating"
for v in def exp def exp; do
  if [ $v = exp ]; then export ZIGMA_B200_LIB=$EXP; else unset ZIGMA_B200_LIB; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-train --no-configs > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/bench_$v.json").read().strip().splitlines()[-1])
    print("$v", d["ms_per_step"], "ms/step", d["value"], d["unit"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["roofline"]["kernel"][:40], d["clocks"])
except Exception as ex:
    print("bench parse failed", ex)
P
done
echo done1
# ---- the build of record (default library): smoke, the full default bench line, launch list
unset ZIGMA_B200_LIB
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default flags)"; timeout 1200 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02b_bench_n1.json; tail -3 gpurun_out/r02b_bench_n1.err
echo "== launch list (ncu, eager launches)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-ref-cuda --no-train --no-configs > gpurun_out/launchlist_bench.log 2>&1; echo "ncu list rc=$?"
echo done2
