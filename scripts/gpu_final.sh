#!/bin/bash
# Build of record: GPU suite, smoke, the default bench line, launch list, one ncu --set full capture of the scan at config 2.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout=900 > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default flags)"; timeout 1200 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/r02b_bench_n1.json; tail -3 gpurun_out/r02b_bench_n1.err
echo "== launch list (ncu, eager launches)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-ref-cuda --no-train --no-configs > gpurun_out/launchlist_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (scan, config 2 layer shape)"
FUSED=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_fwd_ -s 3 -c 1 -f -o gpurun_out/r02g_scan_final python scripts/scan_sweep.py > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"
FUSED=0 timeout 100 python scripts/scan_sweep.py | tail -1
echo "== staggered fairness barrier every k stages (mixed CTAs)"
( for k in 1 2 4 8 16; do ZG_SCAN_WP_SYNC=$k FUSED=0 timeout 100 python scripts/scan_sweep.py | tail -1; done; ZG_SCAN_WP=3 ZG_SCAN_WP_SYNC=4 BS=32 SEQ=4096 EDIM=1536 FUSED=0 timeout 100 python scripts/scan_sweep.py | tail -1; BS=32 SEQ=4096 EDIM=1536 FUSED=0 timeout 100 python scripts/scan_sweep.py | tail -1 ) | tee gpurun_out/scan_sync_sweep.log
timeout 100 python scripts/conv_sweep.py | tail -1
echo done
