#!/bin/bash
# Round 2, third session: GPU suite + smoke + default bench on the build with the positional embedding folded into the first tail,
# a launch list of the TIMED region only (ncu --profile-from-start off: bench.py brackets it with cudaProfilerStart/Stop),
# and three side measurements: conv with the SiLU reciprocal on the FMA pipe, GEMM epilogue with two TMEM loads per wait, memory diagnostics.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu_r02c.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r02c.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default flags)"; timeout 900 python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"; head -c 700 gpurun_out/r02c_bench_n1.json; echo; tail -3 gpurun_out/r02c_bench_n1.err
echo "== launch list (ncu, eager launches of the timed region)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-ref-cuda --no-train --no-configs > gpurun_out/launchlist_bench.log 2>&1; echo "ncu list rc=$?"
python scripts/ncu_list_summary.py gpurun_out/r02c_launches.csv 2>&1 | head -14
echo "== conv / tail"
timeout 100 python scripts/conv_sweep.py | tail -1
ZIGMA_B200_LIB=$PWD/zigma_b200/lib/libzigma_expconvrcp.so timeout 100 python scripts/conv_sweep.py | tail -1 | sed 's/^/[conv rcp on the FMA pipe] /'
echo "== gemm"
timeout 200 python scripts/gemm_bench.py 2>&1 | tail -4
echo "-- epilogue: two TMEM loads per wait"
ZIGMA_B200_LIB=$PWD/zigma_b200/lib/libzigma_expepild2.so timeout 200 python scripts/gemm_bench.py 2>&1 | tail -4
echo "== memory diagnostics"
timeout 100 python scripts/mem_diag.py 2>&1 | tail -3
echo done
