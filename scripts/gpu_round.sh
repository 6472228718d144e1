#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench, launch list.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== debug"; [ -n "$DEBUG_SCRIPT" ] && timeout 300 python $DEBUG_SCRIPT > gpurun_out/debug.log 2>&1; [ -n "$DEBUG_SCRIPT" ] && tail -20 gpurun_out/debug.log
if [ -n "$DO_SWEEP" ]; then echo "== scan sweep"; for n in ${SWEEP_SET:-0 2 3 4}; do ZG_SCAN_NPOLY=$n timeout 300 python scripts/scan_sweep.py 2>&1 | tail -1; done | tee gpurun_out/scan_sweep.log; fi
if [ -n "$DO_MICRO" ]; then echo "== micro"; (if [ -z "$SKIP_GEMM_BENCH" ]; then timeout 300 python scripts/gemm_bench.py; fi; for v in ${CONV_SET:-0 4}; do ZG_CONV_VEC=$v timeout 200 python scripts/conv_sweep.py | tail -1; done; for n in ${TPC2_NPOLY_SET:-0 1 2}; do ZG_SCAN_TPC2_NPOLY=$n timeout 200 python scripts/scan_sweep.py | tail -1; done) 2>&1 | tee gpurun_out/micro.log; fi
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q -rA -p no:cacheprovider --timeout=900 ${PYTEST_EXTRA} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(PASSED|FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -100
if [ -z "$SKIP_BENCH" ]; then
echo "== bench"; timeout 900 python bench.py --steps ${BENCH_STEPS:-20} --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.log | cut -c1-400; tail -5 gpurun_out/bench.err
if [ -n "$BENCH_TCGEN05" ]; then ZIGMA_TCGEN05=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-train > gpurun_out/bench_tcgen05.log 2> gpurun_out/bench_tcgen05.err; tail -1 gpurun_out/bench_tcgen05.log | cut -c1-300; fi
fi
if [ -n "$DO_NCU_LIST" ]; then
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-train --no-graph > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"
fi
if [ -n "$DO_NCU_FULL" ]; then
echo "== ncu full (scan kernel)"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${NCU_KERNEL:-scan_fwd_kernel} -c ${NCU_COUNT:-2} -o gpurun_out/${NCU_OUT:-scan_fwd} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-train --no-graph > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
echo done
