#!/bin/bash
# Round-2 GPU visit: smoke, GPU parity tests, bench (+ other configs), optional ncu launch list / full capture.
mkdir -p gpurun_out; rm -f gpurun_out/parity_log.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
if [ -z "$SKIP_SMOKE" ]; then echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log; fi
if [ -z "$SKIP_TESTS" ]; then echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q -rA -p no:cacheprovider --timeout=900 ${PYTEST_EXTRA} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -30; python scripts/parity_summary.py > gpurun_out/parity_summary.txt 2>&1; fi
if [ -z "$SKIP_BENCH" ]; then
echo "== bench"; timeout 1500 python bench.py --steps ${BENCH_STEPS:-20} --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-6000; tail -5 gpurun_out/bench.err
fi
if [ -n "$DO_REF" ]; then echo "== reference arm"; timeout 1200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-1500; tail -3 gpurun_out/bench_ref.err; fi
if [ -n "$DO_NCU_LIST" ]; then
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-train --no-graph --no-configs > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"
fi
if [ -n "$DO_NCU_FULL" ]; then
echo "== ncu full"; FUSED=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-scan_fwd_tma} -s 3 -c 1 -o gpurun_out/${NCU_OUT:-r02_scan_tma} python scripts/scan_sweep.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
echo done
