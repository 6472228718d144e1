#!/bin/bash
# Round 2, third session, second visit: the cp.async-staged conv forward (ZG_CONV_SMEM=1) -- bit identity against the default kernel,
# timing at the config-2 layer shape for 16 / 32 / 64 tokens per warp; then the full validation (GPU suite, smoke, default bench)
# with the faster conv kernel selected through the environment (the compile-time default follows after the visit).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
echo "== conv tests"; timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "conv" --timeout=200 > gpurun_out/pytest_conv_r02d.log 2>&1; crc=$?; echo "pytest conv rc=$crc"; tail -3 gpurun_out/pytest_conv_r02d.log
echo "== conv sweep"
best_us=$(ZG_CONV_SMEM=0 CONV_SWEEP_JSON=1 timeout 100 python scripts/conv_sweep.py | tee -a gpurun_out/conv_smem_sweep.log | grep '^{' | python -c "import sys,json; print(json.loads(sys.stdin.readline())['conv_us'])")
echo "default kernel: $best_us us"; best=0; best_lch=32
for lch in 16 32 64; do
  us=$(ZG_CONV_SMEM=1 ZG_CONV_SMEM_LCH=$lch CONV_SWEEP_JSON=1 timeout 100 python scripts/conv_sweep.py | tee -a gpurun_out/conv_smem_sweep.log | grep '^{' | python -c "import sys,json; print(json.loads(sys.stdin.readline())['conv_us'])")
  echo "smem-staged kernel, $lch tokens per warp: $us us"
  if [ -n "$us" ] && [ "$crc" = "0" ] && python -c "import sys; sys.exit(0 if float('$us') < 0.97 * float('$best_us') else 1)"; then best=1; best_lch=$lch; best_us=$us; fi
done
echo "selected: ZG_CONV_SMEM=$best ZG_CONV_SMEM_LCH=$best_lch ($best_us us)" | tee gpurun_out/conv_choice.txt
export ZG_CONV_SMEM=$best ZG_CONV_SMEM_LCH=$best_lch
echo "== pytest -m gpu (ZG_CONV_SMEM=$best)"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu_r02d.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r02d.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default flags)"; timeout 900 python bench.py > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; echo "bench rc=$?"; head -c 420 gpurun_out/r02d_bench_n1.json; echo; tail -3 gpurun_out/r02d_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02d_bench_n1.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["clocks"])
for k in d["roofline"].get("other_kernels", []):
    print(k if isinstance(k, str) else {x: (round(v, 4) if isinstance(v, float) else v) for x, v in k.items() if x in ("kernel", "ms_per_launch", "library_ms_per_launch", "frac", "bound")})
PY
echo done
