mkdir -p gpurun_out
bash scripts/gpu_round.sh
for bs in 16 64; do BS=$bs timeout 300 python scripts/bwd_bench.py 2>&1 | grep -v "ours vs reference"; done > gpurun_out/bwd_bench.log 2>&1
for bs in 16 64; do BS=$bs DTYPE=bf16 timeout 600 python scripts/train_bench.py 2>&1 | grep "^{"; done > gpurun_out/train_bench_bs.log 2>&1
BS=16 DTYPE=amp timeout 600 python scripts/train_bench.py 2>&1 | grep "^{" >> gpurun_out/train_bench_bs.log
BS=16 DTYPE=fp32 timeout 600 python scripts/train_bench.py 2>&1 | grep "^{" >> gpurun_out/train_bench_bs.log
BS=16 DTYPE=bf16 PROFILE=1 timeout 600 python scripts/train_bench.py > gpurun_out/train_bench.log 2>&1
