mkdir -p gpurun_out
BS=16 timeout 300 python scripts/bwd_bench.py 2>&1 | grep -E "conv bwd|scan" | tee gpurun_out/bwd_bench_conv.log
