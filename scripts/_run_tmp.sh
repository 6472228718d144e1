mkdir -p gpurun_out
DO_NCU_LIST=1 bash scripts/gpu_round.sh
python scripts/ncu_list_summary.py gpurun_out/launches.csv > gpurun_out/launch_list.txt 2>&1; tail -15 gpurun_out/launch_list.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/bench_ref.err | tail -1 | cut -c1-400 | tee gpurun_out/bench_ref.log
