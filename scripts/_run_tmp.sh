mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train.py -m gpu -q -x -p no:cacheprovider --timeout=600 > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_train.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_n2.log; cut -c1-300 gpurun_out/bench_n2.log
timeout 300 python scripts/train_ddp_bench.py 2>&1 | grep "^{" | tee gpurun_out/train_ddp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/train_ddp_bench.py 2>&1 | grep -E "^\{|Error|error" | tee -a gpurun_out/train_ddp.log
