#!/bin/bash
# round 2, call 2 (2 GPUs): full GPU test suite incl. the world-2 row-partition cases, bench at N=2
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/r2_pytest2.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench rc=$?"; tail -c 5000 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
