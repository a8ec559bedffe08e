#!/bin/bash
# round 2, call 8 (8 GPUs): row-partition tests at 1/2/4/8 ranks, bench at 8 and 4, configs 4 and 5 at 8
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1500 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/r2_pytest_dist8.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/r2_pytest_dist8.log | cut -c1-300
run() { # N port extra...
  N=$1; P=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@"
}
run 8 29521 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench8 rc=$?"; tail -c 4500 gpurun_out/r2_bench_n8.json; tail -2 gpurun_out/r2_bench_n8.err
run 8 29522 --config c4 --steps 3 > gpurun_out/r2_c4_n8.json 2> gpurun_out/r2_c4_n8.err; echo "c4 n8 rc=$?"; tail -c 2500 gpurun_out/r2_c4_n8.json; tail -2 gpurun_out/r2_c4_n8.err
run 8 29523 --config c5 --steps 4 > gpurun_out/r2_c5_n8.json 2> gpurun_out/r2_c5_n8.err; echo "c5 n8 rc=$?"; tail -c 2500 gpurun_out/r2_c5_n8.json; tail -2 gpurun_out/r2_c5_n8.err
run 4 29524 --steps 5 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench4 rc=$?"; tail -c 3000 gpurun_out/r2_bench_n4.json; tail -2 gpurun_out/r2_bench_n4.err
run 4 29525 --config c4 --steps 3 > gpurun_out/r2_c4_n4.json 2> gpurun_out/r2_c4_n4.err; echo "c4 n4 rc=$?"; tail -c 1500 gpurun_out/r2_c4_n4.json
