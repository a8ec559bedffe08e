#!/bin/bash
# round 2, call 1 (1 GPU): full GPU test suite (world-1 row-partition cases included), a short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest1.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
