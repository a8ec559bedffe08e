#!/bin/bash
# round 2, call 10 (1 GPU): direct coarse solve (DMMA), config 5 on one GPU with a warm-up run, launch list + BSR ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_amg.py tests/test_gpu_block.py -m gpu -q > gpurun_out/r2_pytest10.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest10.log | cut -c1-300
timeout 600 python bench.py --config c5 --steps 4 > gpurun_out/r2_c5_n1b.json 2> gpurun_out/r2_c5_n1b.err
echo "c5 n1 rc=$?"; tail -c 1200 gpurun_out/r2_c5_n1b.json; tail -2 gpurun_out/r2_c5_n1b.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'spmv_bsr3_kernel' -s 3 -c 2 -f -o gpurun_out/r2_prof_bsr3 python scripts/bsr_target.py > gpurun_out/r2_ncu_bsr.log 2>&1
echo "ncu bsr rc=$?"; tail -2 gpurun_out/r2_ncu_bsr.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-amg --no-cusparse > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"; tail -2 gpurun_out/r2_ncu_launches.log | cut -c1-200
