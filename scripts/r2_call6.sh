#!/bin/bash
# round 2, call 6 (1 GPU): narrow stream tiles, one-launch sweep of tiny levels (A/B), remaining ncu captures of the setup kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_amg.py tests/test_gpu_block.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_pytest6.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest6.log | cut -c1-250
python scripts/spmv_bench.py > gpurun_out/r2_spmv_schedules.txt 2>&1; echo "spmv_bench rc=$?"; cut -c1-220 gpurun_out/r2_spmv_schedules.txt
python scripts/amg_profile.py timers > gpurun_out/r2_amg_timers_sweep.log 2>&1; grep solve gpurun_out/r2_amg_timers_sweep.log | cut -c1-160
PSB200_SMALL_SWEEP=off python scripts/amg_profile.py timers > gpurun_out/r2_amg_timers_nosweep.log 2>&1; grep solve gpurun_out/r2_amg_timers_nosweep.log | cut -c1-160
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k regex:'agg_assign2_kernel|prolong_rows_kernel|spgemm_symbolic_kernel|spgemm_numeric_kernel|gather_transpose_kernel|cheb_sweep_small' -c 16 \
   -f -o gpurun_out/r2_prof_amg_setup2 python scripts/amg_profile.py setup > gpurun_out/r2_ncu_setup2.log 2>&1
echo "ncu setup2 rc=$?"; tail -2 gpurun_out/r2_ncu_setup2.log
