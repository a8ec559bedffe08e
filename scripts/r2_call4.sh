#!/bin/bash
# round 2, call 4 (1 GPU): hash SpGEMM parity + A/B, ncu captures of the AMG setup / transfer kernels, full bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_amg.py tests/test_gpu_block.py tests/test_gpu_parity.py tests/test_adapter.py tests/test_newton.py -m gpu -q > gpurun_out/r2_pytest4.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest4.log | cut -c1-250
python scripts/amg_profile.py timers > gpurun_out/r2_amg_timers_hash.log 2>&1; echo "timers hash rc=$?"; cut -c1-700 gpurun_out/r2_amg_timers_hash.log | tail -5
PSB200_SPGEMM=sort python scripts/amg_profile.py timers > gpurun_out/r2_amg_timers_sort.log 2>&1; echo "timers sort rc=$?"; cut -c1-700 gpurun_out/r2_amg_timers_sort.log | tail -5
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k regex:'agg_max_kernel|agg_assign2_kernel|prolong_rows_kernel|spgemm_symbolic_kernel|spgemm_numeric_kernel|gather_transpose_kernel' -c 14 \
   -f -o gpurun_out/r2_prof_amg_setup python scripts/amg_profile.py setup > gpurun_out/r2_ncu_setup.log 2>&1
echo "ncu setup rc=$?"; tail -3 gpurun_out/r2_ncu_setup.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k regex:'EpiStore|EpiAddTo|EpiResidual,' -c 6 -f -o gpurun_out/r2_prof_amg_transfer python scripts/amg_profile.py launches > gpurun_out/r2_ncu_transfer.log 2>&1
echo "ncu transfer rc=$?"; tail -3 gpurun_out/r2_ncu_transfer.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench4.json').read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"])
a=d["amg_pcg"]["full"]; print({k:a[k] for k in ("gpu_setup_s","gpu_solve_s","gpu_iters","levels","speedup_setup_plus_solve")}); print(a["gpu_setup_ms_by_level"])
PY
ls -la gpurun_out/*.ncu-rep
