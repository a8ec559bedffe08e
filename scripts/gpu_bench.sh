#!/bin/bash
# One gpurun call: bench (ours + reference arm), launch list, full ncu capture of the hot kernels.
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python scripts/profile_target.py > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'spmv_stream' -s 2 -c 2 -f -o gpurun_out/prof_spmv python scripts/profile_target.py > gpurun_out/ncu_spmv.log 2>&1
echo "ncu spmv rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'vec_kernel' -s 3 -c 2 -f -o gpurun_out/prof_vec python scripts/profile_target.py > gpurun_out/ncu_vec.log 2>&1
echo "ncu vec rc=$?"
ls -la gpurun_out
