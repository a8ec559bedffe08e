#!/bin/bash
# final check of HEAD on one GPU (budget: < 2 GPU-minutes): the new cg1r-restatement test, smoke(), a short headline bench
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_parity.py -q -x -k "restatement or single_reduction_cg_max_iter" > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_final_pytest.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/r2_final_smoke.log
timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu --no-amg --no-cusparse > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_final_bench.json
