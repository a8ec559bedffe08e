#!/bin/bash
# last GPU call of the round (< 2 GPU-minutes left): the Newton driver on the GPU after the host-side changes
# (new Problem virtuals / struct layout, line searches), config-5 problem tests, then smoke()
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_newton.py tests/test_neohookean.py -q -x -m gpu > gpurun_out/r2_final2_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_final2_pytest.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final2_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/r2_final2_smoke.log
