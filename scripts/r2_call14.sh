#!/bin/bash
# round 2, call 14 (1 GPU): final validation of the committed state: full GPU suite, smoke, bench (ours + reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest14.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest14.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke14.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke14.log
timeout 900 python bench.py > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench14.json') if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
a=d["amg_pcg"]["full"]; print({k:a[k] for k in ("gpu_setup_s","gpu_solve_s","gpu_iters","levels","level_spmv_kernels","speedup_setup_plus_solve")}, a["roofline"]["frac"])
print(d["cusparse"])
PY
tail -3 gpurun_out/r2_bench14.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench14_ref.json 2> gpurun_out/r2_bench14_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r2_bench14_ref.json
