#!/bin/bash
# round 2, call 9 (2 GPUs): full suite on the final code, bench at N=2 (cg vs cg1r after the push-side wait removal)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest9.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest9.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 --no-amg > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n2b.json') if l.startswith("{")][-1])
print("value", d["value"], d["config"]["krylov"][:5], "other", d.get("other_krylov"), "e2e", d["e2e"]["value"])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 --no-amg --krylov cg > gpurun_out/r2_bench_n2c.json 2> gpurun_out/r2_bench_n2c.err
echo "bench cg rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n2c.json') if l.startswith("{")][-1])
print("value", d["value"], d["config"]["krylov"][:5], "other", d.get("other_krylov"), "e2e", d["e2e"]["value"], d["roofline"]["profile_ms"])
PY
