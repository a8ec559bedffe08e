#!/bin/bash
# round 2, call 7 (2 GPUs): BSR-3 schedule (parity, bench), new auto plan, C4 at size on 1 and 2 GPUs, full suite
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest7.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest7.log | cut -c1-300
python scripts/spmv_bench.py > gpurun_out/r2_spmv_schedules_b.txt 2>&1; echo "spmv_bench rc=$?"; tail -4 gpurun_out/r2_spmv_schedules_b.txt | cut -c1-330
timeout 900 python bench.py --config c4 --steps 3 > gpurun_out/r2_c4b_119_n1.json 2> gpurun_out/r2_c4b_119_n1.err
echo "c4 119 n1 rc=$?"; python - <<'PY'
import json
for f in ("gpurun_out/r2_c4b_119_n1.json",):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print({k:d[k] for k in ("setup_s","solve_s","iters","levels","spmv_kernel","device_bytes_per_rank_max")}, d["roofline"]["frac"])
PY
tail -3 gpurun_out/r2_c4b_119_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --config c4 --gpus 2 --steps 3 > gpurun_out/r2_c4b_119_n2.json 2> gpurun_out/r2_c4b_119_n2.err
echo "c4 119 n2 rc=$?"; python - <<'PY'
import json
for f in ("gpurun_out/r2_c4b_119_n2.json",):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print({k:d[k] for k in ("setup_s","solve_s","iters","levels","spmv_kernel","partitioned_levels","device_bytes_per_rank_max")}, d["roofline"]["frac"])
PY
tail -3 gpurun_out/r2_c4b_119_n2.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench7.json').read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"])
a=d["amg_pcg"]["full"]; print({k:a[k] for k in ("gpu_setup_s","gpu_solve_s","gpu_iters","levels")}, a["roofline"]["frac"])
PY
