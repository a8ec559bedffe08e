#!/bin/bash
# round 2, call 15 (8 GPUs): final bench line at 8 ranks, fused-push A/B of the partitioned AMG leg, config 4 with the final BSR shape
mkdir -p gpurun_out
run() { N=$1; P=$2; shift 2; timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@"; }
run 8 29571 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n8_final.json') if l.startswith("{")][-1])
print("value", round(d["value"]), d["config"]["krylov"][:5], "other", d.get("other_krylov"), "e2e", round(d["e2e"]["value"]))
a=d["amg_pcg_dist"]; print({k:a.get(k) for k in ("gpu_setup_s","gpu_solve_s","gpu_iters","levels","error")}, "parity", d["parity"].get("ok"))
PY
run 8 29572 --amg-dist-only --fused-push > gpurun_out/r2_amg_n8_fused.json 2> gpurun_out/r2_amg_n8_fused.err; echo "fused rc=$?"; tail -c 900 gpurun_out/r2_amg_n8_fused.json
run 8 29573 --config c4 --steps 3 > gpurun_out/r2_c4_n8_final.json 2> gpurun_out/r2_c4_n8_final.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_c4_n8_final.json') if l.startswith("{")][-1])
print({k:d.get(k) for k in ("setup_s","solve_s","iters","levels","spmv_kernel","device_bytes_per_rank_max")})
PY
