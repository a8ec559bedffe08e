#!/bin/bash
# round 2, call 13 (2 GPUs): fused push after the finalizer / lookup fixes: dist tests (partitioned), C3 and C4 at 2 ranks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -k "partitioned or elasticity or reanalyse" > gpurun_out/r2_pytest13.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest13.log | cut -c1-300
run() { N=$1; P=$2; shift 2; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@"; }
summ() { python - "$1" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
if "amg_pcg_dist" in d:
    a=d["amg_pcg_dist"]; print(sys.argv[1], "value", round(d["value"]), {k:a.get(k) for k in ("gpu_setup_s","gpu_solve_s","gpu_iters","levels","error")}, "parity", d.get("parity",{}).get("ok"))
else:
    print(sys.argv[1], {k:d.get(k) for k in ("setup_s","solve_s","iters","levels","spmv_kernel")})
PY
}
run 2 29561 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_b13_fused.json 2> gpurun_out/r2_b13_fused.err; echo "rc=$?"; summ gpurun_out/r2_b13_fused.json; tail -2 gpurun_out/r2_b13_fused.err
run 2 29563 --config c4 --steps 3 > gpurun_out/r2_c4_13_fused.json 2> gpurun_out/r2_c4_13_fused.err; echo "rc=$?"; summ gpurun_out/r2_c4_13_fused.json
PSB200_FUSED_PUSH=off run 2 29564 --config c4 --steps 3 > gpurun_out/r2_c4_13_unfused.json 2> gpurun_out/r2_c4_13_unfused.err; echo "rc=$?"; summ gpurun_out/r2_c4_13_unfused.json
