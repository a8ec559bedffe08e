#!/bin/bash
# round 2, call 11 (2 GPUs): fused push + boundary-first order (tests, A/B on C3 and C4 at 2 ranks), tile-shape sweeps
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dist.py tests/test_gpu_amg.py tests/test_gpu_block.py tests/test_neohookean.py -m gpu -q -x > gpurun_out/r2_pytest11.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest11.log | cut -c1-300
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
run 2 29541 --steps 3 --warmup 3 > gpurun_out/r2_b11_fused.json 2> gpurun_out/r2_b11_fused.err; echo "rc=$?"; summ gpurun_out/r2_b11_fused.json; tail -2 gpurun_out/r2_b11_fused.err
PSB200_FUSED_PUSH=off run 2 29542 --steps 3 --warmup 3 > gpurun_out/r2_b11_unfused.json 2> gpurun_out/r2_b11_unfused.err; echo "rc=$?"; summ gpurun_out/r2_b11_unfused.json
run 2 29543 --config c4 --steps 3 > gpurun_out/r2_c4_11_fused.json 2> gpurun_out/r2_c4_11_fused.err; echo "rc=$?"; summ gpurun_out/r2_c4_11_fused.json; tail -2 gpurun_out/r2_c4_11_fused.err
PSB200_FUSED_PUSH=off run 2 29544 --config c4 --steps 3 > gpurun_out/r2_c4_11_unfused.json 2> gpurun_out/r2_c4_11_unfused.err; echo "rc=$?"; summ gpurun_out/r2_c4_11_unfused.json
python scripts/spmv_bench.py 128 72 > gpurun_out/r2_spmv_schedules_c.txt 2>&1; echo "spmv rc=$?"; grep -E "stencil33|squared|block_size" gpurun_out/r2_spmv_schedules_c.txt | cut -c1-260
python scripts/bsr_target.py 72 --sweep > gpurun_out/r2_bsr_sweep.txt 2>&1; cat gpurun_out/r2_bsr_sweep.txt | tail -8
