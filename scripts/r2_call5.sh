#!/bin/bash
# round 2, call 5 (2 GPUs): Neo-Hookean problem + Newton with device Hessian, dist tests after the all-gather rework,
# hash-vs-sort SpGEMM hierarchy equality, config 5 at 1 and 2 GPUs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_neohookean.py tests/test_newton.py tests/test_gpu_dist.py tests/test_gpu_amg.py tests/test_gpu_block.py -m gpu -q > gpurun_out/r2_pytest5.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest5.log | cut -c1-300
python scripts/amg_profile.py timers 96 > gpurun_out/r2_t96_hash.log 2>&1; PSB200_SPGEMM=sort python scripts/amg_profile.py timers 96 > gpurun_out/r2_t96_sort.log 2>&1
python - <<'PY'
import json
def lv(p):
    rows=[json.loads(l) for l in open(p) if l.startswith("{")]
    return [r for r in rows if "solve" in r][-1]["amg"]
print("hash", lv("gpurun_out/r2_t96_hash.log")); print("sort", lv("gpurun_out/r2_t96_sort.log"))
PY
timeout 900 python bench.py --config c5 --steps 2 > gpurun_out/r2_c5_n1.json 2> gpurun_out/r2_c5_n1.err
echo "c5 n1 rc=$?"; tail -c 2500 gpurun_out/r2_c5_n1.json; tail -3 gpurun_out/r2_c5_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --config c5 --gpus 2 --steps 2 > gpurun_out/r2_c5_n2.json 2> gpurun_out/r2_c5_n2.err
echo "c5 n2 rc=$?"; tail -c 2500 gpurun_out/r2_c5_n2.json; tail -3 gpurun_out/r2_c5_n2.err
