#!/bin/bash
# round 2, call 16 (1 GPU): config 4 on one GPU with the final BSR tile shape; SpMV schedules with the final auto plan
mkdir -p gpurun_out
timeout 300 python bench.py --config c4 --steps 3 > gpurun_out/r2_c4_n1_final.json 2> gpurun_out/r2_c4_n1_final.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_c4_n1_final.json') if l.startswith("{")][-1])
print({k:d.get(k) for k in ("setup_s","solve_s","iters","levels","spmv_kernel","device_bytes_per_rank_max")}, d["roofline"]["frac"])
PY
python scripts/spmv_bench.py 128 72 > gpurun_out/r2_spmv_schedules_final.txt 2>&1; grep -E "auto=|block_size" gpurun_out/r2_spmv_schedules_final.txt | cut -c1-250
