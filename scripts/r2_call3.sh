#!/bin/bash
# round 2, call 3 (2 GPUs): config 4 shape (block-3 elasticity AMG-PCG) at 72^3 nodes on 1 and 2 GPUs
mkdir -p gpurun_out
timeout 600 python bench.py --config c4 --c4-nodes 72 --steps 3 > gpurun_out/r2_c4_72_n1.json 2> gpurun_out/r2_c4_72_n1.err
echo "c4 n1 rc=$?"; tail -c 2500 gpurun_out/r2_c4_72_n1.json; tail -3 gpurun_out/r2_c4_72_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --config c4 --c4-nodes 72 --gpus 2 --steps 3 > gpurun_out/r2_c4_72_n2.json 2> gpurun_out/r2_c4_72_n2.err
echo "c4 n2 rc=$?"; tail -c 2500 gpurun_out/r2_c4_72_n2.json; tail -3 gpurun_out/r2_c4_72_n2.err
timeout 900 python bench.py --config c4 --steps 3 > gpurun_out/r2_c4_119_n1.json 2> gpurun_out/r2_c4_119_n1.err
echo "c4 119 n1 rc=$?"; tail -c 2500 gpurun_out/r2_c4_119_n1.json; tail -3 gpurun_out/r2_c4_119_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --config c4 --gpus 2 --steps 3 > gpurun_out/r2_c4_119_n2.json 2> gpurun_out/r2_c4_119_n2.err
echo "c4 119 n2 rc=$?"; tail -c 2500 gpurun_out/r2_c4_119_n2.json; tail -3 gpurun_out/r2_c4_119_n2.err
