#!/usr/bin/env bash
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tests/multi_gpu_check.py --bodies 600000 --steps 70 --repartition --fp64 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version" > gpurun_out/r2_repart.log 2>&1; grep -v "^\s*$" gpurun_out/r2_repart.log | head -40
