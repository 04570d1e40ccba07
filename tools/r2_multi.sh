#!/usr/bin/env bash
# multi-GPU validation: tools/r2_multi.sh <ngpus>   (gpurun --gpus N)
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
out=gpurun_out/r2_multi_gpu_check_g$N.log
: > $out
( timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -rfs 2>&1 | tail -5 ) >> $out
for args in "--bodies 200000 --steps 3" "--bodies 200000 --steps 3 --fp64" "--bodies 200000 --steps 2 --no-p2p" "--bodies 1500000 --steps 2" "--bodies 100001 --steps 2 --host-step"; do
  echo "== multi_gpu_check $args" >> $out
  timeout 300 $TR tests/multi_gpu_check.py $args 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -6 >> $out
done
echo "== BH_HOST_PIPELINE_MULTI=1 --host-step" >> $out
BH_HOST_PIPELINE_MULTI=1 timeout 300 $TR tests/multi_gpu_check.py --bodies 100001 --steps 2 --host-step 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -6 >> $out
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_g$N.json 2> gpurun_out/r2_bench_g$N.err
BH_HOST_PIPELINE_MULTI=1 timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-strong --no-cpu-baseline > gpurun_out/r2_bench_g${N}_hostpipe.json 2> gpurun_out/r2_bench_g${N}_hostpipe.err
cat $out; tail -3 gpurun_out/r2_bench_g$N.err; cut -c1-300 gpurun_out/r2_bench_g$N.json
