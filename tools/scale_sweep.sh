#!/bin/bash
# Multi-GPU sweep of the BASELINE.json configurations that shard: c2 (replicas, weak), c4 (knot shards, strong),
# c5 (batch split, strong).  Usage: tools/scale_sweep.sh "8 4" "c2 c4 c5"
set -u
mkdir -p gpurun_out
for n in ${1:-8}; do
  for w in ${2:-c2 c4 c5}; do
    port=$((29500 + n * 10 + ${#w} + RANDOM % 50))
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${w}_n${n}.log 2> gpurun_out/scale_${w}_n${n}.err
    echo "== $w n=$n rc=$?"; tail -1 gpurun_out/scale_${w}_n${n}.log | cut -c1-400
  done
done
