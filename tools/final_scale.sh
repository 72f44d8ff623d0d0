#!/bin/bash
# Final multi-GPU evidence on an N-GPU box: the default bench.py line (c2 headline + c3/c4/c5 under `extra`) exactly as the
# driver launches it, the 2-rank linked-shard tests, and the box's aggregate D2H rate.  Usage: tools/final_scale.sh N
set -u
N=${1:-8}
mkdir -p gpurun_out
if [ "$N" -gt 1 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
  echo "bench N=$N rc=$?"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) \
      tools/d2h_aggregate.py > gpurun_out/r02_d2h_aggregate_${N}gpu.log 2>&1
  echo "d2h N=$N rc=$?"; tail -1 gpurun_out/r02_d2h_aggregate_${N}gpu.log
else
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
  echo "bench N=1 rc=$?"
  python tools/d2h_aggregate.py > gpurun_out/r02_d2h_aggregate_1gpu.log 2>&1; tail -1 gpurun_out/r02_d2h_aggregate_1gpu.log
fi
tail -c 600 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("c2", round(d["value"], 1), d["unit"], "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3))
for k, v in d.get("extra", {}).items():
    print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "shard_matches_single_gpu", "frac", "scaling")})
PY
