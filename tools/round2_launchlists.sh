set -u
for w in c2 c3 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02_ncu_launches_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/l_$w.log 2>&1
  echo "launch list $w rc=$?"
done
