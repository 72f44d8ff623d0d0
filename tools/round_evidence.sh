#!/bin/bash
# One-GPU evidence run: full GPU test suite, bench lines of c2/c4/c5 (both arms for c2), ncu launch lists and
# full-set captures of the dominant kernel of c4 and c5.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
for w in c4 c5; do
  python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_$w.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:bilinear_octet -s 2 -c 1 -f -o gpurun_out/prof_k1_$w \
      python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f_$w.log 2>&1
done
for f in gpurun_out/bench_c2.log gpurun_out/bench_ref.log gpurun_out/bench_c4.log gpurun_out/bench_c5.log; do tail -1 $f | cut -c1-300; done
