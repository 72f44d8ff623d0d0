#!/bin/bash
# The c3 part of tools/round2_evidence.sh (after a change to K7 only).
set -u
mkdir -p gpurun_out
w=c3
python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/r02_bench_$w.err | tail -1 > gpurun_out/r02_bench_$w.json; echo "bench $w rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02_ncu_launches_$w.csv \
    python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/l_$w.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tdb_dmma_kernel -s 2 -c 1 -f -o /tmp/r02_prof_$w \
    python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > /tmp/f_$w.log 2>&1
python tools/profile_summary.py gpurun_out/r02_ncu_launches_$w.csv /tmp/r02_prof_$w.ncu-rep gpurun_out/r02_ncu_summary_$w.json \
    "python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline" > /dev/null
python tools/ncu_summary.py /tmp/r02_prof_$w.ncu-rep 12 > gpurun_out/r02_ncu_hot_instructions_$w.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:tdb_exp_kernel -s 2 -c 1 -f -o /tmp/r02_prof_c3_exp \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > /tmp/f_c3_exp.log 2>&1
python tools/ncu_summary.py /tmp/r02_prof_c3_exp.ncu-rep 12 > gpurun_out/r02_ncu_summary_c3_exp_kernel.txt 2>&1
echo done
