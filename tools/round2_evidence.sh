#!/bin/bash
# Round-2 evidence on ONE GPU: bench lines per workload, ncu launch lists (cold-cache, serialised: shares only) and
# `--set full` captures of the dominant kernels, summarised on the box (the .ncu-rep files are too large to bring back:
# gpurun merges at most 64 MiB); outputs under gpurun_out/r02_*, copied to profiles/ on the build host.
set -u
mkdir -p gpurun_out
for w in c2 c3 c4 c5; do
  python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/r02_bench_$w.err | tail -1 > gpurun_out/r02_bench_$w.json; echo "bench $w rc=$?"
done
declare -A KERN=( [c2]=bilinear_persistent [c3]=tdb_dmma_kernel [c4]=bilinear_octet [c5]=bilinear_octet )
for w in c2 c3 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02_ncu_launches_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_l_$w.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:${KERN[$w]} -s 2 -c 1 -f -o /tmp/r02_prof_$w \
      python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_f_$w.log 2>&1
  python tools/profile_summary.py gpurun_out/r02_ncu_launches_$w.csv /tmp/r02_prof_$w.ncu-rep gpurun_out/r02_ncu_summary_$w.json \
      "python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline" > /dev/null
  python tools/ncu_summary.py /tmp/r02_prof_$w.ncu-rep 12 > gpurun_out/r02_ncu_hot_instructions_$w.txt 2>&1
  echo "ncu $w done"
done
# c3's second kernel, and the assembler at c5
ncu --set full --clock-control none --import-source on -k regex:tdb_exp_kernel -s 2 -c 1 -f -o /tmp/r02_prof_c3_exp \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_f_c3_exp.log 2>&1
python tools/ncu_summary.py /tmp/r02_prof_c3_exp.ncu-rep 12 > gpurun_out/r02_ncu_summary_c3_exp_kernel.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:hessian_assemble -s 2 -c 1 -f -o /tmp/r02_prof_c5_k3 \
    python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_f_c5_k3.log 2>&1
python tools/ncu_summary.py /tmp/r02_prof_c5_k3.ncu-rep 12 > gpurun_out/r02_ncu_summary_c5_hessian_assemble.txt 2>&1
rm -f gpurun_out/r02_ncu_l_*.log gpurun_out/r02_ncu_f_*.log
du -sh gpurun_out; ls gpurun_out | head -40
