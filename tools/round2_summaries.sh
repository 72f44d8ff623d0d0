#!/bin/bash
# Build-host side of tools/round2_evidence.sh: copy the summaries the GPU run brought back into profiles/.
set -u
for w in c2 c3 c4 c5; do
  for f in r02_ncu_summary_$w.json r02_ncu_launches_$w.csv r02_bench_$w.json r02_ncu_hot_instructions_$w.txt; do cp gpurun_out/$f profiles/$f; done
done
cp gpurun_out/r02_ncu_summary_c3_exp_kernel.txt gpurun_out/r02_ncu_summary_c5_hessian_assemble.txt profiles/
python tools/sass_counts.py > profiles/r02_sass_counts.txt
ls profiles | grep r02
