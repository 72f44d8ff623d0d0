#!/bin/bash
# Probe the GPU box: toolchains, CPU, GPU; run FP64 peak microbenchmarks + cuBLAS DGEMM.
mkdir -p gpurun_out
{
echo "== toolchains"; for t in julia go javac node cargo; do printf "%s: " $t; command -v $t || echo "NOT FOUND"; done
echo "== cpu"; nproc; lscpu | grep -E "Model name|Socket|Core|Thread|MHz" ; free -g | head -2
echo "== gpu"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv
echo "== fp64_peak"; ./build/fp64_peak
echo "== torch dgemm"
python - <<'PY'
import torch, time
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"cuBLAS DGEMM {n}^3: {best:.3f} ms  {2*n**3/best*1e-9:.2f} TFLOP/s")
PY
} 2>&1 | tee gpurun_out/probe_box.log
