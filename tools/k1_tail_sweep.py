"""K1 time per interval against the number of intervals: how much the last partial round of work items costs.
c2 shape (persistent kernel, 148 x 12 warps) and c4-shard shapes (octet kernel, 296 CTAs x 4 warps x 8 intervals)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dto_b200 as dto
from dto_b200 import problem_templates as pt

dev = torch.device("cuda")


def run(name, prob, reps=20):
    ev = dto.Evaluator(prob)
    Z = prob.trajectory.datavec
    dZ = torch.from_numpy(Z).to(dev)
    dmu = torch.rand(ev.n_constraints, dtype=torch.float64, device=dev)
    outs = [torch.empty(k, dtype=torch.float64, device=dev) for k in (1, ev.n_vars, ev.n_constraints, ev.nnz_jacobian, ev.nnz_hessian)]
    stream = torch.cuda.ExternalStream(ev.stream)
    step = lambda: ev.eval_all_dev(dZ.data_ptr(), 1.0, dmu.data_ptr(), *[o.data_ptr() for o in outs])
    for _ in range(3):
        step()
    ev.synchronize()
    ev.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps):
            step()
        e1.record(stream)
    ev.synchronize()
    ms = e0.elapsed_time(e1) / reps
    k1, nk = ev.kernel_time_ms()
    nI = prob.trajectory.N - 1
    print(f"{name}: intervals={nI} step={ms*1e3:.1f} us K1={k1/nk*1e3:.1f} us  K1/interval={k1/nk/nI*1e6:.1f} ns  variant={ev.kernel_variant(0)}", flush=True)
    ev.close()


for N in (445, 889, 1777, 1999, 2000, 2369, 3553, 4000, 5329):
    run(f"c2-shape N={N}", pt.quantum_gate_problem(N=N, levels=16, n_drives=4))
for N in (9473, 12501, 18945, 25001, 50001, 100000):
    run(f"c4-shape N={N}", pt.scaled_problem(N=N, state_dim=16, n_controls=2, generator_scale=0.25), reps=10)
for N in (201,):
    pass
