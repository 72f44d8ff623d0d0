"""One rank's share of config c4 on ONE GPU (no exchange): where the strong-scaling loss at 8 ranks comes from.
usage: python tools/c4_size_sweep.py [N ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dto_b200 as dto
from dto_b200 import problem_templates as pt

sizes = [int(a) for a in sys.argv[1:]] or [100000, 50000, 25000, 12500]
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for N in sizes:
    prob = pt.scaled_problem(N=N, state_dim=16, n_controls=2, generator_scale=0.25)
    ev = dto.Evaluator(prob)
    rng = np.random.default_rng(0)
    dZ = torch.from_numpy(prob.trajectory.datavec + 0.01 * rng.standard_normal(ev.n_vars)).to(dev)
    dmu = torch.rand(ev.n_constraints, dtype=torch.float64, device=dev)
    out = [torch.empty(k, dtype=torch.float64, device=dev) for k in (1, ev.n_vars, ev.n_constraints, ev.nnz_jacobian, ev.nnz_hessian)]
    stream = torch.cuda.ExternalStream(ev.stream)
    step = lambda: ev.eval_all_dev(dZ.data_ptr(), 1.0, dmu.data_ptr(), *[o.data_ptr() for o in out])
    for _ in range(3):
        step()
    ev.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            step()
            e1.record(stream)
        ev.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"c4 share N={N}: eval {ms:.4f} ms  ({ms / (N - 1) * 1e6:.2f} ns per interval; x{ms / (N - 1) / (1.42 / 99999):.2f} of the N=100000 rate)", flush=True)
    ev.close()
