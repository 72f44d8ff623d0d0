"""c5 with PER-PROBLEM generators: 4096 independent 8-state problems, each with its own drift and drive matrices."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dto_b200 as dto
from dto_b200 import problem_templates as pt

B = 4096
prob = pt.scaled_problem(N=200, state_dim=8, n_controls=2, generator_scale=0.35)
rng = np.random.default_rng(1)
Gs = 0.35 * rng.standard_normal((B, 3, 8, 8))
for shared in (False, True):
    ev = dto.Evaluator(prob, batch=B, batch_G=None if shared else Gs)
    dev = torch.device("cuda")
    Z = np.tile(prob.trajectory.datavec, B) + 0.01 * rng.standard_normal(B * ev.n_vars)
    dZ = torch.from_numpy(Z).to(dev)
    dmu = torch.rand(B * ev.n_constraints, dtype=torch.float64, device=dev)
    outs = [torch.empty(n, dtype=torch.float64, device=dev) for n in (B, B * ev.n_vars, B * ev.n_constraints, B * ev.nnz_jacobian, B * ev.nnz_hessian)]
    stream = torch.cuda.ExternalStream(ev.stream)
    step = lambda: ev.eval_all_dev(dZ.data_ptr(), 1.0, dmu.data_ptr(), *[o.data_ptr() for o in outs])
    step(); ev.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(5):
            step()
        e1.record(stream)
    ev.synchronize()
    print(f"c5 {'shared' if shared else 'per-problem'} generators: variant={ev.kernel_variant(0)} eval={e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
    ev.close()
