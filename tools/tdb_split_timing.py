"""K7 at c3: device time of the Hessian-only call (forward + adjoint tiles) vs the Jacobian-only call (propagator tiles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dto_b200 as dto
from dto_b200 import problem_templates as pt
prob = pt.carrier_problem(N=1000, state_dim=64, n_drives=2)
ev = dto.Evaluator(prob)
dev = torch.device("cuda")
Z = torch.from_numpy(prob.trajectory.datavec.copy()).to(dev)
mu = torch.rand(ev.n_constraints, dtype=torch.float64, device=dev)
g = torch.empty(ev.n_constraints, dtype=torch.float64, device=dev)
jac = torch.empty(ev.nnz_jacobian, dtype=torch.float64, device=dev)
hess = torch.empty(ev.nnz_hessian, dtype=torch.float64, device=dev)
stream = torch.cuda.ExternalStream(ev.stream)
def timed(name, **kw):
    ev.eval_all_dev(Z.data_ptr(), 1.0, mu.data_ptr(), **kw); ev.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(3):
            ev.eval_all_dev(Z.data_ptr(), 1.0, mu.data_ptr(), **kw)
        e1.record(stream)
    ev.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/3:.3f} ms")
timed("residual only", dg=g.data_ptr())
timed("residual + Jacobian", dg=g.data_ptr(), djac=jac.data_ptr())
timed("Hessian only", dhess=hess.data_ptr())
timed("all", dg=g.data_ptr(), djac=jac.data_ptr(), dhess=hess.data_ptr())
