"""Time the K1 interval kernel per requested output set on the c2 workload (device-resident)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dto_b200 as dto

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
prob = dto.problem_templates.quantum_gate_problem(N=N, levels=16, n_drives=4)
ev = dto.Evaluator(prob)
dev = torch.device("cuda")
Z = torch.from_numpy(prob.trajectory.datavec.copy()).to(dev)
mu = torch.rand(ev.n_constraints, dtype=torch.float64, device=dev)
g = torch.empty(ev.n_constraints, dtype=torch.float64, device=dev)
jac = torch.empty(ev.nnz_jacobian, dtype=torch.float64, device=dev)
hess = torch.empty(ev.nnz_hessian, dtype=torch.float64, device=dev)
J = torch.empty(1, dtype=torch.float64, device=dev); grad = torch.empty(ev.n_vars, dtype=torch.float64, device=dev)
stream = torch.cuda.ExternalStream(ev.stream)
cases = {"g": dict(dg=g.data_ptr()), "g+jac": dict(dg=g.data_ptr(), djac=jac.data_ptr()), "hess": dict(dhess=hess.data_ptr(), dmu=mu.data_ptr()),
         "g+jac+hess": dict(dg=g.data_ptr(), djac=jac.data_ptr(), dhess=hess.data_ptr(), dmu=mu.data_ptr()), "obj+grad": dict(dJ=J.data_ptr(), dgrad=grad.data_ptr())}
for name, kw in cases.items():
    for _ in range(3):
        ev.eval_all_dev(Z.data_ptr(), 1.0, **kw)
    ev.synchronize()
    ev.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(10):
            ev.eval_all_dev(Z.data_ptr(), 1.0, **kw)
        e1.record(stream)
    ev.synchronize()
    ms, n = ev.kernel_time_ms()
    print(f"{name:12s} total {e0.elapsed_time(e1)/10*1e3:8.1f} us/eval   K1 {ms/max(n,1)*1e3:8.1f} us")
