import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import dto_b200 as dto
from dto_b200 import problem_templates as pt
prob = pt.carrier_problem(N=150, state_dim=64, n_drives=2)
ev = dto.Evaluator(prob)
print(ev.kernel_variant(0))
Z = prob.trajectory.datavec.copy()
mu = np.random.default_rng(0).random(ev.n_constraints)
bufs = [np.empty(1), np.empty(ev.n_vars), np.empty(ev.n_constraints), np.empty(ev.nnz_jacobian), np.empty(ev.nnz_hessian)]
import time
for i in range(3):
    t=time.time(); ev.eval_all(Z, 1.0, mu, *bufs); print("eval ms", (time.time()-t)*1e3)
