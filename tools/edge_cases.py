import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np
import dto_b200 as dto, dto_oracle as orc
from dto_b200 import problem_templates as pt

def relerr(a,b):
    s=max(np.abs(b).max() if b.size else 0,1e-300); return np.abs(a-b).max()/s if b.size else 0.0

def check(name, prob, sigma=1.3):
    spec=prob.to_spec(); Z0=prob.trajectory.datavec.copy()
    ev=dto.Evaluator(prob)
    rng=np.random.default_rng(1)
    Z=Z0+0.01*rng.standard_normal(Z0.size); mu=rng.random(ev.n_constraints)
    jst,hst=orc.jacobian_structure(spec,Z0),orc.hessian_structure(spec,Z0)
    jr,jc=ev.jacobian_structure(); hr,hc=ev.hessian_lagrangian_structure()
    ok=np.array_equal(jr,jst[0]) and np.array_equal(jc,jst[1]) and np.array_equal(hr,hst[0]) and np.array_equal(hc,hst[1])
    J,grad,g,jac,hess=np.empty(1),np.empty(ev.n_vars),np.empty(ev.n_constraints),np.empty(ev.nnz_jacobian),np.empty(ev.nnz_hessian)
    ev.eval_all(Z,sigma,mu,J,grad,g,jac,hess)
    errs=[relerr(g,orc.eval_constraint(spec,Z)),relerr(jac,orc.eval_constraint_jacobian(spec,Z,jst)),relerr(hess,orc.eval_hessian_lagrangian(spec,Z,sigma,mu,hst)),relerr(grad,orc.eval_objective_gradient(spec,Z)),abs(J[0]-orc.eval_objective(spec,Z))]
    print(name, ev.kernel_variant(0), "struct", ok, ["%.1e"%e for e in errs]); ev.close()

def no_drive(n, N):
    rng=np.random.default_rng(3)
    G0=rng.standard_normal((n,n))/n
    traj=dto.NamedTrajectory({"x":rng.standard_normal((n,N)),"u":rng.standard_normal((1,N)),"dt":np.full(N,0.1)},timestep="dt",controls=("u",))
    # drive matrix identically zero is still a drive; m=0 needs a component of dim 0: use zero-matrix drive
    G=lambda u: G0 + u[0]*np.zeros((n,n))
    return dto.DirectTrajOptProblem(traj, dto.QuadraticRegularizer("u",traj,1.0), dto.BilinearIntegrator(G,"x","u",traj))

check("N=2 n=8", pt.scaled_problem(N=2,state_dim=8,n_controls=2))
check("N=2 readme", pt.readme_problem(N=2))
check("N=3 n=32 m=1", pt.scaled_problem(N=3,state_dim=32,n_controls=1))
check("zero drive n=16", no_drive(16,4))
check("sigma=0", pt.standard_problem(N=5), sigma=0.0)
check("n=64 m=4", pt.scaled_problem(N=3,state_dim=64,n_controls=4,generator_scale=0.4))
check("n=48 m=4", pt.scaled_problem(N=3,state_dim=48,n_controls=4,generator_scale=0.4))
check("n=16 theta big", pt.scaled_problem(N=4,state_dim=16,n_controls=2,generator_scale=12.0))
check("n=32 theta 3", pt.scaled_problem(N=4,state_dim=32,n_controls=3,generator_scale=3.0))
