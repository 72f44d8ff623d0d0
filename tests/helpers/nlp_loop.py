"""A solver loop around the MOI-style callbacks, for tests: the reference hands its Evaluator to Ipopt /
MadNLP (src/solvers/ipopt_solver/solver.jl:68-85, ext/MadNLPSolverExt/solver.jl:81-97); neither solver
exists in this environment, so the convergence check of BASELINE.json (same objective within 1e-6, same
iteration count +-2) is run with SciPy's interior-point `trust-constr` driven once by the CPU oracle and
once by the CUDA evaluator through identical glue.  Bounds come from the trajectory the way
src/constraints/ builds them (initial/final equality bounds, symmetric or (lo, hi) box bounds)."""
from __future__ import annotations

import numpy as np
from scipy.optimize import Bounds, NonlinearConstraint, minimize
from scipy.sparse import coo_matrix


class OracleCallbacks:
    """The five MOI callbacks served by oracle/dto_oracle.py (CPU, tests only)."""

    def __init__(self, prob):
        import dto_oracle as orc

        self.orc, self.spec = orc, prob.to_spec()
        Z0 = prob.trajectory.vec()
        self.n_vars = Z0.size
        self.jstruct = orc.jacobian_structure(self.spec, Z0)
        self.hstruct = orc.hessian_structure(self.spec, Z0)
        n_dyn, n_nl = orc.n_constraints(self.spec)
        self.n_constraints = n_dyn + n_nl
        lo = np.zeros(self.n_constraints)
        off = n_dyn
        for c in self.spec.get("constraints", []):
            d = orc.constraint_dim(self.spec, c)
            if not c["equality"]:
                lo[off:off + d] = -np.inf
            off += d
        self.bounds = (lo, np.zeros(self.n_constraints))

    def objective(self, Z):
        return self.orc.eval_objective(self.spec, Z)

    def gradient(self, Z):
        return self.orc.eval_objective_gradient(self.spec, Z)

    def constraint(self, Z):
        return self.orc.eval_constraint(self.spec, Z)

    def jacobian(self, Z):
        return self.orc.eval_constraint_jacobian(self.spec, Z, self.jstruct)

    def hessian(self, Z, sigma, mu):
        return self.orc.eval_hessian_lagrangian(self.spec, Z, sigma, mu, self.hstruct)


class DeviceCallbacks:
    """The same five callbacks served by libdto_b200.so through the Python mirror of the reference API."""

    def __init__(self, prob):
        import dto_b200 as dto

        self.ev = dto.Evaluator(prob)
        self.n_vars, self.n_constraints = self.ev.n_vars, self.ev.n_constraints
        self.jstruct = self.ev.jacobian_structure()
        self.hstruct = self.ev.hessian_lagrangian_structure()
        self.bounds = self.ev.constraint_bounds()

    def objective(self, Z):
        return self.ev.eval_objective(Z)

    def gradient(self, Z):
        g = np.empty(self.n_vars)
        self.ev.eval_objective_gradient(g, Z)
        return g

    def constraint(self, Z):
        g = np.empty(self.n_constraints)
        self.ev.eval_constraint(g, Z)
        return g

    def jacobian(self, Z):
        v = np.empty(self.jstruct[0].size)
        self.ev.eval_constraint_jacobian(v, Z)
        return v

    def hessian(self, Z, sigma, mu):
        v = np.empty(self.hstruct[0].size)
        self.ev.eval_hessian_lagrangian(v, Z, sigma, mu)
        return v

    def close(self):
        self.ev.close()


def variable_bounds(traj):
    """Box bounds of the decision vector: initial/final values pin the first/last knot, `bounds` boxes every
    knot in between (src/constraints/linear/bounds_constraint.jl semantics: bounds skip pinned knots)."""
    N, z = traj.N, traj.dim
    lo, hi = np.full((z, N), -np.inf), np.full((z, N), np.inf)
    for name, b in traj.bounds.items():
        r = traj.components[name]
        if np.isscalar(b):
            l, u = -abs(float(b)) * np.ones(len(r)), abs(float(b)) * np.ones(len(r))
        else:
            l, u = (np.broadcast_to(np.asarray(x, float), (len(r),)) for x in b)
        lo[r, :], hi[r, :] = l[:, None], u[:, None]
    for name, v in traj.initial.items():
        lo[traj.components[name], 0] = hi[traj.components[name], 0] = np.asarray(v, float)
    for name, v in traj.final.items():
        lo[traj.components[name], -1] = hi[traj.components[name], -1] = np.asarray(v, float)
    g = traj.global_dim  # global variables are unbounded here
    return (np.concatenate([lo.reshape(-1, order="F"), np.full(g, -np.inf)]),
            np.concatenate([hi.reshape(-1, order="F"), np.full(g, np.inf)]))


def solve(prob, cb, max_iter=100, gtol=1e-8, xtol=1e-10):
    """minimize J(Z) s.t. lo <= g(Z) <= 0, box bounds; returns (Z, info)."""
    n, m = cb.n_vars, cb.n_constraints
    jr, jc = cb.jstruct[0] - 1, cb.jstruct[1] - 1
    hr, hc = cb.hstruct[0] - 1, cb.hstruct[1] - 1
    off = hr != hc

    def sym(v):
        return coo_matrix((np.concatenate([v, v[off]]), (np.concatenate([hr, hc[off]]), np.concatenate([hc, hr[off]]))), shape=(n, n)).tocsr()

    con = NonlinearConstraint(cb.constraint, cb.bounds[0], cb.bounds[1],
                              jac=lambda Z: coo_matrix((cb.jacobian(Z), (jr, jc)), shape=(m, n)).tocsr(),
                              hess=lambda Z, v: sym(cb.hessian(Z, 0.0, v)))
    lo, hi = variable_bounds(prob.trajectory)
    Z0 = np.clip(prob.trajectory.vec(), lo, hi)
    res = minimize(cb.objective, Z0, jac=cb.gradient, hess=lambda Z: sym(cb.hessian(Z, 1.0, np.zeros(m))), method="trust-constr",
                   constraints=[con], bounds=Bounds(lo, hi), options={"maxiter": max_iter, "gtol": gtol, "xtol": xtol, "verbose": 0})
    return res.x, {"objective": float(res.fun), "iterations": int(res.nit), "violation": float(res.constr_violation),
                   "status": int(res.status), "nfev": int(res.nfev)}
