"""Generate tests/golden/*.npz.  NOT reference output: neither Julia nor the reference's packages exist in
this environment (parity unpinned, see oracle/dto_oracle.py).  These are ORACLE-generated regression vectors:
they pin the oracle against accidental change (CPU test) and give the GPU tests a fixed, committed set of
inputs/outputs that does not depend on scipy's version on the box.  The one literal numeric fixture the
reference holds (test/test_utils.jl:55-111, the 15x5 `named_trajectory_type_1` matrix) is stored as input data.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import dto_b200 as dto  # noqa: E402
import dto_oracle as orc  # noqa: E402
from dto_b200 import problem_templates as pt  # noqa: E402

# test/test_utils.jl:57-73 (rows: U~ (8), a (2), da (2), dda (2), dt (1); 5 knots)
TYPE1 = np.array([
    [1.0, 0.957107, 0.853553, 0.75, 0.707107],
    [0.0, 0.103553, 0.353553, 0.603553, 0.707107],
    [0.0, 0.103553, 0.146447, 0.103553, 1.38778e-17],
    [0.0, -0.25, -0.353553, -0.25, -1.52656e-16],
    [0.0, 0.103553, 0.353553, 0.603553, 0.707107],
    [1.0, 0.75, 0.146447, -0.457107, -0.707107],
    [0.0, -0.25, -0.353553, -0.25, -1.249e-16],
    [0.0, 0.603553, 0.853553, 0.603553, 4.16334e-16],
    [0.0, -0.243953, 0.959151, -0.665253, 0.0],
    [0.0, 0.0139165, 0.668917, 0.625329, 0.0],
    [0.00393491, 0.0240775, -0.00942396, 0.00329391, 0.00941354],
    [-0.00223794, -0.0105816, 0.00328457, 0.0204239, 0.0253415],
    [0.0058186, 0.00686586, -0.00422555, 0.00442631, 0.000319156],
    [-0.00134597, -0.00120682, 0.0114915, 0.00189333, -0.0251649],
    [0.2, 0.2, 0.2, 0.2, 0.2],
])


def type1_problem():
    """named_trajectory_type_1 with the DerivativeIntegrator chain the reference tests on it
    (src/integrators/derivative_integrator.jl:118-123) plus a bilinear integrator on the 8-dim iso state."""
    traj = dto.NamedTrajectory({"U": TYPE1[0:8], "a": TYPE1[8:10], "da": TYPE1[10:12], "dda": TYPE1[12:14], "dt": TYPE1[14:15]},
                               controls=("dda", "dt"), timestep="dt", goal={"U": [0.0, 1.0, 0, 0, 1.0, 0, 0, 0]})
    rng = np.random.default_rng(0)
    G0 = rng.standard_normal((8, 8))
    Gd = [rng.standard_normal((8, 8)) for _ in range(2)]
    ints = [dto.BilinearIntegrator((G0, Gd), "U", "a", traj), dto.DerivativeIntegrator("a", "da", traj), dto.DerivativeIntegrator("da", "dda", traj)]
    J = dto.QuadraticRegularizer("dda", traj, 1.0) + dto.TerminalObjective(dto.SqDist(traj.goal["U"]), "U", traj)
    return dto.DirectTrajOptProblem(traj, J, ints)


CASES = {
    "type1": type1_problem,
    "standard_N10": lambda: pt.standard_problem(N=10, seed=0),
    "readme_N50": lambda: pt.readme_problem(N=50, seed=42),
    "gate_n32_N6": lambda: pt.quantum_gate_problem(N=6, levels=16, n_drives=4, seed=42),
    # global variables: all four global component kinds with the reference's own test functions
    "global_N7": lambda: pt.global_problem(N=7, seed=0),
    # small states (n = 8, 16): the eight-intervals-per-warp kernel, a partial octet
    "scaled_n8_N12": lambda: pt.scaled_problem(N=12, state_dim=8, n_controls=2, generator_scale=0.35),
    "scaled_n16_N11": lambda: pt.scaled_problem(N=11, state_dim=16, n_controls=2, generator_scale=0.25),
}


def evaluate(prob, seed=123):
    spec = prob.to_spec()
    Z = prob.trajectory.vec()
    rng = np.random.default_rng(seed)
    jst, hst = orc.jacobian_structure(spec, Z), orc.hessian_structure(spec, Z)
    nd, nn = orc.n_constraints(spec)
    mu = rng.random(nd + nn)
    sigma = 1.5
    return dict(Z=Z, mu=mu, sigma=sigma, jac_rows=jst[0], jac_cols=jst[1], hess_rows=hst[0], hess_cols=hst[1],
                J=orc.eval_objective(spec, Z), grad=orc.eval_objective_gradient(spec, Z), g=orc.eval_constraint(spec, Z),
                jac=orc.eval_constraint_jacobian(spec, Z, jst), hess=orc.eval_hessian_lagrangian(spec, Z, sigma, mu, hst))


if __name__ == "__main__":
    for name, build in CASES.items():
        if os.path.exists(os.path.join(HERE, f"{name}.npz")) and "--all" not in sys.argv:
            continue  # committed vectors are never regenerated silently
        out = evaluate(build())
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
