"""Block-relative error of the CUDA evaluator against the oracle (exact Frechet derivatives) with and without the
alpha_2 series plan: the shorter series must not cost accuracy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import dto_b200 as dto, dto_oracle as orc
from dto_b200 import problem_templates as pt
from helpers import blocks

cases = {"gate_n32": lambda: pt.quantum_gate_problem(N=12, levels=16, n_drives=4), "scaled_n16": lambda: pt.scaled_problem(N=12, state_dim=16, n_controls=2, generator_scale=0.25),
         "scaled_n8": lambda: pt.scaled_problem(N=12, state_dim=8, n_controls=2, generator_scale=0.35), "n16_theta12": lambda: pt.scaled_problem(N=4, state_dim=16, n_controls=2, generator_scale=12.0),
         "n32_theta3": lambda: pt.scaled_problem(N=4, state_dim=32, n_controls=3, generator_scale=3.0), "gate_n64": lambda: pt.quantum_gate_problem(N=5, levels=32, n_drives=2)}
for name, mk in cases.items():
    prob = mk(); spec = prob.to_spec(); rng = np.random.default_rng(7)
    Z0 = prob.trajectory.vec(); Z = Z0 + 0.05 * rng.standard_normal(Z0.size)
    dts = slice(spec["components"][spec["timestep"]][0], spec["N"] * spec["z"], spec["z"]); Z[dts] = np.abs(Z[dts])
    jst, hst = orc.jacobian_structure(spec, Z0), orc.hessian_structure(spec, Z0)
    nd, nn = orc.n_constraints(spec); mu = rng.random(nd + nn)
    gref, Jref, Href = orc.eval_constraint(spec, Z), orc.eval_constraint_jacobian(spec, Z, jst), orc.eval_hessian_lagrangian(spec, Z, 1.0, mu, hst)
    for plan in ("1", "0"):
        os.environ["DTO_B200_SERIES_PLAN"] = plan
        ev = dto.Evaluator(prob)
        g, J, H = np.empty(ev.n_constraints), np.empty(ev.nnz_jacobian), np.empty(ev.nnz_hessian)
        ev.eval_all(Z, 1.0, mu, None, None, g, J, H)
        print(f"{name:12s} plan={plan} variant={ev.kernel_variant(0):10s} residual {blocks.vec_block_relerr(spec, g, gref):.2e}  jacobian {blocks.jac_block_relerr(spec, jst, J, Jref):.2e}  "
              f"hessian {blocks.hess_block_relerr(spec, hst, H, Href):.2e}", flush=True)
        ev.close()
