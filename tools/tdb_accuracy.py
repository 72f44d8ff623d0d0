"""Block-relative error of the TimeDependentBilinearIntegrator kernels against the oracle's tight variational solve, with the
extrapolation columns sized per interval (default) and with eight columns everywhere (DTO_B200_TDB_TOL=0), and the c3
evaluation time of both: fewer right-hand sides must not cost accuracy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import dto_b200 as dto, dto_oracle as orc
from dto_b200 import problem_templates as pt
from helpers import blocks

tols = sys.argv[1:] or ["1e-14", "0"]
cases = {"c3shape_n64": (lambda: pt.carrier_problem(N=6, state_dim=64, n_drives=2, spline_order=1), 1.0),
         "n16_order1": (lambda: pt.carrier_problem(N=6, state_dim=16, n_drives=2, spline_order=1), 1.0),
         "n16_dt_x3": (lambda: pt.carrier_problem(N=5, state_dim=16, n_drives=2, spline_order=1), 3.0),
         "n16_dt_x8": (lambda: pt.carrier_problem(N=5, state_dim=16, n_drives=2, spline_order=1), 8.0),
         "n8_order0_x20": (lambda: pt.carrier_problem(N=5, state_dim=8, n_drives=2, spline_order=0), 20.0)}
for name, (mk, dtscale) in cases.items():
    prob = mk(); spec = prob.to_spec(); rng = np.random.default_rng(7)
    Z0 = prob.trajectory.vec(); Z = Z0 + 0.02 * rng.standard_normal(Z0.size)
    dts = slice(spec["components"][spec["timestep"]][0], spec["N"] * spec["z"], spec["z"]); Z[dts] = np.abs(Z[dts]) * dtscale
    jst, hst = orc.jacobian_structure(spec, Z0), orc.hessian_structure(spec, Z0)
    nd, nn = orc.n_constraints(spec); mu = rng.random(nd + nn)
    gref, Jref, Href = orc.eval_constraint(spec, Z), orc.eval_constraint_jacobian(spec, Z, jst), orc.eval_hessian_lagrangian(spec, Z, 1.0, mu, hst)
    for tol in tols:
        os.environ["DTO_B200_TDB_TOL"] = tol
        ev = dto.Evaluator(prob)
        g, J, H = np.empty(ev.n_constraints), np.empty(ev.nnz_jacobian), np.empty(ev.nnz_hessian)
        ev.eval_all(Z, 1.0, mu, None, None, g, J, H)
        print(f"{name:14s} tol={tol:6s} variant={ev.kernel_variant(0):9s} residual {blocks.vec_block_relerr(spec, g, gref):.2e}  jacobian {blocks.jac_block_relerr(spec, jst, J, Jref):.2e}  "
              f"hessian {blocks.hess_block_relerr(spec, hst, H, Href):.2e}", flush=True)
        ev.close()
# c3 timing
prob = pt.carrier_problem(N=1000, state_dim=64, n_drives=2)
for tol in tols:
    os.environ["DTO_B200_TDB_TOL"] = tol
    ev = dto.Evaluator(prob)
    dev = torch.device("cuda")
    dZ = torch.from_numpy(prob.trajectory.datavec + 0.01 * np.random.default_rng(0).standard_normal(ev.n_vars)).to(dev)
    dmu = torch.rand(ev.n_constraints, dtype=torch.float64, device=dev)
    out = [torch.empty(k, dtype=torch.float64, device=dev) for k in (1, ev.n_vars, ev.n_constraints, ev.nnz_jacobian, ev.nnz_hessian)]
    stream = torch.cuda.ExternalStream(ev.stream)
    step = lambda: ev.eval_all_dev(dZ.data_ptr(), 1.0, dmu.data_ptr(), *[o.data_ptr() for o in out])
    step(); ev.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(3): step()
        e1.record(stream)
    ev.synchronize()
    print(f"c3 (n=64, N=1000) tol={tol}: {e0.elapsed_time(e1) / 3:.3f} ms per evaluation", flush=True)
    ev.close()
