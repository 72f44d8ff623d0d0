"""Jacobian-vector product time at c2: matrix-free (series) vs materialise-then-multiply (the reference's way)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dto_b200 as dto
from dto_b200 import problem_templates as pt

prob = pt.quantum_gate_problem(N=2000, levels=16, n_drives=4)
Z = prob.trajectory.datavec.copy()
for mode in ("", "materialize"):
    if mode:
        os.environ["DTO_B200_JVP"] = mode
    ev = dto.Evaluator(prob)
    rng = np.random.default_rng(0)
    w1, w2 = rng.standard_normal(ev.n_vars), rng.standard_normal(ev.n_constraints)
    y1, y2 = np.empty(ev.n_constraints), np.empty(ev.n_vars)
    for name, f in (("J w", lambda: ev.eval_constraint_jacobian_product(y1, Z, w1)), ("J' w", lambda: ev.eval_constraint_jacobian_transpose_product(y2, Z, w2))):
        f(); f()
        t = time.perf_counter()
        for _ in range(20):
            f()
        print(f"{mode or 'matrix-free':12s} {name}: {(time.perf_counter() - t) / 20 * 1e3:.3f} ms (host call, H2D of Z and w, D2H of y included)")
    ev.close()
