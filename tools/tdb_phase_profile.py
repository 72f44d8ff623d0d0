"""Cycle breakdown of one K7 right-hand side (debug build: DTO_EXTRA_NVCC_FLAGS=-DDTO_TDB_PROFILE python directtrajopt.jl_b200/build.py)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dto_b200 as dto
from dto_b200 import problem_templates as pt
prob = pt.carrier_problem(N=149, state_dim=64, n_drives=2)
ev = dto.Evaluator(prob)
lib = ev._lib
Z = prob.trajectory.datavec.copy()
mu = np.random.default_rng(0).random(ev.n_constraints)
bufs = [np.empty(1), np.empty(ev.n_vars), np.empty(ev.n_constraints), np.empty(ev.nnz_jacobian), np.empty(ev.nnz_hessian)]
out = (ctypes.c_ulonglong * 256)()
ev.eval_all(Z, 1.0, mu, *bufs)
lib.dto_debug_tdb_profile(out, 1)
ev.eval_all(Z, 1.0, mu, *bufs)
lib.dto_debug_tdb_profile(out, 0)
a = np.array(list(out), dtype=np.float64).reshape(2, 16, 8)
names = ["wait drift", "assemble+publish", "barrier 1", "products", "barrier 2", "couple+update"]
for y in range(2):
    for w in range(16):
        n = a[y, w, 6]
        if n == 0: continue
        print(f"group {y} warp {w:2d}: " + "  ".join(f"{nm} {a[y, w, i] / n:7.0f}" for i, nm in enumerate(names)) + f"   total {a[y, w, :6].sum() / n:7.0f}  coupling-only {a[y, w, 7] / n:7.0f}  (rhs {int(n)})")
