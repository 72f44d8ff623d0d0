"""The block-wise parity helpers checked against the oracle itself (no GPU): the closed-form layout used by the full-size
sampled checks must land on the entries of the oracle's structures, and a perturbation of one small block must be seen
at its own scale."""
import numpy as np

import dto_oracle as orc
from dto_b200 import problem_templates as pt
from helpers import blocks


def _oracle_outputs(prob, sigma=1.3, seed=0):
    spec, Z0 = prob.to_spec(), prob.trajectory.vec()
    rng = np.random.default_rng(seed)
    Z = Z0 + 0.03 * rng.standard_normal(Z0.size)
    jst, hst = orc.jacobian_structure(spec, Z0), orc.hessian_structure(spec, Z0)
    nd, nn = orc.n_constraints(spec)
    mu = rng.random(nd + nn)
    return spec, Z, mu, jst, hst, orc.eval_constraint(spec, Z), orc.eval_constraint_jacobian(spec, Z, jst), \
        orc.eval_hessian_lagrangian(spec, Z, sigma, mu, hst)


def test_sampled_interval_check_agrees_with_full_oracle():
    for prob in (pt.scaled_problem(N=7, state_dim=8, n_controls=2), pt.quantum_gate_problem(N=5, levels=3, n_drives=2),
                 pt.bilinear_benchmark(N=6)):
        spec, Z, mu, jst, hst, g, jac, hess = _oracle_outputs(prob)
        er, ej, eh = blocks.sampled_interval_check(spec, Z, 1.3, mu, g, jac, hess, range(1, spec["N"] + 1))
        assert max(er, ej, eh) < 1e-12, (er, ej, eh)
        # the closed-form positions are the oracle's structure entries
        z = spec["z"]
        for k in (1, 2, spec["N"] - 1):
            off = 0
            for i, pos in enumerate(blocks.jac_block_positions(spec, k)):
                d = pos.shape[0]
                assert np.array_equal(jst[0][pos] - 1, np.broadcast_to(off + (k - 1) * d + np.arange(d)[:, None], pos.shape))
                assert np.array_equal(jst[1][pos] - 1, np.broadcast_to((k - 1) * z + np.arange(2 * z)[None, :], pos.shape))
                off += d * (spec["N"] - 1)
        cross, diag = blocks.hess_region_positions(spec, 3)
        assert np.array_equal(hst[0][cross] - 1, np.broadcast_to(z + np.arange(z)[:, None], (z, z)))
        assert np.array_equal(hst[1][cross] - 1, np.broadcast_to(2 * z + np.arange(z)[None, :], (z, z)))
        iu = np.triu_indices(z)
        assert np.array_equal(hst[0][diag[iu]] - 1, 2 * z + iu[0]) and np.array_equal(hst[1][diag[iu]] - 1, 2 * z + iu[1])


def test_block_relerr_sees_small_blocks():
    prob = pt.standard_problem(N=6)
    spec, Z, mu, jst, hst, g, jac, hess = _oracle_outputs(prob)
    assert blocks.jac_block_relerr(spec, jst, jac, jac) == 0.0 and blocks.hess_block_relerr(spec, hst, hess, hess) == 0.0
    # d r / d u entries are ~1e-2: an absolute error of 1e-11 there is invisible to the array's max-norm (identity = 1)
    uo, ud = spec["components"]["u"]
    sel = np.flatnonzero(((jst[1] - 1) % spec["z"] == uo) & (jst[0] <= 4) & ((jst[1] - 1) // spec["z"] == 0))
    bad = jac.copy()
    bad[sel[0]] += 1e-11
    whole = np.abs(bad - jac).max() / np.abs(jac).max()
    assert whole <= 1e-10 < blocks.jac_block_relerr(spec, jst, bad, jac)
    bad = hess.copy()
    k = np.argmin(np.where(hess != 0, np.abs(hess), np.inf))
    bad[k] += 1e-9 * np.abs(hess).max()
    assert blocks.hess_block_relerr(spec, hst, bad, hess) >= 1e-9
    assert blocks.vec_block_relerr(spec, g, g) == 0.0


def test_sampled_interval_check_with_knot_constraints():
    """Constrained problems (c3 shape): interval blocks are looked up in the structure, constraint Hessians are added."""
    for prob in (pt.standard_problem(N=6), pt.carrier_problem(N=4, state_dim=4, n_drives=2, dt=0.1)):
        spec, Z, mu, jst, hst, g, jac, hess = _oracle_outputs(prob)
        er, ej, eh = blocks.sampled_interval_check(spec, Z, 1.3, mu, g, jac, hess, range(1, spec["N"] + 1), jac_structure=jst)
        assert max(er, ej, eh) < 1e-9, (er, ej, eh)
