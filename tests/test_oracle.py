"""Pin the CPU oracle (oracle/dto_oracle.py) the way the reference pins its own evaluator:
central finite differences of the residual / objective (src/integrators/_integrators.jl:97-242,
src/objectives/_objectives.jl:261-336, src/solvers/evaluator.jl:649-791), plus 50-digit mpmath
matrix exponentials for the Frechet terms, plus the literal fixtures the reference holds."""
import numpy as np
import pytest

import dto_b200 as dto
import dto_oracle as orc
from dto_b200 import problem_templates as pt


def dense(vals, structure, shape):
    M = np.zeros(shape)
    M[structure[0] - 1, structure[1] - 1] = vals
    return M


def fd_jacobian(f, x, h=1e-6):
    f0 = f(x)
    J = np.zeros((f0.size, x.size))
    for i in range(x.size):
        e = np.zeros_like(x)
        e[i] = h
        J[:, i] = (f(x + e) - f(x - e)) / (2 * h)
    return J


PROBLEMS = {
    "readme": lambda: pt.readme_problem(N=6),
    "standard": lambda: pt.standard_problem(N=5),
    "evaluator_test": lambda: pt.evaluator_test_problem(N=5),
    "scaled": lambda: pt.scaled_problem(N=4, state_dim=5, n_controls=2, generator_scale=0.7),
    "gate": lambda: pt.quantum_gate_problem(N=4, levels=3, n_drives=2),
    "linreg": lambda: pt.linear_regularizer_problem(N=5),
    "global": lambda: pt.global_problem(N=5),
    "global_ref_fixture": lambda: pt.global_problem(N=4, with_goal=False),
    "global_goal": lambda: pt.global_goal_problem(N=5),
}


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_oracle_jacobian_vs_finite_differences(name):
    prob = PROBLEMS[name]()
    spec = prob.to_spec()
    Z = prob.trajectory.vec()
    st = orc.jacobian_structure(spec, Z)
    nd, nn = orc.n_constraints(spec)
    J = dense(orc.eval_constraint_jacobian(spec, Z, st), st, (nd + nn, Z.size))
    Jfd = fd_jacobian(lambda z: orc.eval_constraint(spec, z), Z)
    assert np.allclose(J, Jfd, atol=1e-6, rtol=1e-6)  # evaluator.jl:751 tolerance


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_oracle_gradient_and_hessian_vs_finite_differences(name):
    prob = PROBLEMS[name]()
    spec = prob.to_spec()
    Z = prob.trajectory.vec()
    rng = np.random.default_rng(3)
    nd, nn = orc.n_constraints(spec)
    mu = rng.random(nd + nn)
    sigma = 2.0
    g = orc.eval_objective_gradient(spec, Z)
    gfd = fd_jacobian(lambda z: np.array([orc.eval_objective(spec, z)]), Z)[0]
    assert np.allclose(g, gfd, atol=1e-6, rtol=1e-6)
    hst = orc.hessian_structure(spec, Z)
    H = dense(orc.eval_hessian_lagrangian(spec, Z, sigma, mu, hst), hst, (Z.size, Z.size))
    jst = orc.jacobian_structure(spec, Z)

    def lag_grad(z):  # gradient of sigma*J + mu'g from the (already FD-checked) first derivatives
        Jm = dense(orc.eval_constraint_jacobian(spec, z, jst), jst, (nd + nn, Z.size))
        return sigma * orc.eval_objective_gradient(spec, z) + Jm.T @ mu

    Hfd = fd_jacobian(lag_grad, Z, h=1e-5)
    Hfd = np.triu((Hfd + Hfd.T) / 2)
    # the reference drops the (v, dt) QuadraticRegularizer entries when dt sits before v; none of
    # these fixtures does that, so the upper triangles must agree
    assert np.allclose(H, Hfd, atol=2e-6, rtol=1e-5)


def test_second_order_frechet_against_mpmath():
    """mu' L2(A; Bi, Bj) x from block-triangular scipy expm vs 50-digit differentiation."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    rng = np.random.default_rng(0)
    n, m = 3, 2
    G = rng.standard_normal((m + 1, n, n))
    x, xn, mu = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    u = rng.standard_normal(m) * 0.3
    dt = 0.4
    spec = {"N": 2, "z": n + m + 1, "timestep": "dt", "components": {"x": (0, n), "u": (n, m), "dt": (n + m, 1)}}
    it = {"kind": "bilinear", "x": "x", "u": "u", "G": G}
    zk = np.concatenate([x, u, [dt]])
    zk1 = np.concatenate([xn, u, [dt]])
    _, Jb, Hb = orc._bilinear_interval(spec, it, zk, zk1, mu)

    def f_mp(uu, dd):
        A = mp.matrix(n, n)
        for r in range(n):
            for c in range(n):
                A[r, c] = dd * (mp.mpf(G[0][r, c]) + sum(uu[i] * mp.mpf(G[1 + i][r, c]) for i in range(m)))
        E = mp.expm(A, method="taylor")
        w = E * mp.matrix([mp.mpf(v) for v in x])
        return sum(mp.mpf(mu[r]) * (mp.mpf(xn[r]) - w[r]) for r in range(n))

    u_mp = [mp.mpf(v) for v in u]
    h = mp.mpf(10) ** -12

    def d2(i, j):
        def shifted(si, sj):
            uu = list(u_mp)
            dd = mp.mpf(dt)
            for idx, s in ((i, si), (j, sj)):
                if idx < m:
                    uu[idx] += s * h
                else:
                    dd += s * h
            return f_mp(uu, dd)

        return (shifted(1, 1) - shifted(1, -1) - shifted(-1, 1) + shifted(-1, -1)) / (4 * h * h)

    for i in range(m + 1):
        for j in range(i, m + 1):
            ref = float(d2(i, j))
            got = Hb[n + i, n + j]
            assert abs(got - ref) <= 1e-11 * max(1.0, abs(ref)), (i, j, got, ref)


def test_structure_counts_match_published_benchmark():
    """docs/src/benchmarks.md: N=51 bilinear problem has 561 variables, 400 rows, 8800 / 9416 nnz."""
    prob = pt.bilinear_benchmark(N=51)
    spec = prob.to_spec()
    Z = prob.trajectory.datavec
    assert Z.size == 561
    assert sum(orc.n_constraints(spec)) == 400
    assert orc.jacobian_structure(spec, Z)[0].size == 8800
    assert orc.hessian_structure(spec, Z)[0].size == 9416


def test_structure_order_is_column_major_upper():
    prob = pt.standard_problem(N=4)
    spec = prob.to_spec()
    Z = prob.trajectory.datavec
    for r, c in (orc.jacobian_structure(spec, Z), orc.hessian_structure(spec, Z)):
        key = c.astype(np.int64) * (r.max() + 1) + r
        assert np.all(np.diff(key) > 0)
    r, c = orc.hessian_structure(spec, Z)
    assert np.all(r <= c)


def test_quadreg_delta_t_squared_quirk():
    """J = sum 1/2 (dt dv)' R (dt dv): dt enters squared (regularizers.jl:79-90), not linearly as
    the docstring says."""
    prob = pt.readme_problem(N=3)
    spec = prob.to_spec()
    Z = prob.trajectory.vec()
    u = prob.trajectory.u[0]
    assert np.isclose(orc.eval_objective(spec, Z), 0.5 * np.sum((0.1 * u) ** 2))


@pytest.mark.parametrize("name", ["standard", "scaled", "gate"])
def test_c_port_of_reference_algorithm_agrees_with_analytic_oracle(name):
    """Two independent restatements -- forward-mode jets through the truncated Taylor expv (the
    reference's algorithm, oracle/dto_oracle.c) and exact Frechet derivatives from block-triangular
    scipy expm (oracle/dto_oracle.py) -- must agree to 1e-12."""
    import dto_oracle_c as oc

    prob = PROBLEMS[name]()
    spec = prob.to_spec()
    Z = prob.trajectory.datavec
    z = spec["z"]
    it = spec["integrators"][0]
    mu = np.random.default_rng(1).random(it["G"].shape[1])
    for k in range(spec["N"] - 1):
        zk, zk1 = Z[k * z : (k + 1) * z], Z[(k + 1) * z : (k + 2) * z]
        r, J, H = orc._bilinear_interval(spec, it, zk, zk1, mu)
        for all_dirs in (True, False):
            r2, J2, H2 = oc.bilinear_interval(spec, it, zk, zk1, mu, all_dirs=all_dirs)
            assert np.abs(r - r2).max() <= 1e-12 * max(1, np.abs(r).max())
            assert np.abs(J - J2).max() <= 1e-12 * np.abs(J).max()
            assert np.abs(H - H2).max() <= 1e-12 * np.abs(H).max()
