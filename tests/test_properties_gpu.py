"""Size-independent properties at BASELINE.json's full size (config c2: n=32, 4 drives, N=2000), where the
CPU oracle would take minutes: adjoint identity of the two matrix-free products, their agreement with the COO
values, central finite differences of the residual along a direction (the reference's own test method,
src/integrators/_integrators.jl:157-191, applied through the C ABI), finite differences of the Lagrangian
gradient against the Hessian values, determinism of repeated evaluations, agreement of the K1 variants."""
import numpy as np
import pytest

import dto_b200 as dto
from dto_b200 import problem_templates as pt

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    prob = pt.quantum_gate_problem(N=2000, levels=16, n_drives=4)
    ev = dto.Evaluator(prob)
    yield prob, ev
    ev.close()


def _evaluate(ev, Z, sigma, mu):
    J, grad = np.empty(1), np.empty(ev.n_vars)
    g, jac, hess = np.empty(ev.n_constraints), np.empty(ev.nnz_jacobian), np.empty(ev.nnz_hessian)
    ev.eval_all(Z, sigma, mu, J, grad, g, jac, hess)
    return J[0], grad, g, jac, hess


def test_products_adjoint_identity_and_coo_agreement(c2):
    prob, ev = c2
    rng = np.random.default_rng(0)
    Z = prob.trajectory.datavec + 0.01 * rng.standard_normal(ev.n_vars)
    v, w = rng.standard_normal(ev.n_vars), rng.standard_normal(ev.n_constraints)
    Jv, JTw = np.empty(ev.n_constraints), np.empty(ev.n_vars)
    ev.eval_constraint_jacobian_product(Jv, Z, v)
    ev.eval_constraint_jacobian_transpose_product(JTw, Z, w)
    assert abs(w @ Jv - JTw @ v) <= 1e-11 * np.linalg.norm(w) * np.linalg.norm(Jv)
    vals = np.empty(ev.nnz_jacobian)
    ev.eval_constraint_jacobian(vals, Z)
    r, c = ev.jacobian_structure()
    ref = np.zeros(ev.n_constraints)
    np.add.at(ref, r - 1, vals * v[c - 1])
    assert np.abs(Jv - ref).max() <= 1e-11 * np.abs(ref).max()
    refT = np.zeros(ev.n_vars)
    np.add.at(refT, c - 1, vals * w[r - 1])
    assert np.abs(JTw - refT).max() <= 1e-11 * np.abs(refT).max()


def test_finite_differences_at_full_size(c2):
    prob, ev = c2
    rng = np.random.default_rng(1)
    Z = prob.trajectory.datavec + 0.01 * rng.standard_normal(ev.n_vars)
    mu = rng.random(ev.n_constraints)
    v = rng.standard_normal(ev.n_vars)
    v /= np.linalg.norm(v)
    sigma, h = 0.7, 1e-5
    _, grad, g, jac, hess = _evaluate(ev, Z, sigma, mu)
    _, gp, rp, jp, _ = _evaluate(ev, Z + h * v, sigma, mu)
    _, gm, rm, jm, _ = _evaluate(ev, Z - h * v, sigma, mu)
    r, c = ev.jacobian_structure()
    Jv = np.zeros(ev.n_constraints)
    np.add.at(Jv, r - 1, jac * v[c - 1])
    fd = (rp - rm) / (2 * h)
    assert np.abs(fd - Jv).max() <= 1e-7 * max(np.abs(Jv).max(), 1.0)
    # gradient of the Lagrangian sigma*J + mu'g along v against the (upper-triangle) Hessian values
    def lag_grad(gr, jv):
        out = sigma * gr.copy()
        np.add.at(out, c - 1, jv * mu[r - 1])
        return out
    fdH = (lag_grad(gp, jp) - lag_grad(gm, jm)) / (2 * h)
    hr, hc = ev.hessian_lagrangian_structure()
    Hv = np.zeros(ev.n_vars)
    np.add.at(Hv, hr - 1, hess * v[hc - 1])
    off = hr != hc
    np.add.at(Hv, hc[off] - 1, hess[off] * v[hr[off] - 1])
    assert np.abs(fdH - Hv).max() <= 1e-6 * max(np.abs(Hv).max(), 1.0)


def test_repeated_evaluations_are_bit_identical(c2):
    """The persistent kernel hands out work dynamically; results must not depend on who computed what."""
    prob, ev = c2
    rng = np.random.default_rng(2)
    Z = prob.trajectory.datavec + 0.01 * rng.standard_normal(ev.n_vars)
    mu = rng.random(ev.n_constraints)
    a = _evaluate(ev, Z, 1.0, mu)
    _evaluate(ev, Z + 0.1, 1.0, mu)
    b = _evaluate(ev, Z, 1.0, mu)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(y))


@pytest.mark.parametrize("n,m,N,scale", [(8, 2, 13, 0.35), (8, 1, 9, 0.5), (8, 3, 10, 0.4), (8, 4, 8, 0.4), (16, 2, 19, 0.3), (16, 1, 8, 0.3),
                                         (8, 2, 21, 9.0), (16, 2, 11, 6.0)])
def test_octet_variant_agrees_with_persistent(monkeypatch, n, m, N, scale):
    """Small states run eight intervals per warp (one tile per jet component, bilinear_octet.cu).  Same series, different
    association of the products: agreement with the one-interval-per-warp kernel far below the parity tolerance,
    including partial octets (N - 1 not a multiple of 8), multi-stage series (scale 6, 9) and batches."""
    prob = pt.scaled_problem(N=N, state_dim=n, n_controls=m, generator_scale=scale)
    res = {}
    for pin in ("", "persistent"):
        if pin:
            monkeypatch.setenv("DTO_B200_KERNEL", pin)
        ev = dto.Evaluator(prob, batch=3)
        assert ev.kernel_variant(0) == (pin or "octet")
        Z = np.tile(prob.trajectory.datavec, 3) + 0.01 * np.random.default_rng(4).standard_normal(3 * ev.n_vars)
        mu = np.random.default_rng(5).random(3 * ev.n_constraints)
        bufs = [np.empty(3), np.empty(3 * ev.n_vars), np.empty(3 * ev.n_constraints), np.empty(3 * ev.nnz_jacobian), np.empty(3 * ev.nnz_hessian)]
        ev.eval_all(Z, 1.0, mu, *bufs)
        assert all(np.isfinite(b).all() for b in bufs)
        res[pin] = bufs
        ev.close()
    for x, y in zip(res[""], res["persistent"]):
        x, y = np.asarray(x), np.asarray(y)
        assert np.abs(x - y).max() <= 1e-12 * max(np.abs(y).max(), 1.0)


def test_k1_variants_agree(monkeypatch):
    prob = pt.quantum_gate_problem(N=40, levels=16, n_drives=4)
    rng = np.random.default_rng(3)
    res = {}
    for pin in ("", "dmma", "generic"):
        if pin:
            monkeypatch.setenv("DTO_B200_KERNEL", pin)
        ev = dto.Evaluator(prob)
        assert ev.kernel_variant(0) == (pin or "persistent")
        Z = prob.trajectory.datavec + 0.01 * np.random.default_rng(4).standard_normal(ev.n_vars)
        mu = np.random.default_rng(5).random(ev.n_constraints)
        res[pin] = _evaluate(ev, Z, 1.0, mu)
        ev.close()
    for pin in ("dmma", "generic"):
        for x, y in zip(res[""], res[pin]):
            x, y = np.asarray(x), np.asarray(y)
            assert np.abs(x - y).max() <= 1e-12 * max(np.abs(y).max(), 1.0), pin
