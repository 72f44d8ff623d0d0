"""CPU checks of the two error models the kernels size their work with (no GPU, no library call):

* the series plan of the bilinear kernels (csrc/series_plan.cu): Taylor length T = min(T(theta1), T(d2) + 1) with
  d2 = ||A^2||_1^(1/2), T(x) = smallest T with x^T / T! <= 2^-53 -- the truncated series must reach double precision;
* the extrapolation columns of the time-dependent kernels (csrc/dto_internal.h: tdb_item_steps): the smallest K with
  theta^(2K+1) / (2^K K!)^2 <= tol for the Gragg-Bulirsch-Stoer scheme with n_k = 2, 4, .., 2K.

Both are restated here in NumPy / mpmath exactly as the device code states them."""
import math

import mpmath as mp
import numpy as np
import pytest


def taylor_terms(x):
    T = 1
    while x ** T / math.factorial(T) > 2.0 ** -53 and T < 60:
        T += 1
    return T


def plan_terms(A):
    theta1 = np.abs(A).sum(axis=0).max()
    d2 = math.sqrt(np.abs(A @ A).sum(axis=0).max())
    t1, t2 = taylor_terms(theta1), taylor_terms(d2) + 1
    return (t2 if (d2 < theta1 and t2 < t1 and theta1 <= t2) else t1), t1


def cases():
    rng = np.random.default_rng(0)
    out = []
    for n in (6, 10):
        H = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        H = (H + H.conj().T) / 2
        G = np.block([[H.imag, H.real], [-H.real, H.imag]])  # the real isomorphism of -iH (config c2's generators)
        for scale in (0.3, 1.0, 3.5):
            out.append(("iso", G * scale / np.abs(G).sum(axis=0).max()))
        N = np.triu(rng.standard_normal((2 * n, 2 * n)), 1)  # strongly non-normal
        for scale in (1.0, 3.9):
            out.append(("triu", N * scale / np.abs(N).sum(axis=0).max()))
        D = rng.standard_normal((2 * n, 2 * n))
        out.append(("dense", D * 2.0 / np.abs(D).sum(axis=0).max()))
    return out


@pytest.mark.parametrize("kind,A", cases())
def test_series_plan_reaches_double_precision(kind, A):
    mp.mp.dps = 60
    n = A.shape[0]
    T, T1 = plan_terms(A)
    assert T <= T1
    x = np.random.default_rng(1).standard_normal(n)
    Am, xm = mp.matrix(A.tolist()), mp.matrix(x.tolist())
    exact = mp.expm(Am) * xm
    term, acc = xm, xm
    for k in range(1, T + 1):  # terms 0..T, in 60 digits: only the truncation is measured
        term = (Am * term) / k
        acc = acc + term
    err = max(abs(acc[i] - exact[i]) for i in range(n))
    scale = max(abs(x).max(), max(abs(exact[i]) for i in range(n)))
    # the bound: (theta1 / d2) d2^T / T! <= 2^-53 (1 + ...) in the 1-norm, relative to ||x||_1 <= n max|x|
    assert err <= 2.0 ** -52 * n * scale, (kind, T, T1, float(err))


def gbs(Gfun, y0, K):
    """One macro step over tau in [0, 1] of y' = G(tau) y: modified midpoint with n_k = 2(k+1) substeps, Gragg smoothing,
    polynomial extrapolation in h^2 with the closed-form weights the kernels use."""
    acc = 0.0 * y0
    for k in range(K):
        nk = 2 * (k + 1)
        h = 1.0 / nk
        w = 1.0
        for l in range(K):
            if l != k:
                w *= (4.0 * (k + 1) ** 2) / (4.0 * (k + 1) ** 2 - 4.0 * (l + 1) ** 2)
        zp, zc = y0, y0 + h * (Gfun(0.0) @ y0)
        for q in range(1, nk):
            zp, zc = zc, zp + 2 * h * (Gfun(q * h) @ zc)
        zn = zp + 2 * h * (Gfun(1.0) @ zc)  # one step beyond, for the smoothing
        acc = acc + w * 0.25 * (zp + 2 * zc + zn)
    return acc


def columns(theta, tol=1e-14, kmax=8):
    pw, den = theta ** 7, 2304.0
    for K in range(3, kmax):
        if pw <= tol * den:
            return K
        pw *= theta * theta
        den *= 4.0 * (K + 1) ** 2
    return kmax


@pytest.mark.parametrize("theta", [0.05, 0.2, 0.5, 1.0])
def test_extrapolation_columns_meet_the_error_model(theta):
    """The K the kernels pick for a step of size theta reaches ~1e-14 relative; one column less than the model allows is
    measurably worse (the model is not vacuous)."""
    import scipy.linalg as sla

    rng = np.random.default_rng(3)
    n = 8
    H = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H = (H + H.conj().T) / 2
    G0 = np.block([[H.imag, H.real], [-H.real, H.imag]])
    G1 = rng.standard_normal((2 * n, 2 * n))
    G1 = (G1 - G1.T) / 2
    w = 2.0
    n1 = lambda M: np.abs(M).sum(axis=0).max()
    s = theta / (n1(G0) + n1(G1) + w)  # theta = |dt| (||G0||_1 + ||G1||_1 + omega), as tdb_item_steps
    Gfun = lambda tau: s * (G0 + math.cos(w * s * tau) * G1)
    y0 = rng.standard_normal(2 * n)
    # reference: the same scheme with 2 x 8 macro steps of 10 columns each would be circular; use a tight RK solve instead
    from scipy.integrate import solve_ivp

    ref = solve_ivp(lambda t, y: Gfun(t) @ y, (0.0, 1.0), y0, method="DOP853", rtol=1e-13, atol=1e-15).y[:, -1]
    K = columns(theta)
    err = np.abs(gbs(Gfun, y0, K) - ref).max() / np.abs(ref).max()
    assert err <= 5e-13, (theta, K, err)  # the reference solve itself is good to ~1e-13


@pytest.mark.parametrize("theta", [0.2, 0.5, 1.0])
def test_error_model_against_a_rotation(theta):
    """Where theta IS the growth rate (a plane rotation: ||G||_1 = spectral radius = theta) the measured error of K columns
    follows theta^(2K+1) / (2^K K!)^2 -- below it, and within two orders of magnitude of it until round-off takes over (which is
    also why more columns than the model asks for do not buy accuracy)."""
    G = theta * np.array([[0.0, 1.0], [-1.0, 0.0]])
    y0 = np.array([1.0, 0.3])
    c, s_ = math.cos(theta), math.sin(theta)
    exact = np.array([[c, s_], [-s_, c]]) @ y0
    for K in range(3, 9):
        model = theta ** (2 * K + 1) / (2.0 ** K * math.factorial(K)) ** 2
        err = np.abs(gbs(lambda tau: G, y0, K) - exact).max()
        assert err <= 2.0 * model + 3e-14, (K, err, model)  # the alternating extrapolation weights amplify round-off to ~1e-14 at K = 8
        if model > 1e-13:
            assert err >= model / 100.0, (K, err, model)


def test_bench_counts_the_executed_right_hand_sides(monkeypatch):
    """bench.py's host restatement of tdb_item_steps: config c3 (theta ~ 0.2 per interval) runs 5 extrapolation columns = 35
    right-hand sides per interval; DTO_B200_TDB_TOL=0 restores eight columns = 80 (count C3 of BASELINE.md section 3)."""
    import os, sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    prob = bench.build_problem("c3", 42)
    monkeypatch.delenv("DTO_B200_TDB_TOL", raising=False)
    assert bench.tdb_rhs_per_interval(prob, prob.trajectory.datavec) == 35.0
    monkeypatch.setenv("DTO_B200_TDB_TOL", "0")
    assert bench.tdb_rhs_per_interval(prob, prob.trajectory.datavec) == 80.0
    assert bench.canonical_flops_per_tdb_interval(64, 2, 1, rhs=80) == pytest.approx(100.3e6, rel=2e-3)
