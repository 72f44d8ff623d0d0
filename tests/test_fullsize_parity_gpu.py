"""Oracle checks AT BASELINE.json's full sizes (configs c2, c4, c5), where evaluating the whole problem with the CPU
oracle would take hours: a seeded sample of (problem, knot interval) pairs -- always including the first and last
interval, the octet tails and the tail of the persistent kernel's work queue -- is recomputed with the oracle
(exact Frechet derivatives from block-triangular matrix exponentials) and compared with the corresponding residual
rows, Jacobian blocks and Hessian knot regions of the CUDA evaluator at 1e-10 relative to the block norm
(tests/helpers/blocks.py).  Both boundaries are checked: host pointers (dto_eval_all) and device pointers
(dto_eval_all_dev, outputs read back with torch)."""
import numpy as np
import pytest

import dto_b200 as dto
from dto_b200 import problem_templates as pt
from helpers import blocks

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _sample_knots(N, rng, count):
    """1-based knots whose interval (k) and Hessian region are checked; N itself checks the last knot's region."""
    fixed = {1, 2, 3, 8, 9, N - 9, N - 8, N - 2, N - 1, N}
    fixed |= {int(k) for k in rng.integers(1, N + 1, size=count)}
    return sorted(k for k in fixed if 1 <= k <= N)


def _eval_host(ev, Z, sigma, mu, batch=1):
    J, grad = np.empty(batch), np.empty(batch * ev.n_vars)
    g, jac, hess = np.empty(batch * ev.n_constraints), np.empty(batch * ev.nnz_jacobian), np.empty(batch * ev.nnz_hessian)
    ev.eval_all(Z, sigma, mu, J, grad, g, jac, hess)
    return J, grad, g, jac, hess


def _eval_dev(ev, Z, sigma, mu, batch=1):
    import torch

    dev = torch.device("cuda", torch.cuda.current_device())
    dZ, dmu = torch.from_numpy(np.ascontiguousarray(Z).reshape(-1)).to(dev), torch.from_numpy(np.ascontiguousarray(mu).reshape(-1)).to(dev)
    outs = [torch.full((n,), float("nan"), dtype=torch.float64, device=dev)
            for n in (batch, batch * ev.n_vars, batch * ev.n_constraints, batch * ev.nnz_jacobian, batch * ev.nnz_hessian)]
    torch.cuda.synchronize()
    ev.eval_all_dev(dZ.data_ptr(), sigma, dmu.data_ptr(), *[o.data_ptr() for o in outs])
    ev.synchronize()
    return [o.cpu().numpy() for o in outs]


def _check(spec, Z, sigma, mu, outs, ks, what):
    _, _, g, jac, hess = outs
    assert np.isfinite(g).all() and np.isfinite(jac).all() and np.isfinite(hess).all(), what
    er, ej, eh = blocks.sampled_interval_check(spec, Z, sigma, mu, g, jac, hess, ks)
    assert er <= TOL and ej <= TOL and eh <= TOL, (what, er, ej, eh)


def test_c2_full_size_sampled_intervals():
    """config c2: n = 32, 4 drives, N = 2000 (persistent kernel: work-queue tail = the last intervals)."""
    prob = pt.quantum_gate_problem(N=2000, levels=16, n_drives=4)
    spec = prob.to_spec()
    ev = dto.Evaluator(prob)
    assert ev.kernel_variant(0) == "persistent"
    rng = np.random.default_rng(21)
    Z = prob.trajectory.vec() + 0.02 * rng.standard_normal(ev.n_vars)
    mu = rng.random(ev.n_constraints)
    ks = _sample_knots(spec["N"], rng, 64)
    host = _eval_host(ev, Z, 1.3, mu)
    _check(spec, Z, 1.3, mu, host, ks, "c2 host")
    devo = _eval_dev(ev, Z, 1.3, mu)
    _check(spec, Z, 1.3, mu, devo, ks, "c2 dev")
    for a, b in zip(host, devo):
        assert np.array_equal(a, b)
    ev.close()


def test_c4_full_size_sampled_intervals():
    """config c4: one trajectory of N = 100 000 knots, n = 16 (octet kernel: the last octet is partial)."""
    prob = pt.scaled_problem(N=100000, state_dim=16, n_controls=2, generator_scale=0.25)
    spec = prob.to_spec()
    ev = dto.Evaluator(prob)
    assert ev.kernel_variant(0) == "octet"
    rng = np.random.default_rng(22)
    Z = prob.trajectory.vec() + 0.02 * rng.standard_normal(ev.n_vars)
    mu = rng.random(ev.n_constraints)
    N = spec["N"]
    ks = sorted(set(_sample_knots(N, rng, 64)) | {N - 1 - ((N - 1) % 8), N - ((N - 1) % 8), 50000, 50001})
    devo = _eval_dev(ev, Z, 0.9, mu)
    _check(spec, Z, 0.9, mu, devo, ks, "c4 dev")
    host = _eval_host(ev, Z, 0.9, mu)
    _check(spec, Z, 0.9, mu, host, ks, "c4 host")
    ev.close()


@pytest.mark.parametrize("per_problem_G", [False, True])
def test_c5_full_batch_sampled_intervals(per_problem_G):
    """config c5: 4096 independent 8-state problems of N = 200 in one launch, shared generators and 4096 different
    systems (per-problem generators: the octet kernel reads fragment-ordered copies from global memory)."""
    B = 4096
    prob = pt.scaled_problem(N=200, state_dim=8, n_controls=2, generator_scale=0.35)
    spec = prob.to_spec()
    rng = np.random.default_rng(23)
    it = prob.integrators[0]
    bG = None
    if per_problem_G:
        bG = it.G[None] * (1.0 + 0.2 * rng.standard_normal((B,) + it.G.shape))
    ev = dto.Evaluator(prob, batch=B, batch_G=bG)
    assert ev.kernel_variant(0) == "octet"
    Z = np.tile(prob.trajectory.vec(), (B, 1)) + 0.02 * rng.standard_normal((B, ev.n_vars))
    mu = rng.random((B, ev.n_constraints))
    N = spec["N"]
    problems = sorted({0, 1, B - 1, B - 2} | {int(b) for b in rng.integers(0, B, size=20)})
    for name, outs in (("dev", _eval_dev(ev, Z, 1.1, mu, batch=B)), ("host", _eval_host(ev, Z, 1.1, mu, batch=B))):
        g, jac, hess = (outs[2].reshape(B, -1), outs[3].reshape(B, -1), outs[4].reshape(B, -1))
        assert np.isfinite(jac).all() and np.isfinite(hess).all() and np.isfinite(g).all()
        for b in problems:
            sp = spec
            if per_problem_G:
                sp = dict(spec)
                sp["integrators"] = [dict(spec["integrators"][0], G=bG[b])] + list(spec["integrators"][1:])
            ks = sorted({1, 2, N - 1, N, 192, 193, 199} | {int(k) for k in rng.integers(1, N + 1, size=3)})
            er, ej, eh = blocks.sampled_interval_check(sp, Z[b], 1.1, mu[b], g[b], jac[b], hess[b], ks)
            assert er <= TOL and ej <= TOL and eh <= TOL, (name, b, er, ej, eh)
    ev.close()
