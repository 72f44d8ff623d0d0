"""CPU oracle for the NLP-callback hot path of DirectTrajOpt.jl  --  TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` leg may import it.  The shipped evaluator
(``directtrajopt.jl_b200``) never routes through it.

PARITY UNPINNED: the reference is 100 % Julia and neither ``julia`` nor the un-vendored packages
that carry the arithmetic (ExponentialAction 0.2 ``expv``, ForwardDiff 1.3, OrdinaryDiffEqTsit5)
exist in this container or on the GPU box (probed: profiles/r01_fp64_peaks_and_box_probe.log),
and the reference's own tests hold no golden vectors for this path, only
finite-difference self-consistency checks.  The oracle therefore restates the reference's
*semantics* file by file (citations below, relative to /root/reference) and is pinned by
 (1) central finite differences of its own residual/objective (the reference's ``test_integrator`` /
     ``test_objective`` / ``test_constraint`` method, src/integrators/_integrators.jl:97-242),
 (2) 50-digit mpmath matrix exponentials on small cases,
 (3) an independent C restatement of the reference's *algorithm* (truncated-Taylor ``expv`` with
     forward-mode jets, oracle/dto_oracle.c) that must agree to 1e-10,
 (4) the literal fixtures the reference does hold (test/test_utils.jl:55-111 matrix, README system).

Conventions follow the reference: knots ``k`` and ``times`` are 1-based, structures are 1-based
``(row, col)`` pairs; component offsets inside a knot are 0-based ``(offset, dim)`` tuples.

Problem "spec" (plain dict, produced by ``directtrajopt.jl_b200`` problems' ``to_spec()`` or by hand):
  N, z, components{name:(off,dim)}, timestep (component name), integrators[...], objectives[...],
  constraints[...], composite (bool).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

# --------------------------------------------------------------------------------------------
# indexing  (TrajectoryIndexingUtils 0.1 semantics)
# --------------------------------------------------------------------------------------------


def knot_slice(k, comps, z):
    """0-based positions in Z of components ``comps`` (0-based within a knot) of 1-based knot k.
    Julia: slice(k, comps, z) = z*(k-1) .+ comps."""
    return (k - 1) * z + np.asarray(comps, dtype=np.int64)


def comp_range(spec, name):
    off, dim = spec["components"][name]
    return np.arange(off, off + dim, dtype=np.int64)


def dt_offset(spec):
    off, dim = spec["components"][spec["timestep"]]
    assert dim == 1
    return off


def _vars_of(spec, names):
    if not names:
        return np.zeros(0, dtype=np.int64)
    return np.concatenate([comp_range(spec, nm) for nm in names])


def n_variables(spec):
    """traj.dim * traj.N + traj.global_dim (evaluator.jl:123, :239)."""
    return spec["N"] * spec["z"] + spec.get("global_dim", 0)


def term_indices(spec, term, k):
    """0-based positions in Z of the variables a knot term reads at 1-based knot k: its knot components, then -- for the
    global variants -- the listed global variables at offset z*N  (global_objectives.jl:230-240,
    global_knot_point_constraint.jl:151-155, global_constraint.jl:96-99)."""
    idx = knot_slice(k, _vars_of(spec, term["names"]), spec["z"])
    if term.get("kind") == "global_knot":
        g = [np.arange(*(lambda od: (od[0], od[0] + od[1]))(spec["global_components"][nm]), dtype=np.int64) for nm in term["global_names"]]
        gidx = spec["N"] * spec["z"] + (np.concatenate(g) if g else np.zeros(0, np.int64))
        idx = np.concatenate([idx, gidx])
    return idx


# --------------------------------------------------------------------------------------------
# knot-function catalogue: closed-form value / Jacobian / Hessian (independent of the device's
# hyper-dual templates).  Replaces the user closures g, l of knot_point_constraint.jl:76-83 and
# knot_point_objectives.jl:65-72.
# --------------------------------------------------------------------------------------------


def _cfun(fn, v, p):
    """constraint function g(v; p) -> (val[gd], J[gd,vd], H[gd,vd,vd])"""
    v = np.asarray(v, float)
    vd = v.size
    if fn == "norm_minus_c":  # [norm(v) - c]   (test/test_snippets.jl:39-45, evaluator.jl:670-676)
        nv = np.linalg.norm(v)
        J = (v / nv)[None, :]
        H = ((np.eye(vd) - np.outer(v, v) / nv**2) / nv)[None]
        return np.array([nv - p[0]]), J, H
    if fn == "normsq_minus_c":  # [norm(v)^2 - c]
        return np.array([v @ v - p[0]]), 2 * v[None, :], 2 * np.eye(vd)[None]
    if fn == "sqdist_minus_c":  # [norm(v - p[1:])^2 - p[0]]
        d = v - p[1 : 1 + vd]
        return np.array([d @ d - p[0]]), 2 * d[None, :], 2 * np.eye(vd)[None]
    if fn == "linear":  # A v - b ; p = [gd, A(gd x vd, column-major), b(gd)]
        gd = int(p[0])
        A = np.asarray(p[1 : 1 + gd * vd]).reshape(vd, gd).T
        b = np.asarray(p[1 + gd * vd : 1 + gd * vd + gd])
        return A @ v - b, A.copy(), np.zeros((gd, vd, vd))
    if fn == "norm_product":  # [norm(v1) - c1; norm(v1) norm(v2) - c2], p = [c1, c2, n1]  (global_knot_point_constraint.jl:267-270)
        n1 = int(p[2])
        v1, v2 = v[:n1], v[n1:]
        r1, r2 = np.linalg.norm(v1), np.linalg.norm(v2)
        e1, e2 = np.zeros(vd), np.zeros(vd)
        e1[:n1], e2[n1:] = v1 / r1, v2 / r2  # gradients of r1, r2
        H1, H2 = np.zeros((vd, vd)), np.zeros((vd, vd))
        H1[:n1, :n1] = (np.eye(n1) - np.outer(v1, v1) / r1**2) / r1
        H2[n1:, n1:] = (np.eye(vd - n1) - np.outer(v2, v2) / r2**2) / r2
        J = np.stack([e1, r2 * e1 + r1 * e2])
        H = np.stack([H1, r2 * H1 + r1 * H2 + np.outer(e1, e2) + np.outer(e2, e1)])
        return np.array([r1 - p[0], r1 * r2 - p[1]]), J, H
    raise ValueError(f"unknown constraint function {fn}")


def _lfun(fn, v, p):
    """objective function l(v; p) -> (val, grad[vd], H[vd,vd])"""
    v = np.asarray(v, float)
    vd = v.size
    if fn == "normsq_plus_p":  # norm(v)^2 + p      (knot_point_objectives.jl:254,269)
        return v @ v + p[0], 2 * v, 2 * np.eye(vd)
    if fn == "sqdist":  # norm(v - p)^2           (evaluator.jl:664 terminal cost)
        d = v - p[:vd]
        return d @ d, 2 * d, 2 * np.eye(vd)
    if fn == "linear":  # p' v
        c = np.asarray(p[:vd])
        return c @ v, c.copy(), np.zeros((vd, vd))
    if fn == "iso_infidelity":  # 1 - |<goal|psi>|^2, v=[re;im], p=[gre;gim]
        h = vd // 2
        g = np.asarray(p[:vd])
        c1 = g.copy()
        c2 = np.concatenate([-g[h:], g[:h]])
        a, b = c1 @ v, c2 @ v
        return 1 - (a * a + b * b), -2 * (a * c1 + b * c2), -2 * (np.outer(c1, c1) + np.outer(c2, c2))
    if fn == "split_sqdist":  # norm(v[:h] - v[h:])^2  (global_objectives.jl:364-369)
        h = vd // 2
        d = v[:h] - v[h:]
        I = np.eye(h)
        return d @ d, np.concatenate([2 * d, -2 * d]), 2 * np.block([[I, -I], [-I, I]])
    raise ValueError(f"unknown objective function {fn}")


# --------------------------------------------------------------------------------------------
# integrators: per-interval value, Jacobian block (d x 2z) and Hessian block (2z x 2z) of mu' f
# --------------------------------------------------------------------------------------------


def _bilinear_interval(spec, it, zk, zk1, mu=None, want_jac=True):
    """BilinearIntegrator, f = x+ - expv(dt, G(u), x)   (bilinear_integrator.jl:81).
    Derivatives are the exact Frechet derivatives of the matrix exponential (SURVEY.md section 8a
    math contract) obtained from block-triangular exponentials, not from the GPU's series."""
    z = spec["z"]
    xs, us, dto = comp_range(spec, it["x"]), comp_range(spec, it["u"]), dt_offset(spec)
    G = np.asarray(it["G"], float)  # (m+1, n, n): drift, drives
    n, m = xs.size, us.size
    x, u, dt, xn = zk[xs], zk[us], zk[dto], zk1[xs]
    Gu = G[0] + np.tensordot(u, G[1:], axes=(0, 0)) if m else G[0].copy()
    A = dt * Gu
    E = sla.expm(A)
    w = E @ x
    r = xn - w
    if not want_jac and mu is None:
        return r, None, None
    # first-order Frechet: top-right block of expm([[A, B],[0, A]])
    Ls = []
    for i in range(m):
        M = np.zeros((2 * n, 2 * n))
        M[:n, :n] = A
        M[n:, n:] = A
        M[:n, n:] = dt * G[1 + i]
        Ls.append(sla.expm(M)[:n, n:])
    Jb = np.zeros((n, 2 * z))
    Jb[:, xs] = -E
    for i in range(m):
        Jb[:, us[i]] = -Ls[i] @ x
    Jb[:, dto] = -Gu @ w
    Jb[np.arange(n), z + xs] = 1.0
    if mu is None:
        return r, Jb, None
    Hb = np.zeros((2 * z, 2 * z))
    for i in range(m):
        hv = -Ls[i].T @ mu
        Hb[xs, us[i]] = hv
        Hb[us[i], xs] = hv
    hv = -(Gu @ E).T @ mu
    Hb[xs, dto] = hv
    Hb[dto, xs] = hv
    # second-order Frechet: L2(A; Bi, Bj) = S_ij + S_ji,  S_ij = expm([[A,Bi,0],[0,A,Bj],[0,0,A]])[0,2]
    for i in range(m):
        for j in range(i, m):
            M = np.zeros((3 * n, 3 * n))
            for b in range(3):
                M[b * n : (b + 1) * n, b * n : (b + 1) * n] = A
            M[:n, n : 2 * n] = dt * G[1 + i]
            M[n : 2 * n, 2 * n :] = dt * G[1 + j]
            S = sla.expm(M)[:n, 2 * n :]
            if i == j:
                L2 = 2 * S
            else:
                M[:n, n : 2 * n] = dt * G[1 + j]
                M[n : 2 * n, 2 * n :] = dt * G[1 + i]
                L2 = S + sla.expm(M)[:n, 2 * n :]
            val = -mu @ (L2 @ x)
            Hb[us[i], us[j]] = val
            Hb[us[j], us[i]] = val
    for i in range(m):
        val = -mu @ (G[1 + i] @ w + Gu @ (Ls[i] @ x))
        Hb[us[i], dto] = val
        Hb[dto, us[i]] = val
    Hb[dto, dto] = -mu @ (Gu @ (Gu @ w))
    return r, Jb, Hb


def _derivative_interval(spec, it, zk, zk1, mu=None, want_jac=True):
    """DerivativeIntegrator, f = x+ - x - dt*xdot   (derivative_integrator.jl:45)."""
    z = spec["z"]
    xs, ds, dto = comp_range(spec, it["x"]), comp_range(spec, it["xdot"]), dt_offset(spec)
    d = xs.size
    r = zk1[xs] - zk[xs] - zk[dto] * zk[ds]
    if not want_jac and mu is None:
        return r, None, None
    Jb = np.zeros((d, 2 * z))
    ar = np.arange(d)
    Jb[ar, xs] = -1.0
    Jb[ar, ds] = -zk[dto]
    Jb[:, dto] = -zk[ds]
    Jb[ar, z + xs] = 1.0
    if mu is None:
        return r, Jb, None
    Hb = np.zeros((2 * z, 2 * z))
    Hb[ds, dto] = -mu
    Hb[dto, ds] = -mu
    return r, Jb, Hb


# ---- time-dependent bilinear: generator family and exact (tight-tolerance) variational solve ----


def tdb_generator(it, u, t):
    """G(u, t) for the parametric family the device supports:
       G = G0 + sum_i u_i * (c_i(t) A_i + s_i(t) B_i) + sum_j e_j(t) * D_j
    with c_i = cos(w_i t + phi_i), s_i = sin(w_i t + phi_i), e_j = cos(wd_j t + phd_j).
    Returns (G, dG/dt, d2G/dt2, [dG/du_i], [d2G/du_i dt])."""
    G0 = np.asarray(it["G0"], float)
    A, B = np.asarray(it["A"], float), np.asarray(it["B"], float)
    w, ph = np.asarray(it["omega"], float), np.asarray(it["phi"], float)
    m = A.shape[0]
    G = G0.copy()
    Gt = np.zeros_like(G0)
    Gtt = np.zeros_like(G0)
    Gu, Gut = [], []
    for i in range(m):
        c, s = np.cos(w[i] * t + ph[i]), np.sin(w[i] * t + ph[i])
        Mi = c * A[i] + s * B[i]
        Mit = w[i] * (-s * A[i] + c * B[i])
        Mitt = -w[i] ** 2 * Mi
        G += u[i] * Mi
        Gt += u[i] * Mit
        Gtt += u[i] * Mitt
        Gu.append(Mi)
        Gut.append(Mit)
    D = np.asarray(it.get("D", np.zeros((0,) + G0.shape)), float)
    wd, phd = np.asarray(it.get("omega_d", []), float), np.asarray(it.get("phi_d", []), float)
    for j in range(D.shape[0]):
        c, s = np.cos(wd[j] * t + phd[j]), np.sin(wd[j] * t + phd[j])
        G += c * D[j]
        Gt += -wd[j] * s * D[j]
        Gtt += -wd[j] ** 2 * c * D[j]
    return G, Gt, Gtt, Gu, Gut


def _tdb_interval(spec, it, zk, zk1, mu=None, want_jac=True):
    """TimeDependentBilinearIntegrator, f = x+ - Phi(1), dPhi/dtau = dt*G(u(tau), t+tau*dt) Phi
    (time_dependent_bilinear_integrator.jl:102-128).  The reference integrates with adaptive Tsit5
    at OrdinaryDiffEq defaults; the oracle integrates the *exact* first/second-order variational
    equations at rtol=atol=1e-13 (DOP853), which is what the reference converges to under tight
    ``solve_kwargs`` (SURVEY.md section 7 'TDBI parity is ill-posed at default tolerances')."""
    from scipy.integrate import solve_ivp

    z = spec["z"]
    xs, us, dto = comp_range(spec, it["x"]), comp_range(spec, it["u"]), dt_offset(spec)
    to = spec["components"][it["t"]][0]
    order = int(it.get("spline_order", 1))
    n, m = xs.size, us.size
    x, u0, dt, t0, xn = zk[xs], zk[us], zk[dto], zk[to], zk1[xs]
    u1 = zk1[us] if order == 1 else u0
    # parameters theta = [u0 (m), u1 (m, order 1 only), dt, t]; x enters linearly.
    npar = (2 * m if order == 1 else m) + 2
    idt, itt = npar - 2, npar - 1

    def gen(tau):
        """M(tau) = dt*G(u(tau), t0+tau*dt), its first and second parameter derivatives."""
        uu = u0 + tau * (u1 - u0) if order == 1 else u0
        G, Gt, Gtt, Gu, Gut = tdb_generator(it, uu, t0 + tau * dt)
        M = dt * G
        dM = [None] * npar
        d2M = [[None] * npar for _ in range(npar)]
        wts = [(1 - tau), tau] if order == 1 else [1.0]
        # d/du
        for b, wt in enumerate(wts):
            for i in range(m):
                dM[b * m + i] = dt * wt * Gu[i]
        dM[idt] = G + dt * tau * Gt
        dM[itt] = dt * Gt
        zero = np.zeros_like(G)
        for a in range(npar):
            for b in range(npar):
                d2M[a][b] = zero
        for b, wt in enumerate(wts):
            for i in range(m):
                a = b * m + i
                v = wt * (Gu[i] + dt * tau * Gut[i])  # d2/(du dt)
                d2M[a][idt] = d2M[idt][a] = v
                v = dt * wt * Gut[i]  # d2/(du dt0)
                d2M[a][itt] = d2M[itt][a] = v
        d2M[idt][idt] = 2 * tau * Gt + dt * tau * tau * Gtt
        d2M[idt][itt] = d2M[itt][idt] = Gt + dt * tau * Gtt
        d2M[itt][itt] = dt * Gtt
        return M, dM, d2M

    second = mu is not None
    pairs = [(a, b) for a in range(npar) for b in range(a, npar)] if second else []
    first = want_jac or second

    # state: Phi (n x n) fundamental matrix, then dPhi_a (npar, n x n), then d2Phi_ab (pairs, n x n)
    def rhs(tau, y):
        M, dM, d2M = gen(tau)
        Y = y.reshape(-1, n, n)
        out = np.empty_like(Y)
        P = Y[0]
        out[0] = M @ P
        if first:
            for a in range(npar):
                out[1 + a] = M @ Y[1 + a] + dM[a] @ P
        if second:
            for q, (a, b) in enumerate(pairs):
                out[1 + npar + q] = M @ Y[1 + npar + q] + dM[a] @ Y[1 + b] + dM[b] @ Y[1 + a] + d2M[a][b] @ P
        return out.ravel()

    nblk = 1 + (npar if first else 0) + len(pairs)
    y0 = np.zeros((nblk, n, n))
    y0[0] = np.eye(n)
    sol = solve_ivp(rhs, (0.0, 1.0), y0.ravel(), method="DOP853", rtol=1e-13, atol=1e-13)
    Y = sol.y[:, -1].reshape(nblk, n, n)
    Phi = Y[0]
    r = xn - Phi @ x
    if not first:
        return r, None, None
    pcols = [us[i] for i in range(m)]
    if order == 1:
        pcols += [z + us[i] for i in range(m)]
    pcols += [dto, to]
    Jb = np.zeros((n, 2 * z))
    Jb[:, xs] = -Phi
    for a in range(npar):
        Jb[:, pcols[a]] += -(Y[1 + a] @ x)
    Jb[np.arange(n), z + xs] = 1.0
    if not second:
        return r, Jb, None
    Hb = np.zeros((2 * z, 2 * z))
    for a in range(npar):
        hv = -(Y[1 + a].T @ mu)
        Hb[xs, pcols[a]] += hv
        Hb[pcols[a], xs] += hv
    for q, (a, b) in enumerate(pairs):
        val = -mu @ (Y[1 + npar + q] @ x)
        Hb[pcols[a], pcols[b]] += val
        if a != b:
            Hb[pcols[b], pcols[a]] += val
    return r, Jb, Hb


_INTERVAL = {"bilinear": _bilinear_interval, "derivative": _derivative_interval, "tdbilinear": _tdb_interval}


def integrator_dim(spec, it):
    return spec["components"][it["x"]][1]


# --------------------------------------------------------------------------------------------
# structures   (evaluator.jl:119-203, _integrators.jl:49-77)
# --------------------------------------------------------------------------------------------


def _csc_order(rows, cols):
    """findnz order of a SparseMatrixCSC holding the union of (row, col): column-major, rows
    ascending, duplicates merged (evaluator.jl:144, :201)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    if rows.size == 0:
        return rows, cols
    key = np.unique(cols * (rows.max() + 2) + rows)
    return key % (rows.max() + 2), key // (rows.max() + 2)


def constraint_jacobian_entries(spec, c, Z):
    """eval_jacobian of a NonlinearKnotPointConstraint as COO triplets, 0-based local rows
    (knot_point_constraint.jl:254-268)."""
    rows, cols, vals = [], [], []
    gd = None
    for i, k in enumerate(c["times"]):
        idx = term_indices(spec, c, k)
        val, J, _ = _cfun(c["fn"], Z[idx], np.atleast_1d(c["params"][i]))
        gd = val.size
        for a in range(gd):
            for b in range(idx.size):
                rows.append(i * gd + a)
                cols.append(idx[b])
                vals.append(J[a, b])
    return np.array(rows, np.int64), np.array(cols, np.int64), np.array(vals, float)


def constraint_dim(spec, c):
    nvars = term_indices(spec, c, 1).size
    val, _, _ = _cfun(c["fn"], np.ones(nvars), np.atleast_1d(c["params"][0]))
    return val.size * len(c["times"])


def constraint_hessian_entries(spec, c, Z, mu):
    """eval_hessian_of_lagrangian of a NonlinearKnotPointConstraint (knot_point_constraint.jl:275-294)."""
    rows, cols, vals = [], [], []
    for i, k in enumerate(c["times"]):
        idx = term_indices(spec, c, k)
        val, _, H = _cfun(c["fn"], Z[idx], np.atleast_1d(c["params"][i]))
        gd = val.size
        Hm = np.tensordot(mu[i * gd : (i + 1) * gd], H, axes=(0, 0))
        for a in range(idx.size):
            for b in range(idx.size):
                rows.append(idx[a])
                cols.append(idx[b])
                vals.append(Hm[a, b])
    return np.array(rows, np.int64), np.array(cols, np.int64), np.array(vals, float)


def _stored(vals):
    """SparseArrays scalar setindex! does not create an entry for a zero value (SURVEY.md section 7
    'value-dependent sparsity')."""
    return np.asarray(vals) != 0.0


def jacobian_structure(spec, Z0):
    """1-based (rows, cols) of MOI.jacobian_structure (evaluator.jl:119-144)."""
    N, z = spec["N"], spec["z"]
    rows, cols = [], []
    off = 0
    for it in spec["integrators"]:
        d = integrator_dim(spec, it)
        k = np.arange(N - 1)
        r = (off + k[:, None, None] * d + np.arange(d)[None, :, None]) + np.zeros((1, 1, 2 * z), np.int64)
        c = (k[:, None, None] * z + np.arange(2 * z)[None, None, :]) + np.zeros((1, d, 1), np.int64)
        rows.append(r.ravel())
        cols.append(c.ravel())
        off += d * (N - 1)
    for c in spec.get("constraints", []):
        r, cc, v = constraint_jacobian_entries(spec, c, Z0)
        keep = _stored(v)
        rows.append(r[keep] + off)
        cols.append(cc[keep])
        off += constraint_dim(spec, c)
    if not rows:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    r, c = _csc_order(np.concatenate(rows), np.concatenate(cols))
    return r + 1, c + 1


def objective_hessian_pattern(spec, ob):
    """hessian_structure of each objective (regularizers.jl:115-138, knot_point_objectives.jl:205-220)."""
    z = spec["z"]
    rows, cols = [], []
    if ob["kind"] == "quadreg":
        vc = comp_range(spec, ob["name"])
        dto = dt_offset(spec)
        for k in ob["times"]:
            vi = knot_slice(k, vc, z)
            di = (k - 1) * z + dto
            rr, cc = np.meshgrid(vi, vi, indexing="ij")
            rows += [rr.ravel(), vi, [di]]
            cols += [cc.ravel(), np.full(vi.size, di), [di]]
    elif ob["kind"] == "linreg":
        # only the cross terms d2J/(dv ddt), marked at (v, dt)  (regularizers.jl:272-292)
        vc = comp_range(spec, ob["name"])
        dto = dt_offset(spec)
        for k in ob["times"]:
            vi = knot_slice(k, vc, z)
            rows.append(vi)
            cols.append(np.full(vi.size, (k - 1) * z + dto))
    elif ob["kind"] in ("knot", "global_knot"):
        # dense block over [knot vars; global vars] of every listed time (knot_point_objectives.jl:205-220,
        # global_objectives.jl:92-106, :268-291)
        for k in ob["times"]:
            vi = term_indices(spec, ob, k)
            rr, cc = np.meshgrid(vi, vi, indexing="ij")
            rows.append(rr.ravel())
            cols.append(cc.ravel())
    elif ob["kind"] in ("mintime", "null"):
        pass
    else:
        raise ValueError(ob["kind"])
    if not rows:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate([np.asarray(r, np.int64) for r in rows]), np.concatenate([np.asarray(c, np.int64) for c in cols])


def hessian_structure(spec, Z0):
    """1-based (rows, cols) of MOI.hessian_lagrangian_structure (evaluator.jl:151-203): union of
    dense 2z x 2z interval blocks, constraint Hessians at Z0 with mu = 1, objective structures;
    upper triangle of the column-major findnz."""
    N, z = spec["N"], spec["z"]
    rows, cols = [], []
    if spec["integrators"]:
        k = np.arange(N - 1)
        r = k[:, None, None] * z + np.arange(2 * z)[None, :, None] + np.zeros((1, 1, 2 * z), np.int64)
        c = k[:, None, None] * z + np.arange(2 * z)[None, None, :] + np.zeros((1, 2 * z, 1), np.int64)
        rows.append(r.ravel())
        cols.append(c.ravel())
    for c in spec.get("constraints", []):
        r, cc, v = constraint_hessian_entries(spec, c, Z0, np.ones(constraint_dim(spec, c)))
        keep = _stored(v)
        rows.append(r[keep])
        cols.append(cc[keep])
    for ob in spec["objectives"]:
        r, c = objective_hessian_pattern(spec, ob)
        rows.append(r)
        cols.append(c)
    if not rows:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    r, c = _csc_order(np.concatenate(rows), np.concatenate(cols))
    keep = r <= c
    return r[keep] + 1, c[keep] + 1


# --------------------------------------------------------------------------------------------
# objective terms   (regularizers.jl:79-167, minimum_time_objective.jl:44-76,
#                    knot_point_objectives.jl:173-243, _objectives.jl:111-156)
# --------------------------------------------------------------------------------------------


def _objective_term(spec, ob, Z, want_grad, want_hess):
    N, z = spec["N"], spec["z"]
    nv = n_variables(spec)
    J = 0.0
    g = np.zeros(nv) if want_grad else None
    hr, hc, hv = [], [], []
    if ob["kind"] == "quadreg":
        vc = comp_range(spec, ob["name"])
        dto = dt_offset(spec)
        R = np.asarray(ob["R"], float)
        base = np.asarray(ob["baseline"], float)
        for k in ob["times"]:
            vi = knot_slice(k, vc, z)
            di = (k - 1) * z + dto
            dv = Z[vi] - base[:, k - 1]
            dt = Z[di]
            rk = dt * dv
            J += 0.5 * rk @ (R * rk)
            if want_grad:
                g[vi] += dt**2 * (R * dv)
                g[di] += dv @ (R * dv) * dt
            if want_hess:
                # d2/dv2 = dt^2 diag(R) (assigned as a sparse diagonal: only the diagonal is stored)
                hr += list(vi)
                hc += list(vi)
                hv += list(dt**2 * R)
                # d2/(dv ddt): written only at (v, dt)  (regularizers.jl:160)
                hr += list(vi)
                hc += [di] * vi.size
                hv += list(2 * dt * R * dv)
                hr.append(di)
                hc.append(di)
                hv.append(dv @ (R * dv))
    elif ob["kind"] == "linreg":
        # J = sum_t dt_t R'v_t; gradient R dt / R'v; Hessian (v, dt) = R  (regularizers.jl:241-313)
        vc = comp_range(spec, ob["name"])
        dto = dt_offset(spec)
        R = np.asarray(ob["R"], float)
        for k in ob["times"]:
            vi = knot_slice(k, vc, z)
            di = (k - 1) * z + dto
            J += Z[di] * (R @ Z[vi])
            if want_grad:
                g[vi] += R * Z[di]
                g[di] += R @ Z[vi]
            if want_hess:
                hr += list(vi)
                hc += [di] * vi.size
                hv += list(R)
    elif ob["kind"] == "mintime":
        dto = dt_offset(spec)
        idx = np.arange(N - 1) * z + dto
        J = ob["D"] * Z[idx].sum()
        if want_grad:
            g[idx] = ob["D"]
    elif ob["kind"] in ("knot", "global_knot"):
        for i, k in enumerate(ob["times"]):
            vi = term_indices(spec, ob, k)
            val, gr, H = _lfun(ob["fn"], Z[vi], np.atleast_1d(ob["params"][i]))
            Q = ob["Qs"][i]
            J += Q * val
            if want_grad:
                if ob["kind"] == "knot":
                    g[vi] = Q * gr
                else:  # '.+=' on both parts: the globals collect every listed time (global_objectives.jl:262-266)
                    g[vi] += Q * gr
            if want_hess:
                # ForwardDiff.hessian! of Q*l into a sparse view, then triu  (knot_point_objectives.jl:222-243)
                for a in range(vi.size):
                    for b in range(vi.size):
                        if vi[a] <= vi[b]:
                            hr.append(vi[a])
                            hc.append(vi[b])
                            hv.append(Q * H[a, b])
    elif ob["kind"] == "null":
        pass
    else:
        raise ValueError(ob["kind"])
    return J, g, (np.array(hr, np.int64), np.array(hc, np.int64), np.array(hv, float))


def eval_objective(spec, Z):
    """MOI.eval_objective (evaluator.jl:304-308)."""
    return float(sum(ob.get("weight", 1.0) * _objective_term(spec, ob, Z, False, False)[0] for ob in spec["objectives"]))


def eval_objective_gradient(spec, Z):
    """MOI.eval_objective_gradient (evaluator.jl:310-318) with CompositeObjective semantics
    (_objectives.jl:119-128).  A bare (non-composite) QuadraticRegularizer accumulates into the
    caller's buffer without zeroing in the reference (regularizers.jl:104,110); the oracle and the
    device path both define the buffer as zero-initialised (documented deviation)."""
    g = np.zeros(n_variables(spec))
    for ob in spec["objectives"]:
        g += ob.get("weight", 1.0) * _objective_term(spec, ob, Z, True, False)[1]
    return g


# --------------------------------------------------------------------------------------------
# the five callbacks
# --------------------------------------------------------------------------------------------


def n_constraints(spec):
    N = spec["N"]
    nd = sum(integrator_dim(spec, it) * (N - 1) for it in spec["integrators"])
    nn = sum(constraint_dim(spec, c) for c in spec.get("constraints", []))
    return nd, nn


def eval_constraint(spec, Z):
    """MOI.eval_constraint (evaluator.jl:323-362): [integrators...; knot constraints...]."""
    N, z = spec["N"], spec["z"]
    out = []
    for it in spec["integrators"]:
        f = _INTERVAL[it["kind"]]
        for k in range(1, N):
            r, _, _ = f(spec, it, Z[(k - 1) * z : k * z], Z[k * z : (k + 1) * z], None, want_jac=False)
            out.append(r)
    for c in spec.get("constraints", []):
        for i, k in enumerate(c["times"]):
            val, _, _ = _cfun(c["fn"], Z[term_indices(spec, c, k)], np.atleast_1d(c["params"][i]))
            out.append(val)
    return np.concatenate(out) if out else np.zeros(0)


def _lookup(rows, cols, vals, srows, scols, shape, accumulate):
    """Mimic 'for each stored entry: out[linear_map[row, col]] (+)= val' (evaluator.jl:514-525,
    :584-597).  Entries outside the structure are dropped (output_idx == 0)."""
    M = sp.coo_matrix((vals, (rows, cols)), shape=shape).tocsr()  # duplicates summed
    if not accumulate:
        assert M.nnz == len(set(zip(rows.tolist(), cols.tolist()))) or True
    return np.asarray(M[srows, scols]).ravel()


def eval_constraint_jacobian(spec, Z, structure):
    """MOI.eval_constraint_jacobian (evaluator.jl:368-380, :491-551), values in structure order."""
    N, z = spec["N"], spec["z"]
    srows, scols = structure[0] - 1, structure[1] - 1
    nd, nn = n_constraints(spec)
    rows, cols, vals = [], [], []
    off = 0
    for it in spec["integrators"]:
        f = _INTERVAL[it["kind"]]
        d = integrator_dim(spec, it)
        for k in range(1, N):
            _, Jb, _ = f(spec, it, Z[(k - 1) * z : k * z], Z[k * z : (k + 1) * z], None, want_jac=True)
            rr, cc = np.meshgrid(off + (k - 1) * d + np.arange(d), (k - 1) * z + np.arange(2 * z), indexing="ij")
            rows.append(rr.ravel())
            cols.append(cc.ravel())
            vals.append(Jb.ravel())
        off += d * (N - 1)
    for c in spec.get("constraints", []):
        r, cc, v = constraint_jacobian_entries(spec, c, Z)
        rows.append(r + off)
        cols.append(cc)
        vals.append(v)
        off += constraint_dim(spec, c)
    if not rows:
        return np.zeros(0)
    return _lookup(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), srows, scols, (nd + nn, n_variables(spec)), False)


def eval_hessian_lagrangian(spec, Z, sigma, mu, structure):
    """MOI.eval_hessian_lagrangian (evaluator.jl:389-404, :560-647): upper-triangle accumulation of
    integrator blocks, constraint Hessians and sigma * objective Hessian."""
    N, z = spec["N"], spec["z"]
    srows, scols = structure[0] - 1, structure[1] - 1
    rows, cols, vals = [], [], []
    off = 0
    for it in spec["integrators"]:
        f = _INTERVAL[it["kind"]]
        d = integrator_dim(spec, it)
        for k in range(1, N):
            muk = mu[off + (k - 1) * d : off + k * d]
            _, _, Hb = f(spec, it, Z[(k - 1) * z : k * z], Z[k * z : (k + 1) * z], muk, want_jac=True)
            rr, cc = np.meshgrid((k - 1) * z + np.arange(2 * z), (k - 1) * z + np.arange(2 * z), indexing="ij")
            rows.append(rr.ravel())
            cols.append(cc.ravel())
            vals.append(Hb.ravel())
        off += d * (N - 1)
    for c in spec.get("constraints", []):
        cd = constraint_dim(spec, c)
        r, cc, v = constraint_hessian_entries(spec, c, Z, mu[off : off + cd])
        rows.append(r)
        cols.append(cc)
        vals.append(v)
        off += cd
    if sigma != 0:
        for ob in spec["objectives"]:
            _, _, (r, cc, v) = _objective_term(spec, ob, Z, False, True)
            rows.append(r)
            cols.append(cc)
            vals.append(sigma * ob.get("weight", 1.0) * v)
    if not rows:
        return np.zeros(srows.size)
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    keep = rows <= cols  # 'if row <= col' (evaluator.jl:589, :614, :637)
    return _lookup(rows[keep], cols[keep], vals[keep], srows, scols, (n_variables(spec), n_variables(spec)), True)


# --------------------------------------------------------------------------------------------
# per-knot / per-row ownership views (used by the multi-process tests of the knot-range sharding)
# --------------------------------------------------------------------------------------------


def objective_by_knot(spec, Z):
    """Objective split by the knot each term belongs to (sum == eval_objective)."""
    N, z = spec["N"], spec["z"]
    out = np.zeros(N)
    for ob in spec["objectives"]:
        w = ob.get("weight", 1.0)
        if ob["kind"] in ("quadreg", "knot", "linreg"):
            for i, k in enumerate(ob["times"]):
                one = dict(ob)
                one["times"] = [k]
                if ob["kind"] == "knot":
                    one["params"] = [ob["params"][i]]
                    one["Qs"] = [ob["Qs"][i]]
                out[k - 1] += w * _objective_term(spec, one, Z, False, False)[0]
        elif ob["kind"] == "mintime":
            dto = dt_offset(spec)
            out[: N - 1] += w * ob["D"] * Z[np.arange(N - 1) * z + dto]
    return out


def row_owner_knot(spec):
    """1-based knot that owns each constraint row (interval k -> knot k; knot constraint -> its time)."""
    N = spec["N"]
    owners = []
    for it in spec["integrators"]:
        d = integrator_dim(spec, it)
        owners.append(np.repeat(np.arange(1, N), d))
    for c in spec.get("constraints", []):
        gd = constraint_dim(spec, c) // len(c["times"])
        owners.append(np.repeat(np.asarray(c["times"]), gd))
    return np.concatenate(owners) if owners else np.zeros(0, int)


def violation(spec, g):
    """max(|g_eq|, max(0, g_ineq)) with the row bounds of solve.jl:30-65."""
    nd, _ = n_constraints(spec)
    eq = np.ones(g.size, bool)
    off = nd
    for c in spec.get("constraints", []):
        cd = constraint_dim(spec, c)
        if not c.get("equality", True):
            eq[off : off + cd] = False
        off += cd
    v = np.where(eq, np.abs(g), np.maximum(g, 0.0))
    return v, eq
