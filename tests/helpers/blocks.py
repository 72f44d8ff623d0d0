"""Block-wise parity helpers (test infrastructure).

BASELINE.md section 2 states the value tolerance as 1e-10 relative *to the block norm*: an identity entry of the
Jacobian must not set the scale for a d r / d u column of magnitude 1e-2.  The blocks are

  Jacobian   one (integrator, interval) row block x the columns of one trajectory component of one knot
             (d r_k/d x_k = -E, d r_k/d u_k, d r_k/d dt_k, d r_k/d x_{k+1} = I, ...); a knot-constraint row x the
             columns of one component.  A block whose reference values are all zero (structural zeros that the
             reference still stores, _integrators.jl:57) is held to the scale of its whole row block.
  Hessian    one knot region: the diagonal block of knot k, and its cross block with knot k-1 (evaluator.jl:151-203);
             all-zero regions are held to the scale of the whole array.

`sampled_interval_check` is the full-size variant (configs whose structures have 10^7..10^8 entries): it recomputes
single knot intervals with the oracle and finds their entries in the value arrays through the closed-form layout of
SURVEY.md section 8a (problems without knot constraints), independently of the library's own structure arrays.
"""
import numpy as np

import dto_oracle as orc


def _group_max(groups, vals):
    order = np.argsort(groups, kind="stable")
    g = groups[order]
    starts = np.flatnonzero(np.r_[True, g[1:] != g[:-1]])
    return g[starts], np.maximum.reduceat(vals[order], starts)


def _component_of_column(spec):
    """component id of every column of one knot"""
    z = spec["z"]
    comp = np.full(z, -1, np.int64)
    for ci, (off, dim) in enumerate(spec["components"].values()):
        comp[off:off + dim] = ci
    assert (comp >= 0).all()
    return comp, len(spec["components"])


def _row_blocks(spec):
    """row block id of every constraint row: (integrator, interval), then one per knot-constraint row"""
    N = spec["N"]
    out, nxt = [], 0
    for it in spec["integrators"]:
        d = orc.integrator_dim(spec, it)
        out.append(nxt + np.repeat(np.arange(N - 1), d))
        nxt += N - 1
    for c in spec.get("constraints", []):
        cd = orc.constraint_dim(spec, c)
        out.append(nxt + np.arange(cd))
        nxt += cd
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def jac_block_relerr(spec, structure, got, ref):
    """max over Jacobian blocks of max|got - ref| / max|ref| (block definition in the module docstring)."""
    if ref.size == 0:
        return 0.0
    rows, cols = structure[0] - 1, structure[1] - 1
    z, N = spec["z"], spec["N"]
    comp, nc = _component_of_column(spec)
    knot = np.minimum(cols // z, N)  # global columns: one group each beyond the knots
    colgrp = np.where(cols < N * z, knot * nc + comp[np.minimum(cols % z, z - 1)], N * nc + (cols - N * z))
    rb = _row_blocks(spec)[rows]
    ncg = int(colgrp.max()) + 1
    err, mag = np.abs(got - ref), np.abs(ref)
    gid, gerr = _group_max(rb * ncg + colgrp, err)
    _, gmag = _group_max(rb * ncg + colgrp, mag)
    pid, pmag = _group_max(rb, mag)
    parent = pmag[np.searchsorted(pid, gid // ncg)]
    scale = np.where(gmag > 0, gmag, np.where(parent > 0, parent, 1.0))
    return float((gerr / scale).max())


def hess_block_relerr(spec, structure, got, ref):
    """max over Hessian knot regions (diagonal block of a knot, cross block with the previous knot, global columns)."""
    if ref.size == 0:
        return 0.0
    rows, cols = structure[0] - 1, structure[1] - 1
    z, N = spec["z"], spec["N"]
    ck, rk = np.minimum(cols // z, N), np.minimum(rows // z, N)
    grp = ck * 2 + (rk != ck)
    err, mag = np.abs(got - ref), np.abs(ref)
    _, gerr = _group_max(grp, err)
    _, gmag = _group_max(grp, mag)
    whole = mag.max()
    scale = np.where(gmag > 0, gmag, whole if whole > 0 else 1.0)
    return float((gerr / scale).max())


def vec_block_relerr(spec, got, ref):
    """constraint residual: per row block"""
    if ref.size == 0:
        return 0.0
    rb = _row_blocks(spec)
    _, gerr = _group_max(rb, np.abs(got - ref))
    _, gmag = _group_max(rb, np.abs(ref))
    whole = np.abs(ref).max()
    return float((gerr / np.where(gmag > 0, gmag, whole if whole > 0 else 1.0)).max())


# ---- closed-form layout (no knot constraints, no globals) --------------------------------------------------------
def jac_block_positions(spec, k):
    """0-based positions in the Jacobian value array of the d_i x 2z blocks of interval k (1-based), one array per
    integrator, shape (d_i, 2z): column (k', l) starts at z*Dsum*(2(k'-1)-1) [k' >= 2] + l*Dsum*([k'>=2] + [k'<=N-1])
    and holds [I1 prev | I1 own | I2 prev | I2 own | ...]."""
    z, N = spec["z"], spec["N"]
    ds = [orc.integrator_dim(spec, it) for it in spec["integrators"]]
    D = sum(ds)
    out = []
    doff = 0
    for d in ds:
        pos = np.empty((d, 2 * z), np.int64)
        for half, kk in ((0, k), (1, k + 1)):  # columns of knot k (own rows), knot k+1 (previous-interval rows)
            both = kk >= 2 and kk <= N - 1
            per = D * ((1 if kk >= 2 else 0) + (1 if kk <= N - 1 else 0))
            start = 0 if kk == 1 else z * D * (2 * (kk - 1) - 1)
            if half == 0:
                inner = 2 * doff + d if both else doff  # own rows of knot kk's column
            else:
                inner = 2 * doff if both else doff      # previous-interval rows
            for l in range(z):
                pos[:, half * z + l] = start + l * per + inner + np.arange(d)
        out.append(pos)
        doff += d
    return out


def hess_region_positions(spec, k):
    """0-based positions of knot k's region: (cross[z, z] or None, diag[z, z] with -1 below the diagonal)."""
    z = spec["z"]
    tri = z * (z + 1) // 2
    base = 0 if k == 1 else tri + (k - 2) * (z * z + tri)
    cross = None
    diag = np.full((z, z), -1, np.int64)
    nc = z if k >= 2 else 0
    if k >= 2:
        cross = np.empty((z, z), np.int64)
    for l in range(z):
        cs = base + l * nc + l * (l + 1) // 2
        if k >= 2:
            cross[:, l] = cs + np.arange(z)
        diag[: l + 1, l] = cs + nc + np.arange(l + 1)
    return cross, diag


def sampled_interval_check(spec, Z, sigma, mu, g, jac, hess, ks, jac_structure=None):
    """Recompute intervals `ks` (1-based) of one problem with the oracle and compare residual rows, Jacobian blocks
    (per component column block) and the Hessian regions of knot k (diagonal + cross block).  Returns the largest
    block-relative errors (residual, Jacobian, Hessian).  Problems with knot constraints have no closed-form Jacobian
    layout: pass the 1-based `jac_structure` (rows, cols) and the interval blocks are looked up in it; the constraints'
    own Hessian terms are added to the expected regions."""
    assert not spec.get("global_dim", 0)
    assert jac_structure is not None or not spec.get("constraints")
    z, N = spec["z"], spec["N"]
    lookup = None
    if jac_structure is not None:
        import scipy.sparse as sp

        nnz = jac_structure[0].size
        lookup = sp.csr_matrix((np.arange(1, nnz + 1, dtype=np.int64), (jac_structure[0] - 1, jac_structure[1] - 1)))
    con_h = None
    if spec.get("constraints"):
        nd, _ = orc.n_constraints(spec)
        off, parts = nd, []
        for c in spec["constraints"]:
            cd = orc.constraint_dim(spec, c)
            parts.append(orc.constraint_hessian_entries(spec, c, Z, mu[off:off + cd]))
            off += cd
        con_h = [np.concatenate([p[i] for p in parts]) for i in range(3)]
    comp, nc = _component_of_column(spec)
    ds = [orc.integrator_dim(spec, it) for it in spec["integrators"]]
    offs = np.concatenate([[0], np.cumsum([d * (N - 1) for d in ds])])
    e_r = e_j = e_h = 0.0
    iu = np.triu_indices(z)

    def interval(i, k):
        it = spec["integrators"][i]
        zk, zk1 = Z[(k - 1) * z:k * z], Z[k * z:(k + 1) * z]
        muk = mu[offs[i] + (k - 1) * ds[i]: offs[i] + k * ds[i]]
        return orc._INTERVAL[it["kind"]](spec, it, zk, zk1, muk, want_jac=True)

    for k in ks:
        Hdiag = np.zeros((z, z))
        Hcross = np.zeros((z, z))
        pos = None
        if k <= N - 1 and lookup is None:
            pos = jac_block_positions(spec, k)
        elif k <= N - 1:
            pos = []
            for i in range(len(ds)):
                blk = lookup[offs[i] + (k - 1) * ds[i]: offs[i] + k * ds[i], (k - 1) * z:(k + 1) * z].toarray() - 1
                assert (blk >= 0).all()
                pos.append(blk)
        for i in range(len(ds)):
            if k <= N - 1:
                r, Jb, Hb = interval(i, k)
                gr = g[offs[i] + (k - 1) * ds[i]: offs[i] + k * ds[i]]
                e_r = max(e_r, np.abs(gr - r).max() / max(np.abs(r).max(), 1e-300))
                got = jac[pos[i]]
                blk = max(np.abs(Jb).max(), 1e-300)
                for half in (0, 1):
                    for c in range(nc):
                        sel = half * z + np.flatnonzero(comp == c)
                        sc = np.abs(Jb[:, sel]).max()
                        e_j = max(e_j, np.abs(got[:, sel] - Jb[:, sel]).max() / (sc if sc > 0 else blk))
                Hdiag += Hb[:z, :z]
            if k >= 2:
                _, _, Hp = interval(i, k - 1)
                Hdiag += Hp[z:, z:]
                Hcross += Hp[:z, z:]
        if sigma != 0:
            for ob in spec["objectives"]:
                one = dict(ob)
                if ob["kind"] in ("quadreg", "linreg", "knot"):
                    if k not in list(ob["times"]):
                        continue
                    idx = list(ob["times"]).index(k)
                    one["times"] = [k]
                    if ob["kind"] == "knot":
                        one["params"], one["Qs"] = [ob["params"][idx]], [ob["Qs"][idx]]
                elif ob["kind"] in ("mintime", "null"):
                    continue
                else:
                    raise AssertionError(ob["kind"])
                _, _, (hr, hc, hv) = orc._objective_term(spec, one, Z, False, True)
                keep = hr <= hc
                np.add.at(Hdiag, (hr[keep] - (k - 1) * z, hc[keep] - (k - 1) * z), sigma * ob.get("weight", 1.0) * hv[keep])
        if con_h is not None:
            hr, hc, hv = con_h
            keep = (hc // z == k - 1) & (hr <= hc)
            np.add.at(Hdiag, (hr[keep] - (k - 1) * z, hc[keep] - (k - 1) * z), hv[keep])
        cross, diag = hess_region_positions(spec, k)
        whole = max(np.abs(Hdiag[iu]).max(), np.abs(Hcross).max(), 1e-300)
        gd = hess[diag[iu]]
        sc = np.abs(Hdiag[iu]).max()
        e_h = max(e_h, np.abs(gd - Hdiag[iu]).max() / (sc if sc > 0 else whole))
        if cross is not None:
            sc = np.abs(Hcross).max()
            e_h = max(e_h, np.abs(hess[cross] - Hcross).max() / (sc if sc > 0 else whole))
    return e_r, e_j, e_h
