#!/usr/bin/env python
"""The clique-DECOMPOSED SDP of Chordal-DeepSDP, built from the hand-off a query path delivers and solved here.

TEST INFRASTRUCTURE (like everything under oracle/): not imported by the product.

What the reference builds (src/Methods/chordal_sdp.jl:96-153, findEllipsoid / setupReach!):
    variables   [gamma_in; gamma_out; gamma_ac1 (bounded); gamma_ac2 (sector)]  >= 0        (:124-139, creation order)
    blocks      one symmetric Z_k, -Z_k PSD, per clique (C_k, parts, D_k) of makeCliques     (setupZs!, :19-57)
                DoubleDecomp: for 1 < k < p two blocks Y_k1, Y_k2 embedded at Z_k[D_k1, D_k1], Z_k[D_k2, D_k2]  (:25-46)
    equalities  Z(gamma) .== Zksum = sum_k E_k' Z_k E_k                                     (setupZksum!, :60-93; :150)
    objective   gamma_out                                                                    (NnSdp.jl:46)
The GPU library hands over exactly the data of the equalities (nnsdp_affine_get: entries of the cover's upper
triangle, z0, COO triplets with duplicates to be summed, variables in the order above) and the cliques
(nnsdp_cliques: C_k, |C_k1|, D_k1, D_k2).  `problem_from_handoff` turns that into the block LMI -- every cover entry
equals the sum of the block entries that sit on it, entries of Z outside every block must carry no coefficient --
and `solve` finds its optimum with a block log-barrier method (no SDP solver is installed).  By Agler's theorem the
optimum equals that of the dense LMI Z(gamma) <= 0 (oracle/sdp_crosscheck.py) when, and only when, the cliques and
the entry numbering / duplicate summation / D_k embedding of the hand-off are right: a misplaced entry makes the
problem infeasible or moves the optimum.

    python oracle/sdp_decomposed.py <handoff.npz> [single|double] [out.json]

The .npz files are written on the GPU box by tools/dump_handoff.py (library outputs only); the solutions are
committed under tests/golden/ and re-checked WITHOUT the solver by tests/test_oracle_cpu.py.
"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.linalg import lu_factor
from scipy.linalg import lu_solve as _lu_solve

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


# ------------------------------------------------------------------------------------------------
# hand-off -> block LMI
# ------------------------------------------------------------------------------------------------
def blocks_of(cliques, mode):
    """Global 0-based index sets of the PSD blocks, in the order setupZs! creates them."""
    out = []
    for Ck, _, Dks in cliques:
        Ck = np.asarray(Ck, dtype=np.int64)
        if mode == "single" or len(Dks) == 1:
            out.append(Ck - 1)
        else:
            assert len(Dks) == 2
            for D in Dks:
                out.append(Ck[np.asarray(D, dtype=np.int64) - 1] - 1)
    return out


def problem_from_handoff(h, cliques, mode="single"):
    """h: dict with nvar, nent, var_out, ent_row, ent_col (1-based), z0, coo_ent, coo_var (1-based), coo_val.
    Returns the block LMI in the variables x = [gamma (those with a coefficient or a cost); splits]:
    for block j, Z_j(x) = Z0_j + sum_i x[V_j[i]] * A_j[i]   (dense, symmetric), to be kept negative definite."""
    nvar, nent = int(h["nvar"]), int(h["nent"])
    er, ec = np.asarray(h["ent_row"]) - 1, np.asarray(h["ent_col"]) - 1
    assert np.all(er <= ec)
    A = sp.coo_matrix((h["coo_val"], (np.asarray(h["coo_ent"]) - 1, np.asarray(h["coo_var"]) - 1)), shape=(nent, nvar)).tocsr()
    A.sum_duplicates()                       # "duplicate (entry, variable) pairs are to be summed"
    z0 = np.asarray(h["z0"], dtype=float)
    c = np.zeros(nvar)
    c[int(h["var_out"])] = 1.0               # obj_func = gamma_out[1]
    colmax = np.asarray(abs(A).max(axis=0).todense()).ravel()
    keep = (colmax > 1e-13 * colmax.max()) | (c != 0)
    kidx = np.nonzero(keep)[0]
    A = A[:, kidx].tocsr()
    ng = len(kidx)
    blocks = blocks_of(cliques, mode)
    Zdim = int(max(ec.max(), max(b.max() for b in blocks))) + 1
    # owners of every cover entry
    member = np.zeros((len(blocks), Zdim), dtype=bool)
    for j, B in enumerate(blocks):
        member[j, B] = True
    own = member[:, er] & member[:, ec]                       # (nblocks, nent)
    nown = own.sum(0)
    # entries of Z that no block holds must be structurally zero (chordal_sdp.jl:150 would read 0 == 0 there)
    orphan = nown == 0
    assert np.all(z0[orphan] == 0.0) and abs(A[orphan]).sum() == 0.0, "Z has a coefficient outside every block"
    primary = np.where(nown > 0, len(blocks) - 1 - np.argmax(own[::-1], axis=0), -1)   # the last owner
    # split variables: one per (entry, non-primary owner)
    split_id = -np.ones(own.shape, dtype=np.int64)
    ns = 0
    for j in range(len(blocks)):
        sel = own[j] & (primary != j)
        split_id[j, sel] = ng + ns + np.arange(sel.sum())
        ns += int(sel.sum())
    n = ng + ns
    loc = [-np.ones(Zdim, dtype=np.int64) for _ in blocks]
    for j, B in enumerate(blocks):
        loc[j][B] = np.arange(len(B))
    out_blocks = []
    for j, B in enumerate(blocks):
        m = len(B)
        ents = np.nonzero(own[j])[0]
        lr, lc = loc[j][er[ents]], loc[j][ec[ents]]
        assert len(ents) == m * (m + 1) // 2, "a block must lie inside the cover"
        rows, cols, vals = [], [], []
        Z0 = np.zeros((m, m))
        pe = ents[primary[ents] == j]                          # entries this block is the primary owner of
        if len(pe):
            sub = A[pe].tocoo()
            f = loc[j][er[pe]] * m + loc[j][ec[pe]]
            rows.append(f[sub.row]); cols.append(sub.col); vals.append(sub.data)
            Z0[loc[j][er[pe]], loc[j][ec[pe]]] = z0[pe]
            for j2 in range(len(blocks)):                      # minus what the other owners hold
                if j2 == j:
                    continue
                s = split_id[j2, pe]
                ok = s >= 0
                rows.append(f[ok]); cols.append(s[ok]); vals.append(-np.ones(ok.sum()))
        se = ents[primary[ents] != j]
        if len(se):
            f = loc[j][er[se]] * m + loc[j][ec[se]]
            rows.append(f); cols.append(split_id[j, se]); vals.append(np.ones(len(se)))
        G = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(m * m, n)).tocsc()
        V = np.nonzero(np.diff(G.indptr))[0]                      # variables with a coefficient in this block
        Aj = np.asarray(G[:, V].todense()).T.reshape(len(V), m, m)
        Aj = Aj + np.transpose(Aj, (0, 2, 1)) - Aj * np.eye(m)[None]      # upper triangle -> symmetric
        Z0 = Z0 + Z0.T - np.diag(np.diag(Z0))
        out_blocks.append({"idx": B, "V": V, "A": Aj, "Z0": Z0})
    cx = np.concatenate([c[kidx], np.zeros(ns)])
    return {"blocks": out_blocks, "c": cx, "ng": ng, "ns": ns, "keep": keep, "nvar": nvar,
            "split_id": split_id, "own": own, "ent_row": er, "ent_col": ec, "Zdim": Zdim}


def block_matrices(prob, x):
    return [b["Z0"] + np.tensordot(x[b["V"]], b["A"], 1) for b in prob["blocks"]]


def lambda_max(prob, x):
    return max(float(np.linalg.eigvalsh(Z).max()) for Z in block_matrices(prob, x))


# ------------------------------------------------------------------------------------------------
# block log-barrier
# ------------------------------------------------------------------------------------------------
def barrier(prob, x, U, gap, shift_var=None, stop=None, mu=5.0, max_newton=300, verbose=False):
    """min c'x s.t. 0 < x[:ng] < U, Z_j(x) - s I < 0 for every block (s = x[shift_var] or 0).
    Damped Newton on the barrier; the multipliers are scaled by the iterate, the split variables are free."""
    blocks, c, ng = prob["blocks"], prob["c"], prob["ng"]
    n = len(c)
    msum = sum(len(b["idx"]) for b in blocks)

    def mats(x):
        s = x[shift_var] if shift_var is not None else 0.0
        return [s * np.eye(len(b["idx"])) - (b["Z0"] + np.tensordot(x[b["V"]], b["A"], 1)) for b in blocks]

    def feasible(x):
        g = x[:ng]
        if np.any(g <= 0) or np.any(g >= U):
            return False
        for M in mats(x):
            try:
                np.linalg.cholesky(M)
            except np.linalg.LinAlgError:
                return False
        return True

    t, it = 1.0, 0
    while True:
        for _ in range(max_newton):
            g = t * c.copy()
            H = np.zeros((n, n))
            for b, M in zip(blocks, mats(x)):
                Mi = np.linalg.inv(M)
                Mi = 0.5 * (Mi + Mi.T)
                V, A = b["V"], b["A"]
                nv, m = len(V), M.shape[0]
                g[V] += A.reshape(nv, -1) @ Mi.ravel()
                T = (Mi @ A @ Mi).reshape(nv, -1)
                Hj = T @ A.reshape(nv, -1).T
                H[np.ix_(V, V)] += 0.5 * (Hj + Hj.T)
                if shift_var is not None:          # the shift enters every block as -s I  (A_s = -I)
                    g[shift_var] -= np.trace(Mi)
                    hv = -(T @ np.eye(m).ravel())
                    H[V, shift_var] += hv
                    H[shift_var, V] += hv
                    H[shift_var, shift_var] += float(np.sum(Mi * Mi))
            gam = x[:ng]
            g[:ng] += -1.0 / gam + 1.0 / (U - gam)
            H[np.arange(ng), np.arange(ng)] += 1.0 / gam ** 2 + 1.0 / (U - gam) ** 2
            sc = np.ones(n)
            sc[:ng] = gam
            Hs = sc[:, None] * H * sc[None, :]
            gs = sc * g
            Hs[np.arange(n), np.arange(n)] += 1e-14 * np.abs(np.diag(Hs)).max()
            du = -np.linalg.solve(Hs, gs)
            lam = np.sqrt(max(-gs @ du, 0.0))
            it += 1
            if verbose:
                print(f"    t {t:.1e} it {it} lam {lam:.3e} obj {c @ x:.8f}", flush=True)
            if lam < 1e-4:
                break
            step = 1.0 if lam < 0.25 else 1.0 / (1.0 + lam)
            dx = sc * du
            while not feasible(x + step * dx) and step > 1e-14:
                step *= 0.5
            if step <= 1e-14:
                break
            x = x + step * dx
            if stop is not None and stop(x):
                return x, it
        if (2 * ng + msum) / t < gap:
            return x, it
        t *= mu


def primal_dual(prob, y, U, tol=1e-9, max_iter=200, verbose=False, stop=None):
    """Primal-dual path following (HKM direction) from a strictly feasible y for
         min c'y   s.t.  S_j(y) = -Z_j(y) >= 0 (blocks),  y[:ng] >= 0,  U - y[:ng] >= 0.
    Multipliers X_j >= 0 (blocks), xl, xu >= 0 (bounds); the dual objective -sum <C_j, X_j> - U sum xu bounds the
    optimum from below once the dual residual c + sum_j A_j^*(X_j) - xl + xu vanishes, so the returned gap is a
    certificate and not an estimate.  Returns (y, info)."""
    blocks, c, ng = prob["blocks"], prob["c"], prob["ng"]
    n = len(c)

    def S_of(y):
        return [-(b["Z0"] + np.tensordot(y[b["V"]], b["A"], 1)) for b in blocks]

    S = S_of(y)
    sl, su = y[:ng].copy(), U - y[:ng]
    msum = sum(M.shape[0] for M in S) + 2 * ng
    mu = (abs(c @ y) + 1.0) / msum
    X = [mu * np.linalg.inv(M) for M in S]
    xl, xu = mu / sl, mu / su

    def max_step(M, dM):
        """largest a with M + a dM >= 0 (M > 0)"""
        L = np.linalg.cholesky(M)
        Li = np.linalg.inv(L)
        w = np.linalg.eigvalsh(Li @ dM @ Li.T).min()
        return np.inf if w >= 0 else -1.0 / w

    info, best, stall = {}, None, 0
    for it in range(max_iter):
        gap = sum(float(np.sum(Xj * Sj)) for Xj, Sj in zip(X, S)) + xl @ sl + xu @ su
        rd = c.copy()                                   # dual residual
        for b, Xj in zip(blocks, X):
            rd[b["V"]] += b["A"].reshape(len(b["V"]), -1) @ Xj.ravel()
        rd[:ng] += -xl + xu
        pobj = float(c @ y)
        dobj = -sum(float(np.sum(Xj * (-b["Z0"]))) for b, Xj in zip(blocks, X)) - U * xu.sum()
        rdn = float(np.abs(rd).max())
        info = {"iter": it, "pobj": pobj, "dobj": dobj, "gap": gap, "rd": rdn}
        if verbose:
            print(f"    it {it:3d} pobj {pobj:.10f} dobj {dobj:.10f} gap {gap:.2e} rd {rdn:.2e}", flush=True)
        if stop is not None and stop(y):
            return y, info
        if rdn < 1e-6 and (best is None or pobj - dobj < best["pobj"] - best["dobj"]):
            best = dict(info, y=y.copy(), X=[Xj.copy() for Xj in X], xl=xl.copy(), xu=xu.copy())
            stall = 0
        else:
            stall += 1
        if best is not None and stall >= 6:             # no better bracket for six iterations: the iterates sit on
            break                                       # the boundary of the (degenerate) optimal face
        if gap < tol * (1 + abs(pobj)) and rdn < tol * (1 + np.abs(c).max()):
            break
        if best is not None and rdn > 1e-4:             # numerical breakdown near the (degenerate) optimum: stop
            break
        mu = gap / msum
        try:
            for M in S + X:
                np.linalg.cholesky(M)
        except np.linalg.LinAlgError:                   # rounding pushed an iterate out of the cone: stop here
            break
        Si = [np.linalg.inv(M) for M in S]
        Mat = np.zeros((n, n))
        for b, Xj, Sij in zip(blocks, X, Si):
            V, A = b["V"], b["A"]
            nv = len(V)
            T = (Sij @ A @ Xj)                         # S^-1 A_k X
            Mj = A.reshape(nv, -1) @ np.transpose(T, (0, 2, 1)).reshape(nv, -1).T
            Mat[np.ix_(V, V)] += 0.5 * (Mj + Mj.T)
        Mat[np.arange(ng), np.arange(ng)] += xl / sl + xu / su
        dscale = 1.0 / np.sqrt(np.maximum(np.diag(Mat), 1e-300))   # Jacobi scaling: multipliers and splits differ by orders
        lu = lu_factor(dscale[:, None] * Mat * dscale[None, :])
        lu_solve = lambda f, r: dscale * _lu_solve(f, dscale * r)

        def direction(sigma_mu, corr=None):
            rhs = -rd.copy()
            # targets: X + dX = sigma_mu S^-1 + sum_k dy_k S^-1 A_k X  (+ correction)
            for b, Xj, Sij in zip(blocks, X, Si):
                rhs[b["V"]] -= b["A"].reshape(len(b["V"]), -1) @ (sigma_mu * Sij - Xj).ravel()
            rhs[:ng] -= -(sigma_mu / sl - xl) + (sigma_mu / su - xu)
            if corr is not None:
                rhs -= corr
            dy = lu_solve(lu, rhs)
            for _ in range(2):                          # iterative refinement of the Schur system
                dy += lu_solve(lu, rhs - Mat @ dy)
            dS = [-np.tensordot(dy[b["V"]], b["A"], 1) for b in blocks]
            dX = []
            for Xj, Sij, dSj in zip(X, Si, dS):
                G = sigma_mu * Sij - Xj - Sij @ dSj @ Xj
                dX.append(0.5 * (G + G.T))
            dsl, dsu = dy[:ng], -dy[:ng]
            dxl = sigma_mu / sl - xl - xl / sl * dsl
            dxu = sigma_mu / su - xu - xu / su * dsu
            return dy, dS, dX, dsl, dsu, dxl, dxu

        def steps(dS, dX, dsl, dsu, dxl, dxu):
            ap = min([max_step(M, d) for M, d in zip(S, dS)] + [np.inf])
            for v, d in ((sl, dsl), (su, dsu)):
                neg = d < 0
                if np.any(neg):
                    ap = min(ap, float(np.min(-v[neg] / d[neg])))
            ad = min([max_step(M, d) for M, d in zip(X, dX)] + [np.inf])
            for v, d in ((xl, dxl), (xu, dxu)):
                neg = d < 0
                if np.any(neg):
                    ad = min(ad, float(np.min(-v[neg] / d[neg])))
            return ap, ad

        # predictor (sigma = 0) gives the centring parameter (Mehrotra's rule); the step itself is the centred one
        dy, dS, dX, dsl, dsu, dxl, dxu = direction(0.0)
        ap, ad = steps(dS, dX, dsl, dsu, dxl, dxu)
        ap, ad = min(1.0, ap), min(1.0, ad)
        gap_aff = sum(float(np.sum((Xj + ad * dXj) * (Sj + ap * dSj))) for Xj, dXj, Sj, dSj in zip(X, dX, S, dS)) \
            + (xl + ad * dxl) @ (sl + ap * dsl) + (xu + ad * dxu) @ (su + ap * dsu)
        sigma = min(0.9, max(0.05, (gap_aff / gap) ** 2)) if gap > 0 else 0.3
        dy, dS, dX, dsl, dsu, dxl, dxu = direction(sigma * mu)
        ap, ad = steps(dS, dX, dsl, dsu, dxl, dxu)
        ap, ad = min(1.0, 0.9 * ap), min(1.0, 0.9 * ad)
        y = y + ap * dy
        S = S_of(y)
        sl, su = y[:ng].copy(), U - y[:ng]
        X = [Xj + ad * d for Xj, d in zip(X, dX)]
        xl, xu = xl + ad * dxl, xu + ad * dxu
    if best is not None:
        return best["y"], best
    return y, info


def solve(prob, gamma_start=None, U=1e4, tol=1e-9, verbose=False):
    """Phase I by the barrier (a common shift s of every block driven below zero) from gamma_start (or ones), then
    the optimum by the primal-dual method."""
    ng, ns = prob["ng"], prob["ns"]
    n = ng + ns
    x = np.ones(n)
    x[ng:] = 0.0
    if gamma_start is not None:
        x[:ng] = np.clip(np.asarray(gamma_start)[prob["keep"]], 1e-6, None)
    U = max(U, 4 * x[:ng].max())
    # phase I: min s  s.t.  Z_j(y) - s I <= 0 -- the same primal-dual method on the problem with one more free
    # variable (coefficient -I in every block), started at s = lambda_max + 1 and stopped as soon as s < 0 by a margin
    p1 = {"c": np.concatenate([np.zeros(n), [1.0]]), "ng": ng, "ns": ns + 1, "blocks": [
        {"idx": b["idx"], "Z0": b["Z0"], "V": np.concatenate([b["V"], [n]]),
         "A": np.concatenate([b["A"], -np.eye(len(b["idx"]))[None]], 0)} for b in prob["blocks"]]}
    xs = np.concatenate([x, [lambda_max(prob, x) + 1.0]])
    scale = max(1.0, max(float(np.abs(b["Z0"]).max()) for b in prob["blocks"]))
    xs, info1 = primal_dual(p1, xs, U, tol=1e-9, verbose=verbose, stop=lambda z: z[n] < -1e-6 * scale)
    it1 = info1.get("iter", 0)
    assert xs[n] < 0, "phase I did not find a strictly feasible point"
    x = xs[:n]
    assert lambda_max(prob, x) < 0
    x, info = primal_dual(prob, x, U, tol=tol, verbose=verbose)
    gamma = np.ones(prob["nvar"])
    gamma[prob["keep"]] = x[:ng]
    return {"obj": float(prob["c"] @ x), "dual_obj": info["dobj"], "gap": info["gap"], "dual_residual": info["rd"], "x": x,
            "gamma": gamma, "newton": it1 + info["iter"], "lambda_max": lambda_max(prob, x), "U": U,
            "X": info.get("X"), "xl": info.get("xl"), "xu": info.get("xu")}


def certificate(prob, x, X, xl, xu, U):
    """Solver-free check of a stored solution.  Returns (primal objective, lambda_max over the blocks, dual objective,
    max |dual residual|, rd' x): x is feasible iff lambda_max <= 0 and 0 <= x[:ng] <= U, so the optimum is AT MOST the
    primal objective; with X_j, xl, xu >= 0 weak duality gives  c'y >= dual objective + rd'y  for every feasible y."""
    c, ng = prob["c"], prob["ng"]
    rd = c.copy()
    dobj = -U * float(np.sum(xu))
    for b, Xj in zip(prob["blocks"], X):
        rd[b["V"]] += b["A"].reshape(len(b["V"]), -1) @ Xj.ravel()
        dobj += float(np.sum(Xj * b["Z0"]))
    rd[:ng] += -xl + xu
    return float(c @ x), lambda_max(prob, x), dobj, float(np.abs(rd).max()), float(rd @ x)


# ------------------------------------------------------------------------------------------------
# A certificate for the decomposed problem from a solve of the DENSE one (Agler's theorem, made constructive).
#   * primal: Z(gamma*) <= 0 with the sparsity of the cover has a zero-fill factorisation  -Z = L D L'  in the natural
#     order (eliminating an index of block x_k only touches later indices of clique k: the cliques of makeCliques,
#     /root/reference/src/Methods/chordal_cliques.jl:13-59, are a perfect elimination ordering with the common tail
#     x_K, 1 last).  The columns of L are grouped by the clique that owns their index, Z_k = -sum_i d_i l_i l_i': every
#     Z_k is negative semidefinite, lies inside C_k x C_k, and they add up to Z.  The split variables of the block LMI
#     are read off the Z_k.
#   * dual: a multiplier X >= 0 of the dense LMI restricted to the blocks, X_j = X[B_j, B_j], is a multiplier of the
#     decomposed one with the same dual objective and the same dual residual on gamma (and none on the splits).
# Both sides are then checked by certificate() exactly like an interior-point solution: nothing here is trusted.
# Used where the block interior-point method runs out of digits (W10-D20: bracket 2e-3).
# ------------------------------------------------------------------------------------------------
def dense_from_handoff(h, keep):
    """(Z0, A) of the dense LMI  Z0 + sum_v x_v A_v <= 0  over the kept variables, symmetric Zdim x Zdim."""
    er, ec = np.asarray(h["ent_row"]) - 1, np.asarray(h["ent_col"]) - 1
    Zdim = int(ec.max()) + 1
    nvar, nent = int(h["nvar"]), int(h["nent"])
    A = sp.coo_matrix((h["coo_val"], (np.asarray(h["coo_ent"]) - 1, np.asarray(h["coo_var"]) - 1)), shape=(nent, nvar)).tocsc()
    A.sum_duplicates()
    kidx = np.nonzero(keep)[0]
    Z0 = np.zeros((Zdim, Zdim))
    Z0[er, ec] = np.asarray(h["z0"], dtype=float)
    Z0 = Z0 + Z0.T - np.diag(np.diag(Z0))
    Ad = np.zeros((len(kidx), Zdim, Zdim))
    for i, v in enumerate(kidx):
        col = A[:, v].tocoo()
        Ad[i, er[col.row], ec[col.row]] = col.data
        Ad[i] = Ad[i] + Ad[i].T - np.diag(np.diag(Ad[i]))
    return Z0, Ad


def chordal_split(Z, blocks, eps):
    """Z (negative definite, pattern inside the union of blocks x blocks, blocks in clique order) -> [Z_k] with
    sum_k embed(Z_k) = Z and every Z_k <= -eps I: zero-fill LDL' of -(Z + eps diag(multiplicity)) in the natural order,
    columns grouped by the last block that holds their index."""
    n = Z.shape[0]
    mult = np.zeros(n)
    for B in blocks:
        mult[B] += 1.0
    M = -(Z + eps * np.diag(mult))
    owner = -np.ones(n, dtype=np.int64)            # the LAST block that contains the index: an index of x_b lies in the
    for j in range(len(blocks)):                   # cliques b-2 (if within beta of the layer start), b-1 and b, and its
        owner[blocks[j]] = j                       # later neighbours (in the filled graph) all lie in clique b
    loc = []
    for B in blocks:
        l = -np.ones(n, dtype=np.int64)
        l[B] = np.arange(len(B))
        loc.append(l)
    out = [-eps * np.eye(len(B)) for B in blocks]
    W = M.copy()
    for i in range(n):
        d = W[i, i]
        assert d > 0, ("not negative definite", i, d)
        l = W[i:, i] / d
        nz = i + np.nonzero(l)[0]
        j = owner[i]
        assert np.all(loc[j][nz] >= 0), ("fill outside the clique", i)
        li = loc[j][nz]
        out[j][np.ix_(li, li)] -= d * np.outer(l[nz - i], l[nz - i])
        W[i:, i:] -= d * np.outer(l, l)
        W[i, i:] = 0.0
        W[i:, i] = 0.0
    return out


def certificate_from_dense(h, cliques, mode="single", U=1e4, gap=1e-7, verbose=False):
    """Solves the dense LMI of the hand-off by the barrier method of sdp_crosscheck.py and turns the result into a
    primal point and multipliers of the decomposed problem (see above).  Returns the dict solve() returns."""
    import sdp_crosscheck as sc

    prob = problem_from_handoff(h, cliques, mode)
    ng, ns = prob["ng"], prob["ns"]
    Z0, A = dense_from_handoff(h, prob["keep"])
    c = prob["c"][:ng]
    n, m = A.shape[0], Z0.shape[0]
    A1 = np.concatenate([A, -np.eye(m)[None], np.eye(m)[None]], 0)          # phase I as in sdp_crosscheck.solve
    c1 = np.zeros(n + 2)
    c1[n], c1[n + 1] = 1.0, -1.0
    s0 = np.linalg.eigvalsh(Z0 + np.tensordot(np.ones(n), A, 1)).max()
    z0 = np.concatenate([np.ones(n), [max(s0, 0.0) + 2.0, 1.0]])
    U = max(U, 4 * z0.max())
    # barrier_solve multiplies t by 5 from 1 until (2n + m) / t < gap: the parameter of its last centring step
    t = 1.0
    while (2 * n + m) / t >= gap:
        t *= 5.0
    if os.environ.get("SDP_LOAD_STATE"):          # the two points of an earlier run (SDP_SAVE_STATE)
        st = np.load(os.environ["SDP_LOAD_STATE"])
        gc, g, Xd, U, it1, it2 = st["gc"], st["g"], st["Xd"], float(st["U"]), 0, 0
    else:
        z, _, it1 = sc.barrier_solve(Z0, A1, c1, z0, U=U, gap=1e-3, stop=lambda z: z[n] - z[n + 1] < -1e-3)
        g = z[:n]
        gc, _, it2 = sc.barrier_solve(Z0, A, c, g, U=U, gap=1e-3)            # a centred point well inside the cone
        tc = 1.0
        while (2 * n + m) / tc >= 1e-3:
            tc *= 5.0
        if verbose:
            print(f"  dense barrier: centred at t = {tc:.3g} after {it1 + it2} Newton steps, objective {float(c @ gc):.8f}",
                  flush=True)
        g, Xd, it3 = sc.barrier_solve(Z0, A, c, gc, U=U, gap=gap, t0=tc)     # on from there (not from t = 1 again)
        if verbose:
            print(f"  dense barrier: gap {gap:g} after {it3} more steps, objective {float(c @ g):.10f}", flush=True)
        it2 += it3
        if os.environ.get("SDP_SAVE_STATE"):      # an hour of Newton steps at W10-D20: keep the two points
            np.savez(os.environ["SDP_SAVE_STATE"], gc=gc, g=g, Xd=Xd, U=U, t=t)
    # a small step back towards the centred point: the LMI then holds with a margin the factorisation below can afford
    # (lambda_max(Z) ~ -1e-12 at the barrier's last iterate is below its rounding), for ~1e-6 of objective; the
    # smallest step that leaves every block negative definite by half of the shift (and by 1e-11) is taken
    dense_obj = float(c @ g)
    g_star = g
    blocks = [b["idx"] for b in prob["blocks"]]
    mmax = max(np.bincount(np.concatenate(blocks)))
    for back in (3e-3, 1e-2, 3e-2):
        g = g_star + back * (gc - g_star)
        Zg = Z0 + np.tensordot(g, A, 1)
        lam = float(np.linalg.eigvalsh(Zg).max())
        assert lam < 0, lam
        eps = -lam / (4.0 * mmax)
        Zk = chordal_split(Zg, blocks, eps)
        worst = max(float(np.linalg.eigvalsh(0.5 * (M + M.T)).max()) for M in Zk)
        if verbose:
            print(f"  step back {back:g}: lambda_max(Z) {lam:.2e}, blocks {worst:.2e} (shift {eps:.1e}), "
                  f"objective {float(c @ g):.10f}", flush=True)
        if worst <= -0.5 * eps and worst <= -1e-11:      # a margin that survives another BLAS's summation order
            break
    # the split variables: what every non-primary owner of a cover entry holds
    x = np.zeros(ng + ns)
    x[:ng] = g
    er, ec, split_id = prob["ent_row"], prob["ent_col"], prob["split_id"]
    for j, B in enumerate(blocks):
        l = -np.ones(prob["Zdim"], dtype=np.int64)
        l[B] = np.arange(len(B))
        sel = np.nonzero(split_id[j] >= 0)[0]
        x[split_id[j, sel]] = Zk[j][l[er[sel]], l[ec[sel]]]
    X = [Xd[np.ix_(B, B)] for B in blocks]
    # multipliers of the box: the barrier's dual estimates, then the residual on gamma is pushed into them
    xl, xu = 1.0 / (t * g), 1.0 / (t * (U - g))
    rd = c.copy()
    for b, Xj in zip(prob["blocks"], X):
        V = b["V"]
        r = b["A"].reshape(len(V), -1) @ Xj.ravel()
        rd[V[V < ng]] += r[V < ng]
    rd += -xl + xu
    # the residual on gamma goes into the box multipliers, which stay >= 0: rd > 0 raises xl (free: the lower bound is
    # 0); rd < 0 first lowers xl as far as it goes, only the rest raises xu (which costs U per unit in the dual objective)
    xl += np.maximum(rd, 0.0)
    take = np.minimum(xl, np.maximum(-rd, 0.0))
    xl -= take
    xu += np.maximum(-rd, 0.0) - take
    pobj, lmax, dobj, rdn, rdx = certificate(prob, x, X, xl, xu, U)
    if verbose:
        print(f"  dense barrier: {it1 + it2} Newton steps, objective {pobj:.10f}, lambda_max(Z) {lam:.2e}; blocks: "
              f"lambda_max {lmax:.2e}, dual objective {dobj:.10f}, dual residual {rdn:.1e}", flush=True)
    gamma = np.ones(prob["nvar"])
    gamma[prob["keep"]] = g
    return {"obj": pobj, "dual_obj": dobj, "gap": pobj - dobj, "dual_residual": rdn, "x": x, "gamma": gamma,
            "newton": it1 + it2, "lambda_max": lmax, "U": U, "X": X, "xl": xl, "xu": xu, "dense_obj": dense_obj}


# ------------------------------------------------------------------------------------------------
# the hand-off in the library's format, from the ORACLE (for CPU development and as a second source)
# ------------------------------------------------------------------------------------------------
def handoff_from_oracle(net, beta, x1min, x1max, qc_out, intv_info=None):
    import nnsdp_oracle as o

    Z0, Zv = o.affine_structure(net, beta, x1min, x1max, qc_out, intv_info=intv_info)
    cliques = o.make_cliques(net, beta)
    er, ec = o.cover_upper_entries(net, cliques)
    nent, nvar = len(er), len(Zv)
    ce, cv, cval = [], [], []
    for v, M in enumerate(Zv):
        vals = M[er - 1, ec - 1]
        nz = np.nonzero(vals)[0]
        ce.append(nz + 1); cv.append(np.full(len(nz), v + 1)); cval.append(vals[nz])
    n1 = net.xdims[0]
    has_out = not isinstance(qc_out, o.QcSafety)
    return {"nvar": nvar, "nent": nent, "var_in": 0, "var_out": n1, "var_bnd": n1 + (1 if has_out else 0),
            "ent_row": er, "ent_col": ec, "z0": Z0[er - 1, ec - 1], "coo_ent": np.concatenate(ce),
            "coo_var": np.concatenate(cv), "coo_val": np.concatenate(cval)}, cliques


def cliques_from_npz(d):
    """[(Ck, parts, Dks)] from the flat arrays of nnsdp_cliques stored by tools/dump_handoff.py."""
    ck_off, ck_idx, ck1, d_off, d_idx = (d[k] for k in ("ck_off", "ck_idx", "ck1_len", "d_off", "d_idx"))
    out = []
    for k in range(len(ck1)):
        Ck = ck_idx[ck_off[k]:ck_off[k + 1]]
        parts = [Ck[:ck1[k]]] + ([Ck[ck1[k]:]] if ck1[k] < len(Ck) else [])
        Ds = [d_idx[d_off[2 * k]:d_off[2 * k + 1]]]
        if d_off[2 * k + 2] > d_off[2 * k + 1]:
            Ds.append(d_idx[d_off[2 * k + 1]:d_off[2 * k + 2]])
        out.append((Ck, parts, Ds))
    return out


def main():
    path = sys.argv[1]
    mode = sys.argv[2] if len(sys.argv) > 2 else "single"
    d = np.load(path)
    cliques = cliques_from_npz(d)
    t0 = time.time()
    prob = problem_from_handoff(d, cliques, mode)
    print(f"{os.path.basename(path)} [{mode}]: {len(prob['blocks'])} blocks, {prob['ng']} multipliers + {prob['ns']} splits, "
          f"build {time.time() - t0:.1f} s", flush=True)
    start = None
    if "--start" in sys.argv:   # multipliers of the dense optimum (tests/golden/scale_W10_D10_optimum.json) as a start
        res = json.load(open(sys.argv[sys.argv.index("--start") + 1]))
        start = np.array(res["oracle_optimum"][str(int(d["beta"]))]["gamma"])
    if "--from-dense" in sys.argv:   # certificate built from a barrier solve of the dense LMI (see certificate_from_dense)
        r = certificate_from_dense(d, cliques, mode, verbose="-v" in sys.argv)
    else:
        r = solve(prob, gamma_start=start, verbose="-v" in sys.argv)
    print(f"  optimum {r['obj']:.8f}  lambda_max over blocks {r['lambda_max']:.2e}  {r['newton']} Newton steps, "
          f"{time.time() - t0:.0f} s", flush=True)
    print(f"  dual objective {r['dual_obj']:.8f}  gap {r['gap']:.1e}  dual residual {r['dual_residual']:.1e}", flush=True)
    if len(sys.argv) > 3 and not sys.argv[3].startswith("-"):
        extra = {"dense_obj": r["dense_obj"]} if "dense_obj" in r else {}
        np.savez_compressed(sys.argv[3], obj=r["obj"], dual_obj=r["dual_obj"], lambda_max=r["lambda_max"], U=r["U"],
                            iterations=r["newton"], mode=mode, x=r["x"], gamma=r["gamma"],
                            X=np.concatenate([Xj.ravel() for Xj in r["X"]]), xl=r["xl"], xu=r["xu"], **extra)


if __name__ == "__main__":
    main()
