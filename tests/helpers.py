"""Shared builders for the parity tests: seeded random nets / queries in oracle and ABI form."""
import numpy as np

import nnsdp_oracle as o


def rand_net(xdims, seed, sigma=None):
    rng = np.random.default_rng(seed)
    W = max(xdims[1:-1])
    if sigma is None:
        sigma = 2.0 / np.sqrt(W * np.log(max(W, 3)))  # scripts/make_networks.jl:44 rule
    return o.random_network(xdims, sigma, rng)


def rand_out(net, kind, rng):
    n1, nK1 = net.xdims[0], net.xdims[-1]
    if kind == "safety":
        A = rng.standard_normal((n1 + nK1 + 1, n1 + nK1 + 1))
        return o.QcSafety(S=A + A.T), None
    if kind == "hplaneS":
        return o.QcSafety(S=o.hplaneS(rng.standard_normal(nK1), 0.7, net)), None
    if kind == "hplane":
        return o.QcReachHplane(rng.standard_normal(nK1)), rng.random(1)
    if kind == "circle":
        return o.QcReachCircle(rng.standard_normal(nK1)), rng.random(1)
    if kind == "ellipsoid":
        return o.QcReachEllipsoid(rng.standard_normal((nK1, nK1)), rng.standard_normal(nK1)), rng.random(1)
    raise ValueError(kind)


def rand_query(net, beta, rng, kind="safety", radius=0.05, centre=None):
    n1, ac = net.xdims[0], net.acdim
    c = rng.uniform(0.5, 1.5, n1) if centre is None else np.asarray(centre, dtype=float)
    qc_out, gout = rand_out(net, kind, rng)
    return o.NumericQuery(
        x1min=c - radius, x1max=c + radius, gin=rng.random(n1), gbnd=rng.random(ac),
        gsec=rng.random(o.sector_lambda_dim(ac, beta) + 2 * ac), qc_out=qc_out, gout=gout)


def to_numeric_batch(nb, net, queries, with_bounds=None):
    """oracle NumericQuery list -> nnsdp_b200.NumericBatch (all queries share one output kind)."""
    q0 = queries[0].qc_out
    kw = {}
    if isinstance(q0, o.QcSafety):
        kw = dict(out_kind=nb.OUT_SAFETY, out_S=np.stack([q.qc_out.S for q in queries]))
    elif isinstance(q0, o.QcReachHplane):
        kw = dict(out_kind=nb.OUT_HPLANE, out_vec=np.stack([q.qc_out.normal for q in queries]),
                  gamma_out=np.stack([q.gout for q in queries]))
    elif isinstance(q0, o.QcReachCircle):
        kw = dict(out_kind=nb.OUT_CIRCLE, out_vec=np.stack([q.qc_out.yc for q in queries]),
                  gamma_out=np.stack([q.gout for q in queries]))
    elif isinstance(q0, o.QcReachEllipsoid):
        kw = dict(out_kind=nb.OUT_ELLIPSOID, out_vec=np.stack([q.qc_out.yc for q in queries]),
                  out_invP=np.stack([q.qc_out.invP for q in queries]),
                  gamma_out=np.stack([q.gout for q in queries]))
    if with_bounds is not None:
        kw.update(ymin=np.stack([b.acymin for b, _ in with_bounds]), ymax=np.stack([b.acymax for b, _ in with_bounds]),
                  smin=np.stack([s.smin for _, s in with_bounds]), smax=np.stack([s.smax for _, s in with_bounds]))
    return nb.NumericBatch(
        x1min=np.stack([q.x1min for q in queries]), x1max=np.stack([q.x1max for q in queries]),
        gamma_in=np.stack([q.gin for q in queries]), gamma_bnd=np.stack([q.gbnd for q in queries]),
        gamma_sec=np.stack([q.gsec for q in queries]), **kw)


def relerr(a, b):
    """normwise relative error (Frobenius / max-abs scale), the 1e-12 criterion for FP64 blocks."""
    a = np.asarray(a)
    b = np.asarray(b)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale
