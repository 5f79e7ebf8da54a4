"""CPU oracle for the Chordal-DeepSDP constraint-construction hot path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  The product path (``nn-sdp_b200/``) never calls it and has no
CPU fallback.

PARITY UNPINNED for the Julia part: the reference (AntonXue/nn-sdp) is Julia + JuMP +
MOSEK, none of which exist in this environment, and the reference ships no tests,
golden vectors or fixtures for this path.  Two pieces
ARE pinned against reference artefacts: the CROWN bounds (the reference's Python
dependency auto_LiRPA runs here; see intervals_crown below) and the clique cover /
sparsity pattern (the reference's plot_sparsity notebook).  This restatement
follows the Julia sources line by line (citations below, relative to
/root/reference) and is cross-checked by (1) a second, independent closed-form
derivation in this file, (2) the structural sparsity statement in the reference's
``experiments/plot_sparsity.ipynb`` cell 5 (nnz(Z) lies inside quickRawZ, and the clique
cover equals quickRawZ + quickEK exactly: index sets and sparsity ARE pinned by it), (3) the reference's constructor
``@assert``s, which are reproduced here as assertions, and (4) hand-worked
examples in ``tests/``.

Two forms are provided:

* *literal*: selector matrices ``E``/``Ec``, ``R' * Q * R`` with scipy.sparse, the
  same operations in the same order the Julia code performs when ``gamma`` is a
  numeric ``Vector{Float64}`` (the reference does this at
  scripts/test_acas.jl:81-85).
* *closed form*: per-variable coefficient formulas,
  vectorised numpy.  This is the strongest fair CPU baseline and is what
  ``bench.py`` times as ``cpu_baseline`` (kind "port").

All index sets are 1-based Int64 (Julia convention) on the API surface so they
compare bit-exactly with what ``makeCliques`` returns.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

# --------------------------------------------------------------------------
# src/MyMath.jl:26-52  --  e, E, Ec
# --------------------------------------------------------------------------


def e(i: int, dim: int) -> sp.csr_matrix:
    """i-th basis vector (1-based) as a dim x 1 sparse column.  MyMath.jl:26-31."""
    assert 1 <= i <= dim
    return sp.csr_matrix(([1.0], ([i - 1], [0])), shape=(dim, 1))


def E(i: int, dims: Sequence[int]) -> sp.csr_matrix:
    """i-th block selector, dims[i] x sum(dims).  MyMath.jl:34-43."""
    dims = [int(d) for d in dims]
    assert 1 <= i <= len(dims)
    width = sum(dims)
    low = sum(dims[: i - 1])  # 0-based start
    n = dims[i - 1]
    rows = np.arange(n)
    return sp.csr_matrix((np.ones(n), (rows, low + rows)), shape=(n, width))


def Ec(elems: Sequence[int], N: int) -> sp.csr_matrix:
    """Clique selector (rows e_i' for i in elems).  MyMath.jl:45-52."""
    elems = np.asarray(elems, dtype=np.int64)
    assert len(elems) >= 1
    assert np.array_equal(elems, np.unique(elems))  # sorted + unique
    assert 1 <= elems[0] and elems[-1] <= N
    n = len(elems)
    return sp.csr_matrix((np.ones(n), (np.arange(n), elems - 1)), shape=(n, N))


# --------------------------------------------------------------------------
# src/MyNeuralNetwork/MyNeuralNetwork.jl:12-27  --  FeedFwdNet
# --------------------------------------------------------------------------


@dataclass
class FeedFwdNet:
    """ReLU feed-forward net.  Ms[k] = [W_k b_k], xdims[k+1] x (xdims[k]+1)."""

    xdims: List[int]
    Ms: List[np.ndarray]
    zdims: List[int] = field(default_factory=list)
    K: int = 0

    def __post_init__(self):
        self.xdims = [int(x) for x in self.xdims]
        self.Ms = [np.asarray(M, dtype=np.float64) for M in self.Ms]
        if not self.zdims:
            self.zdims = self.xdims[:-1] + [1]  # MyNeuralNetwork.jl:17
        assert len(self.xdims) >= 3  # :18
        self.K = len(self.Ms)  # :22
        assert len(self.xdims) == self.K + 1  # :23
        for k in range(self.K):  # :26
            assert self.Ms[k].shape == (self.xdims[k + 1], self.xdims[k] + 1)

    @property
    def Zdim(self) -> int:
        return sum(self.zdims)

    @property
    def acdim(self) -> int:
        return sum(self.xdims[1:-1])


def eval_feed_fwd_net(ffnet: FeedFwdNet, x: np.ndarray) -> np.ndarray:
    """MyNeuralNetwork.jl:40-46 (ReLU)."""
    xk = np.asarray(x, dtype=np.float64)
    for Mk in ffnet.Ms[:-1]:
        xk = np.maximum(Mk @ np.append(xk, 1.0), 0.0)
    return ffnet.Ms[-1] @ np.append(xk, 1.0)


def random_network(xdims: Sequence[int], sigma: float, rng: np.random.Generator) -> FeedFwdNet:
    """Utils.jl:23-27 / scripts/make_networks.jl:22-26 (numpy PCG64 stream, not Julia's)."""
    xdims = [int(x) for x in xdims]
    Ms = []
    for k in range(len(xdims) - 1):
        W = sigma * rng.standard_normal((xdims[k + 1], xdims[k]))
        b = sigma * rng.standard_normal(xdims[k + 1])
        Ms.append(np.concatenate([W, b[:, None]], axis=1))
    return FeedFwdNet(xdims=xdims, Ms=Ms)


def load_nnet(path: str) -> FeedFwdNet:
    """Minimal .nnet reader (format: exts/NNet/utils/readNNet.py:18-78;
    FeedFwdNet construction: src/MyNeuralNetwork/network_files.jl loadFromNnet)."""
    with open(path, "r") as f:
        line = f.readline()
        while line.startswith("//"):
            line = f.readline()
        rec = line.split(",")
        num_layers = int(rec[0])
        sizes = [int(t) for t in f.readline().split(",")[: num_layers + 1]]
        for _ in range(5):  # obsolete flag, mins, maxes, means, ranges
            f.readline()
        Ms = []
        for k in range(num_layers):
            nin, nout = sizes[k], sizes[k + 1]
            W = np.zeros((nout, nin))
            for i in range(nout):
                W[i, :] = [float(t) for t in f.readline().strip().split(",")[:-1]][:nin]
            b = np.zeros(nout)
            for i in range(nout):
                b[i] = float(f.readline().strip().split(",")[0])
            Ms.append(np.concatenate([W, b[:, None]], axis=1))
    return FeedFwdNet(xdims=sizes, Ms=Ms)


# --------------------------------------------------------------------------
# src/Intervals/intervals_easy.jl:2-37, intervals_auto_lirpa.jl:55-62
# --------------------------------------------------------------------------


@dataclass
class IntervalsInfo:
    """Intervals.jl:16-32.  x_intvs has K+1 (min,max) pairs, acx_intvs K-1."""

    ffnet: FeedFwdNet
    x_intvs: List[Tuple[np.ndarray, np.ndarray]]
    acx_intvs: List[Tuple[np.ndarray, np.ndarray]]

    def __post_init__(self):
        K = self.ffnet.K
        assert len(self.x_intvs) == K + 1
        assert all(len(p[0]) == len(p[1]) for p in self.x_intvs)
        assert all(self.ffnet.xdims[k] == len(self.x_intvs[k][0]) for k in range(K + 1))
        assert len(self.acx_intvs) == K - 1
        assert all(len(p[0]) == len(p[1]) for p in self.acx_intvs)
        assert all(self.ffnet.xdims[k + 1] == len(self.acx_intvs[k][0]) for k in range(K - 1))


def _ibp_layer(Mk: np.ndarray, xmin: np.ndarray, xmax: np.ndarray):
    """intervals_easy.jl:22-24: two gemv's per bound plus the bias."""
    Wk, bk = Mk[:, :-1], Mk[:, -1]
    Wp, Wn = np.maximum(Wk, 0.0), np.minimum(Wk, 0.0)
    ymin = (Wp @ xmin) + (Wn @ xmax) + bk
    ymax = (Wp @ xmax) + (Wn @ xmin) + bk
    return ymin, ymax


def intervals_worst_case(x1min, x1max, ffnet: FeedFwdNet) -> IntervalsInfo:
    """intervalsWorstCase, intervals_easy.jl:2-37 (ReLU branch)."""
    x1min = np.asarray(x1min, dtype=np.float64)
    x1max = np.asarray(x1max, dtype=np.float64)
    assert len(x1min) == len(x1max) == ffnet.xdims[0]
    x_intvs = [(x1min, x1max)]
    acx_intvs = []
    xkmin, xkmax = x1min, x1max
    for k, Mk in enumerate(ffnet.Ms, start=1):
        ykmin, ykmax = _ibp_layer(Mk, xkmin, xkmax)
        if k == ffnet.K:
            xkmin, xkmax = ykmin, ykmax
        else:
            acx_intvs.append((ykmin, ykmax))
            xkmin, xkmax = np.maximum(ykmin, 0.0), np.maximum(ykmax, 0.0)
        x_intvs.append((xkmin, xkmax))
    return IntervalsInfo(ffnet=ffnet, x_intvs=x_intvs, acx_intvs=acx_intvs)


def preact_from_x(x_intvs, ffnet: FeedFwdNet):
    """One-step pre-activation IBP given post-activation bounds.
    intervals_auto_lirpa.jl:55-62 (and :77-83)."""
    acx = []
    for k in range(ffnet.K - 1):
        xkmin, xkmax = x_intvs[k]
        ykmin, ykmax = _ibp_layer(ffnet.Ms[k], np.asarray(xkmin), np.asarray(xkmax))
        assert np.all(ykmin <= ykmax)  # :60
        acx.append((ykmin, ykmax))
    return acx


# --------------------------------------------------------------------------
# src/Qc/activ_sector.jl:63-72 -- makeSectorMinMax (ReLU)
# --------------------------------------------------------------------------

SECTOR_EPS = 1e-4  # activ_sector.jl:65


def make_sector_min_max(acxmin, acxmax):
    acxmin = np.asarray(acxmin, dtype=np.float64)
    acxmax = np.asarray(acxmax, dtype=np.float64)
    assert len(acxmin) == len(acxmax)
    smin = np.zeros(len(acxmin))
    smax = np.ones(len(acxmax))
    smin[acxmin > SECTOR_EPS] = 1.0  # :67,70
    smax[acxmax < -SECTOR_EPS] = 0.0  # :68,71
    return smin, smax


# --------------------------------------------------------------------------
# QC descriptors (src/Qc/input.jl:3-8, activ_bounded.jl:3-10, activ_sector.jl:2-20,
# output.jl:3-31)
# --------------------------------------------------------------------------


@dataclass
class QcInputBox:
    x1min: np.ndarray
    x1max: np.ndarray

    def __post_init__(self):
        self.x1min = np.asarray(self.x1min, dtype=np.float64)
        self.x1max = np.asarray(self.x1max, dtype=np.float64)
        assert len(self.x1min) == len(self.x1max)

    @property
    def vardim(self):
        return len(self.x1min)


@dataclass
class QcActivBounded:
    acydim: int
    acymin: np.ndarray
    acymax: np.ndarray

    def __post_init__(self):
        self.acymin = np.asarray(self.acymin, dtype=np.float64)
        self.acymax = np.asarray(self.acymax, dtype=np.float64)
        assert self.acydim == len(self.acymin) == len(self.acymax)
        assert np.all(self.acymin <= self.acymax)  # activ_bounded.jl:8

    @property
    def vardim(self):
        return self.acydim


def sector_lambda_dim(acxdim: int, beta: int) -> int:
    """_lambda_dim = sum((acxdim-beta):acxdim), activ_sector.jl:18."""
    return sum(range(acxdim - beta, acxdim + 1))


def sector_pairs(acxdim: int, beta: int) -> np.ndarray:
    """ijs of activ_sector.jl:29, 1-based, shape (npairs, 2), i ascending then j."""
    out = []
    for i in range(1, acxdim):
        for j in range(i + 1, min(acxdim, i + beta) + 1):
            out.append((i, j))
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)


@dataclass
class QcActivSector:
    acxdim: int
    beta: int
    smin: np.ndarray
    smax: np.ndarray
    base_smin: float = 0.0
    base_smax: float = 1.0

    def __post_init__(self):
        self.smin = np.asarray(self.smin, dtype=np.float64)
        self.smax = np.asarray(self.smax, dtype=np.float64)
        assert self.acxdim == len(self.smin) == len(self.smax)  # :11
        assert 0 <= self.beta  # :12
        assert self.base_smin <= self.base_smax  # :13
        assert np.all(self.smin <= self.smax)  # :14 (lexicographic in Julia; elementwise holds)
        assert np.all(self.base_smin <= self.smin)  # :15
        assert np.all(self.smax <= self.base_smax)  # :16

    @property
    def lam_dim(self):
        return sector_lambda_dim(self.acxdim, self.beta)

    @property
    def vardim(self):  # ReLU: :19
        return self.lam_dim + 2 * self.acxdim


@dataclass
class QcSafety:
    S: np.ndarray  # (n1 + nK1 + 1)^2 symmetric

    vardim = 0


@dataclass
class QcReachHplane:
    normal: np.ndarray
    vardim = 1


@dataclass
class QcReachCircle:
    yc: np.ndarray
    vardim = 1


@dataclass
class QcReachEllipsoid:
    invP: np.ndarray
    yc: np.ndarray
    vardim = 1


def hplaneS(normal, h, ffnet: FeedFwdNet) -> np.ndarray:
    """src/Utils/qc.jl:27-38."""
    n1, nK1 = ffnet.xdims[0], ffnet.xdims[-1]
    normal = np.asarray(normal, dtype=np.float64)
    assert len(normal) == nK1
    S = np.zeros((n1 + nK1 + 1, n1 + nK1 + 1))
    S[n1 : n1 + nK1, -1] = normal
    S[-1, n1 : n1 + nK1] = normal
    S[-1, -1] = -2.0 * h
    return S


def scaleS(S, alphas, ffnet: FeedFwdNet) -> np.ndarray:
    """src/Qc/output.jl:109-124."""
    assert len(alphas) == ffnet.K
    a = float(np.prod(alphas))
    n1, nK1 = ffnet.xdims[0], ffnet.xdims[-1]
    S = np.array(S, dtype=np.float64)
    i1, i2, i3 = slice(0, n1), slice(n1, n1 + nK1), slice(n1 + nK1, n1 + nK1 + 1)
    out = np.zeros_like(S)
    out[i1, i1] = S[i1, i1]
    out[i1, i2] = S[i1, i2] / a
    out[i1, i3] = S[i1, i3]
    out[i2, i2] = S[i2, i2] / a**2
    out[i2, i3] = S[i2, i3] / a
    out[i3, i3] = S[i3, i3]
    out[i2, i1] = out[i1, i2].T
    out[i3, i1] = out[i1, i3].T
    out[i3, i2] = out[i2, i3].T
    return out


def make_qc_activs_intvs(ffnet: FeedFwdNet, x1min, x1max, beta: int, intv_info: Optional[IntervalsInfo] = None):
    """makeQcActivsIntvs, src/Qc/activ.jl:45-67.  The reference's default interval
    method is CROWN through Python (not reproducible here); callers pass
    ``intv_info`` or get IBP (IntervalsWorstCase)."""
    if intv_info is None:
        intv_info = intervals_worst_case(x1min, x1max, ffnet)
    acdim = ffnet.acdim
    acymin = np.concatenate([p[0] for p in intv_info.x_intvs[1:-1]])  # :54
    acymax = np.concatenate([p[1] for p in intv_info.x_intvs[1:-1]])  # :55
    qc_bounded = QcActivBounded(acydim=acdim, acymin=acymin, acymax=acymax)
    sec_min = np.concatenate([p[0] for p in intv_info.acx_intvs])  # :59
    sec_max = np.concatenate([p[1] for p in intv_info.acx_intvs])  # :60
    smin, smax = make_sector_min_max(sec_min, sec_max)
    qc_sector = QcActivSector(acxdim=acdim, beta=beta, smin=smin, smax=smax)
    return [qc_bounded, qc_sector]


# --------------------------------------------------------------------------
# LITERAL FORM: makeZin / makeQ / makeA,b,B / makeZac / makeSide / makeZout
# --------------------------------------------------------------------------


def _col(v) -> sp.csr_matrix:
    return sp.csr_matrix(np.asarray(v, dtype=np.float64).reshape(-1, 1))


def makeZin(gin, qc: QcInputBox, ffnet: FeedFwdNet) -> sp.csr_matrix:
    """src/Qc/input.jl:19-42 (box branch :22-26, :36-40)."""
    gin = np.asarray(gin, dtype=np.float64)
    assert len(gin) == qc.vardim
    G = sp.diags(gin)
    P11 = -2.0 * G
    P12 = _col(G @ (qc.x1min + qc.x1max))
    P22 = sp.csr_matrix([[-2.0 * (qc.x1min @ (G @ qc.x1max))]])
    P = sp.bmat([[P11, P12], [P12.T, P22]], format="csr")
    E1 = E(1, ffnet.zdims)
    Ea = E(ffnet.K + 1, ffnet.zdims)
    Ein = sp.vstack([E1, Ea], format="csr")
    return (Ein.T @ P @ Ein).tocsr()


def makeQ_bounded(gac, qc: QcActivBounded) -> sp.csr_matrix:
    """src/Qc/activ_bounded.jl:13-23."""
    gac = np.asarray(gac, dtype=np.float64)
    assert len(gac) == qc.vardim
    n = qc.acydim
    D = sp.diags(gac)
    Z = sp.csr_matrix((n, n))
    z = sp.csr_matrix((n, 1))
    Q22 = -2.0 * D
    Q23 = _col(D @ (qc.acymin + qc.acymax))
    Q33 = sp.csr_matrix([[-2.0 * (qc.acymin @ (D @ qc.acymax))]])
    return sp.bmat([[Z, Z, z], [Z.T, Q22, Q23], [z.T, Q23.T, Q33]], format="csr")


def makeQ_sector(gac, qc: QcActivSector) -> sp.csr_matrix:
    """src/Qc/activ_sector.jl:23-60 (ReLU branch)."""
    gac = np.asarray(gac, dtype=np.float64)
    assert len(gac) == qc.vardim
    n, beta = qc.acxdim, qc.beta
    lam = gac[:n]  # :26
    if beta > 0:
        ijs = sector_pairs(n, beta)  # :29
        npairs = len(ijs)
        rows = np.repeat(np.arange(npairs), 2)
        cols = (ijs - 1).reshape(-1)
        vals = np.tile([1.0, -1.0], npairs)
        Delta = sp.csr_matrix((vals, (rows, cols)), shape=(npairs, n))  # :30-31
        assert n + npairs == qc.lam_dim  # :32
        v = gac[n : n + npairs]  # :34
        T = (Delta.T @ sp.diags(v) @ Delta).tocsr()  # :35
    else:
        T = sp.csr_matrix((n, n))
    bmin, bmax = qc.base_smin, qc.base_smax
    smin, smax = qc.smin, qc.smax
    Q11 = -2.0 * sp.diags(smin * smax * lam) - 2.0 * (bmin * bmax * T)  # :42
    Q12 = sp.diags((smin + smax) * lam) + (bmin + bmax) * T  # :43
    Q22 = -2.0 * T  # :45
    ld = qc.lam_dim
    eta = gac[ld : ld + n]  # :51-53
    nu = gac[ld + n : ld + 2 * n]  # :52,54
    Q13 = _col(-smin * eta - smax * nu)  # :55
    Q23 = _col(eta + nu)  # :56
    Q33 = sp.csr_matrix((1, 1))
    return sp.bmat([[Q11, Q12, Q13], [Q12.T, Q22, Q23], [Q13.T, Q23.T, Q33]], format="csr")


def makeA(ffnet: FeedFwdNet) -> sp.csr_matrix:
    """src/Qc/activ.jl:7-13."""
    edims = ffnet.zdims[:-1]
    fdims = edims[1:]
    A = None
    for k in range(1, ffnet.K):
        Wk = sp.csr_matrix(ffnet.Ms[k - 1][:, :-1])
        term = E(k, fdims).T @ Wk @ E(k, edims)
        A = term if A is None else A + term
    return A.tocsr()


def makeb(ffnet: FeedFwdNet) -> np.ndarray:
    """src/Qc/activ.jl:16-19."""
    return np.concatenate([M[:, -1] for M in ffnet.Ms[:-1]])


def makeB(ffnet: FeedFwdNet) -> sp.csr_matrix:
    """src/Qc/activ.jl:22-27."""
    edims = ffnet.zdims[:-1]
    fdims = edims[1:]
    B = None
    for j in range(1, ffnet.K):
        term = E(j, fdims).T @ E(j + 1, edims)
        B = term if B is None else B + term
    return B.tocsr()


def makeR(ffnet: FeedFwdNet) -> sp.csr_matrix:
    """R of src/Qc/activ.jl:33-39."""
    A = makeA(ffnet)
    b = _col(makeb(ffnet))
    B = makeB(ffnet)
    nB, mB = B.shape
    return sp.bmat(
        [[A, b], [B, sp.csr_matrix((nB, 1))], [sp.csr_matrix((1, mB)), sp.csr_matrix([[1.0]])]],
        format="csr",
    )


def makeZac(gac, qc, ffnet: FeedFwdNet, R: Optional[sp.csr_matrix] = None) -> sp.csr_matrix:
    """src/Qc/activ.jl:30-41: Zac = R' * Q * R."""
    if isinstance(qc, QcActivBounded):
        Q = makeQ_bounded(gac, qc)
    elif isinstance(qc, QcActivSector):
        Q = makeQ_sector(gac, qc)
    else:
        raise ValueError(f"unrecognized qc: {qc}")
    if R is None:
        R = makeR(ffnet)
    return (R.T @ Q @ R).tocsr()


def makeSide(ffnet: FeedFwdNet) -> sp.csr_matrix:
    """src/Qc/output.jl:34-49."""
    xd, K = ffnet.xdims, ffnet.K
    WK = sp.csr_matrix(ffnet.Ms[K - 1][:, :-1])
    bK = _col(ffnet.Ms[K - 1][:, -1])
    n1, nK, nK1 = xd[0], xd[K - 1], xd[K]
    return sp.bmat(
        [
            [sp.identity(n1, format="csr"), sp.csr_matrix((n1, nK)), sp.csr_matrix((n1, 1))],
            [sp.csr_matrix((nK1, n1)), WK, bK],
            [sp.csr_matrix((1, n1)), sp.csr_matrix((1, nK)), sp.csr_matrix([[1.0]])],
        ],
        format="csr",
    )


def _Eout(ffnet: FeedFwdNet) -> sp.csr_matrix:
    zd, K = ffnet.zdims, ffnet.K
    return sp.vstack([E(1, zd), E(K, zd), E(K + 1, zd)], format="csr")


def reach_S(gout, qc, ffnet: FeedFwdNet) -> np.ndarray:
    """The S of src/Qc/output.jl:64-98 for reach QCs (dense, small)."""
    gout = np.atleast_1d(np.asarray(gout, dtype=np.float64))
    assert len(gout) == 1
    n1, nK1 = ffnet.xdims[0], ffnet.xdims[-1]
    S = np.zeros((n1 + nK1 + 1, n1 + nK1 + 1))
    i2 = slice(n1, n1 + nK1)
    if isinstance(qc, QcReachHplane):
        normal = np.asarray(qc.normal, dtype=np.float64)
        assert len(normal) == nK1
        S23 = normal  # :74
        S33 = -2.0 * gout[0]  # :75
    elif isinstance(qc, QcReachCircle):
        yc = np.asarray(qc.yc, dtype=np.float64)
        assert len(yc) == nK1
        S[i2, i2] = np.eye(nK1)  # :82
        S23 = -yc  # :83
        S33 = yc @ yc - gout[0]  # :84
    elif isinstance(qc, QcReachEllipsoid):
        yc = np.asarray(qc.yc, dtype=np.float64)
        invP = np.asarray(qc.invP, dtype=np.float64)
        assert len(yc) == nK1
        S[i2, i2] = invP.T @ invP  # :91
        S23 = -invP.T @ yc  # :92
        S33 = yc @ yc - gout[0]  # :93 (reference quirk: not yc' invP' invP yc)
    else:
        raise ValueError(f"unrecognized qc: {qc}")
    S[i2, -1] = S23
    S[-1, i2] = S23
    S[-1, -1] = S33
    return S


def makeZout(qc, ffnet: FeedFwdNet, gout=None) -> sp.csr_matrix:
    """src/Qc/output.jl:52-61 (safety) and :64-106 (reach)."""
    if isinstance(qc, QcSafety):
        S = np.asarray(qc.S, dtype=np.float64)
    else:
        S = reach_S(gout, qc, ffnet)
    Eo = _Eout(ffnet)
    R = makeSide(ffnet)
    return (Eo.T @ R.T @ sp.csr_matrix(S) @ R @ Eo).tocsr()


def assemble_Z_literal(ffnet, qc_input, qc_out, qc_activs, gin, gacs, gout=None) -> np.ndarray:
    """Z = Zin + Zout + sum(Zacs), src/Methods/chordal_sdp.jl:114,145 with numeric gamma
    (the recomputation at scripts/test_acas.jl:81-85).  Returns dense Zdim x Zdim."""
    R = makeR(ffnet)
    Z = makeZin(gin, qc_input, ffnet) + makeZout(qc_out, ffnet, gout)
    for g, qc in zip(gacs, qc_activs):
        Z = Z + makeZac(g, qc, ffnet, R=R)
    return np.asarray(Z.todense())


# --------------------------------------------------------------------------
# src/Methods/chordal_cliques.jl:13-59 -- makeCliques
# --------------------------------------------------------------------------


def make_cliques(ffnet: FeedFwdNet, beta: int):
    """Returns a list of (Ck, [Ck1, Ck2] or [Cp], [Dk1(, Dk2)]) of 1-based int64 arrays.
    ``beta`` is the sector QC's beta (0 when no sector QC is used, :18-19)."""
    xd, K = ffnet.xdims, ffnet.K

    def S(k):
        return 0 if k == 0 else sum(xd[:k])

    p = 1
    for i in range(1, K + 1):  # :22-27
        if S(i + 1) + beta >= S(K - 1):
            p = i
            break
    cliques = []
    for k in range(1, p):  # :31
        Ck1 = np.arange(S(k - 1) + 1, S(k + 1) + beta + 1, dtype=np.int64)  # :33
        Ck2 = np.arange(S(K - 1) + 1, S(K) + 1 + 1, dtype=np.int64)  # :34
        assert Ck1[-1] <= Ck2[0]  # :35
        Ck = np.concatenate([Ck1, Ck2])
        Ckdim = len(Ck)
        if k == 1:  # :40-42
            Dk1 = np.arange(1, Ckdim + 1, dtype=np.int64)
            cliques.append((Ck, [Ck1, Ck2], [Dk1]))
        else:  # :45-51
            nk, nk1 = ffnet.zdims[k - 1], ffnet.zdims[k]
            Dk1 = np.concatenate([np.arange(1, nk + nk1 + beta + 1, dtype=np.int64), np.array([Ckdim], dtype=np.int64)])
            Dk2 = np.arange(nk + nk1 + 1, Ckdim + 1, dtype=np.int64)
            cliques.append((Ck, [Ck1, Ck2], [Dk1, Dk2]))
    Cp = np.arange(S(p - 1) + 1, S(K) + 1 + 1, dtype=np.int64)  # :55
    Dp1 = np.arange(1, len(Cp) + 1, dtype=np.int64)
    cliques.append((Cp, [Cp], [Dp1]))
    return cliques


def clique_blocks(Z: np.ndarray, cliques) -> List[np.ndarray]:
    """Z[Ck, Ck] = Ec(Ck) * Z * Ec(Ck)' for every clique (what each PSD variable Zk of
    src/Methods/chordal_sdp.jl:19-57 must reproduce when summed by setupZksum!)."""
    out = []
    for Ck, _, _ in cliques:
        idx = Ck - 1
        out.append(np.ascontiguousarray(Z[np.ix_(idx, idx)]))
    return out


def zksum_pattern(ffnet: FeedFwdNet, cliques) -> np.ndarray:
    """Boolean Zdim x Zdim: where setupZksum! (src/Methods/chordal_sdp.jl:60-93) can place a
    clique variable.  Follows the slice+inject code, including the trailing-block rule
    for the last clique (:71-72)."""
    Zdim = ffnet.Zdim
    pat = np.zeros((Zdim, Zdim), dtype=bool)
    for k, (Ck, parts, _) in enumerate(cliques):
        n = len(Ck)
        if k == len(cliques) - 1:
            pat[Zdim - n :, Zdim - n :] = True
        else:
            assert len(parts) == 2
            c1, c2 = parts[0] - 1, parts[1] - 1
            pat[np.ix_(c1, c1)] = True
            pat[np.ix_(c1, c2)] = True
            pat[np.ix_(c2, c1)] = True
            pat[np.ix_(c2, c2)] = True
    return pat


def structural_pattern_notebook(xdims: Sequence[int], beta: int) -> np.ndarray:
    """quickRawZ of experiments/plot_sparsity.ipynb cell 5: quickEM(beta) | quickE1K | quickEa."""
    xdims = [int(x) for x in xdims]
    K = len(xdims) - 1
    N = sum(xdims[:-1])

    def S(k):
        return sum(xdims[:k])

    pat = np.zeros((N + 1, N + 1), dtype=bool)
    ii = np.arange(1, N + 1)
    for k in range(1, K):  # quickEM
        m = (S(k - 1) + 1 <= ii) & (ii <= S(k + 1) + beta)
        pat[:N, :N] |= np.outer(m, m)
    a = (1 <= ii) & (ii <= xdims[0])  # quickE1K
    b = (S(K - 1) + 1 <= ii) & (ii <= N)
    pat[:N, :N] |= np.outer(a, b) | np.outer(b, a)
    pat[N, :] = True  # quickEa
    pat[:, N] = True
    return pat


def chordal_extension_pattern_notebook(xdims: Sequence[int], beta: int) -> np.ndarray:
    """quickRawZ(beta) overlaid with quickEK of experiments/plot_sparsity.ipynb cells 5, 16-22 (the figures
    Zbeta*.png): the x_K rows and columns are filled in.  The union of Ck x Ck over makeCliques equals it."""
    xdims = [int(x) for x in xdims]
    K = len(xdims) - 1
    N = sum(xdims[:-1])
    pat = structural_pattern_notebook(xdims, beta)
    lastblk = np.arange(1, N + 1) >= sum(xdims[:K - 1]) + 1   # quickEK
    pat[:N, :N] |= lastblk[:, None] | lastblk[None, :]
    return pat


# --------------------------------------------------------------------------
# CLOSED FORM, vectorised numpy
# --------------------------------------------------------------------------


@dataclass
class SectorSplit:
    lam: np.ndarray
    v: np.ndarray
    eta: np.ndarray
    nu: np.ndarray


def split_sector_gamma(gsec, acdim: int, beta: int) -> SectorSplit:
    """Layout of the sector multiplier vector, activ_sector.jl:26,34,51-54."""
    gsec = np.asarray(gsec, dtype=np.float64)
    ld = sector_lambda_dim(acdim, beta)
    assert len(gsec) == ld + 2 * acdim
    return SectorSplit(lam=gsec[:acdim], v=gsec[acdim:ld], eta=gsec[ld : ld + acdim], nu=gsec[ld + acdim :])


def band_T(v: np.ndarray, acdim: int, beta: int) -> np.ndarray:
    """T = sum v_ij (e_i-e_j)(e_i-e_j)' in band storage: Tb[t, i] = T[i, i+t], t=0..beta."""
    Tb = np.zeros((beta + 1, acdim))
    if beta == 0:
        return Tb
    ijs = sector_pairs(acdim, beta) - 1
    i, j = ijs[:, 0], ijs[:, 1]
    np.add.at(Tb[0], i, v)
    np.add.at(Tb[0], j, v)
    Tb[j - i, i] = -v
    return Tb


def out_S_blocks(qc_out, ffnet: FeedFwdNet, gout=None):
    n1, nK1 = ffnet.xdims[0], ffnet.xdims[-1]
    S = np.asarray(qc_out.S, dtype=np.float64) if isinstance(qc_out, QcSafety) else reach_S(gout, qc_out, ffnet)
    i1, i2 = slice(0, n1), slice(n1, n1 + nK1)
    return S[i1, i1], S[i1, i2], S[i1, -1], S[i2, i2], S[i2, -1], S[-1, -1]


def assemble_Z_closed_form(ffnet, qc_input, qc_out, qc_bounded, qc_sector, gin, gbnd, gsec, gout=None) -> np.ndarray:
    """Dense Z via per-variable coefficients; independent of the literal path above."""
    xd, K = ffnet.xdims, ffnet.K
    n1 = xd[0]
    Zdim, acdim = ffnet.Zdim, ffnet.acdim
    a = Zdim - 1
    off = np.concatenate([[0], np.cumsum(xd[:-1])])  # 0-based start of block b (b = 0..K-1), off[K] = a
    Z = np.zeros((Zdim, Zdim))
    gin = np.asarray(gin, dtype=np.float64)
    gbnd = np.asarray(gbnd, dtype=np.float64)

    # Zin (input.jl:22-26)
    i1 = np.arange(n1)
    Z[i1, i1] += -2.0 * gin
    t = gin * (qc_input.x1min + qc_input.x1max)
    Z[i1, a] += t
    Z[a, i1] += t
    Z[a, a] += -2.0 * np.sum(gin * qc_input.x1min * qc_input.x1max)

    # bounded QC (activ_bounded.jl:19-21 through R'QR)
    ie = n1 + np.arange(acdim)
    Z[ie, ie] += -2.0 * gbnd
    t = gbnd * (qc_bounded.acymin + qc_bounded.acymax)
    Z[ie, a] += t
    Z[a, ie] += t
    Z[a, a] += -2.0 * np.sum(gbnd * qc_bounded.acymin * qc_bounded.acymax)

    # sector QC
    beta = qc_sector.beta
    s = split_sector_gamma(gsec, acdim, beta)
    p = qc_sector.smin * qc_sector.smax
    q = qc_sector.smin + qc_sector.smax
    d11 = -2.0 * p * s.lam
    c13 = -qc_sector.smin * s.eta - qc_sector.smax * s.nu
    c23 = s.eta + s.nu
    Tb = band_T(s.v, acdim, beta)
    # dense banded M = diag(q lam) + T and T (acdim x acdim) only for modest sizes; block-wise otherwise
    bias = makeb(ffnet)
    # Abar rows: rho_j.  Work per layer.
    nstart = np.concatenate([[0], np.cumsum(xd[1:-1])])  # neuron offset of layer L=k+1 (k=1..K-1) -> nstart[k-1]
    # Gram + affine parts of Abar' diag(d11) Abar and Abar' c13
    for k in range(1, K):  # W_k maps block k -> layer k+1
        W = ffnet.Ms[k - 1][:, :-1]
        bk = ffnet.Ms[k - 1][:, -1]
        js = slice(nstart[k - 1], nstart[k - 1] + xd[k])
        rb = slice(off[k - 1], off[k - 1] + xd[k - 1])
        dj = d11[js]
        if np.any(dj != 0.0):
            Z[rb, rb] += W.T @ (dj[:, None] * W)
        u = dj * bk + c13[js]
        t = W.T @ u
        Z[rb, a] += t
        Z[a, rb] += t
        Z[a, a] += np.sum(dj * bk * bk) + 2.0 * np.sum(bk * c13[js])
    # Abar' M Bbar + sym, Bbar'(-2T)Bbar, c23
    Z[ie, a] += c23
    Z[a, ie] += c23
    # band loops over offsets t = -beta..beta : M[j, c] with c = j + t
    qlam = q * s.lam
    layer_of = np.concatenate([np.full(xd[k], k) for k in range(1, K)])  # neuron j is row of W_k, k=layer_of[j]
    local_of = np.concatenate([np.arange(xd[k]) for k in range(1, K)])
    for t in range(-beta, beta + 1):
        if t >= 0:
            j = np.arange(0, acdim - t)
            m = Tb[t, j].copy()
        else:
            j = np.arange(-t, acdim)
            m = Tb[-t, j + t].copy()
        c = j + t
        if t == 0:
            Tdiag = m.copy()
            m = m + qlam
        # -2T at (eps_j, eps_c)
        tt = Tb[abs(t), np.minimum(j, c)]
        Z[n1 + j, n1 + c] += -2.0 * tt
        # affine row: Z[a, eps_c] += b_j M[j,c]  (and symmetric)
        np.add.at(Z[a], n1 + c, bias[j] * m)
        np.add.at(Z[:, a], n1 + c, bias[j] * m)
        # rho_j eps_c' * M[j,c]: for each layer k, rows of block k get W_k[jl, :] * m
        for k in range(1, K):
            sel = layer_of[j] == k
            if not np.any(sel):
                continue
            jj, cc, mm = j[sel], c[sel], m[sel]
            W = ffnet.Ms[k - 1][:, :-1]
            rb = slice(off[k - 1], off[k - 1] + xd[k - 1])
            contrib = (W[local_of[jj], :] * mm[:, None]).T  # n_k x len(jj)
            # columns n1 + cc are distinct within one t
            Z[rb, n1 + cc] += contrib
            Z[n1 + cc, rb] += contrib.T

    # Zout (output.jl:52-106): Eout' R' S R Eout
    S11, S12, S13, S22, S23, S33 = out_S_blocks(qc_out, ffnet, gout)
    WK = ffnet.Ms[K - 1][:, :-1]
    bK = ffnet.Ms[K - 1][:, -1]
    rK = slice(off[K - 1], off[K - 1] + xd[K - 1])
    r1 = slice(0, n1)
    Z[r1, r1] += S11
    t = S12 @ WK
    Z[r1, rK] += t
    Z[rK, r1] += t.T
    t = S12 @ bK + S13
    Z[r1, a] += t
    Z[a, r1] += t
    Z[rK, rK] += WK.T @ S22 @ WK
    t = WK.T @ (S22 @ bK + S23)
    Z[rK, a] += t
    Z[a, rK] += t
    Z[a, a] += bK @ S22 @ bK + 2.0 * (bK @ S23) + S33
    return Z


# --------------------------------------------------------------------------
# Query-level helpers used by tests and bench (numeric-gamma "query")
# --------------------------------------------------------------------------


@dataclass
class NumericQuery:
    """One numeric-gamma query: everything the hot path consumes for one (box, output spec)."""

    x1min: np.ndarray
    x1max: np.ndarray
    gin: np.ndarray
    gbnd: np.ndarray
    gsec: np.ndarray
    qc_out: object
    gout: Optional[np.ndarray] = None


def run_query(ffnet: FeedFwdNet, beta: int, query: NumericQuery, form: str = "closed", intv_info=None):
    """bounds -> QCs -> Z -> clique blocks for one query.  Returns dict."""
    if intv_info is None:
        intv_info = intervals_worst_case(query.x1min, query.x1max, ffnet)
    qc_bounded, qc_sector = make_qc_activs_intvs(ffnet, query.x1min, query.x1max, beta, intv_info)
    qc_input = QcInputBox(query.x1min, query.x1max)
    if form == "literal":
        Z = assemble_Z_literal(
            ffnet, qc_input, query.qc_out, [qc_bounded, qc_sector], query.gin, [query.gbnd, query.gsec], query.gout
        )
    else:
        Z = assemble_Z_closed_form(
            ffnet, qc_input, query.qc_out, qc_bounded, qc_sector, query.gin, query.gbnd, query.gsec, query.gout
        )
    cliques = make_cliques(ffnet, beta)
    return {
        "intv": intv_info,
        "qc_bounded": qc_bounded,
        "qc_sector": qc_sector,
        "Z": Z,
        "cliques": cliques,
        "blocks": clique_blocks(Z, cliques),
    }


# --------------------------------------------------------------------------
# vnnlib -> CNF of (QcInputBox, QcSafety) pairs: exts/vnnlib_parser.jl:3-216 (read_statements,
# update_rv_tuple!, read_vnnlib_simple) and experiments/vnnlib_utils.jl:18-56 (loadVnnlibCnf).
# PARITY UNPINNED (Julia; the ACAS property files bench/acas/*.vnnlib are not shipped).  One deliberate
# difference: the reference merges alternatives with equal boxes in a Dict and iterates its values (hash order);
# here boxes keep the order of first appearance.
# --------------------------------------------------------------------------
import re as _re


def vnnlib_statements(path: str) -> List[str]:
    lines = [ln.strip() for ln in open(path).read().split("\n")]
    assert len(lines) > 0
    depth, out, cur = 0, [], ""
    for line in lines:
        if ";" in line:
            line = line[: line.index(";")].strip()
        if not line:
            continue
        depth += line.count("(") - line.count(")")
        assert depth >= 0
        cur += ("" if not cur else " ") + line
        if depth == 0:
            out.append(cur)
            cur = ""
    if cur:
        out.append(cur)
    out = [" ".join(s.split()) for s in out]
    return [s.replace("( ", "(").replace(") ", ")") for s in out]


def _vnnlib_update(rv, op, first, second, n_in, n_out):
    box, mat, rhs = rv
    if first.startswith("X_"):
        idx = int(first[2:])
        assert not second.startswith("X") and not second.startswith("Y")
        assert 0 <= idx < n_in
        if op == "<=":
            box[idx][1] = min(float(second), box[idx][1])
        else:
            box[idx][0] = max(float(second), box[idx][0])
        assert box[idx][0] <= box[idx][1]
        return
    if op == ">=":
        first, second = second, first
    row, b = np.zeros(n_out), 0.0
    if first.startswith("Y_") and second.startswith("Y_"):
        row[int(first[2:])] = 1
        row[int(second[2:])] = -1
    elif first.startswith("Y_"):
        row[int(first[2:])] = 1
        b = float(second)
    else:
        assert second.startswith("Y_")
        row[int(second[2:])] = -1
        b = -1 * float(first)
    mat.append(row)
    rhs.append(b)


def read_vnnlib_simple(path: str, n_in: int, n_out: int):
    """[(box, [(mat, rhs), ...]), ...] with box = [[lo, hi]] * n_in."""
    import copy

    simple = _re.compile(r"^\(assert \((<=|>=) (\S+) (\S+)\)\)$")
    declare = _re.compile(r"^\(declare-const (X|Y)_(\S+) Real\)$")
    rv = [([[-np.inf, np.inf] for _ in range(n_in)], [], [])]
    for line in vnnlib_statements(path):
        if declare.match(line):
            continue
        m = simple.match(line)
        if m:
            for t in rv:
                _vnnlib_update(t, *m.groups(), n_in, n_out)
            continue
        tokens = line.replace("(", " ").replace(")", " ").split()[2:]   # skip 'assert' and 'or'
        conjuncts = " ".join(tokens).split("and")[1:]
        old, rv = rv, []
        for t in old:
            for c in conjuncts:
                cp = copy.deepcopy(t)
                rv.append(cp)
                ct = c.split()
                for i in range(len(ct) // 3):
                    _vnnlib_update(cp, ct[3 * i], ct[3 * i + 1], ct[3 * i + 2], n_in, n_out)
    merged = {}
    for box, mat, rhs in rv:
        key = str(box)
        merged.setdefault(key, (box, []))[1].append((mat, rhs))
    final = []
    for box, specs in merged.values():          # insertion order (the reference: Dict hash order)
        assert all(np.isfinite(r[0]) and np.isfinite(r[1]) for r in box)
        final.append((box, specs))
    return final


def load_vnnlib_cnf(path: str, ffnet: FeedFwdNet):
    """CnfSpec: list of disjunctive clauses, each a list of (QcInputBox, QcSafety)."""
    n_in, n_out = ffnet.xdims[0], ffnet.xdims[-1]
    cnf = []
    for box, specs in read_vnnlib_simple(path, n_in, n_out):
        qc_in = QcInputBox(np.array([b[0] for b in box]), np.array([b[1] for b in box]))
        for mat, rhs in specs:
            A = np.stack(mat)
            clause = []
            for i in range(len(rhs)):
                eps = 1e-4
                clause.append((qc_in, QcSafety(S=hplaneS(-A[i, :], -rhs[i] - eps, ffnet))))
            cnf.append(clause)
    return cnf


# --------------------------------------------------------------------------
# Affine structure of Z in the multipliers (what JuMP holds as AffExpr entries,
# src/Methods/chordal_sdp.jl:96-153): Z(gamma) = Z0 + sum_v gamma_v Z_v
# --------------------------------------------------------------------------


def affine_structure(ffnet: FeedFwdNet, beta: int, x1min, x1max, qc_out, intv_info=None):
    """Z0 and the per-variable coefficient matrices Z_v, obtained by evaluating the LITERAL assembly at
    gamma = 0 and at the unit vectors (Z is affine, so Z_v = Z(e_v) - Z(0) exactly up to rounding).
    Variable order = creation order of setupSafety!/setupReach!: [gamma_in; (gamma_out); gamma_ac1; gamma_ac2].
    Small nets only (nvar literal assemblies)."""
    if intv_info is None:
        intv_info = intervals_worst_case(x1min, x1max, ffnet)
    qcs = make_qc_activs_intvs(ffnet, x1min, x1max, beta, intv_info)
    qc_in = QcInputBox(x1min, x1max)
    n1, ac = ffnet.xdims[0], ffnet.acdim
    nsec = qcs[1].vardim
    has_out = not isinstance(qc_out, QcSafety)
    nvar = n1 + (1 if has_out else 0) + ac + nsec

    def Z(g):
        o = 0
        gin = g[o:o + n1]; o += n1
        gout = g[o:o + 1] if has_out else None
        o += 1 if has_out else 0
        gb = g[o:o + ac]; o += ac
        gs = g[o:o + nsec]
        return assemble_Z_literal(ffnet, qc_in, qc_out, qcs, gin, [gb, gs], gout)

    Z0 = Z(np.zeros(nvar))
    Zv = []
    for v in range(nvar):
        e_v = np.zeros(nvar)
        e_v[v] = 1.0
        Zv.append(Z(e_v) - Z0)
    return Z0, Zv


def cover_upper_entries(ffnet: FeedFwdNet, cliques):
    """1-based (row, col) of the upper triangle of the clique cover, column-major."""
    Zdim = ffnet.Zdim
    cover = np.zeros((Zdim, Zdim), dtype=bool)
    for Ck, _, _ in cliques:
        cover[np.ix_(Ck - 1, Ck - 1)] = True
    rows, cols = [], []
    for c in range(Zdim):
        r = np.nonzero(cover[: c + 1, c])[0]
        rows.append(r + 1)
        cols.append(np.full(len(r), c + 1))
    return np.concatenate(rows), np.concatenate(cols)


# --------------------------------------------------------------------------
# CROWN bounds: the reference's DEFAULT interval method (IntervalsAutoLirpa, Intervals.jl:38,44-45 ->
# intervalsAutoLirpaSliced, intervals_auto_lirpa.jl:44-63 -> exts/auto_lirpa_bridge.py:97-112 ->
# auto_LiRPA BoundedModule.compute_bounds(method="CROWN")).
#
# PINNED against the reference's own dependency: the vendored auto_LiRPA (2021) is imported from
# /root/reference/exts (with four name shims for numpy 2 / Python 3.12 / torch 2.11, no numerics) and driven as
# intervalsAutoLirpaSliced drives it by tests/golden/make_crown_golden.py; this restatement equals its float64
# run to 1e-15 and its float32 run (what the reference executes) to float32 rounding on the reference's shipped
# W10-D10 / W5-D5 nets and two seeded random nets (tests/golden/crown_autolirpa.npz,
# tests/test_oracle_cpu.py::test_crown_restatement_equals_the_vendored_auto_lirpa).  It is a float64
# restatement of the algorithm for ReLU MLPs as read from the vendored sources:
#   * bound_general.py:1212-1366  every pre-activation node gets bounds by backward LiRPA with C = I,
#     except the first linear layer, which is bounded by interval arithmetic (:1257-1262);
#   * operators/activation.py:306-323  ReLU relaxation from (l, u): lb_r = min(l, 0), ub_r = max(u, 0),
#     ub_r = max(ub_r, lb_r + 1e-8), upper_d = ub_r / (ub_r - lb_r), upper_b = -lb_r * upper_d;
#     :387-388 "adaptive" lower slope lower_d = (upper_d > 0.5);
#     :440-455  uA <- uA+ * upper_d + uA- * lower_d, ubias += uA+ . upper_b ; lA <- lA+ * lower_d + lA- * upper_d,
#               lbias += lA- . upper_b;
#   * linear layer: A <- A W, bias += A b;  concretisation on the box: A c -/+ |A| r + bias.
# intervalsAutoLirpaSliced bounds x_{k+1} as the OUTPUT of the k-layer prefix followed by an identity layer,
# i.e. through the relaxation of relu_k (so a lower bound may be negative), then lb = min(lb, ub),
# ub = max(lb, ub) (intervals_auto_lirpa.jl:37-39), and finally one IBP step for acx_intvs (:55-62).
# --------------------------------------------------------------------------


def _relu_relaxation(l, u):
    lb_r = np.minimum(l, 0.0)
    ub_r = np.maximum(u, 0.0)
    ub_r = np.maximum(ub_r, lb_r + 1e-8)
    upper_d = ub_r / (ub_r - lb_r)
    upper_b = -lb_r * upper_d
    lower_d = (upper_d > 0.5).astype(np.float64)
    return upper_d, upper_b, lower_d


def _crown_backward(ffnet: FeedFwdNet, pre, lA, uA, lb, ub, j_start, x1min, x1max):
    """Propagate (lA, lb), (uA, ub), linear in x_{j_start+1} (the output of relu_{j_start}), back to the input.
    pre[j] = (l, u) bounds of y_j, j = 1..; j_start = 0 means the functions are already linear in x_1."""
    for j in range(j_start, 0, -1):
        d_u, b_u, d_l = _relu_relaxation(*pre[j])
        up, un = np.maximum(uA, 0.0), np.minimum(uA, 0.0)
        lp, ln = np.maximum(lA, 0.0), np.minimum(lA, 0.0)
        ub = ub + up @ b_u
        lb = lb + ln @ b_u
        uA = up * d_u + un * d_l
        lA = lp * d_l + ln * d_u
        W, b = ffnet.Ms[j - 1][:, :-1], ffnet.Ms[j - 1][:, -1]
        ub = ub + uA @ b
        lb = lb + lA @ b
        uA = uA @ W
        lA = lA @ W
    c, r = 0.5 * (x1min + x1max), 0.5 * (x1max - x1min)
    return lA @ c - np.abs(lA) @ r + lb, uA @ c + np.abs(uA) @ r + ub


def intervals_crown(x1min, x1max, ffnet: FeedFwdNet) -> IntervalsInfo:
    x1min = np.asarray(x1min, dtype=np.float64)
    x1max = np.asarray(x1max, dtype=np.float64)
    K = ffnet.K
    pre = {}  # pre[j] = bounds of y_j = W_j x_j + b_j, j = 1..K
    pre[1] = _ibp_layer(ffnet.Ms[0], x1min, x1max)
    for i in range(2, K + 1):
        W, b = ffnet.Ms[i - 1][:, :-1], ffnet.Ms[i - 1][:, -1]
        pre[i] = _crown_backward(ffnet, pre, W.copy(), W.copy(), b.copy(), b.copy(), i - 1, x1min, x1max)
    x_intvs = [(x1min, x1max)]
    for k in range(1, K):  # slice k: identity o relu_k o (layers 1..k)
        n = ffnet.xdims[k]
        eye = np.eye(n)
        lo, hi = _crown_backward(ffnet, pre, eye.copy(), eye.copy(), np.zeros(n), np.zeros(n), k, x1min, x1max)
        lo = np.minimum(lo, hi)  # intervals_auto_lirpa.jl:37-38
        hi = np.maximum(lo, hi)
        x_intvs.append((lo, hi))
    lo, hi = pre[K]
    lo = np.minimum(lo, hi)
    hi = np.maximum(lo, hi)
    x_intvs.append((lo, hi))
    acx = preact_from_x(x_intvs, ffnet)  # :55-62
    return IntervalsInfo(ffnet=ffnet, x_intvs=x_intvs, acx_intvs=acx)
