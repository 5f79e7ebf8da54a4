"""Host-side mirror of the reference's Julia interface for the constraint-construction path.

Same names, argument meaning and error behaviour as the reference (AntonXue/nn-sdp, Julia), so
the parity tests read like tests of the reference; every function is a thin wrapper that calls
the CUDA library through the C ABI.  Julia is not available in this environment, so this Python
layer stands where the Julia wrapper (nn-sdp_b200/julia/NnSdpB200.jl) would.

Reference                                                  here
---------------------------------------------------------  ------------------------------------
FeedFwdNet            src/MyNeuralNetwork/MyNeuralNetwork.jl:12-27   FeedFwdNet
IntervalsInfo         src/Intervals/Intervals.jl:16-32               IntervalsInfo
makeIntervalsInfo     src/Intervals/Intervals.jl:38-49               makeIntervalsInfo(..., IntervalsB200())
makeSectorMinMax      src/Qc/activ_sector.jl:63-72                   makeSectorMinMax
QcInputBox / makeZin  src/Qc/input.jl:3-8,19-42                      QcInputBox / makeZin
QcActivBounded/Sector src/Qc/activ_bounded.jl, activ_sector.jl       QcActivBounded / QcActivSector
makeQcActivs, makeZac src/Qc/activ.jl:30-72                          makeQcActivs / makeZac
QcSafety, QcReach*    src/Qc/output.jl:3-31                          QcSafety, QcReachHplane, ...
makeZout              src/Qc/output.jl:52-106                        makeZout
hplaneS               src/Utils/qc.jl:27-38                          hplaneS
makeCliques           src/Methods/chordal_cliques.jl:13-59           makeCliques
SafetyQuery/ReachQuery src/Methods/Methods.jl:19-41                  SafetyQuery / ReachQuery
Z = Zin+Zout+sum(Zacs) src/Methods/chordal_sdp.jl:114,145            assembleZ / assembleCliqueBlocks
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from . import core

_default_ctx: Optional[core.Context] = None


def default_context() -> core.Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = core.Context([0])
    return _default_ctx


def set_default_context(ctx: core.Context):
    global _default_ctx
    _default_ctx = ctx


@dataclass
class FeedFwdNet:
    """ReLU feed-forward net, Ms[k] = [W_k b_k] of size xdims[k+1] x (xdims[k]+1)."""

    xdims: List[int]
    Ms: List[np.ndarray]
    zdims: List[int] = field(default_factory=list)
    K: int = 0
    _net: Optional[core.Net] = field(default=None, repr=False, compare=False)

    def __post_init__(self):
        self.xdims = [int(x) for x in self.xdims]
        self.Ms = [np.asarray(M, dtype=np.float64) for M in self.Ms]
        if not self.zdims:
            self.zdims = self.xdims[:-1] + [1]
        assert len(self.xdims) >= 3
        self.K = len(self.Ms)
        assert len(self.xdims) == self.K + 1
        for k in range(self.K):
            assert self.Ms[k].shape == (self.xdims[k + 1], self.xdims[k] + 1)

    def device(self) -> core.Net:
        """The uploaded copy (weights are uploaded once and stay resident)."""
        if self._net is None or self._net.ctx is not default_context():
            self._net = core.Net(default_context(), self.xdims, self.Ms)
        return self._net

    @property
    def Zdim(self):
        return sum(self.zdims)

    @property
    def acdim(self):
        return sum(self.xdims[1:-1])


# ---- Intervals ---------------------------------------------------------------------------------
class IntervalsMethod:
    pass


class IntervalsB200(IntervalsMethod):
    """IntervalsWorstCase (IBP) evaluated on the GPU."""


class IntervalsCrownB200(IntervalsMethod):
    """IntervalsAutoLirpa (CROWN, sliced variant: the reference's default) evaluated on the GPU in FP64."""


@dataclass
class IntervalsInfo:
    ffnet: FeedFwdNet
    x_intvs: List[Tuple[np.ndarray, np.ndarray]]
    acx_intvs: List[Tuple[np.ndarray, np.ndarray]]

    def __post_init__(self):
        K = self.ffnet.K
        assert len(self.x_intvs) == K + 1
        assert all(self.ffnet.xdims[k] == len(self.x_intvs[k][0]) == len(self.x_intvs[k][1]) for k in range(K + 1))
        assert len(self.acx_intvs) == K - 1
        assert all(self.ffnet.xdims[k + 1] == len(self.acx_intvs[k][0]) == len(self.acx_intvs[k][1]) for k in range(K - 1))


def _split(v: np.ndarray, dims: Sequence[int]):
    out, o = [], 0
    for d in dims:
        out.append(v[o:o + d].copy())
        o += d
    return out


def makeIntervalsInfo(x1min, x1max, ffnet: FeedFwdNet, method: IntervalsMethod = None) -> IntervalsInfo:
    method = method or IntervalsB200()
    if not isinstance(method, (IntervalsB200, IntervalsCrownB200)):
        raise ValueError(f"unrecognized method: {method}")
    x1min = np.asarray(x1min, dtype=np.float64)
    x1max = np.asarray(x1max, dtype=np.float64)
    assert len(x1min) == len(x1max) == ffnet.xdims[0]
    fn = core.bounds_crown if isinstance(method, IntervalsCrownB200) else core.bounds_ibp
    r = fn(ffnet.device(), x1min[None, :], x1max[None, :])
    xs_min, xs_max = _split(r["xmin"][0], ffnet.xdims), _split(r["xmax"][0], ffnet.xdims)
    ac_min, ac_max = _split(r["acxmin"][0], ffnet.xdims[1:-1]), _split(r["acxmax"][0], ffnet.xdims[1:-1])
    return IntervalsInfo(ffnet=ffnet, x_intvs=list(zip(xs_min, xs_max)), acx_intvs=list(zip(ac_min, ac_max)))


def makeSectorMinMax(acxmin, acxmax):
    acxmin = np.asarray(acxmin, dtype=np.float64)
    acxmax = np.asarray(acxmax, dtype=np.float64)
    assert len(acxmin) == len(acxmax)
    return core.sector_minmax(default_context(), acxmin, acxmax)


# ---- QCs ---------------------------------------------------------------------------------------
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
        assert np.all(self.acymin <= self.acymax)

    @property
    def vardim(self):
        return self.acydim


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
        assert self.acxdim == len(self.smin) == len(self.smax)
        assert 0 <= self.beta
        assert (self.base_smin, self.base_smax) == (0.0, 1.0), "ReLU sector: base_smin=0, base_smax=1 (activ.jl:62)"
        assert np.all(self.smin <= self.smax)
        assert np.all(self.base_smin <= self.smin) and np.all(self.smax <= self.base_smax)

    @property
    def lam_dim(self):
        return sum(range(self.acxdim - self.beta, self.acxdim + 1))

    @property
    def vardim(self):
        return self.lam_dim + 2 * self.acxdim


@dataclass
class QcSafety:
    S: np.ndarray
    vardim: int = 0


@dataclass
class QcReachHplane:
    normal: np.ndarray
    vardim: int = 1


@dataclass
class QcReachCircle:
    yc: np.ndarray
    vardim: int = 1


@dataclass
class QcReachEllipsoid:
    invP: np.ndarray
    yc: np.ndarray
    vardim: int = 1


def hplaneS(normal, h, ffnet: FeedFwdNet) -> np.ndarray:
    n1, nK1 = ffnet.xdims[0], ffnet.xdims[-1]
    normal = np.asarray(normal, dtype=np.float64)
    S = np.zeros((n1 + nK1 + 1, n1 + nK1 + 1))
    S[n1:n1 + nK1, -1] = normal
    S[-1, n1:n1 + nK1] = normal
    S[-1, -1] = -2.0 * h
    return S


def loadVnnlibCnf(spec_file: str, ffnet: FeedFwdNet):
    """experiments/vnnlib_utils.jl:18-56: the CNF of a vnnlib property as a list of disjunctive clauses, each a list
    of (QcInputBox, QcSafety); parsed and flattened by the library (nnsdp_vnnlib_read)."""
    r = core.read_vnnlib(spec_file, ffnet.xdims[0], ffnet.xdims[-1])
    cnf = [[] for _ in range(r["nclauses"])]
    for i, c in enumerate(r["clause"]):
        cnf[int(c)].append((QcInputBox(x1min=r["x1min"][i], x1max=r["x1max"][i]), QcSafety(S=r["S"][i])))
    return cnf


def makeQcActivs(ffnet: FeedFwdNet, x1min=None, x1max=None, beta: int = None, intv_info: IntervalsInfo = None):
    """makeQcActivsIntvs (src/Qc/activ.jl:45-67) with bounds from the GPU IBP."""
    assert x1min is not None and x1max is not None and isinstance(beta, int)
    if intv_info is None:
        intv_info = makeIntervalsInfo(x1min, x1max, ffnet)
    acdim = ffnet.acdim
    acymin = np.concatenate([p[0] for p in intv_info.x_intvs[1:-1]])
    acymax = np.concatenate([p[1] for p in intv_info.x_intvs[1:-1]])
    qc_bounded = QcActivBounded(acydim=acdim, acymin=acymin, acymax=acymax)
    sec_min = np.concatenate([p[0] for p in intv_info.acx_intvs])
    sec_max = np.concatenate([p[1] for p in intv_info.acx_intvs])
    smin, smax = makeSectorMinMax(sec_min, sec_max)
    qc_sector = QcActivSector(acxdim=acdim, beta=beta, smin=smin, smax=smax)
    return [qc_bounded, qc_sector]


def makeCliques(qcs, ffnet: FeedFwdNet):
    """makeCliques(qcs, ffnet): beta comes from the first QcActivSector in qcs (0 if none)."""
    secs = [q for q in qcs if isinstance(q, QcActivSector)]
    beta = secs[0].beta if secs else 0
    return core.cliques_from_xdims(ffnet.xdims, beta)


# ---- queries and assembly ---------------------------------------------------------------------
@dataclass
class SafetyQuery:
    ffnet: FeedFwdNet
    qc_input: QcInputBox
    qc_safety: QcSafety
    qc_activs: list

    @property
    def qcs(self):
        return [self.qc_input, self.qc_safety] + list(self.qc_activs)


@dataclass
class ReachQuery:
    ffnet: FeedFwdNet
    qc_input: QcInputBox
    qc_reach: object
    qc_activs: list

    @property
    def qcs(self):
        return [self.qc_input, self.qc_reach] + list(self.qc_activs)


def _out_fields(qc_out, gout, ffnet: FeedFwdNet) -> dict:
    nK1 = ffnet.xdims[-1]
    if isinstance(qc_out, QcSafety):
        S = np.asarray(qc_out.S, dtype=np.float64)
        sd = ffnet.xdims[0] + nK1 + 1
        assert S.shape == (sd, sd)
        return dict(out_kind=L.OUT_SAFETY, out_S=S[None])
    gout = np.atleast_1d(np.asarray(gout, dtype=np.float64))
    assert len(gout) == qc_out.vardim == 1
    if isinstance(qc_out, QcReachHplane):
        assert len(qc_out.normal) == nK1
        return dict(out_kind=L.OUT_HPLANE, out_vec=np.asarray(qc_out.normal, dtype=np.float64)[None], gamma_out=gout)
    if isinstance(qc_out, QcReachCircle):
        assert len(qc_out.yc) == nK1
        return dict(out_kind=L.OUT_CIRCLE, out_vec=np.asarray(qc_out.yc, dtype=np.float64)[None], gamma_out=gout)
    if isinstance(qc_out, QcReachEllipsoid):
        assert len(qc_out.yc) == nK1
        return dict(out_kind=L.OUT_ELLIPSOID, out_vec=np.asarray(qc_out.yc, dtype=np.float64)[None],
                    out_invP=np.asarray(qc_out.invP, dtype=np.float64)[None], gamma_out=gout)
    raise ValueError(f"unrecognized qc: {qc_out}")


def _batch_for(ffnet, qc_input, qc_out, qc_bounded, qc_sector, gin, gbnd, gsec, gout, beta) -> core.NumericBatch:
    ac, n1 = ffnet.acdim, ffnet.xdims[0]
    if qc_input is None:
        qc_input = QcInputBox(np.zeros(n1), np.zeros(n1))
        gin = np.zeros(n1)
    if qc_bounded is None:
        qc_bounded = QcActivBounded(ac, np.zeros(ac), np.zeros(ac))
        gbnd = np.zeros(ac)
    if qc_sector is None:
        qc_sector = QcActivSector(ac, beta, np.zeros(ac), np.ones(ac))
        gsec = np.zeros(qc_sector.vardim)
    if qc_out is None:
        sd = n1 + ffnet.xdims[-1] + 1
        qc_out = QcSafety(np.zeros((sd, sd)))
    gin = np.asarray(gin, dtype=np.float64)
    gbnd = np.asarray(gbnd, dtype=np.float64)
    gsec = np.asarray(gsec, dtype=np.float64)
    assert len(gin) == qc_input.vardim
    assert len(gbnd) == qc_bounded.vardim
    assert len(gsec) == qc_sector.vardim
    return core.NumericBatch(
        x1min=qc_input.x1min[None], x1max=qc_input.x1max[None], gamma_in=gin[None], gamma_bnd=gbnd[None],
        gamma_sec=gsec[None], ymin=qc_bounded.acymin[None], ymax=qc_bounded.acymax[None],
        smin=qc_sector.smin[None], smax=qc_sector.smax[None], **_out_fields(qc_out, gout, ffnet))


def makeZin(gin, qc: QcInputBox, ffnet: FeedFwdNet) -> np.ndarray:
    """Numeric-gamma makeZin (src/Qc/input.jl:19-42): dense Zdim x Zdim."""
    if not isinstance(qc, QcInputBox):
        raise ValueError(f"unrecognized qc: {qc}")
    b = _batch_for(ffnet, qc, None, None, None, gin, None, None, None, 0)
    return core.assemble_dense(ffnet.device(), 0, b, Q=1)[0]


def makeZac(gac, qc, ffnet: FeedFwdNet) -> np.ndarray:
    """Numeric-gamma makeZac (src/Qc/activ.jl:30-41) for QcActivBounded or QcActivSector."""
    if isinstance(qc, QcActivBounded):
        b = _batch_for(ffnet, None, None, qc, None, None, gac, None, None, 0)
        return core.assemble_dense(ffnet.device(), 0, b, Q=1)[0]
    if isinstance(qc, QcActivSector):
        b = _batch_for(ffnet, None, None, None, qc, None, None, gac, None, qc.beta)
        return core.assemble_dense(ffnet.device(), qc.beta, b, Q=1)[0]
    raise ValueError(f"unrecognized qc: {qc}")


def makeZout(*args) -> np.ndarray:
    """makeZout(qc::QcSafety, ffnet) or makeZout(gout, qc::QcReach, ffnet) (src/Qc/output.jl:52-106)."""
    if len(args) == 2:
        qc, ffnet = args
        gout = None
        assert isinstance(qc, QcSafety)
    else:
        gout, qc, ffnet = args
    b = _batch_for(ffnet, None, qc, None, None, None, None, None, gout, 0)
    return core.assemble_dense(ffnet.device(), 0, b, Q=1)[0]


def _query_parts(query):
    qc_out = query.qc_safety if isinstance(query, SafetyQuery) else query.qc_reach
    bnd = [q for q in query.qc_activs if isinstance(q, QcActivBounded)]
    sec = [q for q in query.qc_activs if isinstance(q, QcActivSector)]
    assert len(bnd) <= 1 and len(sec) <= 1, "one QcActivBounded and one QcActivSector at most (activ.jl:64)"
    return qc_out, (bnd[0] if bnd else None), (sec[0] if sec else None)


def assembleZ(query, gin, gacs: Sequence[np.ndarray], gout=None) -> np.ndarray:
    """Z = Zin + Zout + sum(Zacs) (src/Methods/chordal_sdp.jl:114,145) for numeric multipliers.
    gacs follows query.qc_activs."""
    qc_out, bnd, sec = _query_parts(query)
    g = {type(q): ga for q, ga in zip(query.qc_activs, gacs)}
    beta = sec.beta if sec else 0
    b = _batch_for(query.ffnet, query.qc_input, qc_out, bnd, sec, gin, g.get(QcActivBounded), g.get(QcActivSector), gout, beta)
    return core.assemble_dense(query.ffnet.device(), beta, b, Q=1)[0]


def assembleCliqueBlocks(query, gin, gacs: Sequence[np.ndarray], gout=None):
    """The blocks Z[Ck, Ck] for the cliques of makeCliques(query.qcs, ffnet), numeric multipliers.
    Returns (cliques, [block_k])."""
    qc_out, bnd, sec = _query_parts(query)
    g = {type(q): ga for q, ga in zip(query.qc_activs, gacs)}
    beta = sec.beta if sec else 0
    b = _batch_for(query.ffnet, query.qc_input, qc_out, bnd, sec, gin, g.get(QcActivBounded), g.get(QcActivSector), gout, beta)
    flat = core.assemble_blocks(query.ffnet.device(), beta, b, Q=1)[0]
    cliques = makeCliques(query.qcs, query.ffnet)
    return cliques, core.split_blocks(flat, cliques)


def eigmaxZ(query, gin, gacs: Sequence[np.ndarray], gout=None, max_iters: int = 300, tol: float = 1e-10) -> float:
    """eigmax(Symmetric(Matrix(Z))) of src/Methods/Methods.jl:116-117 for numeric multipliers, computed
    matrix-free on the device (nnsdp_batch_lambda_max): the acceptance quantity of experiments/acas.jl:71-79."""
    qc_out, bnd, sec = _query_parts(query)
    g = {type(q): ga for q, ga in zip(query.qc_activs, gacs)}
    beta = sec.beta if sec else 0
    batch = _batch_for(query.ffnet, query.qc_input, qc_out, bnd, sec, gin, g.get(QcActivBounded), g.get(QcActivSector), gout, beta)
    b = core.Batch(query.ffnet.device(), beta, Qcap=1, ring=1)
    try:
        b.set_inputs(batch, Q=1)
        b.prepare()
        lam, _ = b.lambda_max(max_iters=max_iters, tol=tol)
        return float(lam[0])
    finally:
        b.close()
