"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): clique index sets and sparsity patterns bit-exact; sector slopes
(values in {0,1}) bit-exact; FP64 bounds and block entries within 1e-12 relative (normwise per
block: the Gram entries cancel, so an entrywise bound is not attainable -- SURVEY.md section 7).
"""
import os

import numpy as np
import pytest

import nnsdp_oracle as o
from helpers import rand_net, rand_query, relerr, to_numeric_batch

pytestmark = pytest.mark.gpu

TOL = 1e-12

NETS = [
    ([2, 3, 3, 2], 1),
    ([2, 3, 2], 0),
    ([2, 3, 2], 2),
    ([3, 3, 3, 3, 4, 3, 3], 2),          # xdims of experiments/plot_sparsity.ipynb
    ([2, 10, 10, 10, 10, 2], 3),
    ([2, 4, 7, 3, 5, 2], 5),              # beta wider than a layer: band crosses two boundaries
    ([5, 50, 50, 50, 50, 50, 50, 5], 2),  # ACAS-shaped (config 4)
    ([2] + [20] * 10 + [2], 1),
    ([2, 70, 130, 64, 3], 2),             # ragged widths, block-split tiles, 2 Gram tiles
]


def _oracle_bounds(net, q):
    info = o.intervals_worst_case(q.x1min, q.x1max, net)
    xmin = np.concatenate([p[0] for p in info.x_intvs])
    xmax = np.concatenate([p[1] for p in info.x_intvs])
    amin = np.concatenate([p[0] for p in info.acx_intvs])
    amax = np.concatenate([p[1] for p in info.acx_intvs])
    return info, xmin, xmax, amin, amax


@pytest.mark.parametrize("xdims,beta", NETS)
def test_bounds_and_sector(ctx, xdims, beta):
    import nnsdp_b200 as nb

    net = rand_net(xdims, seed=11)
    rng = np.random.default_rng(5)
    qs = [rand_query(net, beta, rng, radius=r) for r in (0.0, 0.01, 0.1, 0.5)]
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    r = nb.bounds_ibp(dnet, np.stack([q.x1min for q in qs]), np.stack([q.x1max for q in qs]))
    for i, q in enumerate(qs):
        _, xmin, xmax, amin, amax = _oracle_bounds(net, q)
        scale = max(np.abs(xmax).max(), np.abs(xmin).max(), 1.0)
        assert np.abs(r["xmin"][i] - xmin).max() <= TOL * scale
        assert np.abs(r["xmax"][i] - xmax).max() <= TOL * scale
        assert np.abs(r["acxmin"][i] - amin).max() <= TOL * scale
        assert np.abs(r["acxmax"][i] - amax).max() <= TOL * scale
        assert np.all(r["xmin"][i] <= r["xmax"][i])
        smin, smax = nb.sector_minmax(ctx, r["acxmin"][i], r["acxmax"][i])
        rmin, rmax = o.make_sector_min_max(r["acxmin"][i], r["acxmax"][i])
        assert np.array_equal(smin, rmin) and np.array_equal(smax, rmax)  # bit-exact on equal inputs
        # one-step pre-activation IBP from given x bounds (intervals_auto_lirpa.jl:55-62)
    amin_d, amax_d = nb.preact_from_x(dnet, r["xmin"], r["xmax"])
    for i, q in enumerate(qs):
        info, *_ = _oracle_bounds(net, q)
        ref = o.preact_from_x(info.x_intvs, net)
        scale = max(np.abs(np.concatenate([p[1] for p in ref])).max(), 1.0)
        assert np.abs(amin_d[i] - np.concatenate([p[0] for p in ref])).max() <= TOL * scale
        assert np.abs(amax_d[i] - np.concatenate([p[1] for p in ref])).max() <= TOL * scale


def test_preact_assert(ctx):
    """ykmin <= ykmax is asserted by the reference (intervals_auto_lirpa.jl:60)."""
    import nnsdp_b200 as nb

    net = rand_net([2, 5, 5, 2], seed=1)
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    xmin = np.ones((1, sum(net.xdims)))
    xmax = np.zeros((1, sum(net.xdims)))  # inverted box
    with pytest.raises(nb.NnsdpError) as e:
        nb.preact_from_x(dnet, xmin, xmax)
    assert e.value.code == -5


@pytest.mark.parametrize("kind", ["safety", "hplaneS", "hplane", "circle", "ellipsoid"])
@pytest.mark.parametrize("xdims,beta", NETS)
def test_dense_Z_and_blocks(ctx, xdims, beta, kind):
    import nnsdp_b200 as nb

    net = rand_net(xdims, seed=3)
    rng = np.random.default_rng(17)
    # tight boxes give stably-active neurons (Gram path), wide boxes give unstable ones
    qs = [rand_query(net, beta, rng, kind=kind, radius=r) for r in (0.001, 0.02, 0.3)]
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    batch = to_numeric_batch(nb, net, qs)
    Z = nb.assemble_dense(dnet, beta, batch)
    flat = nb.assemble_blocks(dnet, beta, batch)
    cliques = dnet.cliques(beta)
    ref_cliques = o.make_cliques(net, beta)
    assert len(cliques) == len(ref_cliques)
    for (a, pa, da), (b, pb, db) in zip(cliques, ref_cliques):
        assert a.dtype == np.int64 and np.array_equal(a, b)
        assert all(np.array_equal(x, y) for x, y in zip(pa, pb)) and len(pa) == len(pb)
        assert all(np.array_equal(x, y) for x, y in zip(da, db)) and len(da) == len(db)
    for i, q in enumerate(qs):
        ref = o.run_query(net, beta, q, form="closed")
        assert relerr(Z[i], ref["Z"]) <= TOL
        assert np.array_equal(Z[i], Z[i].T)  # bit-symmetric
        lit = o.run_query(net, beta, q, form="literal")
        assert relerr(Z[i], lit["Z"]) <= TOL
        # sparsity pattern: structural pattern of the notebook contains ours; mask equals the oracle's
        pat = o.structural_pattern_notebook(net.xdims, beta)
        assert not np.any((Z[i] != 0) & ~pat)
        blocks = nb.split_blocks(flat[i], cliques)
        for blk, rb in zip(blocks, ref["blocks"]):
            assert blk.shape == rb.shape
            assert relerr(blk, rb) <= TOL
            thr = 1e-13 * max(np.abs(rb).max(), 1e-300)
            assert np.array_equal(np.abs(blk) > thr, np.abs(rb) > thr)
        # blocks are exactly the restriction of the dense Z the same library produced
        for blk, (Ck, _, _) in zip(blocks, cliques):
            assert np.array_equal(blk, Z[i][np.ix_(Ck - 1, Ck - 1)])


def test_gram_active_path(ctx):
    """Degenerate box (xmin == xmax): every neuron is stably on or off, so the Gram term
    W' diag(-2 lambda) W is exercised on every layer, including multi-tile widths."""
    import nnsdp_b200 as nb

    xdims, beta = [3, 150, 260, 140, 2], 2
    net = rand_net(xdims, seed=9, sigma=0.3)
    rng = np.random.default_rng(2)
    qs = [rand_query(net, beta, rng, kind="safety", radius=0.0) for _ in range(2)]
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    b = nb.Batch(dnet, beta, Qcap=2, ring=2, dense=True)
    b.set_inputs(to_numeric_batch(nb, net, qs))
    b.run()
    ncon, nact = b.gram_stats()
    assert ncon == 2 * (len(xdims) - 2) and nact > 100
    for i, q in enumerate(qs):
        ref = o.run_query(net, beta, q, form="closed")
        Z = b.get_slot(i).reshape(ref["Z"].shape).T
        assert relerr(Z, ref["Z"]) <= TOL
        assert np.array_equal(Z, Z.T)


def test_supplied_bounds_and_shared_inputs(ctx):
    """Caller-supplied QC data (the CROWN route of the reference: bounds come from outside) and
    stride-0 sharing: a reach batch where only the hyperplane normal differs (NnSdp.jl:73-95)."""
    import nnsdp_b200 as nb

    xdims, beta, nq = [2, 20, 20, 20, 20, 2], 2, 16
    net = rand_net(xdims, seed=4)
    rng = np.random.default_rng(8)
    base = rand_query(net, beta, rng, kind="hplane", radius=0.1)
    ref0 = o.run_query(net, beta, base)
    # perturb the bounds so that they are NOT what IBP would give
    bnd, sec = ref0["qc_bounded"], ref0["qc_sector"]
    ymin = bnd.acymin - 0.01 * rng.random(net.acdim)
    ymax = bnd.acymax + 0.01 * rng.random(net.acdim)
    smin = (rng.random(net.acdim) < 0.3).astype(float)
    smax = np.maximum(smin, (rng.random(net.acdim) < 0.7).astype(float))
    thetas = 2 * np.pi * np.arange(nq) / nq
    normals = np.stack([np.cos(thetas), np.sin(thetas)], axis=1)
    gouts = rng.random((nq, 1))
    batch = nb.NumericBatch(
        x1min=base.x1min, x1max=base.x1max, gamma_in=base.gin, gamma_bnd=base.gbnd, gamma_sec=base.gsec,
        ymin=ymin, ymax=ymax, smin=smin, smax=smax, out_kind=nb.OUT_HPLANE, out_vec=normals, gamma_out=gouts)
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    flat = nb.assemble_blocks(dnet, beta, batch, Q=nq)
    cliques = o.make_cliques(net, beta)
    qb = o.QcActivBounded(net.acdim, ymin, ymax)
    qs_ = o.QcActivSector(net.acdim, beta, smin, smax)
    for i in range(nq):
        Zr = o.assemble_Z_closed_form(net, o.QcInputBox(base.x1min, base.x1max), o.QcReachHplane(normals[i]), qb, qs_,
                                      base.gin, base.gbnd, base.gsec, gouts[i])
        for blk, rb in zip(nb.split_blocks(flat[i], cliques), o.clique_blocks(Zr, cliques)):
            assert relerr(blk, rb) <= TOL


def test_tile_classes_do_not_change_results(ctx):
    """The host-side tile flags only skip terms that are structurally zero: evaluating every term
    in every tile (NNSDP_NO_TILE_CLASSES=1) must give bit-identical output."""
    import nnsdp_b200 as nb

    outs = []
    for flag in ("0", "1"):
        os.environ["NNSDP_NO_TILE_CLASSES"] = flag
        try:
            for xdims, beta in ([2, 70, 130, 64, 3], 3), ([2] + [12] * 6 + [2], 4):
                net = rand_net(xdims, seed=21)
                rng = np.random.default_rng(1)
                qs = [rand_query(net, beta, rng, kind="safety", radius=r) for r in (0.0, 0.2)]
                dnet = nb.Net(ctx, net.xdims, net.Ms)
                outs.append(nb.assemble_blocks(dnet, beta, to_numeric_batch(nb, net, qs)))
        finally:
            os.environ["NNSDP_NO_TILE_CLASSES"] = "0"
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[1], outs[3])


def test_batch_ring_and_chunking(ctx):
    """More queries than ring slots, odd ring, host gather through the two-half ring."""
    import nnsdp_b200 as nb

    xdims, beta, nq = [2, 16, 16, 16, 2], 1, 11
    net = rand_net(xdims, seed=5)
    rng = np.random.default_rng(3)
    qs = [rand_query(net, beta, rng, kind="circle", radius=0.05 * (i + 1)) for i in range(nq)]
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    b = nb.Batch(dnet, beta, Qcap=nq, ring=5)
    b.set_inputs(to_numeric_batch(nb, net, qs))
    out = np.zeros((nq, b.per_query))
    b.run(out)
    cliques = o.make_cliques(net, beta)
    for i, q in enumerate(qs):
        ref = o.run_query(net, beta, q)
        for blk, rb in zip(nb.split_blocks(out[i], cliques), ref["blocks"]):
            assert relerr(blk, rb) <= TOL
    ms, launches = b.stage_ms("emit")
    assert launches >= 3 and ms > 0


def test_reference_api_mirror(ctx):
    """makeZin / makeZac / makeZout / makeCliques through the reference-shaped host API."""
    from nnsdp_b200 import reference_api as R

    xdims, beta = [2, 6, 5, 7, 2], 2
    net = rand_net(xdims, seed=13)
    ff = R.FeedFwdNet(net.xdims, net.Ms)
    rng = np.random.default_rng(6)
    x1min, x1max = np.array([0.4, 0.6]), np.array([0.5, 0.9])
    qc_in = R.QcInputBox(x1min, x1max)
    qc_acts = R.makeQcActivs(ff, x1min=x1min, x1max=x1max, beta=beta)
    o_acts = o.make_qc_activs_intvs(net, x1min, x1max, beta)
    assert np.array_equal(qc_acts[1].smin, o_acts[1].smin) and np.array_equal(qc_acts[1].smax, o_acts[1].smax)
    gin, gb, gs = rng.random(2), rng.random(net.acdim), rng.random(qc_acts[1].vardim)
    assert relerr(R.makeZin(gin, qc_in, ff), o.makeZin(gin, o.QcInputBox(x1min, x1max), net).toarray()) <= TOL
    assert relerr(R.makeZac(gb, qc_acts[0], ff), o.makeZac(gb, o_acts[0], net).toarray()) <= TOL
    assert relerr(R.makeZac(gs, qc_acts[1], ff), o.makeZac(gs, o_acts[1], net).toarray()) <= TOL
    S = R.hplaneS([1.0, -0.5], 0.3, ff)
    assert relerr(R.makeZout(R.QcSafety(S), ff), o.makeZout(o.QcSafety(S), net).toarray()) <= TOL
    ell = R.QcReachEllipsoid(rng.standard_normal((2, 2)), rng.standard_normal(2))
    assert relerr(R.makeZout([0.4], ell, ff), o.makeZout(o.QcReachEllipsoid(ell.invP, ell.yc), net, [0.4]).toarray()) <= TOL
    query = R.SafetyQuery(ff, qc_in, R.QcSafety(S), qc_acts)
    cl = R.makeCliques(query.qcs, ff)
    for (a, _, _), (b, _, _) in zip(cl, o.make_cliques(net, beta)):
        assert np.array_equal(a, b)
    Z = R.assembleZ(query, gin, [gb, gs])
    Zr = o.assemble_Z_literal(net, o.QcInputBox(x1min, x1max), o.QcSafety(S), o_acts, gin, [gb, gs])
    assert relerr(Z, Zr) <= TOL
    cl2, blocks = R.assembleCliqueBlocks(query, gin, [gb, gs])
    for blk, (Ck, _, _) in zip(blocks, cl2):
        assert relerr(blk, Zr[np.ix_(Ck - 1, Ck - 1)]) <= TOL


def test_error_paths(ctx):
    import nnsdp_b200 as nb

    net = rand_net([2, 4, 4, 2], seed=1)
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    with pytest.raises(nb.NnsdpError):  # beta > acdim: _lambda_dim assert of the reference
        dnet.sizes(9)
    rng = np.random.default_rng(0)
    q = rand_query(net, 1, rng)
    batch = to_numeric_batch(nb, net, [q])
    batch.ymin = np.ones((1, net.acdim))
    batch.ymax = np.zeros((1, net.acdim))  # acymin <= acymax violated
    batch.smin = np.zeros((1, net.acdim))
    batch.smax = np.ones((1, net.acdim))
    with pytest.raises(nb.NnsdpError) as e:
        nb.assemble_blocks(dnet, 1, batch)
    assert e.value.code == -5
    b = nb.Batch(dnet, 1, Qcap=1, ring=1)
    with pytest.raises(nb.NnsdpError) as e:
        b.prepare()  # before set_inputs
    assert e.value.code == -4
