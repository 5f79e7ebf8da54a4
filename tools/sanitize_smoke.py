#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Covers: IBP, sector slopes, prep, Gram (DMMA), fill strips, RC/CR window tiles (TMA-staged W tiles on layers of
even width, the manual path on odd ones), the band kernel, edge tiles, the sparse host gather (pack kernel), the
affine-coefficient kernels, packed records, the many-query tensor-core paths (interval propagation with two
accumulator sets, the affine-column products of all layers in one launch), CROWN bounds and the matrix-free
lambda_max, on wide nets (128-row tiles) and a small one.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("nn-sdp_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import nnsdp_b200 as nb  # noqa: E402
import nnsdp_oracle as o  # noqa: E402
from helpers import rand_net, rand_query, relerr, to_numeric_batch  # noqa: E402

ctx = nb.Context([0])
worst = 0.0
for xdims, beta, kind in (([3, 150, 260, 140, 2], 2, "ellipsoid"), ([2, 300, 270, 2], 1, "safety"), ([2, 6, 5, 7, 2], 2, "hplane")):
    net = rand_net(xdims, seed=1, sigma=0.1)
    rng = np.random.default_rng(0)
    qs = [rand_query(net, beta, rng, kind=kind, radius=r) for r in (0.0, 0.2, 0.01)]
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    batch = to_numeric_batch(nb, net, qs)
    b = nb.Batch(dnet, beta, Qcap=3, ring=2)
    b.set_inputs(batch)
    out = np.full((3, b.per_query), np.nan)
    b.run(out)
    cliques = o.make_cliques(net, beta)
    for i, q in enumerate(qs):
        ref = o.run_query(net, beta, q)
        for blk, rb in zip(nb.split_blocks(out[i], cliques), ref["blocks"]):
            worst = max(worst, relerr(blk, rb))
    b.close()
    Z = nb.assemble_dense(dnet, beta, batch)
    worst = max(worst, relerr(Z[1], o.run_query(net, beta, qs[1])["Z"]))
    if max(xdims) <= 300:
        A = nb.affine_form(dnet, beta, to_numeric_batch(nb, net, qs[:1]))
        g = np.concatenate([qs[0].gin, qs[0].gout if kind != "safety" else [], qs[0].gbnd, qs[0].gsec])
        zg = A["z0"].copy()
        np.add.at(zg, A["coo_ent"] - 1, A["coo_val"] * g[A["coo_var"] - 1])
        worst = max(worst, np.abs(zg - Z[0][A["ent_row"] - 1, A["ent_col"] - 1]).max() / max(np.abs(Z[0]).max(), 1.0))

# packed records: band kernel + BAND / DIAG cells, TMA on the even layer and the manual CR path on the odd one
for xdims, beta in (([2, 300, 261, 258, 2], 2), ([2, 129, 128, 127, 300, 4], 3)):
    net = rand_net(xdims, seed=4, sigma=0.1)
    rng = np.random.default_rng(1)
    qs = [rand_query(net, beta, rng, kind="safety", radius=r) for r in (0.0, 0.3, 0.004)]
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    batch = to_numeric_batch(nb, net, qs)
    flat = nb.assemble_blocks(dnet, beta, batch)
    rec, present, _ = nb.assemble_packed(dnet, beta, batch)
    for i in range(3):
        assert np.array_equal(nb.packed_unpack(xdims, beta, rec[i], present[i]), flat[i])
    lam = nb.Batch(dnet, beta, Qcap=3, ring=1)
    lam.set_inputs(batch)
    lam.bounds()
    lam.prepare()
    lm, _, _, conv = lam.lambda_max(max_iters=400, full=True)
    lam.close()
    for i in range(3):
        Z = o.run_query(net, beta, qs[i])["Z"]
        ev = np.linalg.eigvalsh(0.5 * (Z + Z.T))
        if conv[i]:
            worst = max(worst, abs(lm[i] - ev[-1]) / max(abs(ev[0]), abs(ev[-1])) * 1e-5)   # Lanczos: 1e-7 of the spectral scale

# many queries on wide layers: tensor-core interval propagation and batched affine products; CROWN bounds
xdims, beta, nq = [2, 256, 250, 200, 2], 2, 130
net = rand_net(xdims, seed=6, sigma=0.15)
rng = np.random.default_rng(2)
qs = [rand_query(net, beta, rng, kind="safety", radius=0.01 * (1 + i % 3)) for i in range(nq)]
dnet = nb.Net(ctx, net.xdims, net.Ms)
b = nb.Batch(dnet, beta, Qcap=nq, ring=2)
b.set_inputs(to_numeric_batch(nb, net, qs))
b.bounds()
b.prepare()
bd, aff = b.get_bounds(), b.get_affine()
b.close()
for i in (0, nq - 1):
    ref = o.run_query(net, beta, qs[i])
    worst = max(worst, relerr(aff[i], ref["Z"][:, -1]))
    iv = o.intervals_worst_case(qs[i].x1min, qs[i].x1max, net)
    xmax = np.concatenate([p[1] for p in iv.x_intvs])
    worst = max(worst, np.abs(bd["xmax"][i] - xmax).max() / max(np.abs(xmax).max(), 1.0))
c = rng.uniform(0.5, 1.5, (3, 2))
r = nb.bounds_crown(dnet, c - 0.05, c + 0.05)
ref = o.intervals_crown(c[1] - 0.05, c[1] + 0.05, net)
xmax = np.concatenate([p[1] for p in ref.x_intvs])
worst = max(worst, 0.1 * np.abs(r["xmax"][1] - xmax).max() / max(np.abs(xmax).max(), 1.0))   # CROWN: 1e-11

assert worst <= 1e-12, worst
print("sanitize_smoke OK, worst rel err %.2e" % worst)
