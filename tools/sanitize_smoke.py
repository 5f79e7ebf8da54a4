#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Covers: IBP, sector slopes, prep, Gram (DMMA), fill strips, RC/CR window tiles, edge tiles, the sparse
host gather (pack kernel) and the affine-coefficient kernels, on a wide net (128-row tiles) and a small one.
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
assert worst <= 1e-12, worst
print("sanitize_smoke OK, worst rel err %.2e" % worst)
