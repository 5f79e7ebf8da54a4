#!/usr/bin/env python
"""Bit-for-bit check that the Gram tile side (NNSDP_GRAM_TILE=64 / 128) does not change the blocks: run twice, compare
the printed digest.  Degenerate boxes make every layer Gram-active."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("nn-sdp_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import nnsdp_b200 as nb
from helpers import rand_net, rand_query, to_numeric_batch

ctx = nb.Context([0])
h = hashlib.sha256()
for xdims in ([2, 300, 270, 2], [3, 150, 260, 140, 2], [2, 1000, 1000, 2]):
    net = rand_net(xdims, seed=3, sigma=0.1)
    rng = np.random.default_rng(5)
    qs = [rand_query(net, 2, rng, kind="hplane", radius=0.0) for _ in range(2)]
    out = nb.assemble_blocks(nb.Net(ctx, net.xdims, net.Ms), 2, to_numeric_batch(nb, net, qs))
    h.update(out.tobytes())
print("digest", h.hexdigest(), "tile", os.environ.get("NNSDP_GRAM_TILE", "auto"))
