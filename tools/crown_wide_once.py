#!/usr/bin/env python
"""One CROWN call on the stress-size net (profiling target for the tensor-core GEMM: ncu -k regex:dgemm_dmma)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("nn-sdp_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import nnsdp_b200 as nb
from helpers import rand_net
ctx = nb.Context([0])
net = rand_net([2] + [1000] * 20 + [2], seed=1)
dnet = nb.Net(ctx, net.xdims, net.Ms)
c = np.random.default_rng(0).uniform(0.5, 1.5, (2, 2))
nb.bounds_crown(dnet, c - 0.05, c + 0.05)
