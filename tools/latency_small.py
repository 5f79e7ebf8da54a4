#!/usr/bin/env python
"""One-shot latency of the host API on the small BASELINE configs (1: W10-D10 single query; 3: 64 reach
directions on W20-D10; 4: ACAS-shaped 5x50, 45 queries)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("nn-sdp_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import nnsdp_b200 as nb
import nnsdp_oracle as o
from helpers import rand_net, rand_query, to_numeric_batch

ctx = nb.Context([0])
for name, xdims, beta, nq, kind in (("config1 W10-D10 beta1, 1 query", [2] + [10] * 10 + [2], 1, 1, "safety"),
                                    ("config3 W20-D10 beta2, 64 directions", [2] + [20] * 10 + [2], 2, 64, "hplane"),
                                    ("config4 ACAS 5x50 beta2, 45 queries", [5] + [50] * 6 + [5], 2, 45, "hplaneS"),
                                    ("scale W20-D100 beta2, 1 query", [2] + [20] * 100 + [2], 2, 1, "ellipsoid")):
    net = rand_net(xdims, seed=1)
    rng = np.random.default_rng(0)
    qs = [rand_query(net, beta, rng, kind=kind, radius=0.05) for _ in range(nq)]
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    batch = to_numeric_batch(nb, net, qs)
    nb.assemble_blocks(dnet, beta, batch)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        nb.assemble_blocks(dnet, beta, batch)
        ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    for q in qs[:4]:
        o.run_query(net, beta, q)
    tcpu = (time.perf_counter() - t0) / min(4, nq)
    print(f"{name:42s} call {1e3 * np.median(ts):8.3f} ms  ({1e3 * np.median(ts) / nq:7.3f} ms/query)   CPU oracle {1e3 * tcpu:8.2f} ms/query")
