#!/usr/bin/env python
"""Device time of the CROWN bounds against IBP at the reference's sizes (and one wide net)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("nn-sdp_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import nnsdp_b200 as nb
import nnsdp_oracle as o
from helpers import rand_net

ctx = nb.Context([0])
for xdims, Q in (([2] + [10] * 10 + [2], 64), ([2] + [20] * 100 + [2], 16), ([2] + [100] * 50 + [2], 16), ([2] + [1000] * 20 + [2], 2)):
    net = rand_net(xdims, seed=1)
    rng = np.random.default_rng(0)
    c = rng.uniform(0.5, 1.5, (Q, 2))
    lo, hi = c - 0.05, c + 0.05
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    for name, fn in (("ibp", nb.bounds_ibp), ("crown", nb.bounds_crown)):
        fn(dnet, lo, hi)
        dt = 1e9
        for _ in range(5):   # best of 5: the launch-bound cases are sensitive to host noise
            t0 = time.perf_counter()
            r = fn(dnet, lo, hi)
            dt = min(dt, time.perf_counter() - t0)
        w = float(np.mean(r["xmax"] - r["xmin"]))
        print(f"W{xdims[1]}-D{len(xdims) - 2} Q={Q:3d} {name:5s} {1e3 * dt:9.2f} ms per call ({1e3 * dt / Q:8.3f} ms/query)  mean interval width {w:.4g}")
    if xdims[1] <= 20:
        t0 = time.perf_counter()
        o.intervals_crown(lo[0], hi[0], net)
        print(f"      numpy oracle CROWN {1e3 * (time.perf_counter() - t0):.1f} ms/query")
