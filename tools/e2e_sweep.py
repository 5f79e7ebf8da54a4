#!/usr/bin/env python
"""e2e (host gather) throughput of the stress workload for a few settings of the gather knobs, same box, interleaved.
   python tools/e2e_sweep.py   (each setting runs in a fresh subprocess: the knobs are read once per process)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path[:0] = [%(root)r, os.path.join(%(root)r, "nn-sdp_b200")]
import numpy as np, bench, nnsdp_b200 as nb
Q = 8
xdims, Ms, beta, inp = bench.make_workload("stress-W1000-D20-beta2-Q1024", 0, Q=Q)
ctx = nb.Context([0]); net = nb.Net(ctx, xdims, Ms); sz = net.sizes(beta)
pin = nb.PinnedBuffer(Q * sz["sum_ck_sq"]); b = nb.Batch(net, beta, Qcap=Q, ring=8)
batch = nb.NumericBatch(out_kind=nb.OUT_SAFETY, **inp)
def step(flags):
    b.set_inputs(batch, Q=Q); b.run(pin.array, flags=flags)
out = {}
for name, flags in (("full", 0), ("prezeroed", nb.RUN_HOST_PREZEROED)):
    step(flags); ts = []
    for _ in range(4):
        t0 = time.perf_counter(); step(flags); ts.append(time.perf_counter() - t0)
    out[name] = Q / min(ts)
print("RESULT", out)
'''
settings = [{"NNSDP_GATHER_DMA_ZERO_PCT": str(p), "NNSDP_HOST_THREADS": str(t)} for t in (8, 16) for p in (0, 10, 20, 35)]
for rep in range(2):
    for s in settings:
        env = dict(os.environ, **s)
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        print(rep, s, line[0] if line else r.stderr[-300:], flush=True)
