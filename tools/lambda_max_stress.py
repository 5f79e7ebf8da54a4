import sys, time, os
sys.path[:0]=['/root/repo','/root/repo/nn-sdp_b200']
import numpy as np, bench, nnsdp_b200 as nb
xdims, Ms, beta, inp = bench.make_workload("stress-W1000-D20-beta2-Q1024", 0, Q=16)
ctx=nb.Context([0]); net=nb.Net(ctx,xdims,Ms)
b=nb.Batch(net,beta,Qcap=16,ring=1)
b.set_inputs(nb.NumericBatch(out_kind=nb.OUT_SAFETY, **inp), Q=16)
b.bounds(); b.prepare(); b.sync()
for iters,tol in ((50,1e-6),(150,1e-9),(300,1e-11)):
    t0=time.perf_counter(); lam,its=b.lambda_max(max_iters=iters,tol=tol); dt=time.perf_counter()-t0
    print(iters,tol,'time %.3f s'%dt, 'lam', lam[:4], 'its', its[:8])
