#!/usr/bin/env python
"""Dumps the LIBRARY's hand-off for the reference's findEllipsoid runs (experiments/scale.jl:26-29,70) -- run on the
GPU box:   python tests/golden/make_handoff.py gpurun_out/handoff

For the reference's shipped nets bench/rand/scale-I2-O2-W10-D10.nnet (beta = 0..7) and -W10-D20.nnet (beta = 2), box
[0.5, 1.5]^2, ellipsoid (P, yc) as in oracle/sdp_crosscheck.py: CROWN bounds from nnsdp_bounds_crown (the reference's
default IntervalsAutoLirpa), then nnsdp_affine_create / nnsdp_affine_get and nnsdp_cliques.  Everything stored comes
out of libnnsdp_b200.so; the .npz files are what oracle/sdp_decomposed.py solves and what
tests/test_gpu_parity.py::test_affine_handoff_equals_the_committed_fixture re-computes.  Also records the wall time of
nnsdp_affine_create + nnsdp_affine_get and nnz (the reference's setup_secs for the same nets are in dump/scale/*.csv).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "nn-sdp_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def ellipsoid_for(name, net_oracle, x1min, x1max):
    import sdp_crosscheck as sc

    if name == "W10-D10":
        res = json.load(open(os.path.join(HERE, "scale_W10_D10_optimum.json")))
        return np.array(res["P"]), np.array(res["yc"])
    return sc.approx_ellipsoid_population(net_oracle, x1min, x1max)


def handoff(nb, ctx, name, beta, reps=3):
    import nnsdp_oracle as o

    path = os.path.join(HERE, f"scale-I2-O2-{name}.nnet")
    xdims, Ms = nb.read_nnet(path)
    net = nb.Net(ctx, xdims, Ms)
    x1min, x1max = np.full(2, 0.5), np.full(2, 1.5)
    P, yc = ellipsoid_for(name, o.load_nnet(path), x1min, x1max)
    invP = np.linalg.inv(P)
    invP = 0.5 * (invP + invP.T)
    r = nb.bounds_crown(net, x1min[None], x1max[None])
    n_in = xdims[0]
    ymin, ymax = r["xmin"][:, n_in:n_in + sum(xdims[1:-1])], r["xmax"][:, n_in:n_in + sum(xdims[1:-1])]
    smin, smax = nb.sector_minmax(ctx, r["acxmin"], r["acxmax"])
    batch = nb.NumericBatch(x1min=x1min[None], x1max=x1max[None], out_kind=nb.OUT_ELLIPSOID, out_vec=yc[None],
                            out_invP=invP[None], gamma_out=np.zeros((1, 1)), ymin=ymin, ymax=ymax, smin=smin, smax=smax)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        h = nb.affine_form(net, beta, batch)
        times.append(time.perf_counter() - t0)
    sz = net.sizes(beta)
    K = len(xdims) - 1
    p = sz["ncliques"]
    arrs = {k: np.zeros(n, dtype=np.int64) for k, n in (("ck_off", p + 1), ("ck_idx", sz["sum_ck"]), ("ck1_len", p),
                                                          ("d_off", 2 * p + 1), ("d_idx", sz["sum_dk"]))}
    import ctypes as C
    import nnsdp_b200._lib as L

    ip = lambda a: a.ctypes.data_as(L.c_i64p)
    L.check(L.lib.nnsdp_cliques(net._h, beta, *[ip(arrs[k]) for k in ("ck_off", "ck_idx", "ck1_len", "d_off", "d_idx")]))
    out = {k: h[k] for k in ("ent_row", "ent_col", "z0", "coo_ent", "coo_var", "coo_val")}
    out.update({k: np.int64(h[k]) for k in ("nvar", "nent", "nnz", "var_in", "var_out", "var_bnd", "var_sec")})
    out.update(arrs)
    out.update(xdims=np.array(xdims), beta=np.int64(beta), x1min=x1min, x1max=x1max, P=P, yc=yc, invP=invP,
               ymin=ymin[0], ymax=ymax[0], smin=smin[0], smax=smax[0])
    return out, {"net": name, "beta": beta, "nvar": int(h["nvar"]), "nent": int(h["nent"]), "nnz": int(h["nnz"]),
                 "affine_create_get_ms": [1e3 * t for t in times]}


def timing_only(nb, ctx, name, beta):
    """nnsdp_affine_create + get on the nets of dump/scale with IBP bounds left to the library (no fixtures)."""
    path = os.path.join(HERE, f"scale-I2-O2-{name}.nnet")
    if os.path.exists(path):
        xdims, Ms = nb.read_nnet(path)
    else:                                            # same shape, seeded random weights (sigma of make_networks.jl:44)
        W, D = (int(s[1:]) for s in name.split("-"))
        xdims = [2] + [W] * D + [2]
        rng = np.random.default_rng(7)
        Ms = [2.0 / np.sqrt(W * np.log(W)) * rng.standard_normal((xdims[k + 1], xdims[k] + 1)) for k in range(len(xdims) - 1)]
    net = nb.Net(ctx, xdims, Ms)
    batch = nb.NumericBatch(x1min=np.full((1, 2), 0.5), x1max=np.full((1, 2), 1.5), out_kind=nb.OUT_ELLIPSOID,
                            out_vec=np.zeros((1, 2)), out_invP=np.eye(2)[None], gamma_out=np.zeros((1, 1)))
    r = nb.bounds_crown(net, np.full((1, 2), 0.5), np.full((1, 2), 1.5))
    n_in, ac = xdims[0], sum(xdims[1:-1])
    smin, smax = nb.sector_minmax(ctx, r["acxmin"], r["acxmax"])
    batch.ymin, batch.ymax, batch.smin, batch.smax = r["xmin"][:, n_in:n_in + ac], r["xmax"][:, n_in:n_in + ac], smin, smax
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        h = nb.affine_form(net, beta, batch)
        times.append(time.perf_counter() - t0)
    return {"net": name, "beta": beta, "nvar": int(h["nvar"]), "nent": int(h["nent"]), "nnz": int(h["nnz"]),
            "affine_create_get_ms": [1e3 * t for t in times]}


def main():
    import nnsdp_b200 as nb

    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "handoff")
    os.makedirs(outdir, exist_ok=True)
    ctx = nb.Context([0])
    log = []
    for name, betas in (("W10-D10", range(8)), ("W10-D20", [2])):
        for beta in betas:
            data, rec = handoff(nb, ctx, name, beta)
            np.savez_compressed(os.path.join(outdir, f"handoff_{name}_beta{beta}.npz"), **data)
            log.append(rec)
            print(rec, flush=True)
    for name in ("W10-D10", "W20-D50", "W20-D100"):
        for beta in (1, 2, 3):
            rec = timing_only(nb, ctx, name, beta)
            rec["timing_only"] = True
            log.append(rec)
            print(rec, flush=True)
    json.dump(log, open(os.path.join(outdir, "affine_timing.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
