#!/usr/bin/env python
"""Generates the fixtures in tests/golden/ (run in the dev container, where /root/reference exists).

    python tests/golden/make_golden.py

What is pinned by what:
  * ``*.weights``      read with the REFERENCE's own reader exts/NNet/utils/readNNet.py (imported by
                       file path from /root/reference), so the fixture weights are the reference's
                       view of its shipped bench/rand nets; the oracle's load_nnet must reproduce them.
  * ``scale-I2-O2-W5-D5.nnet``  a 2 KB data file copied verbatim from bench/rand (data, not source) so
                       the .nnet reader can be tested where /root/reference does not exist.
  * everything else    produced by oracle/nnsdp_oracle.py (LITERAL form = the Julia operations with
                       numeric gamma).  The reference ships no golden vectors for this path and Julia
                       cannot run here, so these pin the oracle against regressions and
                       give the GPU tests fixed vectors -- they are NOT outputs of the Julia code.
Seeds: numpy PCG64, 1234 and 64 hyperplane directions (config 3).
"""
import importlib.util
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nnsdp_oracle as o  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_readNNet", os.path.join(REF, "exts/NNet/utils/readNNet.py"))
ref_read = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_read)


def ref_net(name):
    weights, biases = ref_read.readNNet(os.path.join(REF, "bench/rand", name))
    xdims = [weights[0].shape[1]] + [w.shape[0] for w in weights]
    Ms = [np.concatenate([w, b[:, None]], axis=1) for w, b in zip(weights, biases)]
    return o.FeedFwdNet(xdims=xdims, Ms=Ms)


def pack_net(net):
    d = {"xdims": np.asarray(net.xdims, dtype=np.int64)}
    for k, M in enumerate(net.Ms):
        d[f"M{k}"] = M
    return d


def pack_cliques(cliques):
    d = {"ncliques": np.int64(len(cliques))}
    for k, (Ck, parts, ds) in enumerate(cliques):
        d[f"Ck{k}"] = Ck
        for i, p in enumerate(parts):
            d[f"Ck{k}_part{i}"] = p
        for i, p in enumerate(ds):
            d[f"Dk{k}_{i}"] = p
    return d


def query_outputs(net, beta, q, tag):
    lit = o.run_query(net, beta, q, form="literal")
    clo = o.run_query(net, beta, q, form="closed")
    assert np.abs(lit["Z"] - clo["Z"]).max() <= 1e-13 * np.abs(lit["Z"]).max()
    info = lit["intv"]
    return {
        f"{tag}_x1min": q.x1min, f"{tag}_x1max": q.x1max, f"{tag}_gin": q.gin, f"{tag}_gbnd": q.gbnd,
        f"{tag}_gsec": q.gsec,
        f"{tag}_xmin": np.concatenate([p[0] for p in info.x_intvs]),
        f"{tag}_xmax": np.concatenate([p[1] for p in info.x_intvs]),
        f"{tag}_acxmin": np.concatenate([p[0] for p in info.acx_intvs]),
        f"{tag}_acxmax": np.concatenate([p[1] for p in info.acx_intvs]),
        f"{tag}_smin": lit["qc_sector"].smin, f"{tag}_smax": lit["qc_sector"].smax,
        f"{tag}_Z": lit["Z"],
    }


def config1():
    """BASELINE.json configs[0]: scale-I2-O2-W10-D10, beta = 1, box [0.5,1.5]^2, safety hplaneS([1,0], h)
    and a reach ellipsoid (invP = I, yc = f(centre))."""
    net = ref_net("scale-I2-O2-W10-D10.nnet")
    beta = 1
    rng = np.random.default_rng(1234)
    ac = net.acdim
    x1min, x1max = np.array([0.5, 0.5]), np.array([1.5, 1.5])
    gin, gbnd = rng.random(2), rng.random(ac)
    gsec = rng.random(o.sector_lambda_dim(ac, beta) + 2 * ac)
    gout = rng.random(1)
    h = 1.0
    S = o.hplaneS(np.array([1.0, 0.0]), h, net)
    yc = o.eval_feed_fwd_net(net, 0.5 * (x1min + x1max))
    out = pack_net(net)
    out["beta"] = np.int64(beta)
    out["S"] = S
    out["yc"] = yc
    out["gout"] = gout
    q = o.NumericQuery(x1min=x1min, x1max=x1max, gin=gin, gbnd=gbnd, gsec=gsec, qc_out=o.QcSafety(S=S))
    out.update(query_outputs(net, beta, q, "safety"))
    q = o.NumericQuery(x1min=x1min, x1max=x1max, gin=gin, gbnd=gbnd, gsec=gsec,
                       qc_out=o.QcReachEllipsoid(invP=np.eye(2), yc=yc), gout=gout)
    out.update(query_outputs(net, beta, q, "ellipsoid"))
    # a tight box so some neurons are stably active (Gram term exercised)
    c = np.array([1.0, 1.0])
    q = o.NumericQuery(x1min=c - 1e-3, x1max=c + 1e-3, gin=gin, gbnd=gbnd, gsec=gsec, qc_out=o.QcSafety(S=S))
    out.update(query_outputs(net, beta, q, "tight"))
    out.update(pack_cliques(o.make_cliques(net, beta)))
    np.savez_compressed(os.path.join(HERE, "config1_W10_D10_beta1.npz"), **out)


def config3():
    """BASELINE.json configs[2]: reach-I2-O2-W20-D10, 64 hyperplane directions (src/NnSdp.jl:81-83),
    shared bounds and multipliers, per-direction gamma_out.  Z is stored for two directions, the
    clique blocks of all 64 are pinned through their Frobenius norms and affine columns."""
    net = ref_net("reach-I2-O2-W20-D10.nnet")
    beta = 2
    rng = np.random.default_rng(1234)
    ac = net.acdim
    x1min, x1max = np.array([0.9, 0.9]), np.array([1.1, 1.1])
    gin, gbnd = rng.random(2), rng.random(ac)
    gsec = rng.random(o.sector_lambda_dim(ac, beta) + 2 * ac)
    gout = rng.random(64)
    theta = 2.0 * np.pi * np.arange(64) / 64
    normals = np.stack([np.cos(theta), np.sin(theta)], axis=1)
    out = pack_net(net)
    out.update(beta=np.int64(beta), normals=normals, gout=gout)
    cliques = o.make_cliques(net, beta)
    fro = np.zeros((64, len(cliques)))
    aff = np.zeros((64, net.Zdim))
    for i in range(64):
        q = o.NumericQuery(x1min=x1min, x1max=x1max, gin=gin, gbnd=gbnd, gsec=gsec,
                           qc_out=o.QcReachHplane(normals[i]), gout=gout[i:i + 1])
        r = o.run_query(net, beta, q, form="literal")
        fro[i] = [np.linalg.norm(b) for b in r["blocks"]]
        aff[i] = r["Z"][:, -1]
        if i in (0, 17):
            out.update(query_outputs(net, beta, q, f"dir{i}"))
    out.update(block_fro=fro, affine_cols=aff)
    out.update(pack_cliques(cliques))
    np.savez_compressed(os.path.join(HERE, "config3_reach_W20_D10_beta2.npz"), **out)


def tiny():
    net = ref_net("scale-I2-O2-W5-D5.nnet")
    out = pack_net(net)
    stats = {}
    for name in ("scale-I2-O2-W20-D10.nnet", "reach-I2-O2-W20-D10.nnet"):
        n = ref_net(name)
        stats[name] = float(np.std(np.concatenate([M[:, :-1].ravel() for M in n.Ms[1:-1]])))
    out["std_scale_W20"] = np.float64(stats["scale-I2-O2-W20-D10.nnet"])   # 2/sqrt(W ln W), make_networks.jl:44
    out["std_reach_W20"] = np.float64(stats["reach-I2-O2-W20-D10.nnet"])   # 1/sqrt(2),      make_networks.jl:50
    np.savez_compressed(os.path.join(HERE, "scale_W5_D5_weights.npz"), **out)
    shutil.copyfile(os.path.join(REF, "bench/rand/scale-I2-O2-W5-D5.nnet"), os.path.join(HERE, "scale-I2-O2-W5-D5.nnet"))
    os.chmod(os.path.join(HERE, "scale-I2-O2-W5-D5.nnet"), 0o644)


if __name__ == "__main__":
    config1()
    config3()
    tiny()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
