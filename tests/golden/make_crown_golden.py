#!/usr/bin/env python
"""CROWN interval fixtures computed by the REFERENCE'S OWN dependency: the auto_LiRPA tree vendored under
/root/reference/exts/auto_LiRPA, imported from there and driven exactly as the reference drives it
(exts/auto_lirpa_bridge.py:97-112 find_bounds_output, called per slice by
src/Intervals/intervals_auto_lirpa.jl:12-63 intervalsAutoLirpaSliced).  Run in the dev container:

    python tests/golden/make_crown_golden.py        ->  tests/golden/crown_autolirpa.npz

Nothing of the reference is modified or copied.  The 2021 auto_LiRPA does not import under numpy 2 /
Python 3.12 / torch 2.11 as is; four shims are installed IN THIS PROCESS before the import (names that moved
or were removed upstream, no numerics involved):
  * numpy.lib.arraysetops.isin          -> numpy.isin
  * appdirs.user_data_dir               -> a temp dir (the package only builds a cache path from it)
  * collections.Sequence & co           -> collections.abc
  * torch.onnx.symbolic_helper._set_opset_version and torch._C.Node.__getitem__ (used while tracing the
    torch module into auto_LiRPA's graph)
onnx / onnx2pytorch are absent, so the torch module the bridge obtains from the ONNX file
(nn.Sequential of Linear / ReLU, auto_lirpa_bridge.py:14-43) is built directly from the same weights.

Two precisions are stored: float32 (what the reference runs: the bridge calls .float()) and float64 (the same
code on double tensors).  The oracle's float64 restatement (oracle/nnsdp_oracle.py intervals_crown) and the
device kernels must match the float64 vectors to rounding and the float32 vectors to float32 rounding.
"""
import collections
import collections.abc
import os
import sys
import tempfile
import types
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

_m = types.ModuleType("numpy.lib.arraysetops")
_m.isin = np.isin
sys.modules["numpy.lib.arraysetops"] = _m
_ad = types.ModuleType("appdirs")
_ad.user_cache_dir = _ad.user_data_dir = lambda *a, **k: tempfile.gettempdir()
sys.modules["appdirs"] = _ad
for _n in ("Sequence", "Mapping", "Iterable", "MutableMapping", "Callable"):
    if not hasattr(collections, _n):
        setattr(collections, _n, getattr(collections.abc, _n))
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.onnx.symbolic_helper as _sh  # noqa: E402

if not hasattr(_sh, "_set_opset_version"):
    def _set_opset_version(v):
        from torch.onnx._internal.torchscript_exporter._globals import GLOBALS
        GLOBALS.export_onnx_opset_version = v
    _sh._set_opset_version = _set_opset_version
if not hasattr(torch._C.Node, "__getitem__"):
    torch._C.Node.__getitem__ = lambda self, k: getattr(self, self.kindOf(k))(k)
sys.path.insert(0, os.path.join(REF, "exts"))
from auto_LiRPA import BoundedModule, BoundedTensor  # noqa: E402
from auto_LiRPA.perturbations import PerturbationLpNorm  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nnsdp_oracle as o  # noqa: E402
from helpers import rand_net  # noqa: E402


def torch_module(Ms, dtype):
    layers = []
    for i, Mk in enumerate(Ms):
        lin = nn.Linear(Mk.shape[1] - 1, Mk.shape[0])
        lin.weight.data = torch.tensor(Mk[:, :-1]).to(dtype)
        lin.bias.data = torch.tensor(Mk[:, -1]).to(dtype)
        layers.append(lin)
        if i < len(Ms) - 1:
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


def bounds_output(Ms, x1min, x1max, dtype):
    """find_bounds_output + the min/max post-processing of autoLirpaBoundsOutput"""
    lo = torch.tensor(np.asarray(x1min)).to(dtype).view(1, -1)
    hi = torch.tensor(np.asarray(x1max)).to(dtype).view(1, -1)
    xc = (hi + lo) / 2
    model = BoundedModule(torch_module(Ms, dtype), xc)
    lb, ub = model.compute_bounds(x=(BoundedTensor(xc, PerturbationLpNorm(x_L=lo, x_U=hi)),), method="CROWN")
    lb, ub = lb[0].detach().double().numpy(), ub[0].detach().double().numpy()
    lb = np.minimum(lb, ub)
    ub = np.maximum(lb, ub)
    return lb, ub


def intervals_sliced(net, x1min, x1max, dtype):
    """x_intvs of intervalsAutoLirpaSliced: slice k = layers 1..k followed by [I 0]; last slice = the net"""
    xs = [(np.asarray(x1min, float), np.asarray(x1max, float))]
    for k in range(1, net.K):
        n = net.Ms[k - 1].shape[0]
        xs.append(bounds_output(list(net.Ms[:k]) + [np.hstack([np.eye(n), np.zeros((n, 1))])], x1min, x1max, dtype))
    xs.append(bounds_output(list(net.Ms), x1min, x1max, dtype))
    return xs


CASES = {
    # the box of experiments/scale.jl:27-28 on the reference's shipped nets
    "scale_W10_D10": (lambda: o.load_nnet(os.path.join(REF, "bench/rand/scale-I2-O2-W10-D10.nnet")), [0.5, 0.5], [1.5, 1.5]),
    "scale_W5_D5": (lambda: o.load_nnet(os.path.join(HERE, "scale-I2-O2-W5-D5.nnet")), [0.5, 0.5], [1.5, 1.5]),
    # seeded random nets (tests/helpers.py rand_net): ACAS-shaped and ragged
    "rand_acas_5x50": (lambda: rand_net([5] + [50] * 6 + [5], seed=4), [-0.1] * 5, [0.15] * 5),
    "rand_ragged": (lambda: rand_net([3, 17, 9, 33, 4], seed=8), [0.0, -0.3, 0.2], [0.4, 0.1, 0.25]),
}


def main():
    out = {}
    for name, (mk, lo, hi) in CASES.items():
        net = mk()
        for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
            xs = intervals_sliced(net, lo, hi, dt)
            out[f"{name}.{tag}.lo"] = np.concatenate([x[0] for x in xs])
            out[f"{name}.{tag}.hi"] = np.concatenate([x[1] for x in xs])
        out[f"{name}.x1min"] = np.asarray(lo, float)
        out[f"{name}.x1max"] = np.asarray(hi, float)
        out[f"{name}.xdims"] = np.asarray(net.xdims)
        for k, Mk in enumerate(net.Ms):
            out[f"{name}.M{k}"] = Mk
        info = o.intervals_crown(np.asarray(lo, float), np.asarray(hi, float), net)
        ref_lo = np.concatenate([p[0] for p in info.x_intvs])
        ref_hi = np.concatenate([p[1] for p in info.x_intvs])
        sc = max(np.abs(ref_lo).max(), np.abs(ref_hi).max())
        for tag in ("f64", "f32"):
            d = max(np.abs(out[f"{name}.{tag}.lo"] - ref_lo).max(), np.abs(out[f"{name}.{tag}.hi"] - ref_hi).max())
            print(f"{name:16s} auto_LiRPA {tag} vs oracle intervals_crown: max abs diff {d:.3e} (scale {sc:.3g})")
    np.savez_compressed(os.path.join(HERE, "crown_autolirpa.npz"), **out)
    print("wrote", os.path.join(HERE, "crown_autolirpa.npz"))


if __name__ == "__main__":
    main()
