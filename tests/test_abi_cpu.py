"""CPU tests of the drop-in boundary: libnnsdp_b200.so loads and exports every symbol that
include/nnsdp_b200.h declares; the host-only integer entry points (sizes, makeCliques, emission plan)
are bit-exact against the oracle; compute entry points fail loudly without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import nnsdp_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nnsdp_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nnsdp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import nnsdp_b200._lib as L

    names = _declared()
    assert len(names) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.abspath(L.LIB_PATH)], capture_output=True, text=True,
                         check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    for n in names:
        assert n in exported, f"{n} declared in include/nnsdp_b200.h but not exported"
        assert n in L.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(L.PROTOTYPES) <= set(names)
    assert L.lib.nnsdp_version() >= 100


def test_struct_layouts_match_header():
    import nnsdp_b200._lib as L

    assert C.sizeof(L.Sizes) == 14 * 8
    # nnsdp_query_inputs: 13 (pointer, stride) pairs + (out_kind, reserved)
    assert C.sizeof(L.QueryInputs) == 13 * 16 + 8
    assert L.QueryInputs.out_kind.offset == 9 * 16
    assert L.QueryInputs.out_S.offset == 9 * 16 + 8


def test_no_cpu_fallback():
    """Without a device every compute entry point returns NNSDP_ERR_CUDA with a message."""
    import nnsdp_b200 as nb

    if nb.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(nb.NnsdpError) as e:
        nb.Context([0])
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_argument_and_assert_errors():
    import nnsdp_b200 as nb

    with pytest.raises(nb.NnsdpError) as e:  # length(xdims) >= 3, MyNeuralNetwork.jl:18
        nb.sizes_from_xdims([2, 2], 0)
    assert e.value.code == -5
    with pytest.raises(nb.NnsdpError) as e:  # 0 <= beta, activ_sector.jl:12
        nb.sizes_from_xdims([2, 3, 2], -1)
    assert e.value.code == -5
    with pytest.raises(nb.NnsdpError) as e:
        nb.sizes_from_xdims([2, 0, 2], 1)
    assert e.value.code == -1


def _check_cliques(xdims, beta):
    import nnsdp_b200 as nb

    net = o.FeedFwdNet(xdims=xdims, Ms=[np.zeros((xdims[k + 1], xdims[k] + 1)) for k in range(len(xdims) - 1)])
    ref = o.make_cliques(net, beta)
    got = nb.cliques_from_xdims(xdims, beta)
    sz = nb.sizes_from_xdims(xdims, beta)
    assert len(got) == len(ref) == sz["ncliques"]
    for (a, pa, da), (b, pb, db) in zip(got, ref):
        assert a.dtype == np.int64 and np.array_equal(a, b)
        assert len(pa) == len(pb) and all(np.array_equal(x, y) for x, y in zip(pa, pb))
        assert len(da) == len(db) and all(np.array_equal(x, y) for x, y in zip(da, db))
    assert sz["Zdim"] == net.Zdim and sz["acdim"] == net.acdim and sz["K"] == net.K
    assert sz["lamdim"] == o.sector_lambda_dim(net.acdim, beta)
    assert sz["secdim"] == sz["lamdim"] + 2 * net.acdim
    assert sz["sum_ck"] == sum(len(c[0]) for c in ref)
    assert sz["sum_ck_sq"] == sum(len(c[0]) ** 2 for c in ref)
    assert sz["max_ck"] == max(len(c[0]) for c in ref)
    return sz


@pytest.mark.parametrize("xdims,beta", [
    ([2, 3, 2], 0), ([2, 3, 2], 2), ([3, 3, 3, 3, 4, 3, 3], 2), ([2] + [10] * 10 + [2], 1), ([2] + [20] * 100 + [2], 2),
    ([5] + [50] * 6 + [5], 2), ([2, 4, 7, 3, 5, 2], 5), ([2] + [1000] * 20 + [2], 2), ([2, 70, 130, 64, 3], 7),
])
def test_cliques_bit_exact_named_configs(xdims, beta):
    sz = _check_cliques(xdims, beta)
    if xdims == [2] + [1000] * 20 + [2]:
        assert sz["ncliques"] == 19 and sz["sum_ck_sq"] * 8 == 1330657432 and sz["max_ck"] == 3003
    if xdims == [2] + [20] * 100 + [2]:   # config 2'
        assert sz["ncliques"] == 99 and sz["Zdim"] == 2003


@settings(max_examples=150, deadline=None)
@given(st.lists(st.integers(1, 9), min_size=3, max_size=9), st.integers(0, 12))
def test_cliques_bit_exact_random_shapes(xdims, beta):
    if beta > sum(xdims[1:-1]):
        return
    _check_cliques(xdims, beta)


@pytest.mark.parametrize("xdims", [[3, 3, 3, 3, 4, 3, 3], [2, 4, 7, 3, 5, 2], [2] + [10] * 10 + [2], [5] + [50] * 6 + [5]])
@pytest.mark.parametrize("beta", [0, 1, 2, 4])
def test_clique_cover_is_the_pattern_the_reference_draws(xdims, beta):
    """Pin against a statement by the reference's authors that is not makeCliques itself: the figures of
    experiments/plot_sparsity.ipynb (xdims [3,3,3,3,4,3,3], beta 0/2/4) draw quickRawZ(beta) with the x_K rows
    and columns filled in; the union of Ck x Ck over nnsdp_cliques must be exactly that set."""
    import nnsdp_b200 as nb

    want = o.chordal_extension_pattern_notebook(xdims, beta)
    cover = np.zeros_like(want)
    for Ck, _, _ in nb.cliques_from_xdims(xdims, beta):
        cover[np.ix_(Ck - 1, Ck - 1)] = True
    assert np.array_equal(cover, want)


@pytest.mark.parametrize("xdims,beta", [([2, 3, 2], 1), ([2] + [10] * 10 + [2], 1), ([5] + [50] * 6 + [5], 2),
                                        ([2, 70, 130, 64, 3], 2), ([2] + [1000] * 20 + [2], 2), ([2] + [100] * 50 + [2], 3)])
@pytest.mark.parametrize("dense", [False, True])
def test_emission_plan_covers_output_exactly_once(xdims, beta, dense):
    """Every output entry belongs to exactly one tile (entry counts add up to the output size)."""
    import nnsdp_b200 as nb

    sz = nb.sizes_from_xdims(xdims, beta)
    ps = nb.plan_stats(xdims, beta, dense=dense)
    total = sum(ps["entries"].values())
    assert total == (sz["Zdim"] ** 2 if dense else sz["sum_ck_sq"])
    assert ps["tile_rows"] in (32, 64, 128) and ps["tile_cols"] == 32


def test_shard_ranges_partition():
    from nnsdp_b200.dist import all_ranges, shard_range

    for Q in (1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            rs = all_ranges(Q, world)
            assert rs[0][0] == 0 and sum(n for _, n in rs) == Q
            for (a, n), (b, _) in zip(rs, rs[1:]):
                assert a + n == b
            assert max(n for _, n in rs) - min(n for _, n in rs) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


@pytest.mark.parametrize("xdims,beta,dense", [([2, 300, 280, 320, 290, 2], 2, False), ([3, 270, 300, 260, 3], 1, True),
                                              ([2] + [1000] * 20 + [2], 2, False), ([2] + [10] * 10 + [2], 1, False)])
def test_gather_plan_accounts_for_every_possible_nonzero(xdims, beta, dense):
    """Host-gather invariant: an output entry that the emission plan does not classify as structurally zero is
    covered by a dense cell (copied whole, always or conditionally) or is in the thin list; cells partition
    every matrix; thin entries never lie inside always-dense cells (they would travel twice)."""
    import nnsdp_b200 as nb

    gp = nb.gather_plan(xdims, beta, dense=dense)
    tiles = nb.plan_tiles(xdims, beta, dense=dense)
    sz = nb.sizes_from_xdims(xdims, beta)
    if dense:
        ns = [sz["Zdim"]]
    else:
        ns = [len(c[0]) for c in nb.cliques_from_xdims(xdims, beta)]
    offs = np.concatenate([[0], np.cumsum([n * n for n in ns])])
    cells = gp["cells"]
    # cells partition every matrix
    for m, n in enumerate(ns):
        cm = cells[cells[:, 0] == m]
        assert int((cm[:, 2].astype(np.int64) * cm[:, 4]).sum()) == n * n
    if not gp["usable"]:
        assert max(xdims) < 256 or len(gp["thin"]) * 8 > offs[-1]
        return
    if sum(n * n for n in ns) > 5e7:   # the stress config: check counts only
        assert 400_000 < len(gp["thin"]) < 1_500_000 and np.all(np.diff(gp["thin"]) != 0)
        always = cells[cells[:, 5] == 1]
        assert 8 * int((always[:, 2].astype(np.int64) * always[:, 4]).sum()) > 0.2 * 8 * offs[-1]
        return
    covered = np.zeros(offs[-1], dtype=np.int8)       # 1 = always dense, 2 = conditionally dense
    for m, r0, nr, c0, nc, kind, blk, pure in cells:
        if kind == 0:
            continue
        blkv = covered[offs[m]:offs[m + 1]].reshape(ns[m], ns[m])   # [col, row]
        blkv[c0:c0 + nc, r0:r0 + nr] = 1 if kind == 1 else 2
    thin = np.zeros(offs[-1], dtype=bool)
    thin[gp["thin"]] = True
    assert not np.any(thin & (covered == 1))
    for t in tiles:
        m, r0, nr, c0, nc, prog = t[0], t[1], t[2], t[3], t[4], t[10]
        if prog == 0:      # ZERO
            continue
        sl = np.zeros((ns[m], ns[m]), dtype=bool)
        sl[c0:c0 + nc, r0:r0 + nr] = True
        idx = offs[m] + np.flatnonzero(sl.ravel())
        if prog in (1, 6):   # SAME / DIAG: Gram / S22 values live in conditional cells; only the band must be thin
            ok = (covered[idx] > 0) | thin[idx]
        else:
            ok = (covered[idx] == 1) | thin[idx]
        assert np.all(ok), (t, int((~ok).sum()))


def test_nnet_reader_against_the_references_reader():
    """nnsdp_nnet_read on the shipped fixture == the weights the reference's own exts/NNet/utils/readNNet.py
    produced (tests/golden/scale_W5_D5_weights.npz) == the oracle's reader."""
    import nnsdp_b200 as nb

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    xd, Ms = nb.read_nnet(os.path.join(gold, "scale-I2-O2-W5-D5.nnet"))
    d = np.load(os.path.join(gold, "scale_W5_D5_weights.npz"))
    assert xd == d["xdims"].tolist()
    for k, M in enumerate(Ms):
        assert np.array_equal(M, d[f"M{k}"])
    ref = o.load_nnet(os.path.join(gold, "scale-I2-O2-W5-D5.nnet"))
    assert all(np.array_equal(a, b) for a, b in zip(Ms, ref.Ms))
    with pytest.raises(nb.NnsdpError):
        nb.read_nnet(os.path.join(gold, "does-not-exist.nnet"))


# ---------------------------------------------------------------------------------------------
# vnnlib -> batched safety queries
# ---------------------------------------------------------------------------------------------
VNNLIB_CASES = {
    # shape of ACAS Xu property 1: a box and one output threshold
    "threshold": (5, 5, """; unsafe if COC >= 1500
(declare-const X_0 Real)
(declare-const X_1 Real)
(declare-const X_2 Real)
(declare-const X_3 Real)
(declare-const X_4 Real)
(declare-const Y_0 Real)
(assert (<= X_0 0.679857769))
(assert (>= X_0 0.6))
(assert (<= X_1 0.5))
(assert (>= X_1 -0.5))
(assert (<= X_2 0.5))
(assert (>= X_2 -0.5))
(assert (<= X_3 0.5))
(assert (>= X_3 0.45))
(assert (<= X_4 -0.45))
(assert (>= X_4 -0.5))
(assert (>= Y_0 3.991125645861615))
"""),
    # shape of property 3 / 4: "COC is minimal" as one conjunction, written over several lines
    "minimal": (2, 3, """(declare-const X_0 Real)
(declare-const X_1 Real)
(declare-const Y_0 Real)
(assert (<= X_0 0.6798))  ; upper
(assert (>= X_0 0.6))
(assert (<= X_1
   0.5))
(assert (>= X_1 -0.5))
(assert (or
  (and (<= Y_0 Y_1) (<= Y_0 Y_2))
  (and (>= Y_1 3.25)(<= 1.5 Y_2))
))
"""),
    # disjunction over input boxes times a disjunction over outputs (shape of property 6 / 7), repeated box
    "boxes": (2, 2, """(assert (>= X_1 -1e-1))
(assert (<= X_1 2.5E-1))
(assert (or (and (<= X_0 0.2)(>= X_0 0.0)) (and (<= X_0 0.9)(>= X_0 0.7)) (and (>= X_0 0.0)(<= X_0 0.2))))
(assert (or (and (<= Y_0 Y_1)) (and (<= Y_1 -0.75)(>= Y_0 0.125)(<= Y_0 Y_0))))
"""),
}


@pytest.mark.parametrize("name", sorted(VNNLIB_CASES))
def test_vnnlib_reader_against_the_restated_reference_parser(name, tmp_path):
    import nnsdp_b200 as nb

    n_in, n_out, text = VNNLIB_CASES[name]
    path = str(tmp_path / f"{name}.vnnlib")
    open(path, "w").write(text)
    net = o.FeedFwdNet(xdims=[n_in, 4, n_out], Ms=[np.zeros((4, n_in + 1)), np.zeros((n_out, 5))])
    cnf = o.load_vnnlib_cnf(path, net)
    got = nb.read_vnnlib(path, n_in, n_out)
    assert got["nclauses"] == len(cnf)
    flat = [(c, qi, qs) for c, clause in enumerate(cnf) for qi, qs in clause]
    assert len(flat) == len(got["clause"])
    for i, (c, qi, qs) in enumerate(flat):
        assert got["clause"][i] == c
        assert np.array_equal(got["x1min"][i], qi.x1min) and np.array_equal(got["x1max"][i], qi.x1max)
        S = np.asarray(qs.S)
        assert np.array_equal(got["S"][i], S) and np.array_equal(np.signbit(got["S"][i]), np.signbit(S))
        assert np.array_equal(got["S"][i], got["S"][i].T)
    if name == "boxes":     # 3 input alternatives, two of them the same box, x 2 output alternatives
        assert got["nclauses"] == 6 and len(flat) == 2 * (1 + 3) + (1 + 3)
        assert np.array_equal(got["x1min"][0], [0.0, -0.1]) and np.array_equal(got["x1max"][-1], [0.9, 0.25])


def test_vnnlib_reader_errors(tmp_path):
    import nnsdp_b200 as nb

    def run(text, n_in=1, n_out=1):
        p = str(tmp_path / "e.vnnlib")
        open(p, "w").write(text)
        return nb.read_vnnlib(p, n_in, n_out)

    with pytest.raises(nb.NnsdpError) as e:      # vnnlib_parser.jl:200: every input needs both bounds
        run("(assert (<= X_0 1.0))\n(assert (<= Y_0 0.0))\n")
    assert e.value.code == -5
    with pytest.raises(nb.NnsdpError) as e:      # :60 empty interval
        run("(assert (<= X_0 1.0))\n(assert (>= X_0 2.0))\n")
    assert e.value.code == -5
    with pytest.raises(nb.NnsdpError) as e:      # :52 index range
        run("(assert (<= X_3 1.0))\n")
    assert e.value.code == -5
    with pytest.raises(nb.NnsdpError) as e:
        run("(check-sat)\n")
    assert e.value.code == -1
    with pytest.raises(nb.NnsdpError) as e:
        run("(assert (and (<= Y_0 1.0)))\n")
    assert e.value.code == -1
    with pytest.raises(nb.NnsdpError):
        nb.read_vnnlib(str(tmp_path / "missing.vnnlib"), 1, 1)
    ok = run("(assert (<= X_0 1.0))\n(assert (>= X_0 0.0)) ; both\n(assert (<= Y_0 0.5))\n")
    assert ok["nclauses"] == 1 and ok["S"].shape == (1, 3, 3) and ok["S"][0, 2, 2] == -2 * (-0.5 - 1e-4)


# ---------------------------------------------------------------------------------------------
# the (never executed) Julia wrapper against the header: every ccall names a declared symbol with the right
# number and kinds of arguments
# ---------------------------------------------------------------------------------------------
def _c_kind(p):
    p = re.sub(r"\b(const|struct)\b", " ", p)
    p = re.sub(r"\s+", " ", p).strip()
    m = re.match(r"^([A-Za-z_0-9]+)\s*(\*+)?\s*(?:\*\s*)?[A-Za-z_0-9]*$", p.replace("* *", "**").replace("* ", "*"))
    stars = p.count("*")
    base = p.replace("*", " ").split()[0]
    scalar = {"int32_t": "i32", "int64_t": "i64", "double": "f64", "float": "f32", "char": "char", "uint64_t": "u64"}
    if stars == 0:
        return scalar.get(base, "struct")
    if base == "char" and stars == 1:
        return "str"
    if base in scalar:
        return "p" * stars + "_" + scalar[base]
    return "p" * stars + "_struct"


def _jl_kind(t):
    t = t.strip()
    scalar = {"Int32": "i32", "Int64": "i64", "Float64": "f64", "Float32": "f32", "UInt64": "u64", "Cstring": "str"}
    if t in scalar:
        return scalar[t]
    depth = 0
    while t.startswith("Ptr{") or t.startswith("Ref{"):
        t = t[4:-1]
        depth += 1
    if t in scalar:
        return "p" * depth + "_" + scalar[t]
    return "p" * depth + "_struct"      # Cvoid (opaque handles) and mirrored structs


def test_julia_wrapper_ccalls_match_the_header():
    jl = open(os.path.join(ROOT, "nn-sdp_b200", "julia", "NnSdpB200.jl")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b([a-z_0-9 ]+?\*?)\s*\b(nnsdp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        ret, name, params = m.groups()
        params = params.replace("\n", " ").strip()
        protos[name] = (ret.strip(), [] if params in ("", "void") else [p.strip() for p in params.split(",")])
    calls = []
    for m in re.finditer(r"ccall\(\(:(nnsdp_[a-z0-9_]+), LIB\),\s*([A-Za-z0-9{}]+),\s*\(", jl):
        j = k = m.end()
        depth = 1
        while depth:
            depth += (jl[k] == "(") - (jl[k] == ")")
            k += 1
        parts, cur, d = [], "", 0
        for c in jl[j:k - 1]:
            d += (c in "{(") - (c in "})")
            if c == "," and d == 0:
                parts.append(cur.strip())
                cur = ""
            else:
                cur += c
        if cur.strip():
            parts.append(cur.strip())
        calls.append((m.group(1), m.group(2), parts))
    assert len(calls) >= 20
    for name, ret, parts in calls:
        assert name in protos, f"{name} is not declared in include/nnsdp_b200.h"
        cret, cparams = protos[name]
        assert _jl_kind(ret) == _c_kind(cret), (name, ret, cret)
        assert len(parts) == len(cparams), (name, parts, cparams)
        for jt, cp in zip(parts, cparams):
            assert _jl_kind(jt) == _c_kind(cp), (name, jt, cp)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 4), st.integers(1, 4), st.integers(0, 2**31 - 1))
def test_vnnlib_reader_random_properties(n_in, n_out, seed):
    """Random files in the supported subset (boxes, simple output asserts, up to two disjunctions mixing input and
    output constraints, comments, statements broken over lines): library == restated reference parser, bit for bit."""
    import tempfile

    import nnsdp_b200 as nb

    rng = np.random.default_rng(seed)
    num = lambda v: rng.choice([f"{v:.6g}", f"{v:.3e}", repr(float(v))])      # noqa: E731
    yterm = lambda: f"Y_{rng.integers(n_out)}"                                  # noqa: E731
    lines = [f"(declare-const X_{i} Real)" for i in range(n_in)] + [f"(declare-const Y_{i} Real)" for i in range(n_out)]
    for i in range(n_in):
        lines.append(f"(assert (<= X_{i} {num(rng.uniform(0.5, 1.0))}))" + ("  ; upper" if rng.random() < 0.3 else ""))
        lines.append(f"(assert (>= X_{i} {num(rng.uniform(-1.0, -0.5))}))")

    def ycmp():
        op = rng.choice(["<=", ">="])
        kind = rng.integers(3)
        if kind == 0:
            return f"({op} {yterm()} {yterm()})"
        if kind == 1:
            return f"({op} {yterm()} {num(rng.normal())})"
        return f"({op} {num(rng.normal())} {yterm()})"

    for _ in range(rng.integers(0, 3)):
        lines.append(f"(assert {ycmp()})")
    have_y = False
    for _ in range(rng.integers(0, 3)):
        alts = []
        for _ in range(rng.integers(1, 4)):
            cs = [ycmp() for _ in range(rng.integers(1, 3))]
            if rng.random() < 0.5:
                i = rng.integers(n_in)
                a, b = sorted(rng.uniform(-0.4, 0.4, 2))
                cs += [f"(<= X_{i} {num(b)})", f"(>= X_{i} {num(a)})"]
            sep = rng.choice(["", " ", "\n    "])
            alts.append("(and " + sep.join(cs) + ")")
        have_y = True
        lines.append("(assert (or " + rng.choice(["", " ", "\n  "]).join(alts) + "))")
    if not have_y:
        lines.append(f"(assert {ycmp()})")
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "p.vnnlib")
        open(path, "w").write("\n".join(lines) + "\n")
        net = o.FeedFwdNet(xdims=[n_in, 3, n_out], Ms=[np.zeros((3, n_in + 1)), np.zeros((n_out, 4))])
        try:
            cnf = o.load_vnnlib_cnf(path, net)
        except AssertionError:      # two disjunctions emptied an input interval: the reader's assert (:60)
            with pytest.raises(nb.NnsdpError) as e:
                nb.read_vnnlib(path, n_in, n_out)
            assert e.value.code == -5
            return
        got = nb.read_vnnlib(path, n_in, n_out)
    flat = [(c, qi, qs) for c, clause in enumerate(cnf) for qi, qs in clause]
    assert got["nclauses"] == len(cnf) and len(flat) == len(got["clause"])
    for i, (c, qi, qs) in enumerate(flat):
        assert got["clause"][i] == c
        assert np.array_equal(got["x1min"][i], qi.x1min) and np.array_equal(got["x1max"][i], qi.x1max)
        S = np.asarray(qs.S)
        assert np.array_equal(got["S"][i], S) and np.array_equal(np.signbit(got["S"][i]), np.signbit(S))


# ---------------------------------------------------------------------------------------------
# Panel-ordered work list of the emitter (csrc/plan.cpp, emit_panel_kernel): dense formats of wide nets write the
# clique blocks Z[C_k, C_k] (the blocks setupZksum! scatters, /root/reference/src/Methods/chordal_sdp.jl:60-93) with the
# fill strips and the window tiles of one 32-column panel as consecutive work items of ONE launch.
# ---------------------------------------------------------------------------------------------
PANEL_SHAPES = [([2, 300, 270, 280, 2], 2, 0), ([3, 150, 260, 140, 2], 4, 0), ([2, 129, 128, 127, 300, 4], 3, 0),
                ([2, 300, 270, 280, 2], 2, 1), ([2] + [1000] * 4 + [2], 2, 0)]


@pytest.mark.parametrize("xdims,beta,dense", PANEL_SHAPES)
def test_panel_work_list_covers_the_fill_and_window_entries_once(xdims, beta, dense):
    import nnsdp_b200 as nb

    items = nb.plan_panel(xdims, beta, dense=dense)
    tiles = nb.plan_tiles(xdims, beta, dense=dense)
    F = {n: i for i, n in enumerate(nb.core.TILE_FIELDS)}
    P = {n: i for i, n in enumerate(nb.core.PANEL_FIELDS)}
    assert len(items) > 0
    PROG = dict(ZERO=0, SAME=1, RC=2, CR=3, MIXED=4, GENERAL=5, DIAG=6, AFF=7)
    names = {v: k for k, v in PROG.items()}
    st = nb.plan_stats(xdims, beta, dense=dense)
    # the matrices of the output: offsets and sides from the tile list (mat -> local extents)
    nmat = int(tiles[:, F["mat"]].max()) + 1
    side = [int((tiles[tiles[:, F["mat"]] == m][:, F["row0"]] + tiles[tiles[:, F["mat"]] == m][:, F["nrows"]]).max()) for m in range(nmat)]
    offs = np.concatenate([[0], np.cumsum([n * n for n in side])])
    paint = [np.zeros((n, n), dtype=np.int8) for n in side]
    ent = {k: 0 for k in PROG}
    last_key = None
    for it in items:
        off, ld, r0, nr, c0, nc, prog = (int(it[P[k]]) for k in ("out_off", "ld", "row0", "nrows", "col0", "ncols", "prog"))
        m = int(np.searchsorted(offs, off, side="right") - 1)
        assert offs[m] == off and ld == side[m]
        assert 0 <= r0 and r0 + nr <= side[m] and 0 <= c0 and c0 + nc <= side[m]
        paint[m][r0:r0 + nr, c0:c0 + nc] += 1
        ent[names[prog]] += nr * nc
        key = (off, c0 // 32, c0, r0)          # the sort order: matrix, panel, first column, first row
        assert last_key is None or key >= last_key
        last_key = key
    # every entry of the fill and window classes exactly once, nothing of the edge class
    for k in ("ZERO", "RC", "CR", "AFF"):
        assert ent[k] == st["entries"][k], k
    assert ent["SAME"] + ent["DIAG"] == st["entries"]["SAME"] + st["entries"]["DIAG"]     # merged into DIAG strips
    assert ent["MIXED"] == ent["GENERAL"] == 0
    total = 0
    for m in range(nmat):
        assert paint[m].max() <= 1
        total += int(paint[m].sum())
    assert total + st["entries"]["MIXED"] + st["entries"]["GENERAL"] == sum(n * n for n in side)
    # the entries the panel list leaves out are exactly the tiles of the edge class
    for t in tiles:
        if names[int(t[F["prog"]])] in ("MIXED", "GENERAL"):
            m, r0, nr, c0, nc = (int(t[F[k]]) for k in ("mat", "row0", "nrows", "col0", "ncols"))
            assert paint[m][r0:r0 + nr, c0:c0 + nc].sum() == 0


@pytest.mark.parametrize("xdims,beta,dense", [([2, 100, 100, 100, 2], 2, 0), ([2, 10, 10, 2], 1, 0),
                                              ([2, 300, 270, 280, 2], 2, 2), ([3, 150, 260, 140, 2], 6, 0)])
def test_panel_work_list_is_empty_where_the_kernels_stay_separate(xdims, beta, dense):
    """Narrow layers (inline band), small nets, packed records, beta beyond the window programs."""
    import nnsdp_b200 as nb

    assert len(nb.plan_panel(xdims, beta, dense=dense)) == 0


def test_library_holds_the_blackwell_instructions_the_design_names():
    """SASS of the built library (no device needed): the CR window program stages its W tile by TMA (UTMALDG.2D against
    an mbarrier: SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK...TRYWAIT), the FP64 contractions run on the tensor cores
    (DMMA.8x8x4, the only FP64 tensor instruction of sm_100a) fed by cp.async (LDGSTS), and every kernel is compiled for
    sm_100a only."""
    import shutil
    import subprocess

    import nnsdp_b200._lib as L

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lib = os.path.abspath(L.LIB_PATH)
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, timeout=300).stdout
    assert sass.count("UTMALDG.2D") >= 5                      # emit_window_kernel<0..4> and emit_panel_kernel<0..4>
    assert "SYNCS.ARRIVE.TRANS64" in sass and "TRYWAIT" in sass
    assert sass.count("DMMA.8x8x4") >= 100                    # gram_kernel, dgemm tiles, ibp_dmma_kernel
    assert "LDGSTS" in sass
    archs = set(ln.split("=")[1].strip() for ln in sass.splitlines() if ln.strip().startswith("arch ="))
    assert archs == {"sm_100a"}, archs
    for name in ("emit_panel_kernel", "emit_fill_kernel", "emit_window_kernel", "emit_edge_kernel", "emit_band_kernel",
                 "gram_kernel", "ibp_dmma_kernel", "dgemm_dmma_affine_layers_kernel", "crown_post_kernel"):
        assert name in sass, name
