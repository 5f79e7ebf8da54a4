"""Packed records (NNSDP_FORMAT_PACKED) on the host: the cell table, the emission plan re-addressed into the cells,
and nnsdp_packed_unpack -- all integer / copy work that runs without a device.

The scatter structure the cells must reproduce is the reference's: Z[C_k, C_k] for the cliques of makeCliques
(/root/reference/src/Methods/chordal_cliques.jl:13-59) as setupZksum! scatters them
(/root/reference/src/Methods/chordal_sdp.jl:60-93).  The oracle supplies a dense Z(gamma); a numpy "packer" fills a
record from it by the cell table alone, and unpacking must give back every clique block bit for bit.
"""
import numpy as np
import pytest

import nnsdp_oracle as o
from helpers import rand_net, rand_query

SHAPES = [
    ([2, 3, 3, 2], 1),
    ([3, 3, 3, 3, 4, 3, 3], 2),
    ([2, 10, 10, 10, 10, 2], 3),
    ([2, 4, 7, 3, 5, 2], 5),
    ([5, 50, 50, 50, 50, 50, 50, 5], 2),
    ([2, 70, 130, 64, 3], 2),
    ([2, 100, 100, 100, 100, 2], 0),
    ([2, 100, 100, 100, 100, 2], 1),
    ([2, 60, 48, 200, 64, 2], 4),
    ([2, 300, 260, 2], 3),
    ([2, 300, 300, 300, 2], 2),
]


def _pack_numpy(Z, lay, beta, present):
    rec = np.full(lay["record_doubles"], np.nan)
    for i, c in enumerate(lay["cells"]):
        if not present[i]:
            continue
        r0, c0, nr, nc, off = int(c["row0"]) - 1, int(c["col0"]) - 1, int(c["nrows"]), int(c["ncols"]), int(c["offset"])
        if c["kind"] == 3:  # BAND: band[t + (beta+1) i] = Z[g0+i, g0+i+t]
            assert nr == beta + 1
            b = np.zeros((nc, nr))
            for t in range(beta + 1):
                idx = np.arange(nc - t)
                b[idx, t] = Z[r0 + idx, r0 + idx + t]
            rec[off:off + nr * nc] = b.reshape(-1)
        else:
            blk = Z[r0:r0 + nr, c0:c0 + nc].copy()
            # only the upper triangle is defined: poison what lies strictly below the diagonal of Z
            gr = (r0 + np.arange(nr))[:, None]
            gc = (c0 + np.arange(nc))[None, :]
            blk[gr > gc] = np.nan
            rec[off:off + nr * nc] = blk.reshape(-1, order="F")
    return rec


@pytest.mark.parametrize("xdims,beta", SHAPES)
def test_layout_is_well_formed(xdims, beta):
    import nnsdp_b200 as nb

    lay = nb.packed_layout(xdims, beta)
    cells, rec, alw = lay["cells"], lay["record_doubles"], lay["always_doubles"]
    Zdim = sum(xdims[:-1]) + 1
    assert len(cells) >= 1 and 0 < alw <= rec
    ends = []
    for c in cells:
        assert c["offset"] % 16 == 0                       # 128-byte aligned cells
        assert 1 <= c["row0"] and c["row0"] + (c["ncols"] if c["kind"] == 3 else c["nrows"]) - 1 <= Zdim
        assert 1 <= c["col0"] and c["col0"] + c["ncols"] - 1 <= Zdim
        ends.append((int(c["offset"]), int(c["offset"] + c["nrows"] * c["ncols"]), int(c["always"])))
    ends.sort()
    for (a0, a1, _), (b0, _, _) in zip(ends, ends[1:]):
        assert a1 <= b0                                     # cells do not overlap inside the record
    assert ends[-1][1] <= rec
    # the always-written cells come first
    assert all(e[1] <= alw for e in ends if e[2]) and all(e[0] >= alw for e in ends if not e[2])
    # optional cells are DIAG cells of hidden blocks, each with a BAND cell over the same range
    for c in cells[cells["always"] == 0]:
        assert c["kind"] == nb.CELL_DIAG and c["row0"] == c["col0"] and c["nrows"] == c["ncols"]
        b = cells[(cells["kind"] == nb.CELL_BAND) & (cells["blk"] == c["blk"])]
        assert len(b) == 1 and b[0]["row0"] == c["row0"] and b[0]["ncols"] == c["nrows"] and b[0]["nrows"] == beta + 1


@pytest.mark.parametrize("xdims,beta", SHAPES)
def test_plan_tiles_cover_their_cells(xdims, beta):
    """Every tile of the packed emission plan lies inside its cell, no entry is written by two tiles, every entry
    of an always-written cell is written, and the upper triangle of every DIAG cell is covered."""
    import nnsdp_b200 as nb

    lay = nb.packed_layout(xdims, beta)
    cells = lay["cells"]
    tiles = nb.plan_tiles(xdims, beta, dense=2)
    F = {n: i for i, n in enumerate(nb.core.TILE_FIELDS)}
    paint = [np.zeros((int(c["nrows"]), int(c["ncols"])), dtype=np.int32) for c in cells]
    for t in tiles:
        ci = int(t[F["mat"]])
        c = cells[ci]
        r0, nr, c0, nc = (int(t[F[k]]) for k in ("row0", "nrows", "col0", "ncols"))
        assert 0 <= r0 and r0 + nr <= c["nrows"] and 0 <= c0 and c0 + nc <= c["ncols"]
        assert int(t[F["grow0"]]) == c["row0"] - 1 + r0 and int(t[F["gcol0"]]) == c["col0"] - 1 + c0
        assert not int(t[F["grow0"]]) > int(t[F["gcol0"]]) + nc - 1      # nothing strictly below the diagonal
        paint[ci][r0:r0 + nr, c0:c0 + nc] += 1
    for c, p in zip(cells, paint):
        if c["kind"] == nb.CELL_BAND:
            assert p.sum() == 0                                          # written by the band kernel
            continue
        assert p.max() <= 1
        gr = (int(c["row0"]) + np.arange(int(c["nrows"])))[:, None]
        gc = (int(c["col0"]) + np.arange(int(c["ncols"])))[None, :]
        assert np.all(p[gr <= gc] == 1), (c, "upper-triangle entries not covered")


@pytest.mark.parametrize("xdims,beta", SHAPES)
@pytest.mark.parametrize("kind", ["safety", "hplane"])
def test_unpack_reproduces_every_clique_block(xdims, beta, kind):
    import nnsdp_b200 as nb

    net = rand_net(xdims, seed=3)
    rng = np.random.default_rng(17)
    for radius in (0.0, 0.3):           # radius 0: every ReLU stable (all Gram blocks active); 0.3: few or none
        q = rand_query(net, beta, rng, kind=kind, radius=radius)
        ref = o.run_query(net, beta, q)
        # the oracle's Z is symmetric to rounding only; a record holds the upper triangle, so compare with that mirrored
        Z = np.triu(ref["Z"]) + np.triu(ref["Z"], 1).T
        lay = nb.packed_layout(xdims, beta)
        cells = lay["cells"]
        # a DIAG cell may be absent when its range of Z is zero apart from what other cells hold
        present = np.ones(len(cells), dtype=np.uint8)
        for i, c in enumerate(cells):
            if c["always"]:
                continue
            r0, m = int(c["row0"]) - 1, int(c["nrows"])
            D = Z[r0:r0 + m, r0:r0 + m].copy()
            ii = np.arange(m)
            for t in range(-beta, beta + 1):
                idx = ii[(ii + t >= 0) & (ii + t < m)]
                D[idx, idx + t] = 0.0
            present[i] = 1 if np.any(D != 0.0) else 0
        rec = _pack_numpy(Z, lay, beta, present)
        flat = nb.packed_unpack(xdims, beta, rec, present)
        off = 0
        for (Ck, _, _), blk in zip(ref["cliques"], o.clique_blocks(Z, ref["cliques"])):
            n = len(Ck)
            mine = flat[off:off + n * n].reshape(n, n).T
            assert np.array_equal(mine, blk), (xdims, beta, radius)
            off += n * n
        assert off == flat.size
        Zu = nb.packed_unpack(xdims, beta, rec, present, dense_Z=True).reshape(Z.shape).T
        assert np.array_equal(Zu, Z)
    # the record is much smaller than the dense blocks for wide nets
    sz = nb.sizes_from_xdims(xdims, beta)
    if min(xdims[1:-1]) >= 100 and sz["ncliques"] >= 3:
        assert lay["record_doubles"] < 0.6 * sz["sum_ck_sq"]


def _random_shapes(n=40, seed=5):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        depth = int(rng.integers(2, 6))
        wide = rng.random() < 0.7
        widths = [int(rng.integers(48, 300)) if wide else int(rng.integers(2, 60)) for _ in range(depth)]
        if wide and rng.random() < 0.3:
            widths[int(rng.integers(0, depth))] = int(rng.integers(48, 64))      # a layer just above the split threshold
        xdims = [int(rng.integers(1, 6))] + widths + [int(rng.integers(1, 5))]
        beta = int(rng.integers(0, min(7, sum(widths)) + 1))
        out.append((xdims, beta))
    return out


@pytest.mark.parametrize("xdims,beta", _random_shapes())
def test_random_shapes_layout_coverage_and_unpack(xdims, beta):
    """Seeded random shapes around the planner's thresholds (48-neuron split, 128-row tiles, 256-wide band mode,
    beta up to 7 and beyond MAX_WINDOW_BETA): the cell table is well formed, the tiles cover their cells, and a
    record packed from the oracle's Z unpacks to every clique block."""
    import nnsdp_b200 as nb

    test_layout_is_well_formed(xdims, beta)
    test_plan_tiles_cover_their_cells(xdims, beta)
    net = rand_net(xdims, seed=1)
    rng = np.random.default_rng(2)
    q = rand_query(net, beta, rng, kind="ellipsoid", radius=0.0)
    ref = o.run_query(net, beta, q)
    Z = np.triu(ref["Z"]) + np.triu(ref["Z"], 1).T
    lay = nb.packed_layout(xdims, beta)
    present = np.ones(len(lay["cells"]), dtype=np.uint8)
    flat = nb.packed_unpack(xdims, beta, _pack_numpy(Z, lay, beta, present), present)
    off = 0
    for (Ck, _, _), blk in zip(ref["cliques"], o.clique_blocks(Z, ref["cliques"])):
        n = len(Ck)
        assert np.array_equal(flat[off:off + n * n].reshape(n, n).T, blk)
        off += n * n


def test_packed_api_argument_errors():
    import ctypes as C
    import nnsdp_b200 as nb
    import nnsdp_b200._lib as L

    xd = (L.c_i64 * 4)(2, 50, 50, 2)
    n = L.c_i64(0)
    assert L.lib.nnsdp_packed_layout(3, xd, 1, 0, None, None, None, None) == L.ERR_ARG          # ncells is required
    assert L.lib.nnsdp_packed_layout(3, xd, -1, 0, None, C.byref(n), None, None) == L.ERR_ASSERT  # 0 <= beta
    assert L.lib.nnsdp_packed_layout(3, xd, 1, 0, None, C.byref(n), None, None) == L.OK and n.value > 0
    cells = (L.PackedCell * 1)()
    assert L.lib.nnsdp_packed_layout(3, xd, 1, 1, cells, C.byref(n), None, None) == L.ERR_ARG   # table too small
    rec = np.zeros(8)
    pres = np.zeros(int(n.value), dtype=np.uint8)
    out = np.zeros(8)
    assert L.lib.nnsdp_packed_unpack(3, xd, 1, rec.ctypes.data_as(L.c_dp), pres.ctypes.data_as(C.POINTER(C.c_uint8)),
                                     L.FORMAT_PACKED, out.ctypes.data_as(L.c_dp)) == L.ERR_ARG  # target must be dense
    with pytest.raises(nb.NnsdpError):
        nb.plan_stats([2, 50, 50, 2], 1, dense=3)


def _mask_from_cells(cells, present, beta, Zdim):
    """Upper-triangle support the present cells define (numpy, from the cell table alone)."""
    M = np.zeros((Zdim, Zdim), dtype=bool)
    for i, c in enumerate(cells):
        if not present[i]:
            continue
        r0, c0, nr, nc = int(c["row0"]) - 1, int(c["col0"]) - 1, int(c["nrows"]), int(c["ncols"])
        if c["kind"] == 3:
            for t in range(beta + 1):
                idx = np.arange(nc - t)
                M[r0 + idx, r0 + idx + t] = True
        else:
            gr = (r0 + np.arange(nr))[:, None]
            gc = (c0 + np.arange(nc))[None, :]
            M[r0:r0 + nr, c0:c0 + nc] |= gr <= gc
    return M


@pytest.mark.parametrize("drop_diag", [False, True])
def test_unpack_on_several_host_threads(monkeypatch, drop_diag):
    """Outputs of 2^20 doubles and more are expanded by several host threads, one clique matrix at a time
    (plan.cpp unpack_record).  A record packed from a random symmetric matrix restricted to the cells' support must
    come back as that matrix's clique blocks (scatter of /root/reference/src/Methods/chordal_sdp.jl:60-93), and the
    result must not depend on the number of threads.  With the DIAG cells absent the band and the slivers the
    other cells hold must still be there."""
    import nnsdp_b200 as nb

    xdims, beta = [3, 260, 300, 280, 256, 270, 2], 2
    Zdim = sum(xdims[:-1]) + 1
    lay = nb.packed_layout(xdims, beta)
    cells = lay["cells"]
    present = np.ones(len(cells), dtype=np.uint8)
    if drop_diag:
        for i, c in enumerate(cells):
            if not c["always"]:
                present[i] = 0
        assert present.sum() < len(cells)
    rng = np.random.default_rng(11)
    A = rng.standard_normal((Zdim, Zdim))
    M = _mask_from_cells(cells, present, beta, Zdim)
    Zu = np.triu(A) * M
    Z = Zu + np.triu(Zu, 1).T
    rec = _pack_numpy(Z, lay, beta, present)
    cliques = nb.cliques_from_xdims(xdims, beta)
    assert sum(len(c[0]) ** 2 for c in cliques) >= 1 << 20 and len(cliques) > 1
    outs = []
    for nthreads in ("1", "3", "16"):
        monkeypatch.setenv("NNSDP_HOST_THREADS", nthreads)
        outs.append(nb.packed_unpack(xdims, beta, rec, present))
    monkeypatch.delenv("NNSDP_HOST_THREADS")
    outs.append(nb.packed_unpack(xdims, beta, rec, present))
    buf = np.full(outs[0].size, np.nan)
    assert nb.packed_unpack(xdims, beta, rec, present, out=buf) is buf
    outs.append(buf)
    for flat in outs:
        off = 0
        for ck in cliques:
            C = np.asarray(ck[0]) - 1
            n = len(C)
            assert np.array_equal(flat[off:off + n * n].reshape(n, n).T, Z[np.ix_(C, C)])
            off += n * n
        assert off == flat.size
