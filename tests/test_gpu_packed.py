"""GPU tests of the packed output format (NNSDP_FORMAT_PACKED): the records the device emits, expanded with
nnsdp_packed_unpack, must equal the dense outputs of the same library BIT FOR BIT -- every clique block
Z[C_k, C_k] (the blocks setupZksum! scatters, /root/reference/src/Methods/chordal_sdp.jl:60-93) and the dense Z
(/root/reference/src/Methods/chordal_sdp.jl:114) -- and through them the oracle to 1e-12 (normwise per block).
"""
import numpy as np
import pytest

import nnsdp_oracle as o
from helpers import rand_net, rand_query, relerr, to_numeric_batch

pytestmark = pytest.mark.gpu

TOL = 1e-12

NETS = [
    ([2, 3, 3, 2], 1),
    ([2, 3, 2], 0),
    ([3, 3, 3, 3, 4, 3, 3], 2),
    ([2, 4, 7, 3, 5, 2], 5),
    ([5, 50, 50, 50, 50, 50, 50, 5], 2),
    ([2] + [20] * 10 + [2], 1),
    ([2, 70, 130, 64, 3], 2),
    ([3, 150, 260, 140, 2], 0),
    ([3, 150, 260, 140, 2], 4),
    ([3, 150, 260, 140, 2], 6),
    ([2, 129, 128, 127, 300, 4], 3),
    ([6, 520, 33, 520, 3], 2),
    ([2] + [100] * 6 + [2], 2),
]


def _check_record(nb, rec, present, lay):
    """Always-written cells hold no NaN in their upper triangle; absent cells were not touched."""
    for i, c in enumerate(lay["cells"]):
        off, nr, nc = int(c["offset"]), int(c["nrows"]), int(c["ncols"])
        cell = rec[off:off + nr * nc].reshape(nc, nr).T
        if not present[i]:
            assert np.all(np.isnan(cell))
            continue
        if c["kind"] == nb.CELL_BAND:
            assert not np.any(np.isnan(cell))
            continue
        gr = (int(c["row0"]) + np.arange(nr))[:, None]
        gc = (int(c["col0"]) + np.arange(nc))[None, :]
        assert not np.any(np.isnan(cell[gr <= gc]))


@pytest.mark.parametrize("kind", ["safety", "hplane", "ellipsoid"])
@pytest.mark.parametrize("xdims,beta", NETS)
def test_packed_records_equal_the_dense_outputs_bit_for_bit(ctx, xdims, beta, kind):
    import nnsdp_b200 as nb

    net = rand_net(xdims, seed=17, sigma=0.1 if max(xdims) >= 100 else None)
    rng = np.random.default_rng(23)
    qs = [rand_query(net, beta, rng, kind=kind, radius=r) for r in (0.0, 0.001, 0.02, 0.3)]
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    batch = to_numeric_batch(nb, net, qs)
    flat = nb.assemble_blocks(dnet, beta, batch)
    Z = nb.assemble_dense(dnet, beta, batch)
    lay = nb.packed_layout(xdims, beta)
    b = nb.Batch(dnet, beta, Qcap=len(qs), ring=3, packed=True)
    assert b.per_query == lay["record_doubles"] and b.ncells == len(lay["cells"])
    b.set_inputs(batch)
    rec = np.full((len(qs), b.per_query), np.nan)
    present = np.full((len(qs), b.ncells), 7, dtype=np.uint8)
    b.run_packed(rec, present)
    st = b.packed_stats()
    assert st["d2h_bytes"] <= rec.nbytes and st["present_optional_cells"] == int(present[:, lay["cells"]["always"] == 0].sum())
    b.close()
    assert set(np.unique(present)) <= {0, 1}
    assert np.all(present[:, lay["cells"]["always"] == 1] == 1)
    cliques = dnet.cliques(beta)
    for i, q in enumerate(qs):
        _check_record(nb, rec[i], present[i], lay)
        r = np.nan_to_num(rec[i], nan=12345.0)      # untouched / undefined bytes must not be used by the unpack
        assert np.array_equal(nb.packed_unpack(xdims, beta, r, present[i]), flat[i])
        assert np.array_equal(nb.packed_unpack(xdims, beta, r, present[i], dense_Z=True).reshape(Z[i].shape).T, Z[i])
        ref = o.run_query(net, beta, q)
        for blk, rb in zip(nb.split_blocks(nb.packed_unpack(xdims, beta, r, present[i]), cliques), ref["blocks"]):
            assert relerr(blk, rb) <= TOL
    # radius 0: every ReLU is stable, so at least one hidden layer has an active neuron -> some DIAG cell present
    if np.any(lay["cells"]["always"] == 0) and kind != "hplane":
        assert present[0, lay["cells"]["always"] == 0].any()
    # the one-shot entry point gives the same records where they are defined
    rec2, present2, _ = nb.assemble_packed(dnet, beta, batch)
    assert np.array_equal(present2, present)
    for i in range(len(qs)):
        assert np.array_equal(nb.packed_unpack(xdims, beta, rec2[i], present2[i]), flat[i])


def test_packed_rings_chunks_and_slots(ctx):
    import nnsdp_b200 as nb

    xdims, beta = [2, 300, 270, 280, 2], 2
    net = rand_net(xdims, seed=3, sigma=0.1)
    rng = np.random.default_rng(5)
    qs = [rand_query(net, beta, rng, kind="circle", radius=0.004 * i) for i in range(7)]
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    flat = nb.assemble_blocks(dnet, beta, to_numeric_batch(nb, net, qs))
    lay = nb.packed_layout(xdims, beta)
    for nq, ring in ((7, 3), (7, 1), (1, 4), (5, 8), (7, 2)):
        b = nb.Batch(dnet, beta, Qcap=nq, ring=ring, packed=True)
        b.set_inputs(to_numeric_batch(nb, net, qs[:nq]))
        rec = np.full((nq, b.per_query), np.nan)
        present = np.zeros((nq, b.ncells), dtype=np.uint8)
        b.run_packed(rec, present)
        for i in range(nq):
            _check_record(nb, rec[i], present[i], lay)
            assert np.array_equal(nb.packed_unpack(xdims, beta, np.nan_to_num(rec[i]), present[i]), flat[i])
        # whole-record copy and the device-resident path (records stay in the ring)
        rec_d = np.zeros((nq, b.per_query))
        b.run_packed(rec_d, None, flags=nb.RUN_DENSE_COPY)
        for i in range(nq):
            assert np.array_equal(nb.packed_unpack(xdims, beta, rec_d[i], present[i]), flat[i])
        b.run_packed(None, None)
        b.sync()
        last_chunk0 = ((nq - 1) // max(1, min(ring, nq))) * min(ring, nq) if ring <= nq else 0
        slot = b.get_slot(0)
        assert np.array_equal(nb.packed_unpack(xdims, beta, slot, present[last_chunk0]), flat[last_chunk0])
        b.close()


def test_packed_reach_batch_shares_the_gram_blocks(ctx):
    import nnsdp_b200 as nb

    xdims, beta, nq = [2, 150, 260, 140, 2], 2, 6
    net = rand_net(xdims, seed=9, sigma=0.3)
    rng = np.random.default_rng(2)
    base = rand_query(net, beta, rng, kind="hplane", radius=0.0)
    th = 2 * np.pi * np.arange(nq) / nq
    normals = np.stack([np.cos(th), np.sin(th)], axis=1)
    gouts = rng.random((nq, 1))
    batch = nb.NumericBatch(x1min=base.x1min, x1max=base.x1max, gamma_in=base.gin, gamma_bnd=base.gbnd,
                            gamma_sec=base.gsec, out_kind=nb.OUT_HPLANE, out_vec=normals, gamma_out=gouts)
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    flat = nb.assemble_blocks(dnet, beta, batch, Q=nq)
    b = nb.Batch(dnet, beta, Qcap=nq, ring=4, packed=True)
    b.set_inputs(batch, Q=nq)
    rec = np.zeros((nq, b.per_query))
    present = np.zeros((nq, b.ncells), dtype=np.uint8)
    b.run_packed(rec, present)
    assert b.stage_ms("gram")[1] == 1
    for i in range(nq):
        assert np.array_equal(nb.packed_unpack(xdims, beta, rec[i], present[i]), flat[i])
    b.close()


def test_packed_stress_size_bit_identity(ctx):
    """BASELINE.json configs[4] (W1000-D20, beta = 2): packed records of two queries (one with Gram-active layers)
    against the dense clique blocks of the same library, all 19 blocks, bit for bit."""
    import bench
    import nnsdp_b200 as nb

    xdims, Ms, beta, inp = bench.make_workload("stress-W1000-D20-beta2-Q1024", 0, Q=2)
    inp["x1min"][1] = inp["x1min"][1] * 0 + 1.0 - 1e-4
    inp["x1max"][1] = inp["x1min"][1] + 2e-4
    dnet = nb.Net(ctx, xdims, Ms)
    batch = nb.NumericBatch(out_kind=nb.OUT_SAFETY, **inp)
    lay = nb.packed_layout(xdims, beta)
    assert lay["record_doubles"] * 8 < 0.25 * 1330657432 and lay["always_doubles"] * 8 < 0.125 * 1330657432
    rec, present, _ = nb.assemble_packed(dnet, beta, batch, Q=2)
    opt = lay["cells"]["always"] == 0
    assert present[1, opt].any() and not present[0, opt].all()
    b = nb.Batch(dnet, beta, Qcap=2, ring=1)
    b.set_inputs(batch, Q=2)
    b.bounds()
    b.prepare()
    for q in range(2):
        b.emit(q, 1)
        b.sync()
        dense = b.get_slot(0)
        mine = nb.packed_unpack(xdims, beta, rec[q], present[q])
        assert np.array_equal(mine, dense)
        del dense, mine
    b.close()


def test_packed_multi_device_context(ctx):
    import nnsdp_b200 as nb

    if nb.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    xdims, beta, nq = [2, 300, 270, 2], 2, 7
    net = rand_net(xdims, seed=3, sigma=0.1)
    rng = np.random.default_rng(5)
    qs = [rand_query(net, beta, rng, kind="ellipsoid", radius=0.02 * i) for i in range(nq)]
    batch = to_numeric_batch(nb, net, qs)
    one, p1, _ = nb.assemble_packed(nb.Net(ctx, net.xdims, net.Ms), beta, batch)
    ctx2 = nb.Context([0, 1])
    two, p2, _ = nb.assemble_packed(nb.Net(ctx2, net.xdims, net.Ms), beta, batch)
    assert np.array_equal(p1, p2)
    for i in range(nq):
        assert np.array_equal(nb.packed_unpack(xdims, beta, one[i], p1[i]), nb.packed_unpack(xdims, beta, two[i], p2[i]))


def test_packed_call_sequence_errors(ctx):
    import nnsdp_b200 as nb

    net = rand_net([2, 60, 60, 2], seed=1)
    dnet = nb.Net(ctx, net.xdims, net.Ms)
    rng = np.random.default_rng(0)
    batch = to_numeric_batch(nb, net, [rand_query(net, 1, rng)])
    b = nb.Batch(dnet, 1, Qcap=1, ring=1)                  # dense blocks: the packed calls must refuse
    b.set_inputs(batch)
    with pytest.raises(nb.NnsdpError) as e:
        b.run_packed(np.zeros(4), None)
    assert e.value.code == -4
    with pytest.raises(nb.NnsdpError) as e:
        b.packed_stats()
    assert e.value.code == -4
    b.close()
    pb = nb.Batch(dnet, 1, Qcap=1, ring=1, packed=True)
    with pytest.raises(nb.NnsdpError) as e:                 # no inputs yet
        pb.run_packed(None, None)
    assert e.value.code == -4
    pb.close()
