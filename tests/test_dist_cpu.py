"""world_size-2 gloo test of the N>1 host path: queries are sharded over ranks with no
data-path collective; ranks exchange only the max step time and small counters.  Runs on CPU: each
rank drives the host-only entry points of the C ABI for its shard."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, Q, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (os.path.join(ROOT, "nn-sdp_b200"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import bench
    import nnsdp_b200 as nb
    from nnsdp_b200.dist import gather_objects, max_over_ranks, shard_range

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q0, nq = shard_range(Q, world, rank)
        # every rank builds the same net and its own queries (bench.make_workload seeds by rank)
        xdims, Ms, beta, inp = bench.make_workload("tiny-W10-D10-beta1-Q64", rank, Q=nq)
        xd0, Ms0, _, inp0 = bench.make_workload("tiny-W10-D10-beta1-Q64", 0, Q=nq)
        same_net = all(np.array_equal(a, b) for a, b in zip(Ms, Ms0))
        distinct = rank == 0 or not np.array_equal(inp["x1min"], inp0["x1min"])
        sz = nb.sizes_from_xdims(xdims, beta)
        cl = nb.cliques_from_xdims(xdims, beta)
        rec = {"rank": rank, "q0": q0, "nq": nq, "ncliques": len(cl), "blocks": nq * len(cl),
               "sum_ck_sq": sz["sum_ck_sq"], "same_net": same_net, "distinct": distinct}
        recs = gather_objects(rec)
        t = max_over_ranks(10.0 + rank)
        if rank == 0:
            ret["recs"] = recs
            ret["t"] = t
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_query_sharding_gloo():
    world, Q = 2, 65
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), Q, ret), nprocs=world, join=True)
        recs, t = list(ret["recs"]), ret["t"]
    assert t == 11.0  # max over ranks
    assert [r["rank"] for r in recs] == [0, 1]
    assert recs[0]["q0"] == 0 and recs[0]["nq"] == 33 and recs[1]["q0"] == 33 and recs[1]["nq"] == 32
    assert sum(r["blocks"] for r in recs) == Q * 9
    assert all(r["same_net"] and r["distinct"] for r in recs)
    assert recs[0]["sum_ck_sq"] == recs[1]["sum_ck_sq"] == 8705
