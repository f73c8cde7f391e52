"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch of frames with pwpp_dist exactly like bench.py
does on GPUs (each rank runs the oracle on its shard here), and the union of the shards equals the single-process
result. Also checks the max-over-ranks timing reduction and the stream->rank ownership rule."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (os.path.join(REPO, "patchwork-plusplus_b200"), os.path.join(REPO, "oracle")):
        sys.path.insert(0, p)
    import oracle_py as O
    import pwpp_dist
    import synth
    d = pwpp_dist.Dist(backend="gloo")
    mine = pwpp_dist.strong_shard(total, rank, world)
    counts = []
    for f in mine:
        a = synth.make_frame(321, f).numpy()[::8]  # thinned frames keep the CPU test fast
        o = O.Oracle(arith=O.ARITH_CANON64); o.estimate(a)
        counts.append(len(o.getGroundIndices()))
    d.barrier()
    slowest = d.max_over_ranks(10.0 + rank)       # rank 1 is "slower": the max must be seen by everybody
    allc = d.gather_ints(counts)
    q.put((rank, list(mine), allc, slowest))
    d.barrier()
    d.close()


def test_two_rank_sharding_matches_single_process():
    total, world = 5, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=300) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # shards are disjoint, contiguous, cover everything, and follow the documented split
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4]
    sys.path.insert(0, os.path.join(REPO, "patchwork-plusplus_b200")); sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle_py as O
    import synth
    single = []
    for f in range(total):
        o = O.Oracle(arith=O.ARITH_CANON64); o.estimate(synth.make_frame(321, f).numpy()[::8]); single.append(len(o.getGroundIndices()))
    for rank, mine, allc, slowest in res:
        assert allc == single
        assert slowest == 11.0


def test_shard_rules():
    sys.path.insert(0, os.path.join(REPO, "patchwork-plusplus_b200"))
    import pwpp_dist as D
    for total in (0, 1, 7, 8, 8192):
        for world in (1, 2, 4, 8):
            got = [i for r in range(world) for i in D.strong_shard(total, r, world)]
            assert got == list(range(total))
            sizes = [len(D.strong_shard(total, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert [list(D.weak_shard(3, r, 4)) for r in range(4)] == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10, 11]]
    assert [D.stream_owner(s, 2, 5) for s in range(5)] == [0, 0, 0, 1, 1]
