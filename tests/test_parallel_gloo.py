"""N>1 host logic on CPU: world size 2, gloo backend, rendezvous on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from yolo2_light_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1) the one collective: weight arena broadcast from rank 0
        arena = torch.arange(0, 100003, dtype=torch.int64).to(torch.uint8) if rank == 0 else torch.zeros(100003, dtype=torch.uint8)
        parallel.broadcast_arena(arena, src=0)
        ref = torch.arange(0, 100003, dtype=torch.int64).to(torch.uint8)
        ok_arena = bool(torch.equal(arena, ref))
        # 2) shards: every image exactly once, results gathered in image order
        G = 7
        lo, hi = parallel.shard_range(G, world, rank)
        local = np.stack([np.full((3, 2), float(i), np.float32) for i in range(lo, hi)]) if hi > lo else np.zeros((0, 3, 2), np.float32)
        allr = parallel.gather_to_rank0(local, G)
        ok_gather = True
        if rank == 0:
            ok_gather = allr.shape == (G, 3, 2) and all(float(allr[i, 0, 0]) == i for i in range(G))
        # 3) max-over-ranks timing
        m = parallel.max_over_ranks(10.0 + rank)
        q.put((rank, ok_arena, ok_gather, m, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] for r in res), "arena broadcast mismatch"
    assert all(r[2] for r in res), "gather mismatch"
    assert all(abs(r[3] - 11.0) < 1e-9 for r in res), "max over ranks"
    assert res[0][4] == (0, 4) and res[1][4] == (4, 7)


@pytest.mark.parametrize("G,world", [(16, 1), (16, 8), (128, 8), (7, 3), (3, 8), (0, 4)])
def test_shard_range_partitions_exactly(G, world):
    from yolo2_light_b200 import parallel
    seen = []
    sizes = []
    for r in range(world):
        lo, hi = parallel.shard_range(G, world, r)
        assert 0 <= lo <= hi <= G
        seen += list(range(lo, hi)); sizes.append(hi - lo)
    assert seen == list(range(G))
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)
