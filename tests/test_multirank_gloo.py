"""world_size-2 CPU (gloo) test of the sharding + single all-gather used by the N>1 bench path."""
import os
import socket

import torch
import torch.multiprocessing as mp

from synergynet_b200 import distributed as sd


def test_shard_range_partitions_exactly():
    for total in (1, 7, 1024, 8191):
        for world in (1, 2, 3, 8):
            spans = [sd.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, ret):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sd.init_process_group('gloo')
    lo, hi = sd.shard_range(12, rank, world)
    full = torch.arange(12 * 3 * 68, dtype=torch.float32).view(12, 3, 68)
    out = sd.gather_landmarks(full[lo:hi].clone())
    ret[rank] = bool(torch.equal(out, full))
    torch.distributed.destroy_process_group()


def test_two_rank_gather_reassembles_batch():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]
