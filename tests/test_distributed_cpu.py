"""world_size-2 gloo test of the N>1 path's host logic: round-robin sharding + the single all-gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from revisionllm_b200 import sweep


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sweep.shard_indices(n_total, rank, world)
        # record of segment i is a deterministic function of i, so the gathered table can be checked exactly
        tok = torch.stack([torch.arange(16, dtype=torch.int32) + int(i) * 100 for i in mine])
        spans = torch.tensor([[int(i), int(i) + 1] for i in mine], dtype=torch.int32)
        f = torch.tensor([float(i) for i in mine])
        local = sweep.pack_records(tok, spans, f * 0.5, f * 2.0, f - 3.0)
        allrec = sweep.allgather_records(local, n_total, rank, world)
        q.put((rank, allrec.numpy()))
    finally:
        dist.destroy_process_group()


def test_allgather_records_world2_gloo():
    world, n_total = 2, 7          # uneven shards: rank 0 gets 4, rank 1 gets 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(got[0], got[1])          # every rank ends with the same table
    un = sweep.unpack_records(torch.from_numpy(got[0]))
    assert un["tokens"][:, 0].tolist() == [i * 100 for i in range(n_total)]
    assert un["spans"][:, 0].tolist() == list(range(n_total))
    assert un["cos"].tolist() == [float(i) - 3.0 for i in range(n_total)]
