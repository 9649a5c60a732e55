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


def _worker_ragged(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shards = sweep.shard_balanced(lengths, world)
        mine = shards[rank]
        local = torch.stack([torch.full((sweep.REC_WORDS,), int(i) * 7, dtype=torch.int32) for i in mine]) if len(mine) else \
            torch.empty((0, sweep.REC_WORDS), dtype=torch.int32)
        q.put((rank, sweep.allgather_indexed(local, shards, rank, world).numpy(), [s.tolist() for s in shards]))
    finally:
        dist.destroy_process_group()


def test_balanced_sharding_and_indexed_allgather_world2_gloo():
    """Ragged (VidChapters-shaped) batches: length-balanced bin packing + the one all-gather that restores global order."""
    rng = np.random.default_rng(5)
    lengths = rng.integers(20, 140, size=37).tolist()
    for world in (1, 2, 3, 8):
        shards = sweep.shard_balanced(lengths, world)
        flat = sorted(i for s in shards for i in s.tolist())
        assert flat == list(range(37))                                          # every window exactly once
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)                          # LPT bound
        assert all(s.tolist() == sorted(s.tolist()) for s in shards)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ragged, args=(r, world, port, lengths, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(got[0][1], got[1][1])
    assert got[0][2] == got[1][2]                                               # same assignment computed on every rank
    assert got[0][1][:, 0].tolist() == [i * 7 for i in range(37)]


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


class _FakeSweepModel:
    """generate() stand-in for the sweep drivers on the CPU: the new tokens are a function of each window and prompt."""
    engine = None
    device = torch.device("cpu")

    def generate(self, ids, images=None, attention_mask=None, max_new_tokens=4, **kw):
        am = attention_mask.bool()
        code = torch.stack([im.float().sum() for im in images]) + (ids.clamp(min=0) * am).sum(dim=1).float()
        new = torch.stack([(code * (t + 1)).round().long() % 89 for t in range(max_new_tokens)], dim=1)
        return {"sequences": torch.cat([ids, new], dim=1), "entropies": torch.full((ids.shape[0], max_new_tokens), 0.25)}


def _ragged_inputs():
    g = torch.Generator().manual_seed(9)
    n, L = 19, 10
    frames = [int(x) for x in torch.randint(1, 50, (n,), generator=g)]
    windows = [torch.randint(-3, 4, (f, 8), generator=g).float() for f in frames]
    ids = torch.randint(3, 300, (n, L), generator=g)
    ids[:, 2] = -200
    am = torch.ones(n, L, dtype=torch.bool)
    for i in range(n):
        if i % 3:
            am[i, L - (i % 3):] = False
    return windows, ids, am


def _worker_ragged_sweep(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        windows, ids, am = _ragged_inputs()
        rec = sweep.ragged_sweep(_FakeSweepModel(), windows, ids, am, None, max_new_tokens=4, rank=rank, world=world,
                                 max_tokens_per_batch=120, eos_token_id=None)
        q.put((rank, rec.numpy()))
    finally:
        dist.destroy_process_group()


def test_ragged_sweep_driver_world2_gloo_equals_one_rank():
    """The whole ragged-sweep driver on two ranks (each scores its length-balanced share in token-budget batches, one
    indexed all-gather) returns on every rank the table one rank computes alone."""
    windows, ids, am = _ragged_inputs()
    alone = sweep.ragged_sweep(_FakeSweepModel(), windows, ids, am, None, max_new_tokens=4, rank=0, world=1,
                               max_tokens_per_batch=120, eos_token_id=None).numpy()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ragged_sweep, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(got[0], got[1])
    np.testing.assert_array_equal(got[0], alone)


class _FakeDenseModel:
    engine = None
    device = torch.device("cpu")

    def generate(self, ids, images=None, max_new_tokens=4, **kw):
        code = images.float().sum(dim=(1, 2)) + ids.clamp(min=0).sum(dim=1).float()
        new = torch.stack([(code * (t + 1)).round().long() % 83 for t in range(max_new_tokens)], dim=1)
        ent = (code[:, None] % 7 + 1.0) * torch.arange(1, max_new_tokens + 1)[None] * 0.125
        return {"sequences": torch.cat([ids, new], dim=1), "entropies": ent}


def _worker_stage1(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(4)
        segs = torch.randint(-3, 4, (11, 6, 8), generator=g).to(torch.bfloat16)
        ids = torch.randint(3, 300, (9,), generator=g)
        res = sweep.stage1_sweep(_FakeDenseModel(), segs, ids, None, max_new_tokens=4, rank=rank, world=world, batch=3, eos_token_id=None)
        q.put((rank, res.records.numpy(), res.local_indices.tolist()))
    finally:
        dist.destroy_process_group()


def test_stage1_sweep_driver_world2_gloo_equals_one_rank():
    """sweep.stage1_sweep on two ranks: segment i goes to rank i mod 2, each rank scores its share in batches of 3, one
    all-gather of the fixed-size records - every rank ends with the table a single rank computes."""
    g = torch.Generator().manual_seed(4)
    segs = torch.randint(-3, 4, (11, 6, 8), generator=g).to(torch.bfloat16)
    ids = torch.randint(3, 300, (9,), generator=g)
    alone = sweep.stage1_sweep(_FakeDenseModel(), segs, ids, None, max_new_tokens=4, rank=0, world=1, batch=3, eos_token_id=None)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_stage1, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {r: (rec, idx) for r, rec, idx in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(got[0][0], got[1][0])
    np.testing.assert_array_equal(got[0][0], alone.records.numpy())
    assert got[0][1] == [0, 2, 4, 6, 8, 10] and got[1][1] == [1, 3, 5, 7, 9]
    un = sweep.unpack_records(alone.records)
    assert un["tokens"].shape[0] == 11 and torch.isfinite(un["h_mean"]).all()


class _FakeStage2Model:
    """generate() stand-in for the stage-2 driver: tokens and entropies are a function of each prompt's stacked windows."""
    engine = None
    device = torch.device("cpu")

    def generate(self, ids, images=None, max_new_tokens=4, **kw):
        code = images.float().sum(dim=(1, 2, 3)) + ids.clamp(min=0).sum(dim=1).float()
        new = torch.stack([(code * (t + 1)).round().long() % 83 for t in range(max_new_tokens)], dim=1)
        ent = (code[:, None] % 7 + 1.0) * torch.arange(1, max_new_tokens + 1)[None] * 0.125
        return {"sequences": torch.cat([ids, new], dim=1), "entropies": ent}


def _stage2_inputs():
    g = torch.Generator().manual_seed(12)
    wins = torch.randint(-3, 4, (11, 5, 8), generator=g).to(torch.bfloat16)
    ids = torch.randint(3, 300, (9,), generator=g)
    kw = dict(grounding_windows=list(range(20, 31)), batch=4, zooms=(4, 2, 1), max_new_tokens=5, perm_seed=3,
              answer_number=lambda t: int(t[0]) % 4, eos_token_id=None)
    return wins, ids, kw


def _worker_stage2(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wins, ids, kw = _stage2_inputs()
        calls = sweep.stage2_pass(_FakeStage2Model(), wins, None, ids, rank=rank, world=world, shard_calls=True, **kw)
        q.put((rank, calls))
    finally:
        dist.destroy_process_group()


def test_stage2_calls_dealt_to_two_ranks_equal_one_rank():
    """One query's stage-2 generate() calls (chunks x zoom levels) dealt round-robin to two ranks + the all-gather of their
    records: every rank ends with the list of calls - tokens, entropy statistics, chosen window - that one rank computes alone."""
    wins, ids, kw = _stage2_inputs()
    alone = sweep.stage2_pass(_FakeStage2Model(), wins, None, ids, **kw)
    assert len(alone) == 11 + 6 + 3
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_stage2, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1] == alone


class _FakeIndexedModel:
    """generate() stand-in for the multi-query sweep: tokens are a function of the segment a row shows and of its prompt."""
    engine = None
    device = torch.device("cpu")

    def generate(self, ids, images=None, image_index=None, max_new_tokens=4, **kw):
        code = images.float().sum(dim=(1, 2))[image_index] + ids.clamp(min=0).sum(dim=1).float()
        new = torch.stack([(code * (t + 1)).round().long() % 83 for t in range(max_new_tokens)], dim=1)
        ent = (code[:, None] % 7 + 1.0) * torch.arange(1, max_new_tokens + 1)[None] * 0.125
        return {"sequences": torch.cat([ids, new], dim=1), "entropies": ent}


def _multi_query_inputs():
    g = torch.Generator().manual_seed(14)
    segs = torch.randint(-3, 4, (11, 6, 8), generator=g).to(torch.bfloat16)
    ids = torch.randint(3, 300, (3, 9), generator=g)                    # three queries
    return segs, ids


def _worker_multi_query(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        segs, ids = _multi_query_inputs()
        rec = sweep.stage1_sweep_queries(_FakeIndexedModel(), segs, ids, None, max_new_tokens=4, rank=rank, world=world, batch_segments=2,
                                         eos_token_id=None)
        q.put((rank, rec.numpy()))
    finally:
        dist.destroy_process_group()


def test_multi_query_sweep_world2_gloo_equals_one_rank():
    """sweep.stage1_sweep_queries on two ranks: segment i -> rank i mod 2 with all its queries, the Q records of a segment travel
    together through one all-gather; every rank ends with the [W * Q] table (row = segment * Q + query) one rank computes."""
    segs, ids = _multi_query_inputs()
    alone = sweep.stage1_sweep_queries(_FakeIndexedModel(), segs, ids, None, max_new_tokens=4, batch_segments=4, eos_token_id=None).numpy()
    assert alone.shape == (11 * 3, sweep.REC_WORDS)
    # row = segment * Q + query: the same segment with another query differs, the same query on another segment differs
    tok = sweep.unpack_records(torch.from_numpy(alone))["tokens"][:, :4]
    assert not torch.equal(tok[0], tok[1]) and not torch.equal(tok[0], tok[3])
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_multi_query, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(got[0], got[1])
    np.testing.assert_array_equal(got[0], alone)
