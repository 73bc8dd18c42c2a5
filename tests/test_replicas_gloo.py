"""N>1 host logic on CPU: world_size-2 gloo run of the request deal + the single all-gather of action tokens."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from emmax_b200.replicas import (ACTION_SLOTS, balanced_owner, gather_action_tokens, merge_by_owner, merge_in_request_order, pack_action_tokens,
                                  shard_requests, shard_requests_balanced)  # fmt: skip


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_requests: int, q) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_requests(n_requests, rank, world)
        per = (n_requests + world - 1) // world
        # each request's "action tokens" are a deterministic function of the request id
        toks = torch.zeros((per, 7), dtype=torch.int64)
        for i, r in enumerate(mine):
            toks[i] = 31744 + (torch.arange(7) * 7 + r) % 256
        gathered = gather_action_tokens(pack_action_tokens(toks))
        merged = merge_in_request_order(gathered, world, n_requests)
        q.put((rank, mine, merged.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_replicas_gather_in_request_order():
    world, n_requests = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_requests, q)) for r in range(world)]
    [p.start() for p in procs]
    results = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    want = [[int(31744 + (j * 7 + r) % 256) for j in range(7)] + [0] * (ACTION_SLOTS - 7) for r in range(n_requests)]
    deals = {}
    for rank, mine, merged in results:
        deals[rank] = mine
        assert merged == want  # every replica ends up with every request's action tokens, in request order
    assert deals == {0: [0, 2, 4], 1: [1, 3]}


def test_single_process_is_a_copy():
    t = pack_action_tokens(torch.arange(14).reshape(2, 7))
    assert t.shape == (2, ACTION_SLOTS) and t.dtype == torch.int32
    assert torch.equal(gather_action_tokens(t), t)
    assert shard_requests(10, 3, 4) == [3, 7]


def test_balanced_deal_of_uneven_token_limits():
    """BASELINE.json configs[4]-shaped batch (limits alternating 512 / 128) over 8 replicas: round-robin hands the even ranks every long
    request; the balanced deal gives every replica the same decode work, covers every request exactly once and is the same on every rank."""
    limits = [512, 128] * 64
    world = 8
    rr = [sum(limits[i] for i in shard_requests(len(limits), r, world)) for r in range(world)]
    assert max(rr) == 4 * min(rr)  # the problem being solved
    deals = [shard_requests_balanced(limits, r, world) for r in range(world)]
    assert sorted(i for d in deals for i in d) == list(range(len(limits)))
    loads = [sum(limits[i] for i in d) for d in deals]
    assert max(loads) == min(loads) == sum(limits) // world
    assert all(len(d) == len(limits) // world for d in deals)
    owner = balanced_owner(limits, world)
    assert all(owner[i] == r for r, d in enumerate(deals) for i in d)
    # uneven counts: 5 requests over 2 replicas, one of them long
    owner = balanced_owner([16, 512, 16, 16, 16], 2)
    assert owner == [1, 0, 1, 1, 1]
    assert shard_requests_balanced([], 0, 4) == []


def _worker_balanced(rank: int, world: int, port: int, limits, q) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        owner = balanced_owner(limits, world)
        mine = shard_requests_balanced(limits, rank, world)
        per = max(owner.count(r) for r in range(world))  # blocks padded to the largest share
        toks = torch.zeros((per, 7), dtype=torch.int64)
        for i, r in enumerate(mine):
            toks[i] = 31744 + (torch.arange(7) * 7 + r) % 256
        gathered = gather_action_tokens(pack_action_tokens(toks))
        q.put((rank, mine, merge_by_owner(gathered, owner, world).tolist()))
    finally:
        dist.destroy_process_group()


def test_two_replicas_balanced_deal_gathers_in_request_order():
    world, limits = 2, [16, 512, 16, 16, 16]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_balanced, args=(r, world, port, limits, q)) for r in range(world)]
    [p.start() for p in procs]
    results = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    want = [[int(31744 + (j * 7 + r) % 256) for j in range(7)] + [0] * (ACTION_SLOTS - 7) for r in range(len(limits))]
    for rank, mine, merged in results:
        assert mine == ([1] if rank == 0 else [0, 2, 3, 4])
        assert merged == want
