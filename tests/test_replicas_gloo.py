"""N>1 host logic on CPU: world_size-2 gloo run of the request deal + the single all-gather of action tokens."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from emmax_b200.replicas import ACTION_SLOTS, gather_action_tokens, merge_in_request_order, pack_action_tokens, shard_requests


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
