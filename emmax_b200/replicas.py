"""
Multi-GPU: independent replicas + ONE collective.

Every `generate_actions` request is independent (one image + prompt -> one action), and the whole bf16 model
(~15 GB) fits a single B200 many times over, so the path shards as **replicas only**: one process per GPU, one full
model per process, requests dealt round-robin (rank r serves requests r, r+N, ...; `shard_requests_balanced` for an offline batch
with uneven token limits). The only exchange is an
all-gather of each replica's action tokens (7 ids padded to 8 x int32) per control tick, enqueued on the decode stream
so no host sync is added. The reference has no multi-GPU inference at all (it asserts bs == 1:
/root/reference/prismatic/extern/hf/modeling_prismatic.py:326, :460-463); this is the "bs=1 x 8 replicas" config of
BASELINE.json.
"""

from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist

ACTION_SLOTS = 8  # 7-DoF action tokens padded to 8 x int32 (32 B per replica per tick)


def shard_requests(n_requests: int, rank: int, world: int) -> List[int]:
    """Indices of the requests replica `rank` serves (round-robin / by robot id)."""
    return list(range(rank, n_requests, world))


def shard_requests_balanced(limits: Sequence[int], rank: int, world: int) -> List[int]:
    """Deal for an OFFLINE batch whose token limits are known and uneven (BASELINE.json configs[4]: mixed 128 / 512 new tokens): the
    round-robin deal of `shard_requests` can hand one replica all the long requests (limits alternating with an even world size do exactly
    that). Here the requests go, longest first, to the replica with the least decode work so far (ties: fewest requests, then lowest rank) —
    longest-processing-time-first list scheduling — so every replica's continuous-batching stream (`Engine.serve`) ends at about the same
    launch. A pure function of (limits, world): every rank computes the same deal without communication. Returns this rank's request
    indices in request order; `balanced_owner` gives the whole assignment (for `merge_by_owner`)."""
    owner = balanced_owner(limits, world)
    return [i for i, r in enumerate(owner) if r == rank]


def balanced_owner(limits: Sequence[int], world: int) -> List[int]:
    """owner[i] = replica that serves request i under `shard_requests_balanced`."""
    load, count = [0] * world, [0] * world
    owner = [0] * len(limits)
    for i in sorted(range(len(limits)), key=lambda j: -int(limits[j])):  # stable: ties keep request order
        r = min(range(world), key=lambda k: (load[k], count[k], k))
        owner[i] = r
        load[r] += int(limits[i])
        count[r] += 1
    return owner


def merge_by_owner(gathered: torch.Tensor, owner: Sequence[int], world: int) -> torch.Tensor:
    """Undo an arbitrary deal: `gathered` is the rank-major all-gather of per-replica blocks padded to the same number of rows, replica r's
    rows in the order of its request indices. Returns one row per request, in request order."""
    per = gathered.shape[0] // world
    seen = [0] * world
    idx = []
    for r in owner:
        idx.append(r * per + seen[r])
        seen[r] += 1
    assert max(seen) <= per, "a replica's block is shorter than the number of requests dealt to it"
    return gathered[torch.tensor(idx, dtype=torch.long, device=gathered.device)]


def pack_action_tokens(token_ids: torch.Tensor, action_dim: int = 7) -> torch.Tensor:
    """[b, action_dim] (or [action_dim]) int -> [b, ACTION_SLOTS] int32, zero padded."""
    t = token_ids.reshape(-1, action_dim).to(torch.int32)
    out = torch.zeros((t.shape[0], ACTION_SLOTS), dtype=torch.int32, device=t.device)
    out[:, :action_dim] = t
    return out


def gather_action_tokens(local: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """All-gather the packed action tokens of every replica: [b, 8] -> [world * b, 8] (rank-major).
    Asynchronous w.r.t. the host on NCCL (enqueued on the current stream); world size 1 is a copy."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.clone() if out is None else out.copy_(local)
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1))
    return out


def merge_in_request_order(gathered: torch.Tensor, world: int, n_requests: int) -> torch.Tensor:
    """Undo the round-robin deal: row (r, i) of the rank-major gather is request r + i * world."""
    per = gathered.shape[0] // world
    idx = [r * per + i for i in range(per) for r in range(world) if r + i * world < n_requests]
    return gathered[torch.tensor(idx, device=gathered.device)]
