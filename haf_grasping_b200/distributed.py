"""Multi-GPU host logic: one process per GPU, units sharded across ranks, and the ONE exchange
of the path -- the per-unit best-grasp records -- over torch.distributed (NCCL on GPUs, gloo in the CPU tests).

Two shardings:
  * by cloud (throughput mode, BASELINE configs[4]): `shard_range` gives each rank a contiguous block of clouds;
    no data-path collective, results gathered with `all_gather_records`.
  * by (approach vector, roll) inside one goal (configs[2]): `unit_blocks` gives each rank a contiguous block of the
    A*R units; `merge_unit_tops` replays the reference's sequential rules on the gathered per-unit tops -- strict '>'
    so the earliest unit wins ties (server.cpp:953), early exit at graspval_top when return_only_best
    (server.cpp:362-365) -- so the merged result equals the single-GPU (and the reference's) result exactly.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous block [begin, end) of rank `rank` when n_items are split as evenly as possible"""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def unit_blocks(n_requests: int, n_rolls: int, rank: int, world: int) -> List[Tuple[int, int, int]]:
    """(request index, roll_begin, roll_end) pieces of this rank's contiguous block of the n_requests*n_rolls units"""
    u0, u1 = shard_range(n_requests * n_rolls, rank, world)
    out = []
    for a in range(n_requests):
        b, e = max(u0, a * n_rolls), min(u1, (a + 1) * n_rolls)
        if b < e:
            out.append((a, b - a * n_rolls, e - a * n_rolls))
    return out


def merge_unit_tops(unit_tops: np.ndarray, n_requests: int, n_rolls: int, return_only_best: Sequence[int],
                    graspval_top: Sequence[int]):
    """unit_tops: [n_requests*n_rolls, 3] (row, col, topval), rows of units nobody evaluated hold topval -1000.
    Returns (per-request list of (row, col, roll, topval, rolls_done), overall (request, row, col, roll, topval))."""
    per = []
    for a in range(n_requests):
        best, brow, bcol, broll, done = -1000, -1, -1, -1, 0
        for roll in range(n_rolls):
            if return_only_best[a] and best >= graspval_top[a]:      # server.cpp:362-365
                break
            row, col, val = (int(v) for v in unit_tops[a * n_rolls + roll])
            if val > best:                                             # server.cpp:953, strict >
                best, brow, bcol, broll = val, row, col, roll
            done += 1
        per.append((brow, bcol, broll, best, done))
    win, wtop = -1, -1000
    for a, p in enumerate(per):
        if p[3] > wtop:
            wtop, win = p[3], a
    overall = (win,) + per[win][:4] if win >= 0 else (-1, -1, -1, -1, -1000)
    return per, overall


def all_gather_records(local: np.ndarray, group=None) -> np.ndarray:
    """all_gather of equally-shaped int32 record arrays (rank-major concatenation); identity without a process group"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local.copy()
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(local, np.int32)).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.concatenate([o.cpu().numpy() for o in out], axis=0)


def sharded_goal_search(evaluate: Callable[[int, int, int], np.ndarray], n_requests: int, n_rolls: int,
                        return_only_best: Sequence[int], graspval_top: Sequence[int], rank: int, world: int, group=None):
    """One goal (n_requests approach vectors x n_rolls rolls) over `world` ranks.
    evaluate(request_index, roll_begin, roll_end) -> int array [roll_end-roll_begin, 3] of per-roll tops of THIS rank's
    units (GraspSearch.search with roll_begin / roll_limit on a GPU; the oracle in the CPU tests)."""
    import torch
    import torch.distributed as dist
    tops = np.full((n_requests * n_rolls, 3), -1000, np.int32)
    tops[:, :2] = -1
    for a, rb, re in unit_blocks(n_requests, n_rolls, rank, world):
        tops[a * n_rolls + rb:a * n_rolls + re] = np.asarray(evaluate(a, rb, re), np.int32).reshape(re - rb, 3)
    if world > 1:
        # element-wise MAX over ranks: every unit is evaluated by exactly one rank, the others hold (-1, -1, -1000)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.from_numpy(tops).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        tops = t.cpu().numpy()
    return merge_unit_tops(tops, n_requests, n_rolls, return_only_best, graspval_top) + (tops,)
