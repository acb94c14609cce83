"""Candidate-sharded sweep: one process per GPU, ONE tiny min-loc exchange per sweep.

Candidates are independent units (every column of L^-1 K*^T is), so the index range [0, m) is cut
into `world_size` contiguous slices; each rank generates its slice on its own GPU from
(seed, global index), runs the fused argmin with `index_base` = slice start, and the per-rank
(value, global index) records -- 16 bytes each -- are all-gathered (NCCL over NVLink on the B200 box,
gloo in the CPU tests) and reduced with np.argmin's ordering: NaN first, then value, then lowest
index.  The fitted state (X, L, alpha) is replicated: every rank fits/uploads its own copy.
"""
from typing import Optional, Sequence, Tuple

import numpy as np


def shard_range(m: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slice [start, stop) of range(m) owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(int(m), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def minloc_better(a: Tuple[float, int], b: Tuple[float, int]) -> bool:
    """True if record a = (value, index) beats b under np.argmin's rules (index < 0: empty)."""
    (av, ai), (bv, bi) = a, b
    if ai < 0:
        return False
    if bi < 0:
        return True
    an, bn = av != av, bv != bv
    if an or bn:
        return ai < bi if (an and bn) else an
    if av != bv:
        return av < bv
    return ai < bi


def reduce_minloc(records: Sequence[Tuple[float, int]]) -> Tuple[float, int]:
    best = (0.0, -1)
    for rec in records:
        if minloc_better(rec, best):
            best = rec
    return best


def all_reduce_minloc(value: float, index: int, group=None, device=None) -> Tuple[float, int]:
    """Min-loc over all ranks of `group` (default group if None).  Works with nccl and gloo."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value), int(index)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cpu") if backend == "gloo" else (device or torch.device("cuda", torch.cuda.current_device()))
    # one 16-byte record per rank: the value's bit pattern and the index, both as int64 (NaN-safe)
    rec = torch.tensor([np.float64(value).view(np.int64).item(), int(index)], dtype=torch.int64, device=dev)
    gathered = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(gathered, rec, group=group)
    rows = torch.stack(gathered).cpu().numpy()
    records = [(float(np.int64(r[0]).view(np.float64)), int(r[1])) for r in rows]
    return reduce_minloc(records)


def sharded_argmin(acquisition, seed: int, lowers, uppers, m: int, group=None, incumbent=None, prune: bool = False):
    """Each rank sweeps its slice of the m counter-based candidates; returns (x (d,), value) on every rank."""
    import torch
    import torch.distributed as dist

    from . import _native

    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    start, stop = shard_range(m, rank, world)
    value, index = 0.0, -1
    if stop > start:
        xs = _native.candidates_uniform(seed, start, stop - start, lowers, uppers)
        if incumbent is not None and start == 0:
            xs[0] = torch.as_tensor(incumbent, dtype=torch.float64, device=xs.device)
        index, value = (acquisition.argmin(xs, index_base=start, prune=True) if prune
                        else acquisition.argmin(xs, index_base=start))
    value, index = all_reduce_minloc(value, index, group=group)
    if incumbent is not None and index == 0:
        x = np.asarray(incumbent, dtype=np.float64)
    else:
        x = _native.candidates_uniform(seed, index, 1, lowers, uppers)[0].cpu().numpy()
    return x, value
