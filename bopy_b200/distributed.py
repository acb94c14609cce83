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


_COMMS = {}      # id(process group) -> NativeComm (one NCCL communicator of the library per group)


class NativeComm:
    """The library's own communicator for the min-loc exchange (`bopy_comm_*`, include/bopy_b200.h): an ncclComm_t
    created from a unique id that rank 0 draws and torch.distributed ships to the other ranks once.  The exchange itself
    (`bopy_minloc_allreduce`) is an ncclAllGather of one 16-byte record per rank plus a one-warp kernel on the sweep's
    stream: the winner lands in the caller's device buffers without any host synchronisation."""

    ID_BYTES = 128

    def __init__(self, group=None, device=None):
        import ctypes

        import torch
        import torch.distributed as dist

        from . import _native
        self.lib = _native.load()
        self.device = _native.resolve_device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        uid = ctypes.create_string_buffer(self.ID_BYTES)
        if self.rank == 0:
            _native.check(self.lib.bopy_comm_unique_id(uid, self.ID_BYTES), "bopy_comm_unique_id")
        t = torch.frombuffer(bytearray(uid.raw), dtype=torch.uint8).clone()
        if dist.get_backend(group) != "gloo":
            t = t.to(self.device)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(t, src=src, group=group)
        raw = bytes(t.cpu().numpy().tobytes())
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _native.check(self.lib.bopy_comm_create(ctypes.byref(handle), raw, self.world, self.rank, self.device.index),
                          "bopy_comm_create")
        self._handle = handle

    def minloc_(self, val_dev, idx_dev, nan_policy="first"):
        """In place: (val_dev (1,) fp64, idx_dev (1,) int64) device tensors become the winner over all ranks."""
        import torch

        from . import _native
        with torch.cuda.device(self.device):
            _native.check(self.lib.bopy_minloc_allreduce(self._handle, _native._ptr(val_dev), _native._ptr(idx_dev),
                                                         _native.NAN_POLICY_IDS[nan_policy], _native._stream(self.device)),
                          "bopy_minloc_allreduce")
        return val_dev, idx_dev

    def close(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            try:
                self.lib.bopy_comm_destroy(handle)
            except Exception:
                pass

    def __del__(self):
        self.close()


def native_comm(group=None, device=None):
    """The NativeComm of `group` (created on first use; collective), or None when the group is not an NCCL group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend(group) != "nccl":
        return None
    key = id(group) if group is not None else 0
    if key not in _COMMS:
        _COMMS[key] = NativeComm(group, device)
    return _COMMS[key]


def all_reduce_minloc_device(val_dev, idx_dev, group=None, nan_policy="first"):
    """Min-loc over all ranks on DEVICE buffers (val (1,) fp64, idx (1,) int64), in place, no host synchronisation:
    `bopy_minloc_allreduce` on NCCL groups.  On a gloo group (CPU tests) it falls back to the host exchange."""
    comm = native_comm(group, val_dev.device)
    if comm is not None:
        return comm.minloc_(val_dev, idx_dev, nan_policy)
    v, i = all_reduce_minloc(float(val_dev.item()), int(idx_dev.item()), group=group, nan_policy=nan_policy)
    val_dev.fill_(v)
    idx_dev.fill_(i)
    return val_dev, idx_dev


def all_reduce_minloc(value: float, index: int, group=None, device=None, nan_policy: str = "first") -> Tuple[float, int]:
    """Min-loc over all ranks of `group` (default group if None), host values in and out.  NCCL groups go through the
    library's collective (`bopy_minloc_allreduce`, one D2H of the winner); gloo groups (CPU tests) through
    torch.distributed.all_gather and the same ordering restated in `reduce_minloc`."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value), int(index)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    if backend == "nccl":
        from . import _native
        dev = _native.resolve_device(device)
        rec_v = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        rec_i = torch.tensor([int(index)], dtype=torch.int64, device=dev)
        native_comm(group, dev).minloc_(rec_v, rec_i, nan_policy)
        both = torch.stack([rec_v.view(torch.int64)[0], rec_i[0]]).cpu().numpy()     # ONE device-to-host copy
        return float(both[:1].view(np.float64)[0]), int(both[1])
    dev = torch.device("cpu") if backend == "gloo" else (device or torch.device("cuda", torch.cuda.current_device()))
    # one 16-byte record per rank: the value's bit pattern and the index, both as int64 (NaN-safe)
    rec = torch.tensor([np.float64(value).view(np.int64).item(), int(index)], dtype=torch.int64, device=dev)
    gathered = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(gathered, rec, group=group)
    rows = torch.stack(gathered).cpu().numpy()
    records = [(float(np.int64(r[0]).view(np.float64)), int(r[1])) for r in rows]
    if nan_policy == "skip":
        records = [(v, (-1 if v != v else i)) for v, i in records]
    return reduce_minloc(records)


def sharded_argmin(acquisition, seed: int, lowers, uppers, m: int, group=None, incumbent=None, prune: bool = False,
                   nan_policy: str = "skip"):
    """Each rank sweeps its slice of the m counter-based candidates; returns (x (d,), value) on every rank.

    On NCCL groups nothing touches the host between the sweep and the winner: the rank's (value, index) stay on the
    device, `bopy_minloc_allreduce` (ncclAllGather of 16-byte records + a one-warp kernel on the sweep's stream) replaces
    them by the global winner, and ONE device-to-host copy fetches it.  `nan_policy` as in `AcquisitionFunction.argmin`
    (default 'skip': an optimiser never proposes a NaN point; falls back to 'first' if every candidate is NaN)."""
    import torch
    import torch.distributed as dist

    from . import _native

    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        host_exchange = dist.get_backend(group) == "gloo"
    else:
        rank, world, host_exchange = 0, 1, False
    start, stop = shard_range(m, rank, world)
    if host_exchange:
        # CPU groups (the gloo tests; the device pieces are stand-ins there): host records, torch all_gather
        value, index = 0.0, -1
        if stop > start:
            xs = _native.candidates_uniform(seed, start, stop - start, lowers, uppers)
            if incumbent is not None and start == 0:
                xs[0] = torch.as_tensor(incumbent, dtype=torch.float64, device=xs.device)
            index, value = acquisition.argmin(xs, index_base=start, prune=prune, nan_policy=nan_policy)
        value, index = all_reduce_minloc(value, index, group=group, nan_policy=nan_policy)
    else:
        dev = _native.resolve_device(getattr(getattr(acquisition, "surrogate", None), "device", None))
        if stop > start:
            xs = _native.candidates_uniform(seed, start, stop - start, lowers, uppers, device=dev)
            if incumbent is not None and start == 0:
                xs[0] = torch.as_tensor(incumbent, dtype=torch.float64, device=xs.device)
            got = acquisition.argmin(xs, index_base=start, prune=prune, nan_policy=nan_policy, on_device=True)
            if isinstance(got[0], torch.Tensor):
                idx_d, val_d = got
            else:                   # an acquisition without a device arg-min (foreign surrogate): host values
                idx_d = torch.tensor([int(got[0])], dtype=torch.int64, device=dev)
                val_d = torch.tensor([float(got[1])], dtype=torch.float64, device=dev)
        else:
            idx_d = torch.full((1,), -1, dtype=torch.int64, device=dev)
            val_d = torch.zeros(1, dtype=torch.float64, device=dev)
        if world > 1:
            all_reduce_minloc_device(val_d, idx_d, group=group, nan_policy=nan_policy)
        both = torch.stack([idx_d[0], val_d.view(torch.int64)[0]]).cpu().numpy()          # the one device-to-host copy
        index, value = int(both[0]), float(both[1:].view(np.float64)[0])
    if index < 0 and nan_policy == "skip":     # every candidate of every rank was NaN: np.argmin's answer
        return sharded_argmin(acquisition, seed, lowers, uppers, m, group=group, incumbent=incumbent, prune=False,
                              nan_policy="first")
    if incumbent is not None and index == 0:
        x = np.asarray(incumbent, dtype=np.float64)
    else:
        x = _native.candidates_uniform(seed, index, 1, lowers, uppers)[0].cpu().numpy()
    return x, value
