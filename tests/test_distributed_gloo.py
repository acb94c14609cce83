"""The N>1 path on CPU: world_size-2 gloo group, sharded index ranges, one min-loc exchange.
The device pieces (candidate generator, fused argmin) are replaced by the oracle's numpy restatements."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bopy_b200.distributed import minloc_better, reduce_minloc, shard_range


def test_shard_range_partitions_exactly():
    for m in (1, 7, 128, 100_003, 1 << 24):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(m, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == m
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [e - s for s, e in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_minloc_ordering_is_np_argmin():
    rng = np.random.default_rng(0)
    for _ in range(200):
        v = rng.choice([0.0, -0.0, 1.0, -1.0, np.nan, 2.5], size=6)
        recs = [(float(x), i) for i, x in enumerate(v)]
        rng.shuffle(recs)
        assert reduce_minloc(recs)[1] == int(np.argmin(v))
    assert not minloc_better((1.0, -1), (5.0, 3)) and minloc_better((5.0, 3), (1.0, -1))


class FakeAcquisition:
    """argmin over candidate rows with the numpy oracle of a toy acquisition (a(x) = |x - 0.3|_1)."""

    def argmin(self, xs, index_base=0, **kwargs):
        v = np.abs(xs.numpy() - 0.3).sum(1)
        i = int(np.argmin(v))
        return index_base + i, float(v[i])


def _worker(rank, world, port, m, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bopy_b200._native as native
    from bopy_b200.distributed import all_reduce_minloc, sharded_argmin
    from oracle import gp_oracle as O
    native.candidates_uniform = lambda seed, base, mm, lo, hi, device=None: torch.from_numpy(
        O.candidates_uniform(seed, base, mm, lo, hi))
    x, val = sharded_argmin(FakeAcquisition(), 42, [0.0, 0.0], [1.0, 1.0], m)
    nan_case = all_reduce_minloc(float("nan") if rank == 1 else -5.0, 10 + rank)
    tie_case = all_reduce_minloc(1.25, 100 - rank)
    if rank == 0:
        out.put((x.tolist(), val, nan_case, tie_case))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("m", [1, 1001])
def test_sharded_argmin_over_gloo_matches_single_process(m):
    from oracle import gp_oracle as O
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 29500 + (os.getpid() + m) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x, val, nan_case, tie_case = out.get()
    xs = O.candidates_uniform(42, 0, m, [0.0, 0.0], [1.0, 1.0])
    v = np.abs(xs - 0.3).sum(1)
    assert x == xs[int(np.argmin(v))].tolist() and val == float(v.min())
    assert np.isnan(nan_case[0]) and nan_case[1] == 11          # a NaN beats every number
    assert tie_case == (1.25, 99)                                # equal values: lowest global index wins


def test_minloc_reduction_properties():
    """Order-independence and agreement with np.argmin for random records incl. NaNs, ties, -0.0 and empty slots."""
    from hypothesis import given, settings
    from hypothesis import strategies as st_

    values = st_.sampled_from([0.0, -0.0, 1.5, -2.25, float("nan"), float("inf"), -float("inf")])

    @settings(max_examples=200, deadline=None)
    @given(st_.lists(values, min_size=1, max_size=12), st_.randoms(use_true_random=False))
    def check(vals, rnd):
        recs = [(v, i) for i, v in enumerate(vals)] + [(0.0, -1)]      # plus an empty record
        expected = int(np.argmin(np.array(vals)))
        rnd.shuffle(recs)
        val, idx = reduce_minloc(recs)
        assert idx == expected
        assert (np.isnan(val) and np.isnan(vals[expected])) or val == vals[expected]
        # associativity: reduce in two halves, then reduce the partial results
        half = len(recs) // 2
        assert reduce_minloc([reduce_minloc(recs[:half]), reduce_minloc(recs[half:])])[1] == expected

    check()


class _ToyNative:
    """CPU stand-in for NativeGP in the multi-start optimiser: a(x) = |x - c|^2 with its gradient."""

    c = np.array([0.3, 0.6, 0.9])

    def value_and_grad(self, xt, kind, **kw):
        r = xt - torch.from_numpy(self.c)
        return (r * r).sum(1), 2.0 * r, None, None


class _ToySurrogate:
    native = _ToyNative()

    def supports_gradient(self):
        return True

    def acquisition_segment_argmin(self, kind, xs, seg_len, index_base=0, **kw):
        v = ((xs - torch.from_numpy(_ToyNative.c)) ** 2).sum(1).reshape(-1, seg_len)
        idx = v.argmin(1)
        return v.gather(1, idx[:, None])[:, 0], index_base + torch.arange(len(idx)) * seg_len + idx


class _ToyAcquisition:
    kind = "lcb"
    surrogate = _ToySurrogate()

    def native_args(self):
        return {}


def _multistart_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bopy_b200._native as native
    from bopy_b200.bounds import Bound, Bounds
    from bopy_b200.optimizer import MultiStartOptimizer
    from oracle import gp_oracle as O
    native.candidates_uniform = lambda seed, base, mm, lo, hi, device=None: torch.from_numpy(
        O.candidates_uniform(seed, base, mm, lo, hi))
    native.gather_rows = lambda xs, idx, index_base=0: xs[idx - index_base]
    native.multistart_step = lambda lo, hi, xc, fc, gc, xt, ft, gt, alpha, first: O.multistart_step(
        lo, hi, xc.numpy(), fc.numpy(), gc.numpy(), xt.numpy(), ft.numpy(), gt.numpy(), alpha.numpy(), first)
    opt = MultiStartOptimizer(_ToyAcquisition(), Bounds([Bound(0.0, 1.0)] * 3), n_starts=5, n_candidates=5 * 128,
                              seed=4, distributed=True, method="gradient", iterations=30)
    res = opt.optimize()
    xs, vs = opt.local_minima()
    out.put((rank, res.x_min.tolist(), res.f_min.tolist(), len(vs)))
    dist.barrier()
    dist.destroy_process_group()


def test_multistart_shards_starts_over_gloo_ranks():
    """5 starts over 2 ranks (3 + 2): each rank refines its own starts, one min-loc picks the winner and its owner
    broadcasts the point -- every rank returns the same (x_min, f_min)."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_multistart_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = sorted(out.get() for _ in range(2))
    (r0, x0, f0, n0), (r1, x1, f1, n1) = got
    assert (n0, n1) == (3, 2)
    assert x0 == x1 and f0 == f1
    assert np.allclose(x0, [_ToyNative.c], atol=1e-8) and f0[0] < 1e-14
