"""torchrun --nproc-per-node N tools/multi_gpu_check.py: the sharded optimisers over NCCL against their single-GPU results.
Every rank fits its own replica; candidates / starts are sharded; one min-loc all-gather per sweep."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bopy_b200.acquisition import EI, LCB  # noqa: E402
from bopy_b200.bounds import Bound, Bounds  # noqa: E402
from bopy_b200.optimizer import CandidateSweepOptimizer, MultiStartOptimizer  # noqa: E402
from bopy_b200.surrogate import B200GPSurrogate  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/bopy_b200_nccl.%h.%p.log")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    X, y, gp = bench.make_problem(1024, 6)
    sur = B200GPSurrogate(gp, device=local)
    sur.fit(X, y)
    bounds = Bounds([Bound(0.0, 1.0)] * 6)
    ok = True
    for name, acq in (("ei", EI(sur)), ("lcb", LCB(sur))):
        acq.fit(X, y)
        m = 300_007                                   # not divisible by the world size
        single = CandidateSweepOptimizer(acq, bounds, n_candidates=m, seed=3).optimize()
        for prune in (False, True):
            sharded = CandidateSweepOptimizer(acq, bounds, n_candidates=m, seed=3, distributed=True, prune=prune).optimize()
            same = np.array_equal(single.x_min, sharded.x_min) and np.array_equal(single.f_min, sharded.f_min)
            ok &= same
            if rank == 0:
                print(f"{name}: sweep sharded over {world} GPUs (prune={prune}) == single GPU: {same}", flush=True)
        ms1 = MultiStartOptimizer(acq, bounds, n_starts=64, n_candidates=1 << 16, seed=5, method="gradient").optimize()
        msd = MultiStartOptimizer(acq, bounds, n_starts=64, n_candidates=1 << 16, seed=5, method="gradient",
                                  distributed=True).optimize()
        same = np.allclose(ms1.x_min, msd.x_min, rtol=0, atol=1e-12) and np.allclose(ms1.f_min, msd.f_min, rtol=1e-12, atol=0)
        ok &= bool(same)
        if rank == 0:
            print(f"{name}: multi-start sharded == single GPU: {same} ({ms1.f_min[0]:.9f} vs {msd.f_min[0]:.9f})", flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("ALL OK" if int(t.item()) == 1 else "MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
