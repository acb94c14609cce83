"""The reference CPU path on THIS host: bopy's call sequence on scikit-learn / scipy at chunk sizes
1 (how DIRECT calls it), 64, 256, 1024, and -- labelled not-the-reference -- the diag-only numpy restatement.  Prints the
host's core count, CPU model, BLAS thread pools and library versions beside the numbers."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402
from oracle import reference_path as R  # noqa: E402


def rate(fn, m, budget_s):
    fn()
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        fn()
        best = max(best, m / (time.perf_counter() - t0))
        if time.perf_counter() - t0 > budget_s:
            break
    return best


def main():
    import scipy
    import sklearn
    from threadpoolctl import threadpool_info
    cores, model = bench.host_info()
    print(json.dumps(dict(cores=cores, cpu=model, numpy=np.__version__, scipy=scipy.__version__, sklearn=sklearn.__version__,
                          pools=[{k: p.get(k) for k in ("user_api", "internal_api", "num_threads")} for p in threadpool_info()])))
    shapes = [(256, 2, 32768), (2048, 6, 32768)] + ([(8192, 20, 4096)] if "--large" in sys.argv else [])
    with bench.all_host_threads(), np.errstate(invalid="ignore", divide="ignore"):
        for n, d, m_cpu in shapes:
            X, y, gp = bench.make_problem(n, d)
            gp.fit(X, y)
            st = O.state_from_sklearn(gp)
            eta = float(y.min())
            xs = O.candidates_uniform(bench.SEED_CAND, 0, m_cpu, np.zeros(d), np.ones(d))
            row = dict(n=n, d=d)
            for c in (1, 64, 256, 1024):
                mm = min(m_cpu, 2048 if c == 1 else m_cpu)
                sub = xs[:mm]
                row[f"reference_c{c}"] = rate(lambda: [R.ei(gp, sub[s:s + c], eta) for s in range(0, mm, c)], mm, 20.0)
            row["diag_only_numpy_not_the_reference"] = rate(lambda: O.acquisition_sweep(st, "ei", xs, eta=eta), m_cpu, 20.0)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
