"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) here.

The reference's tests hold no numeric vectors for the hot path, so the
oracle is pinned against the reference itself: for every case below the reference's
``ScipyGPSurrogate`` (bopy/surrogate.py:72-91) is fitted and its ``LCB`` / ``EI`` / ``POI``
(bopy/acquisition.py:67-131) are evaluated on a candidate set; inputs, fitted hyper-parameters
and outputs are frozen.  Runs only in the authoring container (needs /root/reference).

    python tools/make_golden.py            # rewrites tests/golden/
"""
import hashlib
import os
import sys

import numpy as np
import scipy
import sklearn
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from reference_import import import_reference  # noqa: E402

from bopy_b200.benchmark_functions import bohachevsky, branin, forrester, hartmann6  # noqa: E402
from oracle.gp_oracle import candidates_uniform, kernel_from_sklearn  # noqa: E402

bopy = import_reference()
OUT = os.path.join(ROOT, "tests", "golden")


def run_case(name, X, y, gp, Xs, kappas=(2.0, 0.5), store_X=True, X_seed=None, cov_corner=64):
    sur = bopy.surrogate.ScipyGPSurrogate(gp=gp)
    sur.fit(X, y)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean, cov = sur.predict(Xs)
        var = np.diag(cov).copy()
        out = {}
        for kappa in kappas:
            a = bopy.acquisition.LCB(sur, kappa=kappa)
            a.fit(X, y)
            out[f"lcb_{kappa}"] = a(Xs)
        ei = bopy.acquisition.EI(sur)
        ei.fit(X, y)
        poi = bopy.acquisition.POI(sur)
        poi.fit(X, y)
        ei_v, poi_v = ei(Xs), poi(Xs)
    spec = kernel_from_sklearn(sur.gp.kernel_)
    c = min(cov_corner, Xs.shape[0])
    rec = dict(
        y=y, Xs=Xs, mean=mean, var=var, cov_corner=cov[:c, :c].copy(),
        ei=ei_v, poi=poi_v, eta=np.float64(ei._eta),
        argmin_ei=np.int64(np.argmin(ei_v)), argmin_poi=np.int64(np.argmin(poi_v)),
        kernel_kind=np.str_(spec.kind), kernel_nu=np.float64(spec.nu),
        length_scale=spec.length_scale, amplitude=np.float64(spec.amplitude),
        noise_level=np.float64(spec.noise_level),
        alpha_reg=np.float64(gp.alpha), normalize_y=np.bool_(gp.normalize_y),
        y_mean=np.float64(np.ravel(sur.gp._y_train_mean)[0]), y_std=np.float64(np.ravel(sur.gp._y_train_std)[0]),
        alpha_head=sur.gp.alpha_[:16].copy(), L_diag_head=np.diag(sur.gp.L_)[:16].copy(),
        kappas=np.array(kappas),
        versions=np.str_(f"numpy {np.__version__} scipy {scipy.__version__} sklearn {sklearn.__version__}"),
        X_sha256=np.str_(hashlib.sha256(np.ascontiguousarray(X).tobytes()).hexdigest()),
    )
    for kappa in kappas:
        rec[f"lcb_{kappa}"] = out[f"lcb_{kappa}"]
        rec[f"argmin_lcb_{kappa}"] = np.int64(np.argmin(out[f"lcb_{kappa}"]))
    if store_X:
        rec["X"] = X
    else:
        rec["X_seed"] = np.int64(X_seed)
        rec["X_shape"] = np.array(X.shape)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    nan = int(np.isnan(ei_v).sum())
    print(f"{name:34s} n={X.shape[0]:5d} d={X.shape[1]:2d} m={Xs.shape[0]:5d} ls={spec.length_scale[:3]} "
          f"amp={spec.amplitude:.4g} argmin_ei={int(rec['argmin_ei'])} nan_ei={nan} zeros_ei={(ei_v == 0).sum()} "
          f"var[min,max]=[{np.nanmin(var):.3e},{np.nanmax(var):.3e}]")


def main():
    os.makedirs(OUT, exist_ok=True)
    # (i) the reference's own fixtures ---------------------------------------------------
    # tests/test_surrogate.py:11-27
    X = np.linspace(0, 1, 10).reshape(-1, 1)
    y = forrester(X)
    grid = np.linspace(0, 1, 1001).reshape(-1, 1)
    run_case("ref_forrester_matern15_opt", X, y,
             GaussianProcessRegressor(kernel=Matern(nu=1.5), alpha=1e-5, normalize_y=True), grid)
    run_case("ref_forrester_matern15_fixed", X, y,
             GaussianProcessRegressor(kernel=Matern(nu=1.5), alpha=1e-5, normalize_y=True, optimizer=None), grid)
    # tests/test_acquisiton.py:13-25
    X = np.linspace(-np.pi, np.pi, 10).reshape(-1, 1)
    y = np.sin(X).flatten()
    run_case("ref_sin_matern_default", X, y, GaussianProcessRegressor(kernel=Matern()),
             np.linspace(-np.pi, np.pi, 1001).reshape(-1, 1))
    # examples/example_1d.py:68-79 shape (Forrester, RBF GP, LCB) on the sklearn surrogate
    rng = np.random.default_rng(7)
    X = rng.random((6, 1))
    run_case("c1_forrester_rbf_n6", X, forrester(X),
             GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(0.2), alpha=1e-10, normalize_y=True,
                                      optimizer=None), grid)
    # (ii) config-shaped sets ---------------------------------------
    rng = np.random.default_rng(1234)
    lo, hi = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
    X = lo + rng.random((256, 2)) * (hi - lo)
    run_case("c3_branin_n256", X, branin(X),
             GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF([3.0, 3.0]), alpha=1e-6, normalize_y=True,
                                      optimizer=None),
             candidates_uniform(1235, 0, 2048, lo, hi))
    rng = np.random.default_rng(1234)
    X = rng.random((2048, 6))
    run_case("c4_hartmann6_n2048", X, hartmann6(X),
             GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(0.3 * np.ones(6)), alpha=1e-6,
                                      normalize_y=True, optimizer=None),
             candidates_uniform(1235, 0, 1024, np.zeros(6), np.ones(6)))
    rng = np.random.default_rng(1234)
    X = rng.random((8192, 20))
    y = np.sin(3.0 * X[:, :5].sum(1)) + 0.5 * np.cos(2.0 * X[:, 5:].sum(1) / 3.0)
    run_case("c5_rbf_d20_n8192", X, y,
             GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(1.0 * np.ones(20)), alpha=1e-6,
                                      normalize_y=True, optimizer=None),
             candidates_uniform(1235, 0, 256, np.zeros(20), np.ones(20)), store_X=False, X_seed=1234)
    # (iii) edge cases ---------------------------------------------------------------------
    # candidates ON training points with alpha=1e-10 (var ~ 0, can go negative -> sqrt NaN ->
    # scipy NaN -> np.argmin returns the first NaN) and far-field candidates (EI == 0.0 plateau)
    rng = np.random.default_rng(11)
    X = np.sort(rng.random((12, 1)), axis=0)
    y = forrester(X)
    Xs = np.concatenate([X, np.linspace(0, 1, 200).reshape(-1, 1), np.linspace(5, 50, 40).reshape(-1, 1), X[::-1]])
    run_case("edge_on_training_points", X, y,
             GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(0.15), alpha=1e-10, normalize_y=False,
                                      optimizer=None), Xs)
    # alpha=0: posterior variance at the training points is 0 up to rounding -> exact zeros and
    # small negatives -> NaN in EI/POI (scipy scale>0 rule) -> np.argmin returns the first NaN
    X = np.sort(np.random.default_rng(11).random((10, 1)), axis=0)
    Xs = np.concatenate([np.linspace(0, 1, 100).reshape(-1, 1), X, np.linspace(0, 1, 100).reshape(-1, 1)])
    run_case("edge_alpha0_nan", X, forrester(X),
             GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(0.03), alpha=0.0, normalize_y=False,
                                      optimizer=None), Xs)
    rng = np.random.default_rng(12)
    X = rng.random((64, 3)) * np.array([1.0, 2.0, 4.0])
    y = np.sin(X[:, 0] * 5) + X[:, 1] ** 2 - 0.3 * X[:, 2]
    Xs = rng.random((700, 3)) * np.array([1.0, 2.0, 4.0])
    run_case("ard_amp_white", X, y,
             GaussianProcessRegressor(kernel=ConstantKernel(2.5) * RBF([0.2, 0.7, 1.3]) + WhiteKernel(1e-3),
                                      alpha=1e-8, normalize_y=False, optimizer=None), Xs)
    run_case("iso_rbf_bare", X, y,
             GaussianProcessRegressor(kernel=RBF(0.5), alpha=1e-6, normalize_y=True, optimizer=None), Xs)
    rng = np.random.default_rng(13)
    X = rng.random((50, 2))
    y = bohachevsky(X * 2 - 1)
    Xs = rng.random((600, 2))
    for nu in (0.5, 1.5, 2.5):
        run_case(f"matern{int(nu * 10):02d}_d2", X, y,
                 GaussianProcessRegressor(kernel=ConstantKernel(1.7) * Matern([0.3, 0.5], nu=nu), alpha=1e-7,
                                          normalize_y=True, optimizer=None), Xs)
    # a non-multiple-of-tile n and d=1..: ragged sizes
    rng = np.random.default_rng(14)
    X = rng.random((333, 4))
    y = hartmann6(np.concatenate([X, 0.5 * np.ones((333, 2))], axis=1))
    run_case("ragged_n333_d4_opt", X, y,
             GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF(0.4 * np.ones(4)), alpha=1e-6,
                                      normalize_y=True, n_restarts_optimizer=0, random_state=0),
             rng.random((777, 4)))


if __name__ == "__main__":
    main()
