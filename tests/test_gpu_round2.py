"""Round-2 additions on the device: NaN policy of the arg-min, device-side top-k selection, the one-call multi-start
refinement, the library's own min-loc collective (single-rank communicator), the thread-per-candidate kernel for small n,
and the reference-named ScipyGPSurrogate (lazy L_, Matern(nu=inf)).  -m gpu."""
import ctypes
import os

import numpy as np
import pytest
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern

from conftest import golden_state
from oracle import gp_oracle as O
from test_gpu_parity import native_for

pytestmark = pytest.mark.gpu


def test_argmin_nan_policies_on_the_nan_fixture():
    """edge_alpha0_nan: candidates on training points with alpha = 0 have a posterior variance of 0 up to rounding; where it
    comes out <= 0 the EI is NaN (sqrt / scipy's scale > 0 rule; WHICH of them do is rounding noise, in the reference too).
    'first' must return what np.argmin returns on such values (the first NaN), 'skip' what np.nanargmin does."""
    g, st = golden_state("edge_alpha0_nan")
    gp = native_for(st, "f64")
    xs = gp.candidates(g["Xs"])
    assert np.isnan(g["ei"]).sum() == 9
    out = gp.sweep(xs, acq="ei", eta=float(g["eta"]), want_var=True, want_acq=True, want_min=True, index_base=50)
    a, var = out["acq"].cpu().numpy(), out["var"].cpu().numpy()
    assert np.array_equal(np.isnan(a), ~(var > 0)) and np.isnan(a).any()
    on_training_points = np.abs(g["var"]) < 1e-12
    assert (on_training_points | ~np.isnan(a)).all()                   # NaNs only where the reference's variance is ~0 too
    assert int(out["min_idx"].item()) - 50 == int(np.argmin(a)) and np.isnan(float(out["min_val"].item()))
    gp.set_nan_policy("skip")
    out = gp.sweep(xs, acq="ei", eta=float(g["eta"]), want_acq=True, want_min=True, index_base=50)
    assert int(out["min_idx"].item()) - 50 == int(np.nanargmin(a)) and float(out["min_val"].item()) == np.nanmin(a)
    vals, idxs = gp.segment_argmin(xs[:128], 128, "ei", eta=float(g["eta"]))
    assert int(idxs[0].item()) == int(np.nanargmin(a[:128]))
    gp.set_nan_policy("first")
    gp.close()


def test_optimizer_never_proposes_a_nan_candidate():
    """a candidate exactly on a training point with alpha = 0 has variance <= 0 -> NaN acquisition; np.argmin
    would make it the winner.  The optimisers use the 'skip' policy; pruned and plain sweeps then agree."""
    from bopy_b200.acquisition import EI
    from bopy_b200.surrogate import B200GPSurrogate
    rng = np.random.default_rng(3)
    X = rng.random((40, 2))
    y = np.sin(4 * X[:, 0]) + X[:, 1] ** 2
    sur = B200GPSurrogate(GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF([0.3, 0.3]), alpha=0.0, optimizer=None),
                          device_fit=False)
    sur.fit(X, y)
    ei = EI(sur)
    ei.fit(X, y)
    cands = np.concatenate([rng.random((500, 2)), X[:5], rng.random((500, 2))])
    values = ei(cands)
    assert np.isnan(values).any()
    i_first, v_first = ei.argmin(cands)
    assert i_first == int(np.argmin(values)) and np.isnan(v_first)
    i_skip, v_skip = ei.argmin(cands, nan_policy="skip")
    assert i_skip == int(np.nanargmin(values)) and v_skip == np.nanmin(values)
    i_pruned, v_pruned = ei.argmin(cands, nan_policy="skip", prune=True)
    assert (i_pruned, v_pruned) == (i_skip, v_skip)


def test_topk_min_distance_on_the_device_equals_the_host_rule():
    import torch

    from bopy_b200 import _native
    from bopy_b200.optimizer import OneShotBatchOptimizerTopKStrategy
    rng = np.random.default_rng(5)
    x = rng.random((20000, 3))
    a = rng.standard_normal(20000)
    a[rng.integers(0, 20000, 50)] = np.nan
    a[100] = a[7]                                   # an exact tie: the lower index goes first
    for k, dist, scale in ((8, 0.0, None), (6, 0.25, None), (5, 0.3, [1.0, 2.0, 0.5]), (40, 0.45, None)):
        strat = OneShotBatchOptimizerTopKStrategy(min_distance=dist, scale=scale)
        hx, hf = strat.select(x, a, k)
        dx, df = strat.select_on_device(torch.as_tensor(x, device="cuda"), torch.as_tensor(a, device="cuda"), k)
        assert np.array_equal(hx, dx) and np.array_equal(hf, df), (k, dist)
    idx, val = _native.topk_min_distance(torch.as_tensor(x[:3], device="cuda"), torch.as_tensor(a[:3], device="cuda"), 5)
    assert (idx.cpu().numpy()[3:] == -1).all()      # fewer evaluations than picks


def test_multistart_refine_is_the_python_loop_in_one_call():
    from bopy_b200 import _native
    g, st = golden_state("c3_branin_n256")
    gp = native_for(st, "f64")
    lo, hi = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
    starts = gp.candidates(O.candidates_uniform(9, 0, 200, lo, hi))
    xc1, fc1 = gp.multistart_refine(starts, "lcb", lo, hi, 12, kappa=2.0)
    import torch
    xt = starts.clone()
    xc, gc = torch.empty_like(xt), torch.empty_like(xt)
    fc = torch.empty(200, dtype=torch.float64, device=xt.device)
    alpha = torch.ones_like(fc)
    for k in range(13):
        ft, gt, _, _ = gp.value_and_grad(xt, "lcb", kappa=2.0)
        _native.multistart_step(lo, hi, xc, fc, gc, xt, ft, gt, alpha, first=(k == 0))
    assert torch.equal(xc, xc1) and torch.equal(fc, fc1)
    assert (fc1 <= gp.sweep(starts, acq="lcb", kappa=2.0, want_acq=True)["acq"] + 1e-12).all()    # monotone
    gp.close()


def test_library_collective_with_a_single_rank_communicator():
    """bopy_comm_* / bopy_minloc_allreduce through the C ABI (NCCL bound at run time): with one rank the winner is the
    rank's own record; the NaN policies apply.  (2 / 4 / 8 ranks: bench.py under torchrun, tools/multi_gpu_check.py.)"""
    import torch

    from bopy_b200 import _native
    lib = _native.load()
    uid = ctypes.create_string_buffer(128)
    _native.check(lib.bopy_comm_unique_id(uid, 128), "bopy_comm_unique_id")
    comm = ctypes.c_void_p()
    _native.check(lib.bopy_comm_create(ctypes.byref(comm), uid.raw, 1, 0, torch.cuda.current_device()), "bopy_comm_create")
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for value, index, policy, want in ((-2.5, 77, 0, (-2.5, 77)), (float("nan"), 5, 0, (float("nan"), 5)),
                                       (float("nan"), 5, 1, (None, -1))):
        v = torch.tensor([value], dtype=torch.float64, device="cuda")
        i = torch.tensor([index], dtype=torch.int64, device="cuda")
        _native.check(lib.bopy_minloc_allreduce(comm, ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(i.data_ptr()), policy, stream),
                      "bopy_minloc_allreduce")
        assert int(i.item()) == want[1]
        if want[0] is not None:
            assert (np.isnan(want[0]) and np.isnan(v.item())) or v.item() == want[0]
    lib.bopy_comm_destroy(comm)
    assert lib.bopy_comm_unique_id(None, 128) == _native.ERR_BAD_ARG


@pytest.mark.parametrize("name", ["c1_forrester_rbf_n6", "ref_forrester_matern15_fixed", "ref_sin_matern_default",
                                  "edge_on_training_points", "edge_alpha0_nan"])
def test_small_n_kernel_agrees_with_the_blocked_kernels(name):
    """n <= 32: the thread-per-candidate kernel serves every call; the blocked kernels (BOPY_B200_SMALL_N=0) are the
    cross-check, both against the golden vectors elsewhere and against each other here."""
    g, st = golden_state(name)
    assert st.X_train.shape[0] <= 32
    small = native_for(st, "f64")
    os.environ["BOPY_B200_SMALL_N"] = "0"
    try:
        blocked = native_for(st, "f64")
    finally:
        del os.environ["BOPY_B200_SMALL_N"]
    blocked.set_latency_path(0)
    assert small.launch_info(1 << 20)["grid"] == 148 * 8 and blocked.launch_info(1 << 20)["grid"] == 148
    rng = np.random.default_rng(1)
    lo, hi = g["Xs"].min(0), g["Xs"].max(0)
    Xs = np.concatenate([g["Xs"], lo + rng.random((4000, g["Xs"].shape[1])) * (hi - lo)])
    pv = (st.kernel.amplitude + st.kernel.noise_level) * st.y_std ** 2
    for m in (1, 129, len(Xs)):
        a = small.sweep(small.candidates(Xs[:m]), acq="ei", eta=float(g["eta"]), want_mean=True, want_var=True, want_acq=True,
                        want_min=True, index_base=7)
        b = blocked.sweep(blocked.candidates(Xs[:m]), acq="ei", eta=float(g["eta"]), want_mean=True, want_var=True,
                          want_acq=True, want_min=True, index_base=7)
        ma, mb = a["mean"].cpu().numpy(), b["mean"].cpu().numpy()
        va, vb = a["var"].cpu().numpy(), b["var"].cpu().numpy()
        assert np.max(np.abs(ma - mb)) <= 1e-10 * (np.max(np.abs(mb)) + st.y_std)
        assert np.max(np.abs(va - vb)) <= 1e-10 * np.max(np.abs(vb)) + 1e-11 * pv
        acq_small = a["acq"].cpu().numpy()
        assert int(a["min_idx"].item()) - 7 == int(np.argmin(acq_small))       # np.argmin of its own values
    seg_v, seg_i = small.segment_argmin(small.candidates(Xs[:1024]), 256, "lcb", kappa=2.0)
    own = small.sweep(small.candidates(Xs[:1024]), acq="lcb", kappa=2.0, want_acq=True)["acq"].cpu().numpy()
    assert np.array_equal(seg_i.cpu().numpy(), np.array([s * 256 + int(np.argmin(own[s * 256:(s + 1) * 256])) for s in range(4)]))
    small.close()
    blocked.close()


def test_scipy_surrogate_keeps_the_sklearn_object_usable_and_maps_matern_inf():
    from bopy_b200.kernel_spec import UnsupportedKernelError
    from bopy_b200.surrogate import ScipyGPSurrogate
    from sklearn.gaussian_process.kernels import RationalQuadratic
    rng = np.random.default_rng(8)
    X = rng.random((60, 2))
    y = np.cos(5 * X[:, 0]) * X[:, 1]
    sur = ScipyGPSurrogate(GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF([0.3, 0.4]), alpha=1e-6, normalize_y=True,
                                                    optimizer=None))
    sur.fit(X, y)
    assert sur.fitted_on_device and "L_" not in vars(sur.gp)
    xs = rng.random((50, 2))
    mean, var = sur.predict_diag(xs)
    h_mean, h_std = sur.gp.predict(xs, return_std=True)          # the wrapped object still predicts: L_ fetched on first read
    assert "L_" in vars(sur.gp)
    np.testing.assert_allclose(h_mean, mean, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(h_std ** 2, np.maximum(var, 0.0), rtol=1e-6, atol=1e-10)
    host = GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF([0.3, 0.4]), alpha=1e-6, normalize_y=True, optimizer=None).fit(X, y)
    np.testing.assert_allclose(sur.gp.L_, host.L_, rtol=0, atol=1e-9 * np.abs(host.L_).max())
    # Matern(nu=inf) is scikit-learn's RBF formula
    inf = ScipyGPSurrogate(GaussianProcessRegressor(kernel=Matern([0.3, 0.4], nu=np.inf), alpha=1e-6, normalize_y=True, optimizer=None))
    inf.fit(X, y)
    ref_mean, ref_cov = GaussianProcessRegressor(kernel=Matern([0.3, 0.4], nu=np.inf), alpha=1e-6, normalize_y=True,
                                                 optimizer=None).fit(X, y).predict(xs, return_cov=True)
    m2, v2 = inf.predict_diag(xs)
    np.testing.assert_allclose(m2, ref_mean, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(v2, np.diag(ref_cov), rtol=1e-6, atol=1e-10)
    with pytest.raises(UnsupportedKernelError, match="no CPU fallback"):
        ScipyGPSurrogate(GaussianProcessRegressor(kernel=RationalQuadratic(), optimizer=None)).fit(X, y)
