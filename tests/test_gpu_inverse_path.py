"""The inverse path (probe_inv_kernel: calls of a handful of candidates as ONE product with W = L^-1, the DIRECT probe
of bopy/optimizer.py:95-107) against the golden vectors of the unmodified reference, the oracle and the other two
paths.  Everything goes through the C ABI (`bopy_gp_set_inverse_path` selects it).  Run on the B200 box: -m gpu."""
import numpy as np
import pytest

from conftest import golden_names, golden_state
from oracle import gp_oracle as O
from parity_util import check_mean, check_var, prior_var
from test_gpu_parity import cached_native, native_for, select_path

pytestmark = pytest.mark.gpu

WANT = dict(want_mean=True, want_var=True, want_acq=True, want_min=True)
BLOCKED = [n for n in golden_names() if golden_state(n)[1].X_train.shape[0] > 32]   # n <= 32 is small_n_kernel's


def run(gp, xs, path, acq="ei", eta=0.0, **kw):
    select_path(gp, path)
    out = gp.sweep(gp.candidates(xs), acq=acq, eta=eta, index_base=77, **WANT, **kw)
    return {k: out[k].cpu().numpy() for k in ("mean", "var", "acq", "min_val", "min_idx")}


def capacity(gp):
    gp.set_latency_path(1 << 30)
    return gp.set_inverse_path(1)


@pytest.mark.parametrize("name", BLOCKED)
def test_inverse_path_against_the_reference_and_the_other_paths(name):
    g, st, gp = cached_native(name, "f64")
    cap = capacity(gp)
    n = st.X_train.shape[0]
    assert cap == (8 if n <= 3072 else 2)
    eta, pv = float(g["eta"]), prior_var(st)
    for m in sorted({1, 2, 3, 5, 8} & set(range(1, cap + 1))):
        for off in (0, 11):
            xs = g["Xs"][off:off + m]
            inv = run(gp, xs, "inverse", eta=eta)
            # the reference's own numbers, then the oracle
            o_mean, o_var, _, _ = O.acquisition_sweep(st, "ei", xs, eta=eta)
            for ref_mean, ref_var in ((g["mean"][off:off + m], g["var"][off:off + m]), (o_mean, o_var)):
                err, bound = check_mean(inv["mean"], ref_mean, st, "f64")
                assert (err <= bound).all(), (m, off, float(np.max(err / bound)))
                err, bound = check_var(inv["var"], ref_var, st, "f64")
                assert (err <= bound).all(), (m, off, float(np.max(err / bound)))
            # epilogue on its own moments, arg-min over its own values
            with np.errstate(invalid="ignore"):
                own = O.acquisition("ei", inv["mean"], inv["var"], eta=eta)
            assert np.array_equal(np.isnan(inv["acq"]), np.isnan(own))
            ok = ~np.isnan(own)
            np.testing.assert_allclose(inv["acq"][ok], own[ok], rtol=1e-11, atol=1e-13 * max(1.0, np.abs(own[ok]).max(initial=0)))
            assert int(inv["min_idx"][0]) - 77 == int(np.argmin(inv["acq"]))
            assert inv["min_val"][0] == inv["acq"][int(inv["min_idx"][0]) - 77] or np.isnan(inv["min_val"][0])
            # the chained latency path: same K*, same epilogue, another order of the n products
            lat = run(gp, xs, "latency", eta=eta)
            np.testing.assert_allclose(inv["var"], lat["var"], rtol=1e-9, atol=1e-11 * pv)
            err, bound = check_mean(inv["mean"], lat["mean"], st, "f64")
            assert (err <= bound).all()


def test_a_value_does_not_depend_on_the_call_shape():
    g, st, gp = cached_native("c4_hartmann6_n2048", "f64")
    eta = float(g["eta"])
    full = run(gp, g["Xs"][:8], "inverse", eta=eta)
    for m, off in ((1, 0), (1, 7), (2, 3), (3, 5), (4, 4), (5, 0), (7, 1)):
        part = run(gp, g["Xs"][off:off + m], "inverse", eta=eta)
        for k in ("mean", "var", "acq"):
            assert np.array_equal(part[k], full[k][off:off + m], equal_nan=True), (k, m, off)
    # larger calls fall through to probe_kernel and still agree to rounding
    many = run(gp, g["Xs"][:9], "inverse", eta=eta)
    lat = run(gp, g["Xs"][:9], "latency", eta=eta)
    for k in ("mean", "var", "acq"):
        assert np.array_equal(many[k], lat[k], equal_nan=True)
    np.testing.assert_allclose(many["var"][:8], full["var"], rtol=1e-9, atol=1e-11 * prior_var(st))


@pytest.mark.parametrize("kind", ["matern05_d2", "matern15_d2", "matern25_d2", "ard_amp_white"])
@pytest.mark.parametrize("acq", ["lcb", "poi"])
def test_other_kernels_and_acquisitions(kind, acq):
    g, st, gp = cached_native(kind, "f64")
    eta = float(g["eta"])
    ref = g[{"lcb": "lcb_2.0", "poi": "poi"}[acq]]
    spread = float(np.nanmax(ref) - np.nanmin(ref)) or 1.0
    for off in range(0, 40, 8):
        inv = run(gp, g["Xs"][off:off + 8], "inverse", acq=acq, eta=eta, kappa=2.0)
        with np.errstate(invalid="ignore"):
            own = O.acquisition(acq, inv["mean"], inv["var"], eta=eta, kappa=2.0)
        ok = ~np.isnan(own)
        assert np.array_equal(np.isnan(inv["acq"]), ~ok)
        np.testing.assert_allclose(inv["acq"][ok], own[ok], rtol=1e-11, atol=1e-13 * max(1.0, np.abs(own[ok]).max(initial=0)))
        r = ref[off:off + 8]
        resolved = (np.abs(g["var"][off:off + 8]) > 1e-6 * prior_var(st)) & ~np.isnan(r)
        np.testing.assert_allclose(inv["acq"][resolved], r[resolved], rtol=1e-7, atol=1e-7 * spread)


def test_auto_mode_switches_at_the_16th_small_call_and_every_state_change_drops_w():
    g, st = golden_state("c3_branin_n256")
    gp = native_for(st, "f64")
    eta = float(g["eta"])
    x = g["Xs"][5:6]
    lat, inv = run(gp, x, "latency", eta=eta), run(gp, x, "inverse", eta=eta)
    gp.set_state(st.X_train, st.L, st.alpha, st.kernel.length_scale, amplitude=st.kernel.amplitude,
                 noise_level=st.kernel.noise_level, y_mean=st.y_mean, y_std=st.y_std)      # drops W
    gp.set_latency_path(1 << 30)
    assert gp.set_inverse_path(-1) == 8
    xs = gp.candidates(x)
    for call in range(1, 41):
        out = gp.sweep(xs, acq="ei", eta=eta, want_var=True, want_acq=True)
        want = lat if call < 16 else inv
        assert np.array_equal(out["var"].cpu().numpy(), want["var"]), call
        assert np.array_equal(out["acq"].cpu().numpy(), want["acq"], equal_nan=True), call
    # a call that is too large for the path neither uses W nor counts
    big = gp.sweep(gp.candidates(g["Xs"][:64]), acq="ei", eta=eta, want_var=True)["var"].cpu().numpy()
    assert np.array_equal(big, run(gp, g["Xs"][:64], "latency", eta=eta)["var"])
    with pytest.raises(Exception, match="mode must be"):
        gp.set_inverse_path(2)
    gp.close()


def test_off_when_the_latency_path_is_off_and_on_fp32_handles():
    g, st, gp = cached_native("c3_branin_n256", "f64")
    gp.set_latency_path(0)
    assert gp.set_inverse_path(1) == 0
    eta = float(g["eta"])
    swp = run(gp, g["Xs"][:4], "sweep", eta=eta)
    gp.set_latency_path(0)
    gp.set_inverse_path(1)
    out = gp.sweep(gp.candidates(g["Xs"][:4]), acq="ei", eta=eta, **WANT)
    assert np.array_equal(out["var"].cpu().numpy(), swp["var"])


def test_fp32_handles_are_served_in_fp64():
    """An fp32-mode handle keeps the fp64 factor, inv(L_II), X / l and alpha_ -- all the inverse path needs: DIRECT probes of an
    fp32 surrogate get fp64 answers (the fp64 parity bounds hold) instead of a sweep_tc_kernel launch per probe."""
    for name in ("c4_hartmann6_n2048", "c3_branin_n256"):
        g, st, gp = cached_native(name, "f32")
        assert gp.set_latency_path(4096) == 0                    # still no chained latency path on fp32 handles
        assert gp.set_inverse_path(1) == 8
        eta = float(g["eta"])
        for m in (1, 8):
            out = gp.sweep(gp.candidates(g["Xs"][:m]), acq="ei", eta=eta, index_base=5, **WANT)
            mean, var = out["mean"].cpu().numpy(), out["var"].cpu().numpy()
            err, bound = check_mean(mean, g["mean"][:m], st, "f64")
            assert (err <= bound).all()
            err, bound = check_var(var, g["var"][:m], st, "f64")
            assert (err <= bound).all()
            assert int(out["min_idx"][0]) - 5 == int(np.argmin(out["acq"].cpu().numpy()))
        big = gp.sweep(gp.candidates(g["Xs"][:64]), want_var=True)["var"].cpu().numpy()      # larger calls: the fp32-mode sweep
        err, bound = check_var(big, g["var"][:64], st, "f32", name)
        assert (err <= bound).all()
        gp.set_inverse_path(0)


def test_device_fit_then_growth_by_one_point():
    """bopy_gp_fit keeps the factor W is built from; bopy_gp_append (the Kriging believer / one trial more) drops W and
    the next probes see the grown state."""
    from bopy_b200 import _native
    rng = np.random.default_rng(3)
    n, d = 300, 3
    X = rng.random((n + 1, d))
    y = np.sin(4 * X[:, 0]) + X[:, 1] ** 2 - X[:, 2]
    spec = O.KernelSpec(kind="rbf", length_scale=np.array([0.15, 0.2, 0.25]), amplitude=1.3)
    probes = np.vstack([rng.random((5, d)), X[n:n + 1] + 1e-3])

    def fitted(k):
        st = O.fit_state(X[:k], y[:k], spec, 1e-6, True)
        gp = _native.NativeGP(k, d, kernel="rbf", dtype="f64")
        gp.fit(X[:k], (y[:k] - st.y_mean) / st.y_std, spec.length_scale, amplitude=spec.amplitude, noise_level=0.0,
               alpha_reg=1e-6, y_mean=st.y_mean, y_std=st.y_std)
        return st, gp

    st0, gp = fitted(n)
    assert capacity(gp) == 8
    eta = float(y[:n].min())
    before = run(gp, probes, "inverse", eta=eta)
    o_mean, o_var, _, _ = O.acquisition_sweep(st0, "ei", probes, eta=eta)
    err, bound = check_var(before["var"], o_var, st0, "f64")
    assert (err <= bound).all()
    st1 = O.fit_state(X, y, spec, 1e-6, True)
    gp.append(X, (y - st1.y_mean) / st1.y_std, y_mean=st1.y_mean, y_std=st1.y_std)
    after = run(gp, probes, "inverse", eta=eta)
    o_mean, o_var, _, _ = O.acquisition_sweep(st1, "ei", probes, eta=eta)
    err, bound = check_mean(after["mean"], o_mean, st1, "f64")
    assert (err <= 10 * bound).all()        # the appended row is equal to a fresh fit to rounding (tests/test_gpu_fit.py)
    np.testing.assert_allclose(after["var"], o_var, rtol=1e-8, atol=1e-10 * prior_var(st1))
    assert after["var"][-1] < 0.05 * before["var"][-1]      # the probe next to the new point knows about it
    gp.close()


def test_default_surrogate_serves_direct_style_probes_from_w():
    """Public API, library defaults: acquisition(x.reshape(1, -1)) once per probe (bopy/optimizer.py:96-97), 100 probes."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    from bopy_b200.acquisition import EI
    from bopy_b200.surrogate import B200GPSurrogate
    rng = np.random.default_rng(11)
    X = rng.random((700, 3))
    y = np.sin(4 * X[:, 0]) + X[:, 1] ** 2 - X[:, 2]
    gp = GaussianProcessRegressor(kernel=ConstantKernel(1.3) * RBF([0.3, 0.4, 0.5]), alpha=1e-6, normalize_y=True,
                                  optimizer=None)
    probes = rng.random((100, 3))
    got = {}
    for mode in ("auto", True, False):
        sur = B200GPSurrogate(gp, inverse_path=mode)
        sur.fit(X, y)
        ei = EI(sur)
        ei.fit(X, y)
        got[mode] = np.array([ei(p.reshape(1, -1))[0] for p in probes])
    st = O.fit_state(X, y, O.KernelSpec(kind="rbf", length_scale=np.array([0.3, 0.4, 0.5]), amplitude=1.3), 1e-6, True)
    _, _, want, _ = O.acquisition_sweep(st, "ei", probes, eta=float(y.min()))
    for mode in got:
        np.testing.assert_allclose(got[mode], want, rtol=1e-6, atol=1e-9 * np.ptp(want))
    assert np.array_equal(got["auto"][:15], got[False][:15])      # probe_kernel until W exists ...
    assert np.array_equal(got["auto"][15:], got[True][15:])       # ... probe_inv_kernel from the 16th probe on


def test_direct_builds_w_at_its_first_probe_when_its_budget_is_large():
    """DirectOptimizer (bopy/optimizer.py:70-107) knows its probe budget: with maxf well beyond the library's own
    switch-over it asks for W = L^-1 at the first probe; the optimum is the one the chained path finds."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    from bopy_b200.acquisition import LCB
    from bopy_b200.bounds import Bound, Bounds
    from bopy_b200.optimizer import DirectOptimizer
    from bopy_b200.surrogate import B200GPSurrogate
    rng = np.random.default_rng(5)
    X = rng.random((200, 2))
    y = np.sin(5 * X[:, 0]) * np.cos(3 * X[:, 1])
    bounds = Bounds([Bound(0.0, 1.0), Bound(0.0, 1.0)])
    found = {}
    for mode in ("auto", False):
        gp = GaussianProcessRegressor(kernel=ConstantKernel(1.0) * RBF([0.2, 0.2]), alpha=1e-6, normalize_y=True, optimizer=None)
        sur = B200GPSurrogate(gp, inverse_path=mode)
        sur.fit(X, y)
        acq = LCB(sur, kappa=2.0)
        acq.fit(X, y)
        res = DirectOptimizer(acq, bounds, maxf=400).optimize()
        found[mode] = (res.x_min, res.f_min, sur.native.launch_info(1)["launches"])
    assert found["auto"][2] == 1 and found[False][2] == 2      # probe_inv_kernel alone / probe_kernel + the arg-min finalize
    np.testing.assert_allclose(found["auto"][1], found[False][1], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(found["auto"][0], found[False][0], atol=1e-6)


def test_host_buffer_entry_runs_the_same_kernel_without_copies():
    """bopy_acq_eval_host (numpy in, numpy out: the DIRECT probe) hands the candidates over as kernel parameters and reads
    the results from mapped pinned memory: the numbers are those of the device-buffer entry, bit for bit."""
    g, st, gp = cached_native("c4_hartmann6_n2048", "f64")
    eta = float(g["eta"])
    for m in (1, 3, 8, 9):
        xs = np.ascontiguousarray(g["Xs"][20:20 + m])
        dev = run(gp, xs, "inverse", eta=eta)
        a, mu, var = gp.eval_host(xs, acq="ei", eta=eta, want_acq=True, want_mean=True, want_var=True)
        assert np.array_equal(a, dev["acq"], equal_nan=True) and np.array_equal(mu, dev["mean"]) and np.array_equal(var, dev["var"])
        only_acq, none_m, none_v = gp.eval_host(xs, acq="ei", eta=eta)
        assert np.array_equal(only_acq, dev["acq"], equal_nan=True) and none_m is None and none_v is None
        _, mu2, var2 = gp.eval_host(xs, acq=None, want_acq=False, want_mean=True, want_var=True)
        assert np.array_equal(mu2, dev["mean"]) and np.array_equal(var2, dev["var"])


@pytest.mark.parametrize("name", ["ref_forrester_matern15_fixed", "c1_forrester_rbf_n6", "edge_alpha0_nan", "edge_on_training_points"])
def test_host_buffer_entry_on_small_n_handles(name):
    """n <= 32 (the reference's own examples, DIRECT on ~10 observations): small_n_kernel takes up to 8 candidates of a
    host-buffer call as kernel parameters and writes to mapped pinned memory; same numbers as the device-buffer entry."""
    g, st, gp = cached_native(name, "f64")
    eta = float(g["eta"])
    for m in (1, 5, 8, 9, 40):
        xs = np.ascontiguousarray(g["Xs"][3:3 + m])
        dev = run(gp, xs, "latency", eta=eta)
        a, mu, var = gp.eval_host(xs, acq="ei", eta=eta, want_acq=True, want_mean=True, want_var=True)
        assert np.array_equal(a, dev["acq"], equal_nan=True) and np.array_equal(mu, dev["mean"])
        assert np.array_equal(var, dev["var"], equal_nan=True)
        (only_var,) = [o for o in gp.eval_host(xs, acq=None, want_acq=False, want_var=True) if o is not None]
        assert np.array_equal(only_var, dev["var"], equal_nan=True)


@pytest.mark.parametrize("name", ["c4_hartmann6_n2048", "c3_branin_n256", "ref_forrester_matern15_fixed"])
def test_one_point_predict_is_the_posterior_variance_on_the_fast_path(name):
    """Surrogate.predict on ONE point (bopy/surrogate.py:83-92 as a DIRECT objective or the Kriging believer call it): the
    1 x 1 covariance is the variance of the latency / inverse path, not a walk of one thread block over all of L."""
    g, st, gp = cached_native(name, "f64")
    select_path(gp, "inverse")
    pv = prior_var(st)
    for i in (0, 7):
        xs = gp.candidates(g["Xs"][i:i + 2])
        mean2, cov2 = gp.predict_cov(xs)
        mean1, cov1 = gp.predict_cov(xs[:1])
        assert tuple(cov1.shape) == (1, 1) and tuple(mean1.shape) == (1,)
        np.testing.assert_allclose(cov1.cpu().numpy()[0, 0], cov2.cpu().numpy()[0, 0], rtol=1e-9, atol=1e-11 * pv)
        err, bound = check_mean(mean1.cpu().numpy(), mean2.cpu().numpy()[:1], st, "f64")
        assert (err <= bound).all()
        err, bound = check_var(cov1.cpu().numpy()[0], g["var"][i:i + 1], st, "f64")
        assert (err <= bound).all()


def test_surrogate_predict_of_one_point_through_the_public_api():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    from bopy_b200.surrogate import B200GPSurrogate
    rng = np.random.default_rng(2)
    X = rng.random((300, 3))
    y = np.sin(4 * X[:, 0]) + X[:, 1] ** 2 - X[:, 2]
    gp = GaussianProcessRegressor(kernel=ConstantKernel(1.3) * RBF([0.15, 0.2, 0.25]), alpha=1e-6, normalize_y=True, optimizer=None)
    sur = B200GPSurrogate(gp)
    sur.fit(X, y)
    xs = rng.random((6, 3))
    mean_all, cov_all = sur.predict(xs)
    for i in range(6):
        mean_i, cov_i = sur.predict(xs[i:i + 1])
        assert mean_i.shape == (1,) and cov_i.shape == (1, 1)
        np.testing.assert_allclose(mean_i[0], mean_all[i], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(cov_i[0, 0], cov_all[i, i], rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("name", ["c4_hartmann6_n2048", "c3_branin_n256", "ragged_n333_d4_opt", "matern25_d2"])
def test_full_covariance_of_a_small_set_from_the_latency_path(name):
    """bopy_gp_predict_cov (Surrogate.predict, bopy/surrogate.py:83-92) on a few hundred points: V comes from probe_kernel
    (the solve spread over the block rows of L) and agrees with the V-exporting throughput sweep to rounding."""
    g, st, gp = cached_native(name, "f64")
    pv = prior_var(st)
    for m in (2, 9, 33, min(300, len(g["Xs"]))):
        xs = gp.candidates(g["Xs"][:m])
        select_path(gp, "latency")
        mean_l, cov_l = (t.cpu().numpy() for t in gp.predict_cov(xs))
        select_path(gp, "sweep")
        mean_s, cov_s = (t.cpu().numpy() for t in gp.predict_cov(xs))
        np.testing.assert_allclose(cov_l, cov_s, rtol=1e-9, atol=1e-11 * pv)
        err, bound = check_mean(mean_l, mean_s, st, "f64")
        assert (err <= bound).all()
        assert np.array_equal(cov_l, cov_l.T)
        c = g["cov_corner"].shape[0]
        if m >= c:
            np.testing.assert_allclose(cov_l[:c, :c], g["cov_corner"], rtol=1e-9, atol=1e-11 * pv)
