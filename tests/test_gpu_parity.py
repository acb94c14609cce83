"""Parity of the CUDA path against the CPU oracle and the golden vectors of the unmodified reference.
All calls go through the C ABI (bopy_b200/_native.py -> libbopy_b200.so).  Run on the B200 box: -m gpu."""
import numpy as np
import pytest

from conftest import golden_names, golden_state
from oracle import gp_oracle as O
from parity_util import TOL, check_mean, check_var, is_stated_tie, prior_var

pytestmark = pytest.mark.gpu

NAMES = golden_names()
KERNEL_NAME = {("rbf", 1.5): "rbf", ("matern", 0.5): "matern12", ("matern", 1.5): "matern32", ("matern", 2.5): "matern52"}


def native_for(state, dtype):
    from bopy_b200 import _native
    k = state.kernel
    name = "rbf" if k.kind == "rbf" else KERNEL_NAME[("matern", k.nu)]
    n, d = state.X_train.shape
    gp = _native.NativeGP(n, d, kernel=name, dtype=dtype)
    gp.set_state(state.X_train, state.L, state.alpha, k.length_scale, amplitude=k.amplitude,
                 noise_level=k.noise_level, y_mean=state.y_mean, y_std=state.y_std)
    return gp


_NATIVE = {}

# which kernel serves a call: "latency" = probe_kernel for every m it can hold (fp64 handles only),
# "sweep" = the throughput kernel for every m
PATHS = ["latency", "sweep"]
DTYPE_PATHS = [("f64", "latency"), ("f64", "sweep"), ("f32", "sweep")]


def select_path(gp, path):
    """"inverse": calls of a handful of candidates go to probe_inv_kernel (W = L^-1 built at the first one), larger ones to
    probe_kernel; "latency": probe_kernel for every m, whatever the number of calls before."""
    eff = gp.set_latency_path(0 if path == "sweep" else 1 << 30)
    if path in ("latency", "inverse"):
        assert eff >= 4096
    gp.set_inverse_path(1 if path == "inverse" else 0)
    return gp


def cached_native(name, dtype, path="sweep"):
    key = (name, dtype)
    if key not in _NATIVE:
        g, st = golden_state(name)
        _NATIVE[key] = (g, st, native_for(st, dtype))
    g, st, gp = _NATIVE[key]
    return g, st, select_path(gp, path)


@pytest.mark.parametrize("dtype,path", DTYPE_PATHS)
@pytest.mark.parametrize("name", NAMES)
def test_posterior_diag_matches_reference(name, dtype, path):
    g, st, gp = cached_native(name, dtype, path)
    out = gp.sweep(gp.candidates(g["Xs"]), want_mean=True, want_var=True)
    mean, var = out["mean"].cpu().numpy(), out["var"].cpu().numpy()
    o_mean, o_var = O.posterior_diag(st, g["Xs"])
    for ref_mean, ref_var, who in ((g["mean"], g["var"], "golden"), (o_mean, o_var, "oracle")):
        err, bound = check_mean(mean, ref_mean, st, dtype)
        assert (err <= bound).all(), f"mean vs {who}: worst {np.max(err / bound):.3g}x the bound"
        err, bound = check_var(var, ref_var, st, dtype, name)
        assert (err <= bound).all(), f"var vs {who}: worst {np.max(err / bound):.3g}x the bound, max err {err.max():.3g}"


@pytest.mark.parametrize("dtype,path", DTYPE_PATHS)
@pytest.mark.parametrize("acq", ["lcb", "ei", "poi"])
@pytest.mark.parametrize("name", NAMES)
def test_acquisition_and_argmin(name, acq, dtype, path):
    g, st, gp = cached_native(name, dtype, path)
    eta, kappa = float(g["eta"]), 2.0
    xs = gp.candidates(g["Xs"])
    out = gp.sweep(xs, acq=acq, eta=eta, kappa=kappa, want_mean=True, want_var=True, want_acq=True, want_min=True,
                   index_base=1000)
    mean, var, a = (out[k].cpu().numpy() for k in ("mean", "var", "acq"))
    # (1) the epilogue itself: oracle formula on the kernel's OWN moments -> tight, NaNs in the same places
    with np.errstate(invalid="ignore"):
        a_own = O.acquisition(acq, mean, var, eta=eta, kappa=kappa)
    assert np.array_equal(np.isnan(a), np.isnan(a_own))
    ok = ~np.isnan(a)
    scale = max(1.0, float(np.max(np.abs(a_own[ok])))) if ok.any() else 1.0
    np.testing.assert_allclose(a[ok], a_own[ok], rtol=1e-11, atol=1e-13 * scale)
    # (2) the fused arg-min equals np.argmin of the values the kernel produced (first-min / first-NaN rules)
    idx, val = int(out["min_idx"].item()), float(out["min_val"].item())
    assert idx - 1000 == int(np.argmin(a))
    assert (np.isnan(val) and np.isnan(a[idx - 1000])) or val == a[idx - 1000]
    # (3) against the reference: same index, or a stated tie
    ref = g[{"lcb": "lcb_2.0", "ei": "ei", "poi": "poi"}[acq]]
    if not np.isnan(ref).any() and not np.isnan(a).any():
        assert is_stated_tie(ref, int(np.argmin(ref)), idx - 1000, dtype), \
            f"argmin {idx - 1000} vs reference {int(np.argmin(ref))}: {ref[idx - 1000]!r} vs {ref.min()!r}"
    # (4) values against the reference where the variance is resolved
    pv = prior_var(st)
    resolved = np.abs(g["var"]) > (1e-6 if dtype == "f64" else 1e-1) * pv
    if resolved.any() and not np.isnan(ref[resolved]).any():
        spread = float(np.ptp(ref[resolved])) or 1.0
        tol = 1e-7 if dtype == "f64" else 2e-3
        np.testing.assert_allclose(a[resolved], ref[resolved], rtol=tol, atol=tol * spread)


@pytest.mark.parametrize("name", [n for n in NAMES if not n.startswith("c5")])
def test_predict_full_covariance_contract(name):
    g, st, gp = cached_native(name, "f64")
    c = g["cov_corner"].shape[0]
    mean, cov = gp.predict_cov(gp.candidates(g["Xs"][:c]))
    mean, cov = mean.cpu().numpy(), cov.cpu().numpy()
    assert mean.shape == (c,) and cov.shape == (c, c)
    err, bound = check_mean(mean, g["mean"][:c], st, "f64")
    assert (err <= bound).all()
    pv = prior_var(st)
    np.testing.assert_allclose(cov, g["cov_corner"], rtol=1e-9, atol=1e-11 * pv)
    assert np.allclose(cov, cov.T, rtol=0, atol=1e-13 * pv)


def test_candidate_generator_bit_exact():
    from bopy_b200 import _native
    lo, hi = [-5.0, 0.0, 1.0, 0.25, -1e3, 0.0], [10.0, 15.0, 2.0, 0.5, 1e3, 1e-3]
    for base, m in ((0, 1000), (123456789012, 513), (1 << 40, 7)):
        got = _native.candidates_uniform(1235, base, m, lo, hi).cpu().numpy()
        assert np.array_equal(got, O.candidates_uniform(1235, base, m, lo, hi))


@pytest.mark.parametrize("dtype,path", DTYPE_PATHS)
def test_ragged_candidate_counts_and_position_independence(dtype, path):
    g, st, gp = cached_native("ragged_n333_d4_opt", dtype, path)
    Xs = g["Xs"]
    full = gp.sweep(gp.candidates(Xs), acq="ei", eta=float(g["eta"]), want_mean=True, want_var=True, want_acq=True)
    full = {k: full[k].cpu().numpy() for k in ("mean", "var", "acq")}
    for m in (1, 2, 127, 128, 129, 255, 257, 777):
        for off in (0, 37):
            if off + m > len(Xs):
                continue
            part = gp.sweep(gp.candidates(Xs[off:off + m]), acq="ei", eta=float(g["eta"]), want_mean=True,
                            want_var=True, want_acq=True, want_min=True)
            for k in ("mean", "var", "acq"):   # a candidate's arithmetic does not depend on where it sits
                assert np.array_equal(part[k].cpu().numpy(), full[k][off:off + m], equal_nan=True), (k, m, off)
            assert int(part["min_idx"].item()) == int(np.argmin(full["acq"][off:off + m]))


@pytest.mark.parametrize("path", PATHS)
def test_nan_rules_end_to_end(path):
    g, st, gp = cached_native("edge_alpha0_nan", "f64", path)
    out = gp.sweep(gp.candidates(g["Xs"]), acq="ei", eta=float(g["eta"]), want_var=True, want_acq=True, want_min=True)
    var, a = out["var"].cpu().numpy(), out["acq"].cpu().numpy()
    bad = ~(np.sqrt(np.where(var < 0, np.nan, var)) > 0)
    assert np.array_equal(np.isnan(a), bad)
    if bad.any():                                  # np.argmin: the first NaN wins
        assert int(out["min_idx"].item()) == int(np.flatnonzero(bad)[0])
        assert np.isnan(out["min_val"].item())


def test_sharded_argmin_equals_global_argmin():
    """The multi-GPU reduction on one GPU: min-loc of per-slice arg-mins == arg-min of the whole range."""
    from bopy_b200 import _native
    from bopy_b200.distributed import reduce_minloc, shard_range
    g, st, gp = cached_native("c3_branin_n256", "f64", "sweep")
    lo, hi, m = [-5.0, 0.0], [10.0, 15.0], 100_003
    eta = float(g["eta"])
    whole = gp.sweep(_native.candidates_uniform(99, 0, m, lo, hi), acq="ei", eta=eta, want_min=True)
    for world in (2, 8):
        recs = []
        for r in range(world):
            s, e = shard_range(m, r, world)
            o = gp.sweep(_native.candidates_uniform(99, s, e - s, lo, hi), acq="ei", eta=eta, want_min=True, index_base=s)
            recs.append((float(o["min_val"].item()), int(o["min_idx"].item())))
        val, idx = reduce_minloc(recs)
        assert idx == int(whole["min_idx"].item()) and val == float(whole["min_val"].item())


def test_full_size_c4_properties():
    """BASELINE config C4 at full per-GPU size (n=2048, d=6, 2^21 candidates): size-independent properties."""
    from bopy_b200 import _native
    g, st, gp = cached_native("c4_hartmann6_n2048", "f64", "sweep")
    m, eta = 1 << 21, float(g["eta"])
    xs = _native.candidates_uniform(1235, 0, m, np.zeros(6), np.ones(6))
    out = gp.sweep(xs, acq="ei", eta=eta, want_mean=True, want_var=True, want_acq=True, want_min=True)
    a = out["acq"].cpu().numpy()
    var = out["var"].cpu().numpy()
    pv = prior_var(st)
    # EI is a negated expectation of a positive part: <= 0 up to rounding of the two cancelling terms
    assert np.isfinite(a).all() and (a <= 1e-12 * np.abs(a).max()).all()
    assert (var > 0).all() and (var <= pv * (1 + 1e-12)).all()          # 0 < posterior var <= prior var
    assert int(out["min_idx"].item()) == int(np.argmin(a))
    # the first 1024 candidates ARE the golden candidate set (same generator, same seed)
    err, bound = check_var(var[:1024], g["var"], st, "f64")
    assert (err <= bound).all()
    np.testing.assert_array_equal(xs[:1024].cpu().numpy(), g["Xs"])
    # checksum-of-checksums: a strided sub-sweep reproduces the same values bit for bit
    sub = gp.sweep(xs[::4097].contiguous(), acq="ei", eta=eta, want_acq=True)["acq"].cpu().numpy()
    assert np.array_equal(sub, a[::4097])
    # the latency path on the same subset: same arithmetic, another order of a few partial sums
    select_path(gp, "latency")
    lat = gp.sweep(xs[::4097].contiguous(), acq="ei", eta=eta, want_acq=True, want_var=True, want_min=True)
    np.testing.assert_allclose(lat["var"].cpu().numpy(), var[::4097], rtol=0, atol=1e-13 * pv)
    np.testing.assert_allclose(lat["acq"].cpu().numpy(), sub, rtol=1e-9, atol=1e-12 * np.abs(sub).max())
    assert int(lat["min_idx"].item()) == int(np.argmin(lat["acq"].cpu().numpy()))


@pytest.mark.parametrize("path", PATHS)
def test_interpolation_property_at_training_points(path):
    g, st, gp = cached_native("c3_branin_n256", "f64", path)
    out = gp.sweep(gp.candidates(st.X_train), want_mean=True, want_var=True)
    mean, var = out["mean"].cpu().numpy(), out["var"].cpu().numpy()
    assert np.max(np.abs(mean - g["y"])) < 1e-3 * np.ptp(g["y"])       # alpha_reg = 1e-6: near-interpolation
    assert (np.abs(var) < 1e-4 * prior_var(st)).all()


def test_c_abi_error_behaviour():
    import ctypes
    from bopy_b200 import _native
    lib = _native.load()
    h = ctypes.c_void_p()
    assert lib.bopy_gp_create(ctypes.byref(h), 0, 0, 0, 0, 2) == _native.ERR_BAD_ARG       # n < 1
    assert b"n must be" in lib.bopy_last_error()
    assert lib.bopy_gp_create(ctypes.byref(h), 0, 0, 0, 10, 33) == _native.ERR_UNSUPPORTED  # d > 32
    assert lib.bopy_gp_create(ctypes.byref(h), 0, 7, 0, 10, 2) == _native.ERR_BAD_ARG       # dtype
    assert lib.bopy_gp_create(ctypes.byref(h), 0, 0, 0, 10, 2) == _native.OK
    rc = lib.bopy_gp_posterior_acq(h, ctypes.c_void_p(8), 1, 1, 0.0, 2.0, None, None, None, 0, None, None, None)
    assert rc == _native.ERR_NOT_READY                                                       # before set_state
    lib.bopy_gp_destroy(h)
    g, st, gp = cached_native("c3_branin_n256", "f64")
    xs = gp.candidates(g["Xs"][:4])
    with pytest.raises(Exception, match="unknown acquisition"):
        _native.check(lib.bopy_gp_posterior_acq(gp._handle, xs.data_ptr(), 4, 9, 0.0, 2.0, None, None, None, 0, None,
                                                None, None), "posterior_acq")
    with pytest.raises(Exception, match="m must be"):
        _native.check(lib.bopy_gp_posterior_acq(gp._handle, xs.data_ptr(), 0, 1, 0.0, 2.0, None, None, None, 0, None,
                                                None, None), "posterior_acq")


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("n,d", [(1, 1), (2, 32), (127, 3), (128, 1), (129, 5), (257, 32)])
def test_extreme_shapes_against_the_oracle(n, d, path):
    """Smallest / largest supported dimensionality, n around the 128-row block edge, candidates far away."""
    from bopy_b200 import _native
    rng = np.random.default_rng(100 * n + d)
    X = rng.random((n, d))
    y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(n)
    spec = O.KernelSpec(kind="rbf", length_scale=0.5 + rng.random(d), amplitude=1.7)
    st = O.fit_state(X, y, spec, 1e-6, normalize_y=True)
    gp = select_path(native_for(st, "f64"), path)
    xs = np.concatenate([rng.random((300, d)), 50.0 + rng.random((11, d))])     # the last 11 are far-field
    out = gp.sweep(gp.candidates(xs), acq="ei", eta=float(y.min()), want_mean=True, want_var=True, want_acq=True,
                   want_min=True, index_base=1 << 40)
    mean, var, a = (out[k].cpu().numpy() for k in ("mean", "var", "acq"))
    o_mean, o_var, o_a, (o_idx, _) = O.acquisition_sweep(st, "ei", xs, eta=float(y.min()))
    err, bound = check_mean(mean, o_mean, st, "f64")
    # dense 1-D designs make alpha_ huge (|alpha|_1 ~ 1e6): the mean is a cancelling dot product whose value moves by
    # eps * |k*|.|alpha| per ulp of exp() -- the oracle's libm and the kernel's exp_nonpos differ by up to 2 ulp
    dot_cond = st.y_std * (np.abs(O.kernel_cross(st.kernel, xs, st.X_train)) @ np.abs(st.alpha))
    assert (err <= bound + 16 * np.finfo(np.float64).eps * dot_cond).all()
    err, bound = check_var(var, o_var, st, "f64")
    assert (err <= bound).all()
    # far field: the posterior is the prior -> exactly the prior variance, one common acquisition value
    assert np.all(var[-11:] == prior_var(st)) and np.ptp(a[-11:]) == 0.0 and np.ptp(mean[-11:]) == 0.0
    idx = int(out["min_idx"].item()) - (1 << 40)
    assert idx == int(np.argmin(a)) and is_stated_tie(o_a, o_idx, idx, "f64")


def test_c_abi_rejects_unsupported_dimensionality():
    from bopy_b200 import _native
    with pytest.raises(Exception, match="exceeds the supported maximum"):
        _native.NativeGP(10, 33)
