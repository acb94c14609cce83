"""The latency path (probe_kernel: the solve of a small candidate batch spread over the block rows of L) against
the throughput path (sweep_kernel) and the oracle.  Both go through the C ABI; `bopy_gp_set_latency_path`
selects which one serves a call.  Run on the B200 box: -m gpu."""
import numpy as np
import pytest

from conftest import golden_names
from oracle import gp_oracle as O
from parity_util import check_mean, check_var, prior_var
from test_gpu_parity import cached_native, native_for, select_path

pytestmark = pytest.mark.gpu

WANT = dict(want_mean=True, want_var=True, want_acq=True, want_min=True)


def run(gp, xs, path, acq="ei", eta=0.0, **kw):
    select_path(gp, path)
    out = gp.sweep(gp.candidates(xs), acq=acq, eta=eta, index_base=77, **WANT, **kw)
    return {k: out[k].cpu().numpy() for k in ("mean", "var", "acq", "min_val", "min_idx")}


@pytest.mark.parametrize("name", [n for n in golden_names() if not n.startswith("edge_alpha0")])
def test_two_paths_agree_on_the_golden_sets(name):
    g, st, gp = cached_native(name, "f64")
    eta = float(g["eta"])
    lat, swp = run(gp, g["Xs"], "latency", eta=eta), run(gp, g["Xs"], "sweep", eta=eta)
    pv = prior_var(st)
    # same arithmetic, different order of a few partial sums: rounding-level agreement, scaled by conditioning
    dot_cond = st.y_std * (np.abs(O.kernel_cross(st.kernel, g["Xs"], st.X_train)) @ np.abs(st.alpha))
    assert (np.abs(lat["mean"] - swp["mean"]) <= 64 * np.finfo(np.float64).eps * (dot_cond + np.abs(swp["mean"]))).all()
    np.testing.assert_allclose(lat["var"], swp["var"], rtol=0, atol=1e-12 * pv)
    assert int(lat["min_idx"][0]) == 77 + int(np.nanargmin(lat["acq"])) or np.isnan(lat["acq"]).any()


@pytest.mark.parametrize("m", [1, 2, 7, 8, 9, 15, 16, 17, 31, 33, 64, 100, 255, 256, 300, 1000, 1024])
def test_every_batch_shape_against_the_oracle(m):
    """m walks through the 8 / 16 / 32-candidate batch shapes, ragged tails and several batches per CTA group."""
    g, st, gp = cached_native("c4_hartmann6_n2048", "f64")
    xs, eta = g["Xs"][:m], float(g["eta"])
    lat = run(gp, xs, "latency", eta=eta)
    o_mean, o_var, o_a, (o_idx, o_val) = O.acquisition_sweep(st, "ei", xs, eta=eta)
    err, bound = check_mean(lat["mean"], o_mean, st, "f64")
    assert (err <= bound).all()
    err, bound = check_var(lat["var"], o_var, st, "f64")
    assert (err <= bound).all()
    assert int(lat["min_idx"][0]) - 77 == int(np.argmin(lat["acq"]))
    assert lat["min_val"][0] == lat["acq"][int(lat["min_idx"][0]) - 77]
    # a candidate's arithmetic does not depend on the batch shape or on where it sits
    full = run(gp, g["Xs"], "latency", eta=eta)
    for k in ("mean", "var", "acq"):
        assert np.array_equal(lat[k], full[k][:m], equal_nan=True), k


def test_large_n_many_block_rows():
    """n = 8192: 64 block rows -> 64 CTAs per candidate group, two groups, several batches each."""
    g, st, gp = cached_native("c5_rbf_d20_n8192", "f64")
    eta = float(g["eta"])
    for m in (1, 16, 17, len(g["Xs"])):
        lat = run(gp, g["Xs"][:m], "latency", eta=eta)
        err, bound = check_mean(lat["mean"], g["mean"][:m], st, "f64")
        assert (err <= bound).all()
        err, bound = check_var(lat["var"], g["var"][:m], st, "f64")
        assert (err <= bound).all()


def test_repeated_launches_reuse_flags_and_tickets():
    """The hand-off flags carry a per-launch epoch and roles come from a monotonic ticket: hundreds of back-to-back
    launches of changing shape must neither hang nor see a stale flag."""
    g, st, gp = cached_native("c3_branin_n256", "f64")
    eta = float(g["eta"])
    rng = np.random.default_rng(5)
    ref = run(gp, g["Xs"], "latency", eta=eta)
    for it in range(300):
        m = int(rng.integers(1, len(g["Xs"]) + 1))
        off = int(rng.integers(0, len(g["Xs"]) - m + 1))
        out = run(gp, g["Xs"][off:off + m], "latency", eta=eta)
        assert np.array_equal(out["acq"], ref["acq"][off:off + m], equal_nan=True), (it, m, off)
    # interleaved with the throughput path on the same handle (they share the V workspace)
    swp = run(gp, g["Xs"], "sweep", eta=eta)
    again = run(gp, g["Xs"], "latency", eta=eta)
    assert np.array_equal(again["acq"], ref["acq"], equal_nan=True)
    # C3 is over-determined (posterior variances down to 1e-9 of the prior): EI moves by d(var) / (2 sigma)
    np.testing.assert_allclose(swp["acq"], ref["acq"], rtol=1e-6, atol=1e-7 * np.nanmax(np.abs(ref["acq"])))


@pytest.mark.parametrize("kind", ["matern05_d2", "matern15_d2", "matern25_d2", "ard_amp_white"])
@pytest.mark.parametrize("acq", ["lcb", "poi"])
def test_other_kernels_and_acquisitions(kind, acq):
    g, st, gp = cached_native(kind, "f64")
    eta = float(g["eta"])
    lat = run(gp, g["Xs"], "latency", acq=acq, eta=eta, kappa=2.0)
    with np.errstate(invalid="ignore"):
        own = O.acquisition(acq, lat["mean"], lat["var"], eta=eta, kappa=2.0)
    ok = ~np.isnan(own)
    assert np.array_equal(np.isnan(lat["acq"]), ~ok)
    np.testing.assert_allclose(lat["acq"][ok], own[ok], rtol=1e-11, atol=1e-13 * max(1.0, np.abs(own[ok]).max()))
    ref = g[{"lcb": "lcb_2.0", "poi": "poi"}[acq]]
    if not np.isnan(ref).any():
        spread = float(np.ptp(ref)) or 1.0
        resolved = np.abs(g["var"]) > 1e-6 * prior_var(st)
        np.testing.assert_allclose(lat["acq"][resolved], ref[resolved], rtol=1e-7, atol=1e-7 * spread)


def test_fp32_handles_have_no_latency_path():
    g, st, gp = cached_native("c3_branin_n256", "f32")
    assert gp.set_latency_path(4096) == 0


def test_limit_is_clamped_and_reported():
    g, st, gp = cached_native("c3_branin_n256", "f64")
    assert gp.set_latency_path(0) == 0
    assert gp.set_latency_path(100) == 100
    assert gp.set_latency_path(1 << 40) == 148 * 128
    with pytest.raises(Exception, match="max_m must be"):
        gp.set_latency_path(-1)


def test_direct_style_single_point_probes_through_the_public_api():
    """The reference's calling pattern: acquisition(x.reshape(1, -1)) once per probe (bopy/optimizer.py:96-97)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    from bopy_b200.acquisition import EI
    from bopy_b200.surrogate import B200GPSurrogate
    rng = np.random.default_rng(11)
    X = rng.random((700, 3))
    y = np.sin(4 * X[:, 0]) + X[:, 1] ** 2 - X[:, 2]
    gp = GaussianProcessRegressor(kernel=ConstantKernel(1.3) * RBF([0.3, 0.4, 0.5]), alpha=1e-6, normalize_y=True,
                                  optimizer=None)
    sur = B200GPSurrogate(gp, device_fit=False, inverse_path=False)   # every probe on probe_kernel (bit-for-bit check below)
    sur.fit(X, y)
    ei = EI(sur)
    ei.fit(X, y)
    st = O.state_from_sklearn(gp)
    probes = rng.random((40, 3))
    got = np.array([ei(p.reshape(1, -1))[0] for p in probes])
    _, _, want, _ = O.acquisition_sweep(st, "ei", probes, eta=float(y.min()))
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9 * np.ptp(want))
    assert np.array_equal(got, ei(probes))          # one at a time == all at once, bit for bit
    sur_off = B200GPSurrogate(gp, device_fit=False, latency_max_m=0)
    sur_off.fit(X, y)
    ei_off = EI(sur_off)
    ei_off.fit(X, y)
    np.testing.assert_allclose(ei_off(probes), got, rtol=1e-9, atol=1e-12 * np.abs(got).max())


def test_handle_is_reused_while_n_stays_within_its_blocks():
    """A BayesOpt loop adds one point per trial: the handle (310 MB of workspace at n = 2048) is kept as long as
    ceil(n / 128) does not change, and a reused handle gives exactly what a fresh one gives."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel

    from bopy_b200.surrogate import B200GPSurrogate
    rng = np.random.default_rng(21)
    X = rng.random((400, 3))
    y = np.cos(3 * X[:, 0]) + X[:, 1] * X[:, 2]
    xs = rng.random((50, 3))

    def make():
        return B200GPSurrogate(GaussianProcessRegressor(ConstantKernel(1.0) * RBF(0.3 * np.ones(3)), alpha=1e-6,
                                                         normalize_y=True, optimizer=None))
    sur = make()
    sur.fit(X[:300], y[:300])
    first = sur.native
    for n in (301, 384, 290, 257):
        sur.fit(X[:n], y[:n])
        assert sur.native is first and sur.native.n == n
        fresh = make()
        fresh.fit(X[:n], y[:n])
        big = rng.random((5000, 3))                       # throughput path on the reused handle too
        for pts in (xs, big):
            for a, b in zip(sur.predict_diag(pts), fresh.predict_diag(pts)):
                if n == 301:      # 300 -> 301 is a one-row append of the factor: equal to rounding, not bit for bit
                    np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-10 * np.abs(b).max())
                else:
                    assert np.array_equal(a, b)
    sur.fit(X[:385], y[:385])
    assert sur.native is not first and sur.native.n == 385


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_host_buffer_entry_equals_the_device_buffer_entry(dtype):
    g, st, gp = cached_native("ragged_n333_d4_opt", dtype)
    if dtype == "f64":
        select_path(gp, "latency")
    eta = float(g["eta"])
    for m in (1, 5, 300, 2048):
        xs = np.ascontiguousarray(g["Xs"][:m])
        acq, mean, var = gp.eval_host(xs, "ei", eta=eta, want_acq=True, want_mean=True, want_var=True)
        dev = gp.sweep(gp.candidates(xs), acq="ei", eta=eta, want_mean=True, want_var=True, want_acq=True)
        for host, key in ((acq, "acq"), (mean, "mean"), (var, "var")):
            assert np.array_equal(host, dev[key].cpu().numpy(), equal_nan=True), (key, m)
    only_var = gp.eval_host(xs, None, want_acq=False, want_var=True)
    assert only_var[0] is None and only_var[1] is None and np.array_equal(only_var[2], var)
    with pytest.raises(Exception, match="host-buffer entry"):
        gp.eval_host(np.zeros((4097, 4)), "ei")
    with pytest.raises(Exception, match="BOPY_ACQ_NONE"):
        gp.eval_host(xs, None, want_acq=True)
