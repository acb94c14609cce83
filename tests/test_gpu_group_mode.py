"""Group mode of the fp64 throughput sweep (csrc/sweep_group_kernel.cuh): several thread blocks share one candidate tile so
that the solve workspace in flight fits the L2.  The arithmetic and its order are those of the one-tile-per-block kernel,
so every output must be BIT-identical to it, for every group size / slot count / interleave depth, on ragged candidate
counts, for every acquisition, with per-tile records and both NaN policies; and both must match the oracle."""
import numpy as np
import pytest

from conftest import golden_state
from oracle import gp_oracle as O
from parity_util import check_mean, check_var
from test_gpu_parity import native_for


def test_group_schedule_is_a_valid_order_cpu():
    """Host-only: every (tile, row) exactly once, after (tile, row - 1) and after (tile - 2, last row) -- the two
    dependencies a job waits for (V of the previous row; the workspace slot the tile before the previous one used)."""
    from bopy_b200 import _native
    for R in (4, 5, 8, 16, 17, 64):
        for lead in sorted({0, 1, 2, R // 4, R // 2}):
            tiles = 7
            pos = {}
            for j in range(R * tiles):
                k, i = _native.group_schedule(j, R, lead)
                assert 0 <= i < R and k >= 0 and (k, i) not in pos
                pos[(k, i)] = j
            for k in range(tiles - 1):               # the first tiles are complete within the enumerated prefix
                for i in range(R):
                    assert (k, i) in pos
                    if i > 0:
                        assert pos[(k, i - 1)] < pos[(k, i)]
                if k >= 2:
                    assert pos[(k - 2, R - 1)] < pos[(k, 0)]


def _outputs(gp, xs, acq, eta, kappa=2.0, index_base=0):
    out = gp.sweep(xs, acq=acq, eta=eta, kappa=kappa, want_mean=True, want_var=True, want_acq=True, want_min=True,
                   index_base=index_base)
    res = {k: out[k].cpu().numpy().copy() for k in ("mean", "var", "acq")}
    res["min_idx"], res["min_val"] = int(out["min_idx"].item()), float(out["min_val"].item())
    return res


def _same(a, b):
    return all(np.array_equal(a[k], b[k], equal_nan=True) for k in ("mean", "var", "acq")) and a["min_idx"] == b["min_idx"] \
        and (a["min_val"] == b["min_val"] or (np.isnan(a["min_val"]) and np.isnan(b["min_val"])))


SETTINGS = [(-1, -1, -1), (2, 2, 0), (3, 3, 2), (8, 3, -1), (16, 2, 1), (37, 4, 3)]


@pytest.mark.gpu
def test_default_is_group_mode_where_the_workspace_exceeds_l2():
    g, st = golden_state("c4_hartmann6_n2048")
    gp = native_for(st, "f64")
    G, S, lead = gp.set_group_mode()
    assert G >= 2 and S == 3 and lead == 8 and -(-148 // G) * S * 2048 * 128 * 8 <= 96 << 20
    assert gp.set_group_mode(0) == (0, 3, 8)
    gp.close()
    g, st = golden_state("c3_branin_n256")        # two block rows: the one-tile-per-block kernel
    gp = native_for(st, "f64")
    assert gp.set_group_mode()[0] == 0 and gp.set_group_mode(8)[0] == 0
    gp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("acq", ["ei", "lcb", "poi"])
def test_group_mode_is_bit_identical_to_one_tile_per_block_c4(acq):
    g, st = golden_state("c4_hartmann6_n2048")
    gp = native_for(st, "f64")
    gp.set_latency_path(0)
    eta = float(g["eta"])
    m = 148 * 128 + 77                      # more tiles than thread blocks and a ragged last tile
    xs = gp.candidates(O.candidates_uniform(99, 0, m, np.zeros(6), np.ones(6)))
    gp.set_group_mode(0)
    ref = _outputs(gp, xs, acq, eta, index_base=1000)
    small = {mm: _outputs(gp, xs[:mm], acq, eta) for mm in (5, 128, 300)}
    # against the golden vectors of the unmodified reference (the fixture's own candidates)
    xg = gp.candidates(g["Xs"])
    for setting in SETTINGS:
        eff = gp.set_group_mode(*setting)
        assert eff[0] >= 2
        assert _same(_outputs(gp, xs, acq, eta, index_base=1000), ref), setting
        assert _same(_outputs(gp, xs, acq, eta, index_base=1000), ref), (setting, "second launch on the same control block")
        for mm, r in small.items():
            assert _same(_outputs(gp, xs[:mm], acq, eta), r), (setting, mm)
        got = _outputs(gp, xg, "ei", eta)
        err, bound = check_mean(got["mean"], g["mean"], st, "f64")
        assert (err <= bound).all()
        err, bound = check_var(got["var"], g["var"], st, "f64")
        assert (err <= bound).all()
        assert got["min_idx"] == int(g["argmin_ei"])
    gp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,nu,n,d", [("matern", 2.5, 700, 3), ("matern", 0.5, 1024, 2), ("rbf", 0.0, 1500, 20)])
def test_group_mode_other_kernels_and_ragged_n(kind, nu, n, d):
    """Ragged n (padding rows inside the last block row), Matern kernels, d = 20 (the X/l block row aliases the residual
    tile and the candidates are not staged): identical to the one-tile-per-block kernel and equal to the oracle."""
    rng = np.random.default_rng(n)
    X = rng.random((n, d))
    y = np.sin(3.0 * X.sum(1)) + 0.1 * rng.standard_normal(n)
    spec = O.KernelSpec(kind=kind, nu=nu, length_scale=np.full(d, 0.4 if d < 10 else 1.0), amplitude=1.3, noise_level=0.0)
    st = O.fit_state(X, y, spec, 1e-6, True)
    gp = native_for(st, "f64")
    gp.set_latency_path(0)
    m = 2 * 148 * 128 // 3 + 11
    xs_host = O.candidates_uniform(5, 0, m, np.zeros(d), np.ones(d))
    xs = gp.candidates(xs_host)
    eta = float(np.min(y))
    gp.set_group_mode(0)
    ref = _outputs(gp, xs, "ei", eta)
    vals_ref, idx_ref = (t.cpu().numpy().copy() for t in gp.segment_argmin(xs, 512, "ei", eta=eta))
    o_mean, o_var, o_acq, _ = O.acquisition_sweep(st, "ei", xs_host[:512], eta=eta)
    err, bound = check_mean(ref["mean"][:512], o_mean, st, "f64")
    assert (err <= bound).all()
    err, bound = check_var(ref["var"][:512], o_var, st, "f64")
    assert (err <= bound).all()
    for setting in [(2, 2, 1), (3, 3, -1), (5, 2, 0), (-1, -1, -1)]:
        eff = gp.set_group_mode(*setting)
        if eff[0] < 2:
            eff = gp.set_group_mode(4, setting[1], setting[2])
        assert eff[0] >= 2
        assert _same(_outputs(gp, xs, "ei", eta), ref), setting
        vals, idx = (t.cpu().numpy() for t in gp.segment_argmin(xs, 512, "ei", eta=eta))      # per-tile records
        assert np.array_equal(vals, vals_ref, equal_nan=True) and np.array_equal(idx, idx_ref), setting
    gp.close()


@pytest.mark.gpu
def test_group_mode_nan_policies():
    """Candidates on training points of a barely regularised fit give NaN acquisition values; 'first' / 'skip' pick the same
    winner in both kernels (alpha = 0 on a short length scale: the variance at a training point is rounding noise around 0)."""
    rng = np.random.default_rng(11)
    n, d = 600, 2
    X = rng.random((n, d))
    y = np.cos(5 * X[:, 0]) * X[:, 1]
    spec = O.KernelSpec(kind="rbf", nu=0.0, length_scale=np.full(d, 0.02), amplitude=1.0, noise_level=0.0)
    st = O.fit_state(X, y, spec, 0.0, False)
    gp = native_for(st, "f64")
    gp.set_latency_path(0)
    xs_host = np.concatenate([rng.random((700, d)), X[:40], rng.random((900, d))])
    xs = gp.candidates(xs_host)
    res = {}
    for G in (0, 2):
        gp.set_group_mode(G, 2, 1)
        for policy in ("first", "skip"):
            gp.set_nan_policy(policy)
            res[(G, policy)] = _outputs(gp, xs, "ei", float(np.min(y)))
    gp.set_nan_policy("first")
    for policy in ("first", "skip"):
        assert _same(res[(0, policy)], res[(2, policy)]), policy
    a = res[(2, "first")]["acq"]
    assert np.isnan(a).any()
    if True:
        assert res[(2, "first")]["min_idx"] == int(np.argmin(a)) and res[(2, "skip")]["min_idx"] == int(np.nanargmin(a))
    gp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c3_branin_n256", "matern25_d2", "ard_amp_white", "edge_on_training_points"])
def test_warp_autonomous_kernel_parity(name, monkeypatch):
    """The opt-in warp-autonomous kernel for n <= 256 (csrc/sweep_warp_kernel.cuh, BOPY_B200_WARP_KERNEL=1): same results as
    the golden vectors of the unmodified reference and, to rounding, as the default kernel."""
    g, st = golden_state(name)
    if st.X_train.shape[0] > 256:
        pytest.skip("more than two block rows")
    monkeypatch.setenv("BOPY_B200_WARP_KERNEL", "1")
    monkeypatch.setenv("BOPY_B200_SMALL_N", "0")
    wk = native_for(st, "f64")
    monkeypatch.delenv("BOPY_B200_WARP_KERNEL")
    ref = native_for(st, "f64")
    for gp in (wk, ref):
        gp.set_latency_path(0)
    assert wk.launch_info(1 << 20)["grid"] == 296 and ref.launch_info(1 << 20)["grid"] == 148
    d = g["Xs"].shape[1]
    rng = np.random.default_rng(2)
    lo, hi = g["Xs"].min(0), g["Xs"].max(0)
    Xs = np.concatenate([g["Xs"], lo + rng.random((300 * 128 + 37, d)) * (hi - lo)])
    eta = float(g["eta"])
    a = _outputs(wk, wk.candidates(Xs), "ei", eta, index_base=3)
    b = _outputs(ref, ref.candidates(Xs), "ei", eta, index_base=3)
    pv = (st.kernel.amplitude + st.kernel.noise_level) * st.y_std ** 2
    assert np.max(np.abs(a["mean"] - b["mean"])) <= 1e-12 * (np.max(np.abs(b["mean"])) + st.y_std)
    assert np.max(np.abs(a["var"] - b["var"])) <= 1e-12 * pv
    assert a["min_idx"] - 3 == int(np.argmin(a["acq"]))
    m = len(g["Xs"])
    err, bound = check_mean(a["mean"][:m], g["mean"], st, "f64")
    assert (err <= bound).all()
    err, bound = check_var(a["var"][:m], g["var"], st, "f64")
    assert (err <= bound).all()
    vals, idx = (t.cpu().numpy() for t in wk.segment_argmin(wk.candidates(Xs[:1024]), 256, "lcb", kappa=2.0))
    own = _outputs(wk, wk.candidates(Xs[:1024]), "lcb", eta)["acq"]
    assert np.array_equal(idx, np.array([s * 256 + int(np.argmin(own[s * 256:(s + 1) * 256])) for s in range(4)]))
    wk.close()
    ref.close()
