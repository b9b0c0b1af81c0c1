"""CPU tests that pin the oracle (oracle/) -- the checker every GPU parity test relies on.

The reference holds no golden vector for this path (PARITY UNPINNED), so the
oracle is pinned by: an independent NumPy restatement (tests/np_restatement.py), NumPy's FFT, SciPy's
QUADPACK, the invariants of the scheme and a binary128 build of itself.
"""
import math

import numpy as np
import pytest

from np_restatement import NpModel
from oracle import oracle as orc
from oracle.oracle import Oracle, OracleConfig


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def smooth_field(rng, shape, amp, kmax=3):
    """random smooth field: a few low zonal / meridional harmonics"""
    nlat, nlon = shape
    lam = 2 * np.pi * np.arange(nlon) / nlon
    phi = np.pi * (np.arange(nlat) + 0.5) / nlat
    f = np.zeros(shape)
    for k in range(kmax + 1):
        for l in range(1, kmax + 1):
            a, b = rng.standard_normal(2)
            f += np.sin(l * phi)[:, None] * (a * np.cos(k * lam) + b * np.sin(k * lam))[None, :]
    return amp * f / np.abs(f).max()


def generic_state(nlon, nlat, seed=0):
    """A smooth, symmetry-free state: the filter's s1 = sum(d*w) is then well conditioned."""
    rng = np.random.default_rng(seed)
    u = smooth_field(rng, (nlat, nlon), 30.0)
    u[0] = u[-1] = 0.0
    v = smooth_field(rng, (nlat - 1, nlon), 20.0)
    gd = 5.0e4 + smooth_field(rng, (nlat, nlon), 5.0e3)
    ghs = 2.0e3 + smooth_field(rng, (nlat, nlon), 2.0e3)
    for f in (gd, ghs):  # a pole is one point: single-valued scalars there
        f[0] = f[0].mean()
        f[-1] = f[-1].mean()
    return u, v, gd, ghs


# ------------------------------------------------------------------------------------------- FFTPACK
@pytest.mark.parametrize("n", [12, 30, 36, 64, 100, 180, 360, 1440, 3600, 7200])
def test_rfft_matches_numpy(n):
    x = np.random.default_rng(n).standard_normal(n)
    X = orc.rfft_forward(x)
    ref = np.fft.rfft(x) / n
    exp = np.zeros(n)
    exp[0] = ref[0].real
    for k in range(1, (n + 1) // 2):
        exp[2 * k - 1] = 2 * ref[k].real      # a_k, rfftf1.f:87-107
        exp[2 * k] = -2 * ref[k].imag         # b_k
    if n % 2 == 0:
        exp[n - 1] = ref[n // 2].real
    assert np.abs(X - exp).max() < 5e-16 * max(1.0, math.log2(n))
    assert np.abs(orc.rfft_backward(X) - x).max() < 1e-14


def test_rfft_factor_order():
    # rffti1.f:38-58: 4 first, a factor 2 moved to the front
    assert orc.rfft_factors(360) == [2, 4, 3, 3, 5]
    assert orc.rfft_factors(180) == [4, 3, 3, 5]
    assert orc.rfft_factors(3600) == [4, 4, 3, 3, 5, 5]
    with pytest.raises(orc.OracleError):
        orc.rfft_factors(14)


# ------------------------------------------------------------------------------------------- filter
def test_filter_row_map_and_projection():
    cfg = OracleConfig(num_lon=72, num_lat=37, time_step_size=600, zonal_tend_filter_cutoff_wavenumber=[5, 4, 3])
    o = Oracle(cfg)
    ff, fc, hf, hc = o.filter_rows()
    nlat = 37
    # full rows: south 1+k, north nlat-k (1-based) -> 0-based k and nlat-1-k
    assert list(np.nonzero(ff)[0]) == [1, 2, 3, nlat - 4, nlat - 3, nlat - 2]
    assert [fc[j] for j in (1, 2, 3, nlat - 2, nlat - 3, nlat - 4)] == [5, 4, 3, 5, 4, 3]
    # half rows: south k (mask c_k); north flag nlat-k+1 (k=1 out of bounds) but mask nlat-k  (B2)
    assert list(np.nonzero(hf)[0]) == [0, 1, 2, nlat - 3, nlat - 2]
    assert [hc[h] for h in (0, 1, 2)] == [5, 4, 3]
    assert hc[nlat - 2] == 5 and hc[nlat - 3] == 4 and hc[nlat - 4] == 3 and hf[nlat - 4] == 0
    x = np.random.default_rng(1).standard_normal(72)
    y = o.filter_row(False, 1, x)
    # projection onto wavenumbers 0..c plus Re(c+1)  (B3), idempotent
    X = np.fft.rfft(x)
    Y = np.zeros_like(X)
    Y[:6] = X[:6]
    Y[6] = X[6].real
    assert np.abs(y - np.fft.irfft(Y, 72)).max() < 2e-15
    assert np.abs(o.filter_row(False, 1, y) - y).max() < 2e-15


# ------------------------------------------------------------------------------------------- operators
@pytest.mark.parametrize("adv", ["center_diff", "upwind", "weno"])
@pytest.mark.parametrize("pass_", ["all", "fast", "slow"])
def test_space_operators_match_numpy(adv, pass_):
    nlon, nlat = 72, 37
    u, v, gd, ghs = generic_state(nlon, nlat)
    cfg = OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=600, uv_adv_scheme=adv,
                       uv_adv_upwind_lon_beta=0.3, uv_adv_upwind_lat_beta=0.1,
                       zonal_tend_filter_cutoff_wavenumber=[4, 3, 2])
    o = Oracle(cfg)
    o.set_state(u, v, gd, ghs)
    o.run_init()
    npm = NpModel(nlon, nlat, 600, adv=adv, beta_lon=0.3, beta_lat=0.1, cutoff=[4, 3, 2])
    npm.ghs = ghs
    a = o.space_operators(pass_)
    b = npm.tend(npm.make_state(u, v, gd), pass_)
    for x, y in zip(a, b):
        assert rel(x, y) < 1e-13 or np.abs(y).max() == 0.0


def test_update_state_matches_numpy():
    nlon, nlat = 60, 31
    u, v, gd, ghs = generic_state(nlon, nlat, seed=3)
    o = Oracle(OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=600, zonal_tend_filter_cutoff_wavenumber=[4, 4]))
    o.set_state(u, v, gd, ghs)
    o.run_init()
    npm = NpModel(nlon, nlat, 600, cutoff=[4, 4])
    npm.ghs = ghs
    st = npm.make_state(u, v, gd)
    td = npm.tend(st, "all")
    o.space_operators("all")
    got = o.update_state_preview(123.0)
    exp = npm.update(123.0, td, st)
    for x, y in zip(got, exp):
        assert rel(x, y) < 1e-14


@pytest.mark.parametrize("case,adv,diff", [("mountain_zonal_flow", "upwind", False),
                                           ("jet_zonal_flow", "center_diff", True),
                                           ("steady_geostrophic_flow", "center_diff", False)])
def test_step_matches_numpy(case, adv, diff):
    nlon, nlat = 120, 61
    cfg = OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=600, subcycles=6, uv_adv_scheme=adv,
                       uv_adv_upwind_lat_beta=0.1, zonal_tend_filter_cutoff_wavenumber=[4] * 5,
                       use_diffusion=diff, diffusion_coef=1.0e5)
    o = Oracle(cfg)
    o.set_initial_condition(case)
    u, v, gd = o.state()
    ghs = o.ghs()
    o.run_init()
    npm = NpModel(nlon, nlat, 600, 6, adv=adv, beta_lat=0.1, cutoff=[4] * 5, use_diffusion=diff, diffusion_coef=1.0e5)
    npm.ghs = ghs
    st = npm.make_state(u, v, gd)
    for _ in range(3):
        o.step(1)
        st = npm.step(st)
    ou, ov, ogd = o.state()
    assert rel(ou, st[0]) < 1e-12 and rel(ogd, st[2]) < 1e-12
    assert np.abs(ov - st[1]).max() < 1e-11 * max(1.0, np.abs(ou).max())
    assert abs(o.diag()[2] - npm.beta) < 1e-12


@pytest.mark.parametrize("order,split", [(3, "csp2"), (4, "csp2"), (3, "none"), (4, "none")])
def test_runge_kutta_step_matches_numpy(order, split):
    """the specified runge_kutta integrator: oracle against the NumPy second reading, and the
    property it is built for -- the quadratic invariant is kept to rounding on a flat-bottom case"""
    nlon, nlat = 72, 37
    dt = 300.0 if split == "csp2" else 60.0
    cfg = OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=dt, subcycles=4, split_scheme=split,
                       zonal_tend_filter_cutoff_wavenumber=[4] * 3, time_scheme="runge_kutta", time_order=order)
    o = Oracle(cfg)
    o.set_initial_condition("steady_geostrophic_flow")
    u, v, gd = o.state()
    rng = np.random.default_rng(5)
    gd = gd * (1 + 1e-3 * np.cos(3 * np.linspace(0, 2 * np.pi, nlon, endpoint=False))[None, :] * np.cos(np.linspace(-1.5, 1.5, nlat))[:, None] ** 4)
    o.set_state(u, v, gd, o.ghs())
    ghs = o.ghs()
    o.run_init()
    e0 = o.diag()[1]
    npm = NpModel(nlon, nlat, dt, 4, split=split, cutoff=[4] * 3, time_scheme="runge_kutta", time_order=order)
    npm.ghs = ghs
    st = npm.make_state(u, v, gd)
    for _ in range(3):
        o.step(1)
        st = npm.step(st)
    ou, ov, ogd = o.state()
    assert rel(ou, st[0]) < 1e-12 and rel(ogd, st[2]) < 1e-12
    assert np.abs(ov - st[1]).max() < 1e-11 * max(1.0, np.abs(ou).max())
    assert abs(o.diag()[2] - npm.beta) < 1e-12
    assert abs(npm.beta - npm.beta_direct) < 1e-6   # the tendency-product form of beta is -2 <K, phi> / (dt <K, K>)
    assert abs(o.diag()[2] - 1.0) < 1e-3            # beta = 1 + O(dt^2)
    assert abs(o.diag()[1] / e0 - 1) < 5e-14        # energy kept by the beta fix


@pytest.mark.parametrize("smooth,pass_", [(True, "fast"), (False, "fast"), (True, "slow"), (False, "all")])
def test_moving_reduced_tendency_matches_numpy(smooth, pass_):
    """the specified moving reduced tendency: on a row with factor r the tendency becomes the
    average, over the r offsets of the reduced grid, of the reduced-cell means handed back to the fine cells -- written
    here literally (box means for every offset), against the oracle's triangular-kernel form; then the s1 / s2 rescale"""
    nlon, nlat = 72, 37
    factors = [8, 4, 2]
    u, v, gd, ghs = generic_state(nlon, nlat, seed=11)
    base = dict(num_lon=nlon, num_lat=nlat, time_step_size=600)
    o = Oracle(OracleConfig(use_zonal_reduce=True, reduce_adv_lon=True, use_reduce_tend_smooth=smooth,
                            zonal_reduce_factors=factors, **base))
    o0 = Oracle(OracleConfig(**base))      # no filter rows, no reduced rows: the raw tendencies
    for q in (o, o0):
        q.set_state(u, v, gd, ghs)
        q.run_init()
    got, raw = o.space_operators(pass_), o0.space_operators(pass_)
    U, V, _ = o.iap_state()
    wts = (U, V, gd + ghs)

    def moving(row, r):
        out = np.zeros_like(row)
        for off in range(r):
            x = np.roll(row, -off)
            x = np.repeat(x.reshape(-1, r).mean(axis=1), r)
            out += np.roll(x, off)
        return out / r

    for f, (g, t, w) in enumerate(zip(got, raw, wts)):
        exp = t.copy()
        nrows = t.shape[0]
        for k, r in enumerate(factors, start=1):
            rows = (k - 1, nrows - k) if f == 1 else (k, nrows - 1 - k)      # half rows / full rows next to the pole row
            for j in rows:
                if np.abs(t[j]).max() == 0.0:
                    continue
                s1 = np.sum(t[j] * w[j])
                if smooth and not abs(s1) > 1e-16:
                    continue
                y = moving(t[j], r)
                exp[j] = y * s1 / np.sum(y * w[j]) if smooth else y
                assert abs(np.sum(y) - np.sum(t[j])) <= 1e-12 * np.abs(t[j]).sum()   # the reduction conserves the row sum
        assert np.abs(g - exp).max() <= 1e-12 * max(np.abs(exp).max(), 1e-300)
    # without reduce_adv_lon the slow pass is left alone
    o2 = Oracle(OracleConfig(use_zonal_reduce=True, reduce_adv_lon=False, zonal_reduce_factors=factors, **base))
    o2.set_state(u, v, gd, ghs)
    o2.run_init()
    for a, b in zip(o2.space_operators("slow"), o0.space_operators("slow")):
        assert np.array_equal(a, b)


def test_reduced_rows_reject_overlap_with_filter_rows_and_bad_factors():
    with pytest.raises(Exception):
        Oracle(OracleConfig(num_lon=72, num_lat=37, time_step_size=600, use_zonal_reduce=True, zonal_reduce_factors=[5]))
    with pytest.raises(Exception):
        Oracle(OracleConfig(num_lon=72, num_lat=37, time_step_size=600, use_zonal_reduce=True, zonal_reduce_factors=[4],
                            zonal_tend_filter_cutoff_wavenumber=[4]))


def test_weno_step_matches_numpy():
    """WENO advection (src/weno_mod.F90:69-300) through whole csp2 steps, second reading in tests/np_restatement.py"""
    nlon, nlat = 96, 49
    cfg = OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=600, subcycles=4, uv_adv_scheme="weno",
                       zonal_tend_filter_cutoff_wavenumber=[4, 4])
    o = Oracle(cfg)
    o.set_initial_condition("mountain_zonal_flow")
    u, v, gd = o.state()
    ghs = o.ghs()
    o.run_init()
    npm = NpModel(nlon, nlat, 600, 4, adv="weno", cutoff=[4, 4])
    npm.ghs = ghs
    st = npm.make_state(u, v, gd)
    for _ in range(3):
        o.step(1)
        st = npm.step(st)
    ou, ov, ogd = o.state()
    assert rel(ou, st[0]) < 1e-12 and rel(ogd, st[2]) < 1e-12
    assert np.abs(ov - st[1]).max() < 1e-11 * max(1.0, np.abs(ou).max())
    assert abs(o.diag()[2] - npm.beta) < 1e-12


@pytest.mark.parametrize("case", ["steady_geostrophic_flow", "mountain_zonal_flow"])
def test_isp_step_matches_numpy(case):
    """isp_splitting (src/dycore_mod.F90:689-752) incl. its beta * 4 / dt, second reading in tests/np_restatement.py"""
    nlon, nlat = 96, 49
    cfg = OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=600, subcycles=4, split_scheme="isp",
                       zonal_tend_filter_cutoff_wavenumber=[3, 3])
    o = Oracle(cfg)
    o.set_initial_condition(case)
    u, v, gd = o.state()
    ghs = o.ghs()
    o.run_init()
    npm = NpModel(nlon, nlat, 600, 4, split="isp", cutoff=[3, 3])
    npm.ghs = ghs
    st = npm.make_state(u, v, gd)
    for _ in range(2):
        o.step(1)
        st = npm.step(st)
    ou, ov, ogd = o.state()
    assert rel(ou, st[0]) < 1e-12 and rel(ogd, st[2]) < 1e-12
    assert np.abs(ov - st[1]).max() < 1e-11 * max(1.0, np.abs(ou).max())
    assert abs(o.diag()[2] / npm.beta - 1) < 1e-10


@pytest.mark.parametrize("order,coef", [(2, 1.0e5), (4, 1.0e15)])
def test_diffusion_matches_numpy(order, coef):
    """ordinary_diffusion alone (src/diffusion_mod.F90:74-217), order 2 and order 4 (two Laplacian passes, sign -1)"""
    nlon, nlat = 72, 37
    u, v, gd, ghs = generic_state(nlon, nlat, seed=7)
    cfg = OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=600, use_diffusion=True, diffusion_order=order,
                       diffusion_coef=coef, zonal_tend_filter_cutoff_wavenumber=[4, 3])
    o = Oracle(cfg)
    o.set_state(u, v, gd, ghs)
    o.run_init()
    npm = NpModel(nlon, nlat, 600, cutoff=[4, 3], use_diffusion=True, diffusion_order=order, diffusion_coef=coef)
    npm.ghs = ghs
    st = npm.diffusion(600.0, npm.make_state(u, v, gd))
    o.ordinary_diffusion(600.0)
    got = o.state()
    for x, y, x0 in zip(got, st[:3], (u, v, gd)):
        assert rel(x - x0, y - x0) < 1e-11, order     # the increment itself, not the state it is added to
        assert rel(x, y) < 1e-14


def test_unsplit_step_matches_numpy():
    nlon, nlat = 72, 37
    u, v, gd, ghs = generic_state(nlon, nlat, seed=5)
    cfg = OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=120, split_scheme="none",
                       zonal_tend_filter_cutoff_wavenumber=[4, 4, 4])
    o = Oracle(cfg)
    o.set_state(u, v, gd, ghs)
    o.run_init()
    npm = NpModel(nlon, nlat, 120, split="none", cutoff=[4, 4, 4])
    npm.ghs = ghs
    st = npm.make_state(u, v, gd)
    o.step(2)
    st = npm.step(npm.step(st))
    for x, y in zip(o.state(), st[:3]):
        assert rel(x, y) < 1e-12


# ------------------------------------------------------------------------------------------- invariants
def test_antisymmetry_sums_vanish():
    """check_antisymmetry, dycore_mod.F90:794-851: the four sums are zero up to rounding."""
    o = Oracle(OracleConfig(num_lon=120, num_lat=61, time_step_size=600, use_zonal_tend_filter=False))
    u, v, gd, ghs = generic_state(120, 61, seed=7)
    o.set_state(u, v, gd, ghs)
    o.run_init()
    o.space_operators("all")
    sums, scale = o.check_antisymmetry()
    assert np.all(np.abs(sums) < 1e-13 * scale)


def test_energy_conservation_and_beta():
    """centred scheme + qcon_modified conserves total energy to rounding."""
    cfg = OracleConfig(num_lon=120, num_lat=61, time_step_size=600, subcycles=6,
                       zonal_tend_filter_cutoff_wavenumber=[4] * 5)
    o = Oracle(cfg)
    o.set_initial_condition("rossby_haurwitz_wave")
    o.run_init()
    _, e0, _ = o.diag()
    o.step(24)
    m1, e1, beta = o.diag()
    assert abs(e1 - e0) / e0 < 1e-13
    assert 0 < beta - 1.0 < 1e-3
    # the same run in binary128 conserves energy to 1e-25
    oq = Oracle(OracleConfig(num_lon=48, num_lat=25, time_step_size=900, subcycles=4,
                             zonal_tend_filter_cutoff_wavenumber=[4, 4]), kind="quad")
    assert oq.lib.orc_real_bytes() == 16
    oq.set_initial_condition("rossby_haurwitz_wave")
    oq.run_init()
    _, q0, _ = oq.diag()
    oq.step(3)
    assert oq.diag()[1] == q0


def test_steady_geostrophic_flow_is_stationary():
    cfg = OracleConfig(num_lon=120, num_lat=61, time_step_size=600, subcycles=6,
                       zonal_tend_filter_cutoff_wavenumber=[4] * 5)
    o = Oracle(cfg)
    o.set_initial_condition("steady_geostrophic_flow")
    u0, v0, gd0 = o.state()
    o.run_init()
    o.step(24)
    u, v, gd = o.state()
    assert np.abs(gd - gd0).max() / gd0.max() < 1e-3
    assert np.abs(v).max() < 5e-2 and np.abs(u - u0).max() < 0.2


def test_steady_geostrophic_error_converges_at_second_order():
    """A known answer that does not come from restating the code: Williamson et al. 1992 test case 2
    (steady_geostrophic_flow_test_mod.F90:14-16) is an exact steady solution of the continuous equations, so the
    departure from the initial condition after a fixed time is the truncation error of the whole step -- C-grid
    operators, IAP transform, polar filter, predictor-corrector, csp2 -- and must fall by 4 per halving of the mesh and
    the time step (second-order centred differences and a second-order integrator)."""
    errs = []
    for nlon, nlat, dt in [(36, 19, 1200.0), (72, 37, 600.0), (144, 73, 300.0)]:
        o = Oracle(OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=dt, subcycles=6, split_scheme="csp2",
                                zonal_tend_filter_cutoff_wavenumber=[4] * 5))
        o.set_initial_condition("steady_geostrophic_flow")
        u0, v0, gd0 = [x.copy() for x in o.state()]
        o.run_init()
        o.step(int(6 * 3600 / dt))
        u, v, gd = o.state()
        errs.append((np.sqrt(np.mean((gd - gd0) ** 2) / np.mean(gd0 ** 2)), np.abs(u - u0).max(), np.abs(v).max()))
    for coarse, fine in zip(errs, errs[1:]):
        for a, b in zip(coarse, fine):
            assert 3.7 < a / b < 4.3, errs
    assert errs[-1][0] < 1.2e-4


def test_balanced_jet_without_the_bump_stays_put():
    """The zonal jet of Galewsky et al. 2004 (jet_zonal_flow_test_mod.F90:16-66) is in geostrophic balance with the
    height field the IC plugin integrates with QUADPACK; without the height bump the continuous solution is steady.
    The oracle must hold it -- v stays near zero, u and gd near the initial state -- and the residual must shrink with
    the mesh (the jet is 1/7 of a radian wide: 72x37 barely resolves it).  Checks the plugin's quadrature against the
    discrete operators, which no restatement of either alone does."""
    res = []
    for nlon, nlat, dt in [(72, 37, 600.0), (144, 73, 300.0)]:
        o = Oracle(OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=dt, subcycles=6, split_scheme="csp2",
                                zonal_tend_filter_cutoff_wavenumber=[4] * 5))
        o.set_initial_condition("jet_zonal_flow")
        u0, v0, gd0 = [x.copy() for x in o.state()]
        bump = gd0 - gd0.min(axis=1, keepdims=True)
        # 120 m cos(lat) at 45N, centred on (lon, lat) = (90E, 45N); nothing of it on the far side of the globe
        assert abs(bump.max() / 9.80616 - 120.0 * math.cos(math.pi / 4)) < 0.5 and bump.max(axis=0).min() == 0.0
        j, i = np.unravel_index(bump.argmax(), bump.shape)
        assert (i, j) == (nlon // 2, 3 * (nlat - 1) // 4)
        gdz = np.repeat(gd0.min(axis=1, keepdims=True), nlon, axis=1)
        assert np.abs(u0).max() == pytest.approx(80.0, abs=0.5) and not v0.any()
        o.set_state(u0, v0, gdz, None)
        o.run_init()
        o.step(int(24 * 3600 / dt))                                   # one day
        u, v, gd = o.state()
        res.append((np.abs(u - u0).max(), np.abs(v).max(), np.abs(gd - gdz).max() / 9.80616))
        assert np.ptp(u - u.mean(axis=1, keepdims=True)) < 1e-9      # and stays zonally symmetric
    assert res[1][0] < 0.15 and res[1][1] < 0.4 and res[1][2] < 40.0, res       # m/s, m/s, m (of 80 m/s, 10 km)
    assert all(b < 0.7 * a for a, b in zip(res[0], res[1])), res


def test_rossby_haurwitz_wave_moves_at_the_published_phase_speed():
    """A known answer that does not come from restating the code: the wavenumber-4 Rossby-Haurwitz wave of the
    reference's test case (rossby_haurwitz_wave_test_mod.F90, Williamson et al. 1992 test case 6: R = 4, omega = K =
    7.848e-6 1/s) moves eastward without change of shape at nu = (R (3 + R) omega - 2 Omega) / ((1 + R) (2 + R)) in the
    nondivergent limit, 12.2 degrees per day; published shallow-water runs of the case sit within a few per cent of it.
    The oracle's full step (operators, polar filter, predictor-corrector with beta, csp2) must reproduce that: phase of
    zonal wavenumber 4 of the geopotential at mid latitudes after one model day."""
    cfg = OracleConfig(num_lon=180, num_lat=91, time_step_size=240.0, subcycles=6, split_scheme="csp2",
                       zonal_tend_filter_cutoff_wavenumber=[4] * 5)
    o = Oracle(cfg)
    o.set_initial_condition("rossby_haurwitz_wave")
    gd0 = o.state()[2].copy()
    o.run_init()
    nsteps = 360   # one day
    o.step(nsteps)
    gd1 = o.state()[2]
    R, om, Om = 4.0, 7.848e-6, 7.292e-5
    nu = (R * (3.0 + R) * om - 2.0 * Om) / ((1.0 + R) * (2.0 + R))
    expect = math.degrees(nu * nsteps * cfg.time_step_size)          # degrees of longitude per day
    rows = range(55, 71)                                             # 20N .. 50N, where the wave is strongest
    shift = []
    for j in rows:
        c0, c1 = np.fft.rfft(gd0[j])[4], np.fft.rfft(gd1[j])[4]
        dphi = -np.angle(c1 / c0)                                    # pattern f(lambda - s): phase of mode 4 falls by 4 s
        shift.append(math.degrees(dphi) / 4.0)
        assert abs(abs(c1) / abs(c0) - 1.0) < 0.05                   # the wave keeps its amplitude
    got = float(np.mean(shift))
    assert 11.0 < expect < 13.0
    assert abs(got / expect - 1.0) < 0.12, (got, expect)
    assert np.std(shift) < 0.5                                       # the same speed at every latitude: shape preserved


def test_shamir_paldor_rossby_wave_moves_at_its_analytic_phase_speed():
    """A known answer from the linearised equations: the (n, k) = (5, 10) Rossby wave of Shamir & Paldor (2016) on a
    5 km layer (shallow_water_waves_test_mod.F90) is an eigensolution of the linear shallow-water equations on the sphere
    that travels westward at the root C of its cubic dispersion relation (:130-171), 3.0 degrees of longitude per day,
    keeping its shape.  The initial amplitude is 4 m2/s2 on 5e4, so the oracle's full nonlinear step must move the
    pattern at that speed: projection of zonal wavenumber 10 of gd, u and v on the initial pattern after one day."""
    lib = orc.load("strict")
    C = lib.orc_swe_phase_speed(0)
    cfg = OracleConfig(num_lon=180, num_lat=91, time_step_size=300.0, subcycles=6, split_scheme="csp2",
                       zonal_tend_filter_cutoff_wavenumber=[12] * 5)     # the filter keeps wavenumber 10
    o = Oracle(cfg)
    o.set_initial_condition("shallow_water_waves")
    f0 = [x.copy() for x in o.state()]
    o.run_init()
    nsteps = 288
    o.step(nsteps)
    expect = math.degrees(C * nsteps * cfg.time_step_size)
    assert -3.1 < expect < -2.9
    k = 10
    for a0, a1, w in zip(f0, o.state(), (o.table(0), o.table(1), o.table(0))):      # u, v, gd
        c0, c1 = np.fft.rfft(a0, axis=1)[:, k], np.fft.rfft(a1, axis=1)[:, k]
        w = w[:len(c0)]
        z, nrm = np.sum(w * np.conj(c0) * c1), np.sum(w * np.abs(c0) ** 2)
        shift = math.degrees(-np.angle(z) / k)        # pattern f(lambda - s): the phase of mode k falls by k s
        assert abs(shift / expect - 1.0) < 0.06, (shift, expect)
        assert abs(z) / nrm > 0.97                     # the same pattern, not just the same wavenumber
        assert abs(math.sqrt(np.sum(w * np.abs(c1) ** 2) / nrm) - 1.0) < 0.03


def test_pole_rows():
    """u = U = 0 on the pole rows for ever; dgd on a pole row is one zonal constant (cap formula)."""
    cfg = OracleConfig(num_lon=72, num_lat=37, time_step_size=600, subcycles=4,
                       zonal_tend_filter_cutoff_wavenumber=[4, 4])
    o = Oracle(cfg)
    o.set_initial_condition("rossby_haurwitz_wave")
    o.run_init()
    du, dv, dgd = o.space_operators("all")
    assert np.all(du[0] == 0) and np.all(du[-1] == 0)
    assert np.ptp(dgd[0]) == 0 and np.ptp(dgd[-1]) == 0
    o.step(3)
    u, v, gd = o.state()
    iu, _, _ = o.iap_state()
    assert np.all(u[0] == 0) and np.all(u[-1] == 0) and np.all(iu[0] == 0) and np.all(iu[-1] == 0)


def test_reference_behaviour_for_nonzero_u_on_a_pole_row():
    """What gmd_set_state refuses (include/gmd.h: u on the two pole rows must be 0;
    tests/test_gpu_parity.py::test_pole_rows_and_reference_layout checks the refusal), shown on the reference's side: the
    reference would take such a state.  iap_transform makes U(pole) from it, and because the tendency loops never write
    du on a pole row (rows 2..nlat-1, src/dycore_mod.F90:197-364) update_state carries that U(pole) unchanged for ever
    (:626-630) while u(pole) of every later state is never written (the no_pole loop :639-643) and reads 0: an
    inconsistent pair whose frozen U keeps entering the four-point Coriolis / advection averages of the rows next to the
    pole (:477-505).  No IC plugin of the reference produces it (the Shamir-Paldor amplitudes give ~1e-146 there, taken
    as 0), so the product treats it as a caller error instead of carrying a frozen pole wind."""
    kw = dict(num_lon=72, num_lat=37, time_step_size=600.0, subcycles=4, split_scheme="csp2",
              zonal_tend_filter_cutoff_wavenumber=[4, 4, 4])
    o = Oracle(OracleConfig(**kw))
    o.set_initial_condition("mountain_zonal_flow")
    u, v, gd = o.state()
    ghs = o.ghs()
    assert not u[0].any() and not u[-1].any()
    o.run_init()
    o.step(3)
    rng = np.random.default_rng(0)
    u2 = u.copy()
    u2[0], u2[-1] = 10.0 * rng.standard_normal(72), 10.0 * rng.standard_normal(72)
    o2 = Oracle(OracleConfig(**kw))
    o2.set_state(u2, v, gd, ghs)                  # accepted
    o2.run_init()
    U0 = o2.iap_state()[0].copy()
    assert np.abs(U0[0]).max() > 1e3
    o2.step(3)
    a, b, U = o.state(), o2.state(), o2.iap_state()[0]
    assert not b[0][0].any() and not b[0][-1].any()                            # u(pole) reads 0 from the first step on ...
    assert np.array_equal(U[0], U0[0]) and np.array_equal(U[-1], U0[-1])       # ... U(pole) is frozen at what it was given
    assert np.abs(b[1] - a[1]).max() > 1e-3 and np.isfinite(b[2]).all()        # and keeps moving the rows next to the poles


def test_reference_behaviour_when_the_filter_rescaling_denominator_is_zero():
    """The other documented deviation, shown on the reference's side: the
    SMOOTHING block rescales a filtered tendency row by s1 / s2 (src/dycore_mod.F90:212-219) with s2 the inner product
    AFTER the filter.  A row whose tendency is a pure two-grid wave (here: a +-8 m2/s2 zigzag in gd and a +-0.5 m/s zigzag
    in u on the second row from the south pole of a fluid at rest: pressure-gradient and mass-divergence tendencies
    that alternate exactly) is filtered to exact zeros, s2 = 0 while s1 is not, and the reference divides: NaN rows, then
    the fatal 'Total mass is NaN!' of diag_run (src/diag_mod.F90:79-87).  The product leaves such a row unscaled when
    its s2 is exactly 0 instead of aborting."""
    nlon, nlat = 64, 33
    o = Oracle(OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=300.0, split_scheme="none",
                            zonal_tend_filter_cutoff_wavenumber=[4, 4, 4]))
    assert o.filter_rows()[0][2] == 1
    u, v, gd = np.zeros((nlat, nlon)), np.zeros((nlat - 1, nlon)), np.full((nlat, nlon), 5.0e4)
    zigzag = np.where(np.arange(nlon) % 2 == 0, 1.0, -1.0)
    gd[2] += 8.0 * zigzag
    u[2] = 0.5 * zigzag
    o.set_state(u, v, gd, None)
    o.run_init()
    du, dv, dgd = o.space_operators("all")
    assert np.isnan(du[2]).all() and np.isnan(dgd[2]).all()
    assert np.isfinite(np.delete(du, 2, axis=0)).all() and np.isfinite(np.delete(dgd, 2, axis=0)).all() and np.isfinite(dv).all()
    with pytest.raises(orc.OracleError, match="Total mass is NaN"):
        o.step(1)


def test_fast_math_build_is_at_the_noise_floor():
    """The reference is built with -Ofast; our strict and -ffast-math builds of the same source bracket
    what any build of it can reproduce.  One model step must agree to ~1e-11 on a well conditioned case."""
    cfg = OracleConfig(num_lon=120, num_lat=61, time_step_size=600, subcycles=6, uv_adv_scheme="upwind",
                       uv_adv_upwind_lat_beta=0.1, zonal_tend_filter_cutoff_wavenumber=[4] * 3)
    a, b = Oracle(cfg, "strict"), Oracle(cfg, "fast")
    for o in (a, b):
        o.set_initial_condition("mountain_zonal_flow")
        o.run_init()
        o.step(2)
    for x, y in zip(a.state(), b.state()):
        assert np.abs(x - y).max() < 1e-10 * max(1.0, np.abs(y).max())


# ------------------------------------------------------------------------------------------- ICs / golden
def test_jet_profile_matches_quadpack_golden(golden_dir):
    d = np.load(golden_dir / "jet_gd_profile_181.npz")
    got = np.array([orc.jet_gd_profile(float(x)) for x in d["lat"]])
    assert np.abs(got - d["gd"]).max() / d["gd"].max() < 1e-14


def test_jet_profile_matches_scipy_live():
    scipy_integrate = pytest.importorskip("scipy.integrate")
    pi = 4 * np.arctan(1.0)
    omega, radius, g = 2 * pi / 86400.0, 6.37122e6, 9.80616
    lat0 = pi / 7
    lat1 = pi / 2 - lat0
    en = np.exp(-4 / (lat1 - lat0) ** 2)

    def integrand(lat):
        u = 0.0 if (lat <= lat0 or lat >= lat1) else 80.0 / en * np.exp(1 / (lat - lat0) / (lat - lat1))
        return radius * u * (2 * omega * np.sin(lat) + np.tan(lat) / radius * u)

    for lat in (0.3, 0.6, 0.8, 1.0, 1.3, 1.5):
        r, _ = scipy_integrate.quad(integrand, -0.5 * pi, lat, epsabs=1e-10, epsrel=1e-3, limit=500)
        assert abs(orc.jet_gd_profile(lat) - (g * 1e4 - r)) < 1e-9


def test_shallow_water_waves_ic_second_reading():
    """shallow_water_waves_test_mod.F90 read a second time, in NumPy: the phase speed is the smallest-magnitude root
    of the cubic dispersion relation (:130-171, solved there by Cardano's formula, here by numpy.roots), the
    amplitudes :176-271 vectorised, v on half row j with the amplitude of FULL row j (:307-313)."""
    omega, g, a, H0, n, k = 7.29212e-5, 9.80616, 6371220.0, 5.0e3, 5, 10
    sigma = 0.5 + np.sqrt(0.25 + k * k)
    En = g * H0 / a ** 2 * (n + sigma) ** 2
    # Cj = -(D + Delta0 / D) / (3 k^2) are the roots of  k^2 C^3 - En C - (2 omega g H0 / a^2) = 0
    roots = np.roots([k * k, 0.0, -En, -2.0 * omega * g * H0 / a ** 2])
    assert np.abs(roots.imag).max() < 1e-20
    C = -np.abs(roots.real).min()
    lib = orc.load("strict")
    lib.orc_swe_phase_speed.restype = __import__("ctypes").c_double
    for flag, want in ((0, C), (1, roots.real.max()), (-1, roots.real.min())):
        assert abs(lib.orc_swe_phase_speed(flag) / want - 1) < 1e-12
    nlon, nlat = 48, 25
    o = Oracle(OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=100.0))
    o.set_initial_condition("shallow_water_waves")
    u, v, gd = o.state()
    pi = 4 * np.arctan(1.0)
    lat = -0.5 * pi + np.arange(nlat) * (pi / (nlat - 1))
    lat[-1] = 0.5 * pi
    lon = np.arange(nlon) * (2 * pi / nlon)
    sl, cl, tl = np.sin(lat), np.cos(lat), np.tan(lat)
    a3 = sigma * (sigma + 1) * (sigma + 2)
    a4 = a3 * (sigma + 3)
    a5 = a4 * (sigma + 4)
    C5 = (4 * a5 * sl ** 4 - 20 * a4 * sl ** 2 + 15 * a3) * sl / 15
    C5p = (4 * a5 * sl ** 4 - 12 * a4 * sl ** 2 + 3 * a3) * cl / 3
    psi = 1e-8 * cl ** sigma * C5
    dpsi = 1e-8 * cl ** sigma * (-sigma * tl * C5 + C5p)
    Kp = (g * H0 + a ** 2 * C ** 2 * cl ** 2) / (C * cl)
    Km = (g * H0 - a ** 2 * C ** 2 * cl ** 2) / (C * cl)
    vt = np.sqrt(2 * omega * np.abs(Km) / cl ** 2) * psi
    ht = np.sqrt(2 * omega * np.abs(Km) * a ** 2 * H0 ** 2) / Km * (dpsi + tl * (0.5 * Kp / Km - 2 * omega / C) * psi)
    ut = (2 * omega * sl / C) * vt + (g / a / cl / C) * ht
    want_u = ut[:, None] * np.cos(k * (lon + 0.5 * (2 * pi / nlon)))[None, :]
    want_v = (k * vt)[:-1, None] * np.cos(k * lon - 0.5 * pi)[None, :]
    want_gd = g * ht[:, None] * np.cos(k * lon)[None, :] + 5.0e4
    for got, want in ((u, want_u), (v, want_v), (gd, want_gd)):
        assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max()
    assert np.abs(u).max() > 1e-3 and np.abs(gd - 5e4).max() > 0.1   # the wave is there


@pytest.mark.parametrize("name", ["rh_36x19_csp2", "rh_72x37_nosplit", "mz_60x31_upwind", "jz_72x37_diffusion",
                                  "sg_48x25_isp", "mz_48x25_weno", "mz_60x31_rk3_csp2", "sw_72x37_rk4_nosplit",
                                  "jz_72x37_reduce", "mz_60x31_reduce_plain"])
def test_oracle_reproduces_committed_golden(name, golden_dir):
    from golden.make_golden import CASES
    d = np.load(golden_dir / f"case_{name}.npz")
    kw, tc, n = CASES[name]
    o = Oracle(OracleConfig(**kw))
    o.set_initial_condition(tc)
    for x, k in zip(o.state(), ("u0", "v0", "gd0")):
        assert np.array_equal(x, d[k])
    o.run_init()
    o.step(n)
    for x, k in zip(o.state(), ("u1", "v1", "gd1")):
        assert np.abs(x - d[k]).max() <= 1e-12 * max(1.0, np.abs(d[k]).max())
    m, e, b = o.diag()
    assert abs(m - d["mass"][-1]) <= 1e-14 * abs(m) and abs(e - d["energy"][-1]) <= 1e-14 * abs(e)
