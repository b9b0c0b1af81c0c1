"""Regenerates the committed fixtures under tests/golden/.  Run from the repo root:

    python tests/golden/make_golden.py

Two kinds of fixture:

* ``jet_gd_profile_181.npz`` -- the balanced geopotential profile of the Galewsky-jet initial
  condition computed with SciPy's ``quad`` (QUADPACK QAGS, the algorithm of the reference's
  lib/quadpack.f90:1877 ``qags``) with the reference's own tolerances epsabs=1e-10, epsrel=1e-3
  (jet_zonal_flow_test_mod.F90:63).  This is the one vector on the path that comes from the
  reference's third-party algorithm itself rather than from our restatement.
* ``case_*.npz`` -- small end-to-end cases produced by the strict binary64 CPU oracle
  (oracle/liboracle.so): initial state, state after ``nsteps`` model steps and the mass / energy /
  beta series.  The reference ships no golden output for this path and cannot be built here
  (PARITY UNPINNED, oracle/gmd_oracle.h), so these pin the CUDA path to the oracle, not the oracle
  to the reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.oracle import Oracle, OracleConfig  # noqa: E402

CASES = {
    # name: (config kwargs, test case, nsteps)
    "rh_36x19_csp2": (dict(num_lon=36, num_lat=19, time_step_size=1800.0, subcycles=4, split_scheme="csp2",
                           zonal_tend_filter_cutoff_wavenumber=[4, 4]), "rossby_haurwitz_wave", 4),
    "rh_72x37_nosplit": (dict(num_lon=72, num_lat=37, time_step_size=300.0, split_scheme="none",
                              zonal_tend_filter_cutoff_wavenumber=[4, 4, 4]), "rossby_haurwitz_wave", 4),
    "mz_60x31_upwind": (dict(num_lon=60, num_lat=31, time_step_size=1200.0, subcycles=8, split_scheme="csp2",
                             uv_adv_scheme="upwind", uv_adv_upwind_lat_beta=0.1,
                             zonal_tend_filter_cutoff_wavenumber=[4, 4, 4]), "mountain_zonal_flow", 4),
    "jz_72x37_diffusion": (dict(num_lon=72, num_lat=37, time_step_size=900.0, subcycles=6, split_scheme="csp2",
                                zonal_tend_filter_cutoff_wavenumber=[4] * 4, use_diffusion=True,
                                diffusion_coef=1.0e5), "jet_zonal_flow", 4),
    "sg_48x25_isp": (dict(num_lon=48, num_lat=25, time_step_size=1200.0, subcycles=4, split_scheme="isp",
                          zonal_tend_filter_cutoff_wavenumber=[3, 3]), "steady_geostrophic_flow", 3),
    "mz_48x25_weno": (dict(num_lon=48, num_lat=25, time_step_size=600.0, subcycles=4, split_scheme="csp2",
                           uv_adv_scheme="weno", zonal_tend_filter_cutoff_wavenumber=[4, 4]),
                      "mountain_zonal_flow", 3),
    # order-4 diffusion (two Laplacian passes, src/diffusion_mod.F90:107,172-179), unsplit
    "mz_60x32_diff4": (dict(num_lon=60, num_lat=32, time_step_size=600.0, split_scheme="none", use_diffusion=True,
                            diffusion_order=4, diffusion_coef=1.0e15, zonal_tend_filter_cutoff_wavenumber=[4, 4]),
                       "mountain_zonal_flow", 4),
    # runge_kutta: the specified extension of DESIGN.md section 8 (not a reference feature; pins CUDA to the oracle)
    "mz_60x31_rk3_csp2": (dict(num_lon=60, num_lat=31, time_step_size=900.0, subcycles=4, split_scheme="csp2",
                               time_scheme="runge_kutta", time_order=3, zonal_tend_filter_cutoff_wavenumber=[4, 4, 4]),
                          "mountain_zonal_flow", 4),
    "sw_72x37_rk4_nosplit": (dict(num_lon=72, num_lat=37, time_step_size=150.0, split_scheme="none",
                                  time_scheme="runge_kutta", time_order=4, zonal_tend_filter_cutoff_wavenumber=[12, 12, 12]),
                             "shallow_water_waves", 4),
    # moving reduced tendency: the specified extension of DESIGN.md section 8 (keys of the reference's run/namelist.jz_test)
    "jz_72x37_reduce": (dict(num_lon=72, num_lat=37, time_step_size=900.0, subcycles=6, split_scheme="csp2",
                             use_zonal_reduce=True, reduce_adv_lon=True, use_reduce_tend_smooth=True,
                             zonal_reduce_factors=[8, 4, 2, 2], use_diffusion=True, diffusion_coef=1.0e5),
                        "jet_zonal_flow", 4),
    "mz_60x31_reduce_plain": (dict(num_lon=60, num_lat=31, time_step_size=600.0, subcycles=4, split_scheme="csp2",
                                   use_zonal_reduce=True, zonal_reduce_factors=[6, 3, 2]), "mountain_zonal_flow", 4),
}


def make_jet_profile():
    from scipy.integrate import quad
    pi = 4 * np.arctan(1.0)
    omega, radius, g = 2 * pi / 86400.0, 6.37122e6, 9.80616
    u_max, lat0 = 80.0, pi / 7
    lat1 = pi / 2 - lat0
    en = np.exp(-4 / (lat1 - lat0) ** 2)

    def uf(lat):
        return 0.0 if (lat <= lat0 or lat >= lat1) else u_max / en * np.exp(1 / (lat - lat0) / (lat - lat1))

    def integrand(lat):
        u = uf(lat)
        return radius * u * (2 * omega * np.sin(lat) + np.tan(lat) / radius * u)

    nlat = 181
    lats = -0.5 * pi + np.arange(nlat) * (pi / (nlat - 1))
    lats[-1] = 0.5 * pi
    gd = np.empty(nlat)
    gd[0] = g * 1.0e4
    for j in range(1, nlat):
        r, _ = quad(integrand, -0.5 * pi, lats[j], epsabs=1.0e-10, epsrel=1.0e-3, limit=500)
        gd[j] = g * 1.0e4 - r
    np.savez(os.path.join(HERE, "jet_gd_profile_181.npz"), lat=lats, gd=gd)


def make_case(name, kw, test_case, nsteps):
    o = Oracle(OracleConfig(**kw))
    o.set_initial_condition(test_case)
    u0, v0, gd0 = o.state()
    ghs = o.ghs()
    o.run_init()
    series = [o.diag()]
    for _ in range(nsteps):
        o.step(1)
        series.append(o.diag())
    u1, v1, gd1 = o.state()
    series = np.array(series)
    np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), u0=u0, v0=v0, gd0=gd0, ghs=ghs, u1=u1, v1=v1,
                        gd1=gd1, mass=series[:, 0], energy=series[:, 1], beta=series[:, 2], nsteps=nsteps,
                        test_case=test_case, config=repr(kw))


if __name__ == "__main__":
    only = sys.argv[1:]
    if not only:
        make_jet_profile()
    # NB: a Rossby-Haurwitz wave (R=4) must not be combined with a cutoff < 4: the filtered dv row is
    # then orthogonal to V (pure sin 4 lambda), s2 is rounding noise and s1/s2 blows up -- that is the
    # reference's own behaviour (dycore_mod.F90:229-234), not something a fixture should pin.
    for name, (kw, tc, n) in CASES.items():
        if only and name not in only:
            continue
        make_case(name, kw, tc, n)
        print("wrote", name)
