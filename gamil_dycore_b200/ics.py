"""NumPy mirror of the reference's test-case plugins (src/test_cases/barotropic/*_test_mod.F90), compact layout.

Host-side input synthesis for bench.py and the tests (the product's own IC plugins are the C++ ones in
gamil_dycore_b200/host/test_cases.cpp, used by the dycore_test driver).  Mesh as in src/mesh_mod.F90:44-114, i.e. with
cos(pole) = 0 -- the plugins run BEFORE reset_cos_lat_at_poles.
"""
import numpy as np

PI = 4.0 * np.arctan(1.0)
OMEGA = 2.0 * PI / 86400.0
RADIUS = 6.37122e6
G = 9.80616


def mesh(nlon, nlat):
    dlon = 2 * PI / nlon
    dlat = PI / (nlat - 1)
    full_lon = np.arange(nlon) * dlon
    half_lon = full_lon + 0.5 * dlon
    full_lat = -0.5 * PI + np.arange(nlat) * dlat
    full_lat[-1] = 0.5 * PI
    half_lat = full_lat[:-1] + 0.5 * dlat
    fc, fs = np.cos(full_lat), np.sin(full_lat)
    fc[0] = fc[-1] = 0.0
    fs[0], fs[-1] = -1.0, 1.0
    return dict(full_lon=full_lon, half_lon=half_lon, full_lat=full_lat, half_lat=half_lat, full_cos=fc, full_sin=fs,
                half_cos=np.cos(half_lat), half_sin=np.sin(half_lat))


def rossby_haurwitz_wave(nlon, nlat, R=4.0, omg=7.848e-6, gd0=8.0e3 * G):
    """rossby_haurwitz_wave_test_mod.F90:33-85"""
    m = mesh(nlon, nlat)
    cl, sl = m["full_cos"][:, None], m["full_sin"][:, None]
    lon = m["half_lon"][None, :]
    u = RADIUS * omg * (cl + R * cl ** (R - 1) * sl ** 2 * np.cos(R * lon) - cl ** (R + 1) * np.cos(R * lon))
    cl, sl = m["half_cos"][:, None], m["half_sin"][:, None]
    lon = m["full_lon"][None, :]
    v = -RADIUS * omg * (R * cl ** (R - 1) * sl * np.sin(R * lon))
    cl = m["full_cos"][:, None]
    a = 0.5 * omg * (2 * OMEGA + omg) * cl ** 2 + 0.25 * omg ** 2 * (
        (R + 1) * cl ** (2 * R + 2) + (2 * R ** 2 - R - 2) * cl ** (2 * R) - 2 * R ** 2 * cl ** (2 * R - 2))
    b = 2 * (OMEGA + omg) * omg * cl ** R * (R ** 2 + 2 * R + 2 - (R + 1) ** 2 * cl ** 2) / (R + 1) / (R + 2)
    c = 0.25 * omg ** 2 * cl ** (2 * R) * ((R + 1) * cl ** 2 - R - 2)
    gd = gd0 + RADIUS ** 2 * (a + b * np.cos(R * lon) + c * np.cos(2 * R * lon))
    return u, v, gd, np.zeros((nlat, nlon))


def _zonal_flow(nlon, nlat, u0, gd0, ghs):
    m = mesh(nlon, nlat)
    u = np.zeros((nlat, nlon))
    u[1:-1] = u0 * m["full_cos"][1:-1, None]
    v = np.zeros((nlat - 1, nlon))
    gd = gd0 - (RADIUS * OMEGA * u0 + u0 ** 2 * 0.5) * (m["full_sin"][:, None] ** 2) - ghs
    return u, v, gd, ghs


def steady_geostrophic_flow(nlon, nlat):
    """steady_geostrophic_flow_test_mod.F90:20-62 (alpha = 0)"""
    u0 = 2 * PI * RADIUS / (12 * 86400.0)
    return _zonal_flow(nlon, nlat, u0, 2.94e4, np.zeros((nlat, nlon)))


def mountain_zonal_flow(nlon, nlat):
    """mountain_zonal_flow_test_mod.F90:28-98 (smooth_mountain = .false.)"""
    m = mesh(nlon, nlat)
    lon0, lat0, ghs0, R = PI * 1.5, PI / 6.0, 2000.0 * G, PI / 9.0
    dlon = np.abs(m["full_lon"] - lon0)
    dlon = np.minimum(dlon, 2 * PI - dlon)[None, :]
    dd = np.minimum(R, np.sqrt(dlon ** 2 + (m["full_lat"][:, None] - lat0) ** 2))
    ghs = ghs0 * (1.0 - dd / R)
    return _zonal_flow(nlon, nlat, 20.0, 5960.0 * G, ghs)


CASES = {"rossby_haurwitz_wave": rossby_haurwitz_wave, "steady_geostrophic_flow": steady_geostrophic_flow,
         "mountain_zonal_flow": mountain_zonal_flow}
