"""ctypes binding of the CPU oracle (oracle/gmd_oracle.h).

TEST INFRASTRUCTURE ONLY.  Import this from tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- never from the product package
``gamil_dycore_b200``.  PARITY UNPINNED: see the header of oracle/gmd_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

SPLIT = {"none": 0, "": 0, "csp1": 1, "csp2": 2, "isp": 3}
ADV = {"center_diff": 0, "upwind": 1, "weno": 2}
PASS = {"all": 0, "fast": 1, "slow": 2}
TIME = {"predict_correct": 0, "runge_kutta": 1}
IC = {"rossby_haurwitz_wave": 0, "steady_geostrophic_flow": 1, "mountain_zonal_flow": 2, "jet_zonal_flow": 3,
      "shallow_water_waves": 4}


class _Cfg(C.Structure):
    _fields_ = [
        ("num_lon", C.c_int), ("num_lat", C.c_int), ("subcycles", C.c_int),
        ("time_step_size", C.c_double), ("qcon_modified", C.c_int), ("split_scheme", C.c_int),
        ("uv_adv_scheme", C.c_int), ("uv_adv_upwind_lon_beta", C.c_double),
        ("uv_adv_upwind_lat_beta", C.c_double), ("use_zonal_tend_filter", C.c_int),
        ("cutoff", C.c_int * 20), ("use_diffusion", C.c_int), ("diffusion_order", C.c_int),
        ("diffusion_coef", C.c_double), ("time_scheme", C.c_int), ("time_order", C.c_int),
        ("use_zonal_reduce", C.c_int), ("reduce_adv_lon", C.c_int), ("use_reduce_tend_smooth", C.c_int),
        ("zonal_reduce_factors", C.c_int * 20),
    ]


@dataclass
class OracleConfig:
    """Numeric keys of /dycore_params/ with the reference's defaults (params_mod.F90:13-66)."""
    num_lon: int
    num_lat: int
    time_step_size: float
    subcycles: int = 4
    qcon_modified: bool = True
    split_scheme: str = "csp2"
    uv_adv_scheme: str = "center_diff"
    uv_adv_upwind_lon_beta: float = 0.0
    uv_adv_upwind_lat_beta: float = 0.5
    use_zonal_tend_filter: bool = True
    zonal_tend_filter_cutoff_wavenumber: List[int] = field(default_factory=list)
    use_diffusion: bool = False
    diffusion_order: int = 2
    diffusion_coef: float = 0.0
    time_scheme: str = "predict_correct"     # or "runge_kutta" (specified extension)
    time_order: int = 3                      # runge_kutta: 3 (SSP-RK3) or 4 (classical)
    use_zonal_reduce: bool = False           # moving reduced tendency (specified extension)
    reduce_adv_lon: bool = False
    use_reduce_tend_smooth: bool = False
    zonal_reduce_factors: List[int] = field(default_factory=list)

    def to_c(self) -> _Cfg:
        c = _Cfg()
        c.num_lon, c.num_lat, c.subcycles = self.num_lon, self.num_lat, self.subcycles
        c.time_step_size = self.time_step_size
        c.qcon_modified = int(self.qcon_modified)
        c.split_scheme = SPLIT[self.split_scheme]
        c.uv_adv_scheme = ADV[self.uv_adv_scheme]
        c.uv_adv_upwind_lon_beta = self.uv_adv_upwind_lon_beta
        c.uv_adv_upwind_lat_beta = self.uv_adv_upwind_lat_beta
        c.use_zonal_tend_filter = int(self.use_zonal_tend_filter)
        for k in range(20):
            c.cutoff[k] = self.zonal_tend_filter_cutoff_wavenumber[k] if k < len(self.zonal_tend_filter_cutoff_wavenumber) else 0
        c.use_diffusion = int(self.use_diffusion)
        c.diffusion_order = self.diffusion_order
        c.diffusion_coef = self.diffusion_coef
        c.time_scheme = TIME[self.time_scheme]
        c.time_order = self.time_order
        c.use_zonal_reduce = int(self.use_zonal_reduce)
        c.reduce_adv_lon = int(self.reduce_adv_lon)
        c.use_reduce_tend_smooth = int(self.use_reduce_tend_smooth)
        for k in range(20):
            c.zonal_reduce_factors[k] = self.zonal_reduce_factors[k] if k < len(self.zonal_reduce_factors) else 0
        return c


def build(force: bool = False) -> None:
    """Compile the three oracle libraries in place (gcc, a few seconds)."""
    if force or not all(os.path.exists(os.path.join(_HERE, n)) for n in ("liboracle.so", "liboracle_fast.so", "liboracle_quad.so")):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)


def _cpu_has_avx2_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = line.split()
                    return "avx2" in fl and "fma" in fl
    except OSError:
        pass
    return False


_LIBS = {}


def load(kind: str = "strict") -> C.CDLL:
    """kind: 'strict' (binary64, IEEE), 'fast' (-O3 -ffast-math, timing only), 'quad' (binary128)."""
    if kind == "fast" and not _cpu_has_avx2_fma():
        kind = "strict"
    if kind in _LIBS:
        return _LIBS[kind]
    name = {"strict": "liboracle.so", "fast": "liboracle_fast.so", "quad": "liboracle_quad.so"}[kind]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    P, D, I = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.orc_create.argtypes = [C.POINTER(_Cfg), C.POINTER(P)]
    lib.orc_destroy.argtypes = [P]
    lib.orc_destroy.restype = None
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_set_initial_condition.argtypes = [P, C.c_int, D, C.c_int]
    lib.orc_set_state.argtypes = [P, D, D, D, D]
    lib.orc_run_init.argtypes = [P]
    lib.orc_step.argtypes = [P, C.c_int]
    lib.orc_get_state.argtypes = [P, D, D, D]
    lib.orc_get_iap_state.argtypes = [P, D, D, D]
    lib.orc_get_ghs.argtypes = [P, D]
    lib.orc_get_diag.argtypes = [P, D, D, D]
    lib.orc_get_vor_div.argtypes = [P, D, D]
    lib.orc_get_step_count.argtypes = [P]
    lib.orc_space_operators.argtypes = [P, C.c_int, D, D, D]
    lib.orc_check_antisymmetry.argtypes = [P, D]
    lib.orc_update_state_preview.argtypes = [P, C.c_double, D, D, D, D, D, D]
    lib.orc_predict_correct.argtypes = [P, C.c_double, C.c_int]
    lib.orc_ordinary_diffusion.argtypes = [P, C.c_double]
    lib.orc_get_table.argtypes = [P, C.c_int, D]
    lib.orc_get_filter_rows.argtypes = [P, I, I, I, I]
    lib.orc_filter_row.argtypes = [P, C.c_int, C.c_int, D]
    lib.orc_jet_gd_profile.argtypes = [C.c_double]
    lib.orc_jet_gd_profile.restype = C.c_double
    lib.orc_rfft_forward_f64.argtypes = [C.c_int, D]
    lib.orc_rfft_backward_f64.argtypes = [C.c_int, D]
    lib.orc_rfft_factors.argtypes = [C.c_int, I, C.c_int]
    lib.orc_real_bytes.restype = C.c_int
    lib.orc_swe_phase_speed.argtypes = [C.c_int]
    lib.orc_swe_phase_speed.restype = C.c_double
    _LIBS[kind] = lib
    return lib


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleError(RuntimeError):
    pass


class Oracle:
    """One oracle model instance (mirrors dycore_init ... dycore_final of the reference)."""

    def __init__(self, cfg: OracleConfig, kind: str = "strict"):
        self.lib = load(kind)
        self.cfg = cfg
        self.nlon, self.nlat = cfg.num_lon, cfg.num_lat
        h = C.c_void_p()
        c = cfg.to_c()
        ier = self.lib.orc_create(C.byref(c), C.byref(h))
        if ier:
            raise OracleError(self.lib.orc_last_error().decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, ier):
        if ier:
            raise OracleError(self.lib.orc_last_error().decode())

    # shapes
    def full(self):
        return np.zeros((self.nlat, self.nlon))

    def half(self):
        return np.zeros((self.nlat - 1, self.nlon))

    def set_initial_condition(self, test_case: str, params=None):
        p = np.ascontiguousarray(params, dtype=np.float64) if params is not None else None
        self._chk(self.lib.orc_set_initial_condition(self.h, IC[test_case], _dp(p), 0 if p is None else p.size))

    def set_state(self, u, v, gd, ghs=None):
        u, v, gd = (np.ascontiguousarray(a, dtype=np.float64) for a in (u, v, gd))
        ghs = None if ghs is None else np.ascontiguousarray(ghs, dtype=np.float64)
        self._chk(self.lib.orc_set_state(self.h, _dp(u), _dp(v), _dp(gd), _dp(ghs)))

    def run_init(self):
        self._chk(self.lib.orc_run_init(self.h))

    def step(self, n=1):
        self._chk(self.lib.orc_step(self.h, n))

    def state(self):
        u, v, gd = self.full(), self.half(), self.full()
        self.lib.orc_get_state(self.h, _dp(u), _dp(v), _dp(gd))
        return u, v, gd

    def iap_state(self):
        u, v, gd = self.full(), self.half(), self.full()
        self.lib.orc_get_iap_state(self.h, _dp(u), _dp(v), _dp(gd))
        return u, v, gd

    def ghs(self):
        g = self.full()
        self.lib.orc_get_ghs(self.h, _dp(g))
        return g

    def diag(self):
        m, e, b = C.c_double(), C.c_double(), C.c_double()
        self.lib.orc_get_diag(self.h, C.byref(m), C.byref(e), C.byref(b))
        return m.value, e.value, b.value

    def vor_div(self):
        vor, div = self.half(), self.full()
        self.lib.orc_get_vor_div(self.h, _dp(vor), _dp(div))
        return vor, div

    def space_operators(self, pass_: str = "all"):
        du, dv, dgd = self.full(), self.half(), self.full()
        self._chk(self.lib.orc_space_operators(self.h, PASS[pass_], _dp(du), _dp(dv), _dp(dgd)))
        return du, dv, dgd

    def check_antisymmetry(self):
        s = np.zeros(8)
        self.lib.orc_check_antisymmetry(self.h, _dp(s))
        return s[:4], s[4:]

    def update_state_preview(self, dt: float):
        u, v, gd, iu, iv, igd = self.full(), self.half(), self.full(), self.full(), self.half(), self.full()
        self.lib.orc_update_state_preview(self.h, dt, _dp(u), _dp(v), _dp(gd), _dp(iu), _dp(iv), _dp(igd))
        return u, v, gd, iu, iv, igd

    def predict_correct(self, dt: float, pass_: str = "all"):
        self._chk(self.lib.orc_predict_correct(self.h, dt, PASS[pass_]))

    def ordinary_diffusion(self, dt: float):
        self._chk(self.lib.orc_ordinary_diffusion(self.h, dt))

    def table(self, which: int):
        n = self.nlat - 1 if which in (1, 5, 7, 9) else self.nlat
        out = np.zeros(n)
        self.lib.orc_get_table(self.h, which, _dp(out))
        return out

    def filter_rows(self):
        ff, fc = np.zeros(self.nlat, np.int32), np.zeros(self.nlat, np.int32)
        hf, hc = np.zeros(self.nlat - 1, np.int32), np.zeros(self.nlat - 1, np.int32)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self.lib.orc_get_filter_rows(self.h, ip(ff), ip(fc), ip(hf), ip(hc))
        return ff, fc, hf, hc

    def filter_row(self, half: bool, row0: int, x):
        x = np.array(x, dtype=np.float64)
        self.lib.orc_filter_row(self.h, int(half), row0, _dp(x))
        return x


def rfft_forward(x, kind="strict"):
    x = np.array(x, dtype=np.float64)
    ier = load(kind).orc_rfft_forward_f64(x.size, _dp(x))
    if ier:
        raise OracleError(f"rfft1f ier={ier}")
    return x


def rfft_backward(x, kind="strict"):
    x = np.array(x, dtype=np.float64)
    ier = load(kind).orc_rfft_backward_f64(x.size, _dp(x))
    if ier:
        raise OracleError(f"rfft1b ier={ier}")
    return x


def rfft_factors(n, kind="strict"):
    fac = (C.c_int * 32)()
    nf = load(kind).orc_rfft_factors(n, fac, 32)
    if nf < 0:
        raise OracleError(f"rfft1i ier={-nf}")
    return list(fac[:nf])


def jet_gd_profile(lat: float) -> float:
    return load().orc_jet_gd_profile(lat)
