"""gamil_dycore_b200 -- B200-native barotropic shallow-water time step of dongli/gamil-dycore.

This Python module is only the thin ctypes binding of the C ABI in ``include/gmd.h`` (the product is
``libgmd.so``: hand-written sm_100a CUDA behind the reference's ``dycore_init/run/final`` interface,
src/dycore_mod.F90:22-25).  It is used by the test-suite, ``bench.py`` and ``__graft_entry__.py``; the
reference-facing host program is the C++ ``dycore_test`` in ``gamil_dycore_b200/host`` (and the Fortran shim in
``fortran/``).  There is no CPU fallback: every compute call raises ``GmdError`` without a CUDA device.
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
LAYOUT_COMPACT, LAYOUT_REFERENCE = 0, 1
OK, ERR_NAN, ERR_ARG, ERR_CUDA, ERR_COMM, ERR_STATE = range(6)


class _Cfg(C.Structure):
    _fields_ = [
        ("num_lon", C.c_int), ("num_lat", C.c_int), ("subcycles", C.c_int),
        ("time_step_size", C.c_double), ("qcon_modified", C.c_int), ("split_scheme", C.c_int),
        ("uv_adv_scheme", C.c_int), ("uv_adv_upwind_lon_beta", C.c_double),
        ("uv_adv_upwind_lat_beta", C.c_double), ("use_zonal_tend_filter", C.c_int),
        ("cutoff", C.c_int * 20), ("use_diffusion", C.c_int), ("diffusion_order", C.c_int),
        ("diffusion_coef", C.c_double), ("rank", C.c_int), ("nranks", C.c_int), ("device", C.c_int),
        ("polar_band_rows", C.c_int), ("time_scheme", C.c_int), ("time_order", C.c_int),
        ("use_zonal_reduce", C.c_int), ("reduce_adv_lon", C.c_int), ("use_reduce_tend_smooth", C.c_int),
        ("zonal_reduce_factors", C.c_int * 20),
    ]


@dataclass
class Config:
    """Numeric keys of /dycore_params/ with the reference's defaults (src/params_mod.F90:13-66)."""
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
    rank: int = 0
    nranks: int = 1
    device: int = -1
    polar_band_rows: int = 0
    time_scheme: str = "predict_correct"     # or "runge_kutta" (specified extension)
    time_order: int = 3
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
        cw = self.zonal_tend_filter_cutoff_wavenumber
        for k in range(20):
            c.cutoff[k] = cw[k] if k < len(cw) else 0
        c.use_diffusion = int(self.use_diffusion)
        c.diffusion_order = self.diffusion_order
        c.diffusion_coef = self.diffusion_coef
        c.rank, c.nranks, c.device = self.rank, self.nranks, self.device
        c.polar_band_rows = self.polar_band_rows
        c.time_scheme = TIME[self.time_scheme]
        c.time_order = self.time_order
        c.use_zonal_reduce = int(self.use_zonal_reduce)
        c.reduce_adv_lon = int(self.reduce_adv_lon)
        c.use_reduce_tend_smooth = int(self.use_reduce_tend_smooth)
        for k in range(20):
            c.zonal_reduce_factors[k] = self.zonal_reduce_factors[k] if k < len(self.zonal_reduce_factors) else 0
        return c


class GmdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gmd error {code}: {msg}")
        self.code = code


LIB_NAMES = {"fast": "libgmd.so", "strict": "libgmd_strict.so", "trace": "libgmd_trace.so"}


def build(force: bool = False) -> None:
    """Compile libgmd.so / libgmd_strict.so / dycore_test in place with nvcc for sm_100a (no GPU needed)."""
    args = ["make", "-C", _HERE, "-j4"] + (["-B"] if force else [])
    res = subprocess.run(args, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libgmd failed:\n" + res.stdout + res.stderr)


_LIBS = {}


def load(kind: str = "fast") -> C.CDLL:
    """Load the product library.  Raises (never falls back) if it has not been built.
    `kind` may also be a path to an experimental build of the same sources (tools/tune_stage.py)."""
    if kind in _LIBS:
        return _LIBS[kind]
    path = kind if os.path.sep in kind else os.path.join(_HERE, LIB_NAMES[kind])
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)")
    lib = C.CDLL(path)
    P, D, I = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.gmd_config_defaults.argtypes = [C.POINTER(_Cfg)]
    lib.gmd_config_defaults.restype = None
    lib.gmd_create.argtypes = [C.POINTER(_Cfg), C.POINTER(P)]
    lib.gmd_destroy.argtypes = [P]
    lib.gmd_destroy.restype = None
    lib.gmd_last_error.restype = C.c_char_p
    lib.gmd_comm_unique_id.argtypes = [C.c_void_p]
    lib.gmd_comm_init.argtypes = [P, C.c_void_p]
    lib.gmd_peer_export.argtypes = [P, C.c_void_p]
    lib.gmd_peer_connect.argtypes = [P, C.c_void_p, C.c_int]
    lib.gmd_peer_disconnect.argtypes = [P]
    lib.gmd_get_band.argtypes = [P, I, I]
    lib.gmd_get_fused_rows.argtypes = [P, I, I]
    lib.gmd_set_state.argtypes = [P, D, D, D, D, C.c_int]
    lib.gmd_run_init.argtypes = [P]
    lib.gmd_step.argtypes = [P, C.c_int]
    lib.gmd_step_async.argtypes = [P, C.c_int]
    lib.gmd_sync.argtypes = [P]
    lib.gmd_get_state.argtypes = [P, D, D, D, C.c_int]
    lib.gmd_get_iap_state.argtypes = [P, D, D, D, C.c_int]
    lib.gmd_get_diag.argtypes = [P, D, D, D]
    lib.gmd_get_diag_series.argtypes = [P, C.c_int, D, D, D]
    lib.gmd_get_vor_div.argtypes = [P, D, D, C.c_int]
    lib.gmd_get_step_count.argtypes = [P]
    lib.gmd_space_operators.argtypes = [P, C.c_int, D, D, D, C.c_int]
    lib.gmd_predict_correct.argtypes = [P, C.c_double, C.c_int]
    lib.gmd_ordinary_diffusion.argtypes = [P, C.c_double]
    lib.gmd_filter_row.argtypes = [P, C.c_int, C.c_int, D]
    lib.gmd_get_filter_rows.argtypes = [P, I, I, I, I]
    lib.gmd_get_table.argtypes = [P, C.c_int, D]
    lib.gmd_set_stream.argtypes = [P, C.c_void_p]
    lib.gmd_set_graph_mode.argtypes = [P, C.c_int]
    lib.gmd_last_step_ms.argtypes = [P, C.POINTER(C.c_float)]
    lib.gmd_kernel_launches.argtypes = [P]
    lib.gmd_kernel_launches.restype = C.c_longlong
    lib.gmd_algorithmic_bytes_per_column_step.argtypes = [P]
    lib.gmd_algorithmic_bytes_per_column_step.restype = C.c_double
    lib.gmd_time_stage_kernel.argtypes = [P, C.c_int, C.POINTER(C.c_float), D]
    lib.gmd_time_stage_variant.argtypes = [P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), D]
    lib.gmd_trace_begin.argtypes = [P]
    lib.gmd_trace_dump.argtypes = [P, C.c_char_p]
    _LIBS[kind] = lib
    return lib


PEER_BLOB_BYTES = 256   # GMD_PEER_BLOB_BYTES, include/gmd.h

_HOST = None


def load_host() -> C.CDLL:
    """libgmd_host.so: the host side of the drop-in behind a C ABI (include/gmd_host.h) -- the IC plugins."""
    global _HOST
    if _HOST is None:
        path = os.path.join(_HERE, "libgmd_host.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(path)
        D = C.POINTER(C.c_double)
        lib.gmd_host_initial_condition.argtypes = [C.c_char_p, C.c_int, C.c_int, D, D, D, D]
        lib.gmd_host_last_error.restype = C.c_char_p
        _HOST = lib
    return _HOST


def initial_condition(test_case: str, num_lon: int, num_lat: int):
    """(u, v, gd, ghs) of the named test-case plugin (src/test_cases/barotropic/*_test_mod.F90 as rebuilt in
    gamil_dycore_b200/host/test_cases.cpp), compact layout"""
    lib = load_host()
    u, v = np.zeros((num_lat, num_lon)), np.zeros((num_lat - 1, num_lon))
    gd, ghs = np.zeros((num_lat, num_lon)), np.zeros((num_lat, num_lon))
    ier = lib.gmd_host_initial_condition(test_case.encode(), num_lon, num_lat, _dp(u), _dp(v), _dp(gd), _dp(ghs))
    if ier:
        raise GmdError(ier, lib.gmd_host_last_error().decode())
    return u, v, gd, ghs


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def comm_unique_id(kind: str = "fast") -> bytes:
    """NCCL unique id (128 bytes) created on the calling rank; distribute it, then Dycore.comm_init on all ranks."""
    lib = load(kind)
    buf = C.create_string_buffer(128)
    ier = lib.gmd_comm_unique_id(buf)
    if ier:
        raise GmdError(ier, lib.gmd_last_error().decode())
    return buf.raw


class Dycore:
    """One model instance: ``dycore_init`` ... ``dycore_final`` of the reference (src/dycore_mod.F90:60-157)."""

    def __init__(self, cfg: Config, kind: str = "fast"):
        self.lib = load(kind)
        self.cfg = cfg
        self.nlon, self.nlat = cfg.num_lon, cfg.num_lat
        h = C.c_void_p()
        c = cfg.to_c()
        ier = self.lib.gmd_create(C.byref(c), C.byref(h))
        if ier:
            raise GmdError(ier, self.lib.gmd_last_error().decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.gmd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, ier):
        if ier:
            raise GmdError(ier, self.lib.gmd_last_error().decode())

    def full(self):
        return np.zeros((self.nlat, self.nlon))

    def half(self):
        return np.zeros((self.nlat - 1, self.nlon))

    def comm_init(self, uid: bytes):
        self._chk(self.lib.gmd_comm_init(self.h, C.create_string_buffer(uid, 128)))

    def peer_export(self) -> bytes:
        """this rank's peer blob (CUDA-IPC handles of its field slab and signal page), gmd_peer_export"""
        buf = C.create_string_buffer(PEER_BLOB_BYTES)
        self._chk(self.lib.gmd_peer_export(self.h, buf))
        return buf.raw

    def peer_connect(self, blobs):
        """`blobs`: the peer blobs of ALL ranks in rank order (list of bytes); switches halo rows and the
        two-scalar all-reduces to the peer-memory path, gmd_peer_connect"""
        raw = b"".join(blobs)
        self._chk(self.lib.gmd_peer_connect(self.h, C.create_string_buffer(raw, len(raw)), len(blobs)))

    def peer_disconnect(self):
        self._chk(self.lib.gmd_peer_disconnect(self.h))

    def band(self):
        a, b = C.c_int(), C.c_int()
        self._chk(self.lib.gmd_get_band(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def fused_rows(self):
        """rows [a, b) whose predict_correct runs as the fused kernel k_pc; (0, 0) when it is not in use"""
        a, b = C.c_int(), C.c_int()
        self._chk(self.lib.gmd_get_fused_rows(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_state(self, u, v, gd, ghs=None, layout=LAYOUT_COMPACT):
        u, v, gd = (np.ascontiguousarray(a, dtype=np.float64) for a in (u, v, gd))
        ghs = None if ghs is None else np.ascontiguousarray(ghs, dtype=np.float64)
        self._chk(self.lib.gmd_set_state(self.h, _dp(u), _dp(v), _dp(gd), _dp(ghs), layout))

    def run_init(self):
        self._chk(self.lib.gmd_run_init(self.h))

    def step(self, n=1):
        self._chk(self.lib.gmd_step(self.h, n))

    def step_async(self, n=1):
        self._chk(self.lib.gmd_step_async(self.h, n))

    def sync(self):
        self._chk(self.lib.gmd_sync(self.h))

    def state(self):
        u, v, gd = self.full(), self.half(), self.full()
        self._chk(self.lib.gmd_get_state(self.h, _dp(u), _dp(v), _dp(gd), LAYOUT_COMPACT))
        return u, v, gd

    def state_reference_layout(self):
        shp = (self.nlat + 4, self.nlon + 4)
        u, v, gd = np.zeros(shp), np.zeros(shp), np.zeros(shp)
        self._chk(self.lib.gmd_get_state(self.h, _dp(u), _dp(v), _dp(gd), LAYOUT_REFERENCE))
        return u, v, gd

    def iap_state(self):
        u, v, gd = self.full(), self.half(), self.full()
        self._chk(self.lib.gmd_get_iap_state(self.h, _dp(u), _dp(v), _dp(gd), LAYOUT_COMPACT))
        return u, v, gd

    def diag(self):
        m, e, b = C.c_double(), C.c_double(), C.c_double()
        self._chk(self.lib.gmd_get_diag(self.h, C.byref(m), C.byref(e), C.byref(b)))
        return m.value, e.value, b.value

    def diag_series(self, n):
        m, e, b = np.zeros(n), np.zeros(n), np.zeros(n)
        self._chk(self.lib.gmd_get_diag_series(self.h, n, _dp(m), _dp(e), _dp(b)))
        return m, e, b

    def vor_div(self):
        vor, div = self.half(), self.full()
        self._chk(self.lib.gmd_get_vor_div(self.h, _dp(vor), _dp(div), LAYOUT_COMPACT))
        return vor, div

    def step_count(self):
        return self.lib.gmd_get_step_count(self.h)

    def space_operators(self, pass_: str = "all"):
        du, dv, dgd = self.full(), self.half(), self.full()
        self._chk(self.lib.gmd_space_operators(self.h, PASS[pass_], _dp(du), _dp(dv), _dp(dgd), LAYOUT_COMPACT))
        return du, dv, dgd

    def predict_correct(self, dt: float, pass_: str = "all"):
        self._chk(self.lib.gmd_predict_correct(self.h, dt, PASS[pass_]))

    def ordinary_diffusion(self, dt: float):
        self._chk(self.lib.gmd_ordinary_diffusion(self.h, dt))

    def filter_row(self, half: bool, row0: int, x):
        x = np.array(x, dtype=np.float64)
        self._chk(self.lib.gmd_filter_row(self.h, int(half), row0, _dp(x)))
        return x

    def filter_rows(self):
        ff, fc = np.zeros(self.nlat, np.int32), np.zeros(self.nlat, np.int32)
        hf, hc = np.zeros(self.nlat - 1, np.int32), np.zeros(self.nlat - 1, np.int32)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self._chk(self.lib.gmd_get_filter_rows(self.h, ip(ff), ip(fc), ip(hf), ip(hc)))
        return ff, fc, hf, hc

    def table(self, which: int):
        n = self.nlat - 1 if which in (1, 5, 7, 9) else self.nlat
        out = np.zeros(n)
        self._chk(self.lib.gmd_get_table(self.h, which, _dp(out)))
        return out

    def set_stream(self, cuda_stream: int):
        self._chk(self.lib.gmd_set_stream(self.h, C.c_void_p(cuda_stream)))

    def set_graph_mode(self, on: bool):
        self._chk(self.lib.gmd_set_graph_mode(self.h, int(on)))

    def last_step_ms(self) -> float:
        ms = C.c_float()
        self._chk(self.lib.gmd_last_step_ms(self.h, C.byref(ms)))
        return ms.value

    def kernel_launches(self) -> int:
        return int(self.lib.gmd_kernel_launches(self.h))

    def algorithmic_bytes_per_column_step(self) -> float:
        return float(self.lib.gmd_algorithmic_bytes_per_column_step(self.h))

    def time_stage_variant(self, pass_: str, mode: int, reps: int = 20):
        ms, nb = C.c_float(), C.c_double()
        self._chk(self.lib.gmd_time_stage_variant(self.h, PASS[pass_], mode, reps, C.byref(ms), C.byref(nb)))
        return ms.value, nb.value

    def trace_begin(self):
        """start recording the device timeline of the following steps (kind="trace" build only)"""
        self._chk(self.lib.gmd_trace_begin(self.h))

    def trace_end(self):
        """the recorded timeline: {"steps": [{"step", "launches": [{"seq", "kernel", "start_ns", "end_ns", "wait_ns"}]}]}"""
        import json
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".json") as f:
            self._chk(self.lib.gmd_trace_dump(self.h, f.name.encode()))
            return json.load(open(f.name))

    def time_stage_kernel(self, reps: int = 20):
        ms, nb = C.c_float(), C.c_double()
        self._chk(self.lib.gmd_time_stage_kernel(self.h, reps, C.byref(ms), C.byref(nb)))
        return ms.value, nb.value
