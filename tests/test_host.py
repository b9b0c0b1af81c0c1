"""CPU tests of the host side: the C ABI surface, the namelist reader, the log / clock formats, the IC plugins and
the history writer of gamil_dycore_b200/host (mirrors of src/params_mod.F90, log_mod, time_mod, test_cases, history_mod),
and the latitude-band helpers over a 2-process gloo group.  No GPU, no compute call into libgmd."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import gamil_dycore_b200 as gmd
from oracle.oracle import Oracle, OracleConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELF = os.path.join(ROOT, "gamil_dycore_b200", "host_selftest")


@pytest.fixture(scope="module", autouse=True)
def _build_product():
    gmd.build()


def selftest(*args, cwd=None):
    res = subprocess.run([SELF, *map(str, args)], capture_output=True, text=True, cwd=cwd)
    return res.returncode, res.stdout


def write(tmp_path, text, name="namelist"):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


# ------------------------------------------------------------------------------------------------ C ABI
@pytest.mark.parametrize("kind", ["fast", "strict", "trace"])
def test_library_exports_every_declared_symbol(kind):
    hdr = open(os.path.join(ROOT, "include", "gmd.h")).read()
    names = sorted(set(re.findall(r"\b(gmd_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    lib = ctypes.CDLL(os.path.join(ROOT, "gamil_dycore_b200", gmd.LIB_NAMES[kind]))
    for n in names:
        assert hasattr(lib, n), n
    assert lib.gmd_version() == 100


def test_host_library_exports_every_declared_symbol():
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "gmd_host.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(gmd_host_[a-z0-9_]+)\s*\(", hdr)))
    assert names == ["gmd_host_initial_condition", "gmd_host_last_error"]
    lib = ctypes.CDLL(os.path.join(ROOT, "gamil_dycore_b200", "libgmd_host.so"))
    for n in names:
        assert hasattr(lib, n), n
    # error behaviour of the plugin selector (src/dycore_test.F90:29-42: unknown test case -> log_error)
    lib.gmd_host_last_error.restype = ctypes.c_char_p
    z = (ctypes.c_double * 16)()
    assert lib.gmd_host_initial_condition(b"no_such_case", 4, 5, z, z, z, z) == 2
    assert b"no_such_case" in lib.gmd_host_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gmd.GmdError) as ei:
        gmd.Dycore(gmd.Config(num_lon=36, num_lat=19, time_step_size=600.0))
    assert ei.value.code == gmd.ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_config_defaults_match_params_mod():
    lib = gmd.load()
    c = gmd._Cfg()
    lib.gmd_config_defaults(ctypes.byref(c))
    # src/params_mod.F90:15,55-56,59,65-66
    assert (c.subcycles, c.uv_adv_upwind_lon_beta, c.uv_adv_upwind_lat_beta) == (4, 0.0, 0.5)
    assert (c.use_zonal_tend_filter, c.diffusion_order, c.use_diffusion) == (1, 2, 0)
    assert list(c.cutoff) == [0] * 20 and (c.rank, c.nranks, c.device) == (0, 1, -1)


def _c_prototypes(hdr):
    """name -> (return type, [argument types]) of every gmd_* prototype of a header, comments stripped"""
    code = re.sub(r"^\s*#.*$", "", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S), flags=re.M)
    out = {}
    for ret, name, args in re.findall(r"([A-Za-z_][\w\s\*]*?)\b(gmd_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", code):
        args = [a.strip() for a in args.split(",")]
        args = [] if args == ["void"] else [re.sub(r"\s*\b\w+$", "", a).replace("const ", "").replace(" ", "") for a in args]
        out[name] = (ret.replace("const ", "").replace(" ", ""), args)
    return out


def _fortran_interfaces(src):
    """name -> (return kind, [argument kinds]) of every bind(c) interface of fortran/gmd_c.F90, kinds in C spelling"""
    src = re.sub(r"!.*", "", src)
    out = {}
    pat = re.compile(r"^\s*(subroutine|integer\(c_int\) function|type\(c_ptr\) function)\s+(\w+)\s*\(([^)]*)\)\s*"
                     r"bind\(c, name='(\w+)'\)(.*?)^\s*end (?:subroutine|function)", re.S | re.M)
    for kind, fname, args, cname, body in pat.findall(src):
        assert fname == cname
        args = [a.strip() for a in args.split(",") if a.strip()]
        decl = {}
        for line in body.splitlines():
            m = re.match(r"\s*(type\(c_ptr\)|type\(gmd_config\)|integer\(c_int\)|real\(c_double\)|character\(kind=c_char\))"
                         r"((?:,\s*\w+(?:\(\w+\))?)*)\s*::\s*(.*)", line)
            if not m:
                continue
            base, attrs, names = m.groups()
            by_value = "value" in attrs
            for n in re.findall(r"(\w+)(?:\([^)]*\))?", names):
                if base == "type(c_ptr)":
                    decl[n] = "ptr" if by_value else "ptr*"
                elif base == "type(gmd_config)":
                    decl[n] = "gmd_config*"
                elif base == "integer(c_int)":
                    decl[n] = "int" if by_value else "int*"
                elif base == "real(c_double)":
                    decl[n] = "double" if by_value else "double*"
                else:
                    decl[n] = "ptr"            # character buffers travel as void*
        out[cname] = ({"subroutine": "void", "integer(c_int) function": "int", "type(c_ptr) function": "ptr"}[kind],
                      [decl[a] for a in args])
    return out


def test_fortran_shim_matches_the_c_header():
    """fortran/gmd_c.F90 cannot be compiled in this image (no Fortran compiler), so its ISO_C_BINDING interfaces are
    checked textually against include/gmd.h: struct gmd_config field by field (order, type, extent -- bind(c) derived
    types follow the C layout rules), every enum constant it mirrors, and every bound procedure's return and argument
    kinds; and the replacement dycore_mod (fortran/dycore_mod.F90) only calls what gmd_c declares."""
    hdr = open(os.path.join(ROOT, "include", "gmd.h")).read()
    f90 = open(os.path.join(ROOT, "fortran", "gmd_c.F90")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    # struct
    body = re.search(r"typedef struct gmd_config \{(.*?)\} gmd_config;", code, re.S).group(1)
    c_fields = [(t, n, int(d) if d else 1) for t, n, d in re.findall(r"\b(int|double)\s+(\w+)(?:\[(\d+)\])?\s*;", body)]
    fbody = re.search(r"type, bind\(c\) :: gmd_config(.*?)end type gmd_config", f90, re.S).group(1)
    f_fields = [({"integer(c_int)": "int", "real(c_double)": "double"}[t], n, int(d) if d else 1)
                for t, n, d in re.findall(r"^\s*(integer\(c_int\)|real\(c_double\))\s+(\w+)(?:\((\d+)\))?\s*$", fbody, re.M)]
    assert len(c_fields) >= 24 and f_fields == c_fields
    # the ctypes mirror of the Python interface is the third copy of the same struct
    py = [(("int" if f[1] in (ctypes.c_int, ctypes.c_int * 20) else "double"), 20 if f[1] is ctypes.c_int * 20 else 1)
          for f in gmd._Cfg._fields_]
    assert py == [(t, d) for t, _, d in c_fields]
    assert ctypes.sizeof(gmd._Cfg) == sum((4 if t == "int" else 8) * d for t, _, d in c_fields) + _c_padding(c_fields)
    # enum constants
    c_enum = {}
    for grp in re.findall(r"enum \{(.*?)\}", code, re.S):
        for n, v in re.findall(r"(GMD_\w+)\s*=\s*(\d+)", grp):
            c_enum[n] = int(v)
    f_par = {n: int(v) for line in re.findall(r"integer\(c_int\), parameter ::(.*)", f90)
             for n, v in re.findall(r"(GMD_\w+)\s*=\s*(\d+)", line)}
    assert len(f_par) >= 13 and all(c_enum[n] == v for n, v in f_par.items())
    # procedures
    protos = _c_prototypes(hdr)
    ifaces = _fortran_interfaces(f90)
    assert len(ifaces) >= 17
    norm = lambda t: {"gmd_model*": "ptr", "constgmd_model*": "ptr", "void*": "ptr", "char*": "ptr",
                      "gmd_model**": "ptr*"}.get(t, t)
    for name, (fret, fargs) in ifaces.items():
        cret, cargs = protos[name]
        cargs = [norm(a) for a in cargs]
        # host arrays are handed over as c_loc(...) pointers (type(c_ptr), value) or as assumed-size / scalar dummies
        # passed by reference: both are a C double* / int*
        fargs_n = [("double*" if (a == "ptr" and c == "double*") else a) for a, c in zip(fargs, cargs)]
        assert norm(cret) == fret, name
        assert len(fargs) == len(cargs) and fargs_n == cargs, (name, fargs, cargs)
    # the replacement dycore_mod
    dm = re.sub(r"!.*", "", open(os.path.join(ROOT, "fortran", "dycore_mod.F90")).read())
    called = set(re.findall(r"\b(gmd_[a-z_]+)\s*\(", dm))
    assert called and called <= set(ifaces) | {"gmd_error_message"}, called - set(ifaces)
    for sub in ("dycore_init", "dycore_restart", "dycore_run", "dycore_final"):      # src/dycore_mod.F90:22-25
        assert re.search(r"subroutine %s\b" % sub, dm), sub


def _c_padding(fields):
    """bytes of alignment padding of a C struct of int / double members (natural alignment, x86-64 / aarch64)"""
    off = pad = 0
    for t, _, d in fields:
        a = 4 if t == "int" else 8
        if off % a:
            pad += a - off % a
            off += a - off % a
        off += a * d
    if off % 8:
        pad += 8 - off % 8
    return pad


@pytest.mark.parametrize("header", ["gmd.h", "gmd_host.h"])
def test_headers_are_plain_c(header, tmp_path):
    """the boundary is a C ABI: both headers compile as C99 on their own (no C++, CUDA or torch types)"""
    src = write(tmp_path, '#include "%s"\nint main(void) { return 0; }\n' % header, "t.c")
    res = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only",
                          "-I", os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


# ------------------------------------------------------------------------------------------------ namelist
def parse(path):
    rc, out = selftest("parse", path)
    return rc, dict(l.split("=", 1) for l in out.strip().splitlines() if "=" in l), out


@pytest.mark.parametrize("name,expect", [
    ("namelist.rh_test", dict(test_case="rossby_haurwitz_wave", num_lon="360", num_lat="181", time_step_size="240", subcycles="6",
                              split_scheme="csp2", cutoff="4,4,4,4,4" + ",0" * 15)),
    ("namelist.mz_test", dict(test_case="mountain_zonal_flow", num_lon="180", num_lat="90", uv_adv_scheme="upwind",
                              uv_adv_upwind_lat_beta="0.10000000000000001", subcycles="8", restart_period="10 days")),
    ("namelist.jz_test", dict(test_case="jet_zonal_flow", use_diffusion="1", cutoff=",".join(["4"] * 20), run_hours="6")),
    ("namelist.sg_test", dict(test_case="steady_geostrophic_flow", num_lon="3600", num_lat="1801")),
])
def test_shipped_namelists_parse(name, expect):
    rc, kv, out = parse(os.path.join(ROOT, "run", name))
    assert rc == 0, out
    for k, v in expect.items():
        assert kv[k] == v, (k, kv[k])


def test_namelist_defaults_comments_and_unknown_keys(tmp_path):
    text = """! leading comment
&dycore_params
 test_case = 'rossby_haurwitz_wave', case_name='x' ! trailing comment
 num_lon = 36
 num_lat = 19, time_step_size = 1.5d2
 time_scheme = "predict_correct"
 split_scheme = 'none'
 qcon_modified = .false.
 history_periods = '12 hours'
/
"""
    rc, kv, out = parse(write(tmp_path, text))
    assert rc == 0, out
    # defaults of src/params_mod.F90:13-66; restart_period falls back to history_periods (:113)
    assert kv["subcycles"] == "4" and kv["uv_adv_scheme"] == "center_diff" and kv["uv_adv_upwind_lat_beta"] == "0.5"
    assert kv["use_zonal_tend_filter"] == "1" and kv["time_step_size"] == "150" and kv["restart_period"] == "12 hours"
    # the reference's own rh/jz namelists use keys this commit does not define: a namelist read aborts
    bad = text.replace(" num_lon = 36", " days = 3\n num_lon = 36")
    rc, _, out = parse(write(tmp_path, bad, "bad"))
    assert rc == 2 and "Cannot match namelist object name days" in out


def test_namelist_subscripts_and_steps_period(tmp_path):
    """name(k) = ... assigns from element k on, as gfortran does (a round-1 build ignored the subscript); the IO period
    unit 'steps' of src/io_mod.F90:209-210"""
    text = """&dycore_params
 test_case = 'rossby_haurwitz_wave', case_name='x'
 num_lon = 36, num_lat = 19, time_step_size = 150
 zonal_tend_filter_cutoff_wavenumber = 4, 4
 zonal_tend_filter_cutoff_wavenumber(3) = 3, 2
 zonal_tend_filter_cutoff_wavenumber(7) = 1
 history_periods = '8 steps'
/
"""
    rc, kv, out = parse(write(tmp_path, text))
    assert rc == 0, out
    assert kv["cutoff"] == "4,4,3,2,0,0,1" + ",0" * 13
    rc, out = selftest("clock", write(tmp_path, text.replace("/\n", " run_hours = 1\n/\n"), "clk"), "20")
    assert rc == 0, out   # one line per step: the history alert rings every 8 steps = 1200 s
    lines = out.strip().splitlines()
    rings = [k for k, l in enumerate(lines[1:], 1) if l.split()[1] == "1"]
    assert rings and rings[0] == 8
    bad = text.replace("(3) = 3, 2", "(3:4) = 3, 2")
    rc, _, out = parse(write(tmp_path, bad, "bad"))
    assert rc == 2 and "unsupported subscript" in out


def test_reduced_tendency_and_runge_kutta_keys_parse(tmp_path):
    """the keys of the reference's own run/namelist.jz_test:12-19 that its params_mod does not define (use_zonal_reduce,
    reduce_adv_lon, zonal_reduce_factors, use_reduce_tend_smooth) and time_scheme = 'runge_kutta' / time_order
    (params_mod.F90:40-45): the specified extensions of DESIGN.md section 8"""
    text = """&dycore_params
test_case = 'jet_zonal_flow'
case_name = 'jz_test.360x181.pc.dt240.diffused'
run_days = 6
num_lon = 360
num_lat = 181
time_step_size = 240
time_scheme = 'runge_kutta'
time_order = 4
qcon_modified = .true.
split_scheme = 'csp2'
subcycles = 6
use_zonal_reduce = .true.
reduce_adv_lon = .true.
zonal_reduce_factors = 8, 8, 4, 2, 2
use_reduce_tend_smooth = .true.
use_diffusion = .true.
diffusion_coef = 1.0e5
/
"""
    rc, kv, out = parse(write(tmp_path, text))
    assert rc == 0, out
    assert kv["time_scheme"] == "runge_kutta" and kv["time_order"] == "4"
    assert kv["use_zonal_reduce"] == "1" and kv["reduce_adv_lon"] == "1" and kv["use_reduce_tend_smooth"] == "1"
    assert kv["reduce_factors"] == "8,8,4,2,2" + ",0" * 15


# ------------------------------------------------------------------------------------------------ mesh tables
@pytest.mark.parametrize("nlon,nlat,cut,use_filter", [
    (72, 37, [5, 4, 3], True), (360, 181, [4] * 5, True), (180, 90, [4, 4, 4], True), (3600, 1801, [4] * 20, True),
    (7200, 3601, [4] * 20, True), (96, 49, [3, 0, 5], False),
])
def test_product_mesh_tables_are_the_oracles_bit_for_bit(tmp_path, nlon, nlat, cut, use_filter):
    """csrc/gmd_mesh.h (what gmd_create builds on the host and copies to the device: mesh_init, data_init,
    reset_cos_lat_at_poles, filter_init -- src/mesh_mod.F90:44-114, data_mod.F90:26-47, dycore_mod.F90:159-173,
    filter_mod.F90:35-103) against the oracle's tables and filter row map, bit for bit, without a GPU (the same check
    through gmd_get_table is tests/test_gpu_parity.py::test_tables_and_filter_rows_bit_identical)."""
    nml = write(tmp_path, f"""&dycore_params
 test_case = 'rossby_haurwitz_wave', num_lon = {nlon}, num_lat = {nlat}, time_step_size = 60
 use_zonal_tend_filter = {'.true.' if use_filter else '.false.'}
 zonal_tend_filter_cutoff_wavenumber = {', '.join(map(str, cut))}
/
""")
    out = str(tmp_path / "tables.bin")
    rc, txt = selftest("tables", nml, out)
    assert rc == 0, txt
    assert txt.split() == [str(nlon), str(nlat), str(max(cut))]
    raw = open(out, "rb").read()
    o = Oracle(OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=60.0, use_zonal_tend_filter=use_filter,
                            zonal_tend_filter_cutoff_wavenumber=cut))
    o.set_initial_condition("steady_geostrophic_flow")
    o.run_init()                                   # reset_cos_lat_at_poles happens at the head of dycore_run
    off = 0
    for which in range(10):
        n = nlat - 1 if which in (1, 5, 7, 9) else nlat
        got = np.frombuffer(raw, dtype=np.float64, count=n, offset=off)
        off += 8 * n
        assert np.array_equal(got, o.table(which)), which
    maps = [np.frombuffer(raw, dtype=np.int32, count=nlat, offset=off + 4 * nlat * k) for k in range(6)]
    assert len(raw) == off + 6 * 4 * nlat
    ff, fc, hf, hc = o.filter_rows()
    assert np.array_equal(maps[0], ff) and np.array_equal(maps[1], fc)
    assert np.array_equal(maps[2][:nlat - 1], hf) and np.array_equal(maps[3][:nlat - 1], hc)
    assert not maps[4].any() and not maps[5].any()        # no reduced rows configured
    assert ff.sum() == (2 * sum(1 for c in cut if c) if use_filter else 0)


# ------------------------------------------------------------------------------------------------ formats
@pytest.mark.parametrize("x,s", [(2.812376625197987e19, "0.28123766251980E+20"), (5.083437998896658e12, "0.50834379988967E+13"),
                                 (1.0000123, "0.10000123000000E+01"), (-0.5, "-0.5000000000000E+00"), (0.0, "0.00000000000000E+00"),
                                 (9.9999999999999e-5, "0.99999999999999E-04"), (9.99999999999999e-5, "0.10000000000000E-03")])
def test_to_string_real8_is_fortran_E20_14(x, s):
    rc, out = selftest("fmt", repr(x))
    assert rc == 0 and out.strip() == s


def test_clock_and_history_alert(tmp_path):
    nml = os.path.join(ROOT, "run", "namelist.rh_test")   # dt 240 s, history every 6 hours, 1 day
    rc, out = selftest("clock", nml, 95)
    lines = out.strip().splitlines()
    assert lines[0] == "0001-01-01T00:00:00Z 0 90 360"      # no frame at t = 0 (src/time_mod.F90:192-213)
    assert lines[1].split()[0] == "0001-01-01T00:04:00Z"
    rings = [k for k, l in enumerate(lines[1:], 1) if l.split()[1] == "1"]
    assert rings[0] == 90 and lines[90].split()[0] == "0001-01-01T06:00:00Z"
    assert rings == [90]                                     # rings once, re-armed by the next time_advance


# ------------------------------------------------------------------------------------------------ IC plugins
@pytest.mark.parametrize("tc,extra,tol", [("rossby_haurwitz_wave", "", 1e-15), ("steady_geostrophic_flow", "", 1e-15),
                                          ("mountain_zonal_flow", "", 1e-15), ("jet_zonal_flow", "", 1e-9),
                                          ("mountain_zonal_flow", "&mountain_zonal_flow_test_params\n smooth_mountain = .true.\n/\n", 1e-14)])
def test_ic_plugins_match_oracle(tmp_path, tc, extra, tol):
    nlon, nlat = 72, 37
    text = f"&dycore_params\n test_case='{tc}'\n case_name='t'\n num_lon={nlon}\n num_lat={nlat}\n time_step_size=100\n/\n" + extra
    out_bin = str(tmp_path / "ic.bin")
    rc, out = selftest("ic", write(tmp_path, text), out_bin)
    assert rc == 0 and out.startswith("Use "), out
    a = np.fromfile(out_bin)
    nf, nh = nlon * nlat, nlon * (nlat - 1)
    u, v, gd, ghs = a[:nf].reshape(nlat, nlon), a[nf:nf + nh].reshape(nlat - 1, nlon), a[nf + nh:2 * nf + nh].reshape(nlat, nlon), a[2 * nf + nh:].reshape(nlat, nlon)
    o = Oracle(OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=100.0))
    o.set_initial_condition(tc, params=[1.0] if extra else None)
    uo, vo, gdo = o.state()
    for x, y in ((u, uo), (v, vo), (gd, gdo), (ghs, o.ghs())):
        assert np.abs(x - y).max() <= tol * max(np.abs(y).max(), 1.0)
    assert not u[0].any() and not u[-1].any()   # u = 0 on the pole rows: the invariant gmd_set_state checks


def test_shallow_water_waves_plugin_matches_oracle(tmp_path):
    """shallow_water_waves_test_mod.F90 (Shamir-Paldor Rossby wave): the host plugin against the oracle, and the phase
    speed notice the reference logs (:109-110)."""
    nlon, nlat = 72, 37
    text = f"&dycore_params\n test_case='shallow_water_waves'\n case_name='t'\n num_lon={nlon}\n num_lat={nlat}\n time_step_size=100\n/\n"
    out_bin = str(tmp_path / "ic.bin")
    rc, out = selftest("ic", write(tmp_path, text), out_bin)
    assert rc == 0 and out.startswith("Use shallow water waves initial condition."), out
    assert "Phase speed is -" in out
    a = np.fromfile(out_bin)
    nf, nh = nlon * nlat, nlon * (nlat - 1)
    u, v, gd, ghs = a[:nf].reshape(nlat, nlon), a[nf:nf + nh].reshape(nlat - 1, nlon), a[nf + nh:2 * nf + nh].reshape(nlat, nlon), a[2 * nf + nh:]
    o = Oracle(OracleConfig(num_lon=nlon, num_lat=nlat, time_step_size=100.0))
    o.set_initial_condition("shallow_water_waves")
    for x, y in zip((u, v, gd), o.state()):
        assert np.abs(x - y).max() <= 1e-13 * np.abs(y).max()
    assert not ghs.any()
    # pole rows: cos(+-pi/2)^(sigma - 1.5) -- not 0, but below GMD_POLE_U_TINY (include/gmd.h)
    assert 0 < np.abs(u[0]).max() < 1e-100 and 0 < np.abs(u[-1]).max() < 1e-100


def test_unknown_test_case_message(tmp_path):
    text = "&dycore_params\n test_case='nope'\n case_name='t'\n num_lon=36\n num_lat=19\n time_step_size=100\n/\n"
    rc, out = selftest("ic", write(tmp_path, text), str(tmp_path / "x"))
    assert rc == 2 and "Unknown test case nope!" in out   # src/dycore_test.F90:41


# ------------------------------------------------------------------------------------------------ restart file
def test_restart_file_schema_and_round_trip(tmp_path):
    """restart_write / restart_read (src/restart_mod.F90:22-75): schema of the reference, interior data, bit-exact
    round trip through the host's own classic-netCDF reader"""
    from scipy.io import netcdf_file
    nml = os.path.join(ROOT, "run", "namelist.mz_test")
    rc, out = selftest("restart", nml, 120, cwd=str(tmp_path))
    assert rc == 0, out
    name, when, verdict = out.strip().splitlines()
    assert name == "mz_c_u_01.180x90.dt720.r.0001-01-02T00:00:00Z.nc" and when == "0001-01-02T00:00:00Z" and verdict == "identical"
    f = netcdf_file(str(tmp_path / name), "r", mmap=False)
    assert list(f.dimensions.items()) == [("time", None), ("lon", 180), ("lat", 90), ("ilon", 180), ("ilat", 89)]
    assert list(f.variables) == ["time", "lon", "lat", "ilon", "ilat", "u", "v", "gd", "ghs"]
    assert list(f._attributes) == ["dataset", "desc", "author", "restart_time", "elapsed_seconds"]
    assert f.dataset.decode().rstrip() == "restart" and f.desc.decode().rstrip() == "Restart file"
    assert f.restart_time.decode().rstrip() == "0001-01-02T00:00:00Z" and f.elapsed_seconds == 86400.0
    assert f.variables["u"].dimensions == ("time", "lat", "ilon") and f.variables["v"].dimensions == ("time", "ilat", "lon")
    assert f.variables["gd"].long_name == b"geopotential depth" and f.variables["gd"].units == b"m2 s-2"
    o = Oracle(OracleConfig(num_lon=180, num_lat=90, time_step_size=720.0))
    o.set_initial_condition("mountain_zonal_flow")
    u, v, gd = o.state()
    assert np.allclose(f.variables["u"][0], u, rtol=0, atol=1e-12) and np.allclose(f.variables["gd"][0], gd, rtol=1e-15)
    assert np.array_equal(f.variables["ghs"][0], o.ghs())


# ------------------------------------------------------------------------------------------------ history file
def test_history_file_schema(tmp_path):
    from scipy.io import netcdf_file
    nml = os.path.join(ROOT, "run", "namelist.mz_test")
    rc, out = selftest("history", nml, 120, cwd=str(tmp_path))
    assert rc == 0, out
    name = out.strip()
    assert name == "mz_c_u_01.180x90.dt720.h0.0001-01-02T00:00:00Z.nc"    # src/io_mod.F90:417-419, time_mod.F90:128
    raw = open(tmp_path / name, "rb").read(4)
    assert raw == b"CDF\x01"                                                 # classic format (NF90_CLOBBER)
    f = netcdf_file(str(tmp_path / name), "r", mmap=False)
    assert list(f.dimensions.items()) == [("time", None), ("lon", 180), ("lat", 90), ("ilon", 180), ("ilat", 89)]
    assert list(f.variables) == ["time", "lon", "lat", "ilon", "ilat", "u", "v", "gh", "ghs", "vor", "div", "te", "tm"]
    attrs = list(f._attributes)
    assert attrs == ["dataset", "desc", "author", "time_step_size", "time_scheme", "split_scheme", "subcycles"]
    assert f.dataset.decode().rstrip() == "hist0" and len(f.desc) == 256 and f.author.decode().rstrip() == "N/A"
    assert f.time_step_size == 720.0 and f.subcycles == 8 and f.split_scheme.decode().rstrip() == "csp2"
    for v in f.variables.values():
        assert v.data.dtype == np.dtype(">f8")
    assert f.variables["u"].dimensions == ("time", "lat", "lon") and f.variables["vor"].dimensions == ("time", "lat", "ilon")
    assert f.variables["div"].dimensions == ("time", "ilat", "lon") and f.variables["te"].dimensions == ("time",)
    assert f.variables["time"].units == b"days since 0001-01-01T00_00_00" and f.variables["time"][0] == 1.0
    assert f.variables["u"].long_name == b"u wind component" and f.variables["gh"].units == b"m2 s-2"
    assert f.variables["lat"][0] == -90.0 and f.variables["lat"][-1] == 90.0 and abs(f.variables["ilon"][0] - 1.0) < 1e-12
    # A-grid averaging of the C-grid winds (src/history_mod.F90:102-108)
    o = Oracle(OracleConfig(num_lon=180, num_lat=90, time_step_size=720.0))
    o.set_initial_condition("mountain_zonal_flow")
    u, v, gd = o.state()
    ua = 0.5 * (u + np.roll(u, 1, axis=1))
    assert np.allclose(f.variables["u"][0], ua, rtol=0, atol=1e-12)
    assert np.allclose(f.variables["gh"][0], gd + o.ghs(), rtol=1e-15)
    assert f.variables["te"][0] == 2.5 and f.variables["tm"][0] == 1.5
    assert not f.variables["vor"][0][-1].any()


# ------------------------------------------------------------------------------------------------ decomposition
def test_band_arithmetic():
    from gamil_dycore_b200 import parallel
    for nlat, n in ((1801, 8), (3601, 8), (181, 2), (90, 4), (721, 3)):
        rows = [parallel.band(r, n, nlat) for r in range(n)]
        assert rows[0][0] == 0 and rows[-1][1] == nlat
        assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        sizes = [b - a for a, b in rows]
        assert max(sizes) - min(sizes) <= 1
    assert parallel.halo_rows(0, 2, 181) == ([], [91, 92]) and parallel.halo_rows(1, 2, 181) == ([90], [])
    # shorter polar bands (gmd_config.polar_band_rows): contiguous cover, the middle ranks even
    for nlat, n, pbr in ((1801, 8, 113), (1801, 4, 350), (3601, 8, 300), (181, 3, 40)):
        rows = [parallel.band(r, n, nlat, pbr) for r in range(n)]
        assert rows[0] == (0, pbr) and rows[-1] == (nlat - pbr, nlat)
        assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        mid = [b - a for a, b in rows[1:-1]]
        assert max(mid) - min(mid) <= 1 and sum(mid) == nlat - 2 * pbr
    assert parallel.band(1, 2, 181, 40) == parallel.band(1, 2, 181)      # ignored below 3 ranks
    assert parallel.polar_band_rows_for(2, 3600, 1801) == 0
    assert 300 <= parallel.polar_band_rows_for(4, 3600, 1801) <= 450 and 48 <= parallel.polar_band_rows_for(8, 3600, 1801) < 112


def _gloo_worker(rank, world, port, nlat, nlon, q):
    import torch.distributed as dist
    from gamil_dycore_b200 import parallel
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    full, half = rng.standard_normal((nlat, nlon)), rng.standard_normal((nlat - 1, nlon))
    r0, r1 = parallel.band(rank, world, nlat)
    lf, lh = np.zeros_like(full), np.zeros_like(half)
    lf[r0:r1] = full[r0:r1]
    lh[r0:min(r1, nlat - 1)] = half[r0:min(r1, nlat - 1)]
    gf = parallel.gather_field(lf, nlat)
    gh = parallel.gather_field(lh, nlat)
    q.put((rank, bool(np.array_equal(gf, full) and np.array_equal(gh, half))))
    dist.destroy_process_group()


def test_gather_bands_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, 29577, 37, 12, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


class _FakeDycore:
    """stands in for gamil_dycore_b200.Dycore in the wiring test below: records the calls parallel.connect makes"""

    def __init__(self, rank, fail_connect):
        self.rank, self.fail_connect, self.calls = rank, fail_connect, []

    def peer_export(self):
        self.calls.append("export")
        return bytes([self.rank]) * 256

    def peer_connect(self, blobs):
        self.calls.append(("connect", [b[0] for b in blobs]))
        if self.fail_connect:
            raise RuntimeError("cudaIpcOpenMemHandle failed (simulated)")

    def peer_disconnect(self):
        self.calls.append("disconnect")

    def comm_init(self, uid):
        self.calls.append(("comm_init", len(uid)))


def _connect_worker(rank, world, port, fail_rank, q):
    import torch.distributed as dist
    import gamil_dycore_b200 as pkg
    from gamil_dycore_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg.comm_unique_id = lambda kind="fast": b"\x07" * 128     # no NCCL on the CPU box
    d = _FakeDycore(rank, fail_connect=(rank == fail_rank))
    mode = parallel.connect(d)
    strict = None
    if fail_rank >= 0:
        try:
            parallel.connect(_FakeDycore(rank, fail_connect=(rank == fail_rank)), fallback=False)
        except RuntimeError as e:
            strict = "gmd_peer_connect failed" in str(e)
    q.put((rank, mode, d.calls, strict))
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_connect_wiring_and_nccl_fallback_over_gloo_world_size_2(fail_rank):
    """parallel.connect: blobs gathered in rank order; if the peer mapping fails on ANY rank, EVERY rank disconnects and
    falls back to the NCCL bootstrap (or raises with fallback=False)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_connect_worker, args=(r, 2, 29579 + fail_rank, fail_rank, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, mode, calls, strict in res:
        assert calls[0] == "export" and calls[1] == ("connect", [0, 1])
        if fail_rank < 0:
            assert mode == "peer" and len(calls) == 2
        else:
            assert mode == "nccl" and calls[2] == "disconnect" and calls[3] == ("comm_init", 128) and strict is True


def test_pc_ring_schedule():
    """The fused predict_correct kernel (gamil_dycore_b200/csrc/gmd_pc.cuh) passes rows between its three warps through
    shared-memory rings, in lockstep ticks.  Replay its schedule with the depths and lags the header defines: every
    read must see the row written in an EARLIER tick (a barrier in between), and nothing may be overwritten in the tick
    it is read or before."""
    import re
    src = open(os.path.join(ROOT, "gamil_dycore_b200", "csrc", "gmd_pc.cuh")).read()

    def const(name):
        m = re.search(r"constexpr int [^;]*\b" + name + r" = (\d+)", src)
        assert m, name
        return int(m.group(1))
    D1, D2, DAB, DOP, IND, REC = (const(n) for n in ("PC_D1", "PC_D2", "PC_DEPTH_AB", "PC_DEPTH_OP", "PC_IN_DEPTH", "PC_REC_DEPTH"))
    depth = {"A": DAB, "B": DAB, "O": DOP, "P": DOP}
    # five CTAs per SM (GMD_PC_MINB): rings + input ring (7 lines per packet) + row records + barriers, 1 KB reserved per
    # CTA by the hardware, ~0.5 KB static (reduction scratch of the fold) -- within B200's 228 KB per SM
    smem = (3 * (2 * DAB + 2 * DOP) + IND * 7) * 512 + REC * 16 * 8 + IND * 8
    assert 5 * (smem + 1024 + 512) <= 228 * 1024, smem
    for n in range(1, 70):                      # rows of a chunk
        ja, jb = 0, n
        rja0, rjb0 = ja - 2, jb + 4
        ring = {k: {} for k in "AOBP"}          # slot -> (row, tick written)
        inr, rec = {}, {}                       # input ring: slot -> (packet, tick issued); records: slot -> (row, tick)
        def issue(it, t):
            if rja0 + it >= rjb0:
                return
            inr[it % IND] = (it, t)
            for y in ([rja0 - 1, rja0, rja0 + 1] if it == 0 else [rja0 + it + 1]):
                rec[(y - (ja - 3)) % REC] = (y, t)
        for q in range(IND - 1):
            issue(q, -1)
        reads = []
        def rd(k, x, t):
            e = ring[k].get((x - (ja - 2)) % depth[k])
            assert e is not None and e[0] == x and e[1] < t, (n, t, k, x, e)
            reads.append((t, k, (x - (ja - 2)) % depth[k]))
        def rdrec(x, t):
            for y in (x - 1, x, x + 1):
                e = rec.get((y - (ja - 3)) % REC)
                assert e is not None and e[0] == y and e[1] < t, (n, t, "rec", y, e)
        for t in range(n + D2):
            writes, issues = [], []
            x = rja0 + t                        # S1
            if x < rjb0:
                e = inr.get(t % IND)
                assert e is not None and e[0] == t and e[1] < t, (n, t, "packet", e)
                rdrec(x, t)
                writes += [("A", x), ("O", x)]
            if t < D1 or ja - 1 + (t - D1) < jb + 2:   # S2's warp is in its idle prefix or its row loop: it issues
                issues.append(t + IND - 1)
            if t >= D1:                         # S2
                x = ja - 1 + (t - D1)
                if x < jb + 2:
                    if t == D1:
                        for y in (x - 1, x, x + 1):
                            rd("A", y, t)
                    rd("A", x + 2, t), rd("A", x + 1, t), rd("O", x, t)
                    rdrec(x, t)
                    writes += [("B", x), ("P", x)]
            if t >= D2:                         # S3a
                x = ja + (t - D2)
                if x < jb:
                    if t == D2:
                        for y in (x - 1, x, x + 1):
                            rd("B", y, t)
                    rd("B", x + 2, t), rd("B", x + 1, t), rd("P", x, t)
                    rdrec(x, t)
            # a packet issued in tick t lands in the slot S1 read in tick t-1; S1 reads slot t % IND in tick t
            for it in issues:
                assert it % IND != t % IND
                issue(it, t)
            for k, x in writes:
                assert (t, k, (x - (ja - 2)) % depth[k]) not in reads, (n, t, k, x)   # no write into a slot read this tick
                ring[k][(x - (ja - 2)) % depth[k]] = (x, t)
