"""Row-pair diffusion / derive kernels (k_derive2, k_laplace2, k_diff_update2; GMD_EW_ROWS=1) against the
element-indexed ones (GMD_EW_ROWS=0) on one GPU: same state after a few steps, and the device time per step of each.

    python tools/check_ew_rows.py [out.json]

Cases: the 0.25 degree Galewsky jet with filter + order-2 diffusion (the C3 workload of bench.py), an order-4 upwind
mountain flow at 1 degree, and the committed diffusion / WENO fixtures (tests/golden) -- those against the oracle's
stored answer too, with the tolerance of tests/test_gpu_parity.py::test_golden_cases.  Test tooling: uses oracle/ only
through the committed fixtures."""
import ast
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gamil_dycore_b200 as gmd   # noqa: E402


def rel(a, b):
    n = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / n) if n > 0 else float(np.abs(a).max())


def run(kw, ic, nsteps, ew_rows, graph=True):
    os.environ["GMD_EW_ROWS"] = str(ew_rows)
    d = gmd.Dycore(gmd.Config(**kw))
    d.set_graph_mode(graph)
    d.set_state(*ic)
    d.run_init()
    d.step(nsteps)                 # graphs captured, clocks up
    st = d.state()
    d.step(nsteps)
    ms = d.last_step_ms() / nsteps
    return st, d.state(), d.diag(), ms


def main():
    out = {"cases": []}
    ok = True
    # 1. analytic ICs through the product's own plugins
    big = [
        ("jz_0.25deg", "jet_zonal_flow", dict(num_lon=1440, num_lat=721, time_step_size=30.0, subcycles=6, split_scheme="csp2",
                                              zonal_tend_filter_cutoff_wavenumber=[4] * 20, use_diffusion=True, diffusion_coef=6.0e3), 10),
        ("mz_1deg_diff4", "mountain_zonal_flow", dict(num_lon=360, num_lat=181, time_step_size=240.0, subcycles=6, split_scheme="csp2",
                                                      uv_adv_scheme="upwind", zonal_tend_filter_cutoff_wavenumber=[4] * 5,
                                                      use_diffusion=True, diffusion_order=4, diffusion_coef=1.0e13), 10),
    ]
    for name, tc, kw, n in big:
        ic = gmd.initial_condition(tc, kw["num_lon"], kw["num_lat"])
        a0, a1, da, ms0 = run(kw, ic, n, 0)
        b0, b1, db, ms1 = run(kw, ic, n, 1)
        diffs = [rel(x, y) for x, y in zip(b1, a1)]
        bit = all(np.array_equal(x, y) for x, y in zip(b1, a1))
        rec = dict(case=name, steps=2 * n, rel_l2_rows_vs_elements=diffs, bit_identical=bit,
                   ms_per_step_elements=ms0, ms_per_step_rows=ms1, mass_rel=abs(db[0] / da[0] - 1), energy_rel=abs(db[1] / da[1] - 1))
        ok &= max(diffs) <= 1e-14 and np.isfinite(b1[2]).all()
        out["cases"].append(rec)
        print(json.dumps(rec), flush=True)
    # 2. committed fixtures: both forms against the oracle's stored answer
    gdir = os.path.join(ROOT, "tests", "golden")
    for name in ("jz_72x37_diffusion", "mz_60x32_diff4", "mz_48x25_weno"):
        g = np.load(os.path.join(gdir, f"case_{name}.npz"), allow_pickle=True)
        kw = ast.literal_eval(str(g["config"]))
        n = int(g["nsteps"])
        res = {}
        for ew in (0, 1):
            os.environ["GMD_EW_ROWS"] = str(ew)
            d = gmd.Dycore(gmd.Config(**kw))
            d.set_state(g["u0"], g["v0"], g["gd0"], g["ghs"])
            d.run_init()
            d.step(n)
            res[ew] = d.state()
        errs = {ew: [rel(x, g[k]) for x, k in zip(res[ew], ("u1", "v1", "gd1"))] for ew in (0, 1)}
        bit = all(np.array_equal(x, y) for x, y in zip(res[0], res[1]))
        rec = dict(case=f"golden:{name}", steps=n, rel_l2_vs_oracle_elements=errs[0], rel_l2_vs_oracle_rows=errs[1], bit_identical=bit)
        ok &= max(errs[1]) <= max(1e-12, 2.0 * max(errs[0]))
        out["cases"].append(rec)
        print(json.dumps(rec), flush=True)
    out["ok"] = bool(ok)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)
    print("OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    t = time.time()
    rc = main()
    print(f"{time.time() - t:.1f} s")
    sys.exit(rc)
