"""Fused (k_pc) vs three-sweep predict_correct on the same state: where do they differ?  (GPU; debugging aid)"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamil_dycore_b200 as gmd

def run(kw, ic, nsteps, fused, graph):
    os.environ["GMD_FUSED"] = "1" if fused else "0"
    d = gmd.Dycore(gmd.Config(**kw))
    u, v, gd, ghs = gmd.initial_condition(ic, kw["num_lon"], kw["num_lat"])
    d.set_state(u, v, gd, ghs)
    d.set_graph_mode(graph)
    d.run_init()
    out = []
    for n in nsteps:
        d.step(n)
        out.append([a.copy() for a in d.state()] + [d.diag()])
    return out

cases = {
  "c3": (dict(num_lon=1440, num_lat=721, time_step_size=30.0, subcycles=6, split_scheme="csp2",
              zonal_tend_filter_cutoff_wavenumber=[4] * 20, use_diffusion=True, diffusion_coef=6.0e3), "jet_zonal_flow"),
  "c3_nodiff": (dict(num_lon=1440, num_lat=721, time_step_size=30.0, subcycles=6, split_scheme="csp2",
              zonal_tend_filter_cutoff_wavenumber=[4] * 20), "jet_zonal_flow"),
  "c1": (dict(num_lon=360, num_lat=181, time_step_size=240.0, subcycles=6, split_scheme="csp2",
              zonal_tend_filter_cutoff_wavenumber=[4] * 5), "rossby_haurwitz_wave"),
}
which = sys.argv[1:] or list(cases)
for name in which:
    kw, ic = cases[name]
    for graph in (False, True):
        a = run(kw, ic, [1, 1, 3], True, graph)
        b = run(kw, ic, [1, 1, 3], False, graph)
        for k, (x, y) in enumerate(zip(a, b)):
            msg = []
            for f, p, q in zip("u v gd".split(), x[:3], y[:3]):
                diff = np.abs(p - q).max(axis=1)
                scale = np.abs(q).max() + 1e-300
                bad = np.nonzero(diff > 1e-12 * scale)[0]
                msg.append(f"{f}: max {diff.max() / scale:.2e} rows>1e-12: {bad[:8].tolist()}..{bad[-3:].tolist()} n={len(bad)}")
            print(name, "graph" if graph else "direct", "after", [1, 2, 5][k], "steps:", " | ".join(msg),
                  "mass rel", abs(x[3][0] / y[3][0] - 1), "energy rel", abs(x[3][1] / y[3][1] - 1), flush=True)
