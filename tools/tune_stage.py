"""Times every stage-kernel instantiation of one or more builds of libgmd (GPU box).

    python tools/tune_stage.py [lib.so ...]      # default: gamil_dycore_b200/libgmd.so
Prints ms / launch and achieved algorithmic GB/s for fast S1/S2/S3a, slow S1/S2/S3a on the 0.1 degree grid."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gamil_dycore_b200 as gmd  # noqa: E402
import gamil_dycore_b200 as gmd_pkg  # noqa: E402

nlon, nlat = int(os.environ.get("NLON", 3600)), int(os.environ.get("NLAT", 1801))
u, v, gd, ghs = gmd_pkg.initial_condition("steady_geostrophic_flow", nlon, nlat)
libs = sys.argv[1:] or ["fast"]
for lib in libs:
    kind = lib if lib in ("fast", "strict") else os.path.abspath(lib)
    d = gmd.Dycore(gmd.Config(num_lon=nlon, num_lat=nlat, time_step_size=10.0, subcycles=10,
                              zonal_tend_filter_cutoff_wavenumber=[4] * 20), kind=kind)
    d.set_state(u, v, gd, ghs)
    d.run_init()
    out = {}
    for p in ("fast", "slow"):
        for mode, nm in ((0, "S1"), (1, "S2"), (2, "S3a"), (4, "S1lazy")):
            d.time_stage_variant(p, mode, 3)
            ms, nb = d.time_stage_variant(p, mode, 20)
            out[f"{p}.{nm}"] = (round(ms * 1e3, 1), round(nb / ms / 1e6))
        try:   # the fused predict_correct kernel (13 / 11 words per column), where the configuration uses it
            d.time_stage_variant(p, 5, 3)
            ms, nb = d.time_stage_variant(p, 5, 20)
            out[f"{p}.k_pc"] = (round(ms * 1e3, 1), round(nb / ms / 1e6))
        except gmd.GmdError:
            pass
    d.step(3)
    d.step(10)
    print(os.path.basename(lib), json.dumps(out), "step_ms", round(d.last_step_ms() / 10, 3), flush=True)
    d.close()
