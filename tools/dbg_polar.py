import ast, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamil_dycore_b200 as gmd
g = np.load("tests/golden/case_rh_36x19_csp2.npz", allow_pickle=True)
kw = ast.literal_eval(str(g["config"]))
for kind in ("fast",):
    d = gmd.Dycore(gmd.Config(**kw), kind=kind)
    d.set_graph_mode(False)
    d.set_state(g["u0"], g["v0"], g["gd0"], g["ghs"])
    d.run_init()
    print(kind, "diag0", d.diag())
    for p, dt in (("slow", 900.0), ("fast", 450.0), ("fast", 450.0), ("fast", 450.0), ("fast", 450.0), ("slow", 900.0)):
        try:
            d.predict_correct(dt, p)
        except Exception as e:
            print("exc", e)
        U, V, G = d.iap_state()
        bad = [np.argwhere(~np.isfinite(a)) for a in (U, V, G)]
        print(kind, p, "nan counts", [len(b) for b in bad], "rows", [sorted(set(b[:, 0].tolist())) for b in bad])
        if any(len(b) for b in bad):
            break
