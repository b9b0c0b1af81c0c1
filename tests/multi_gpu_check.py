"""Decomposition check, launched with torchrun (one rank per latitude band):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/multi_gpu_check.py [peer|nccl] [out.json]

Every rank builds its latitude band of the same global problem, the bands exchange halos inside libgmd (argument
"peer": NVLink peer memory, the default; "nccl": ncclSend/Recv), and the gathered result is compared
  (a) with the CPU oracle, beside the case's own rounding-noise floor (two oracle runs from a 1e-16-perturbed gd) --
      the same rule the single-GPU parity tests use: rel-L2 <= max(1e-12, 20 x floor), mass/energy <= 1e-13;
  (b) with the SAME GPU code run in one band (decomposition invariance):
      max-norm <= max(1e-13, 5 x floor) -- the bands differ from the single band only in the summation order of the
      two-scalar all-reduces.
Exit code 0 = pass.

With fewer GPUs than ranks (the driver's single-GPU test lease) the ranks SHARE device 0: CUDA IPC maps the
neighbour's slab across the two processes exactly as on two devices, the GPU time-slices between the two contexts
(a kernel spinning on a neighbour's flag is pre-empted at the end of its time slice), and torch.distributed runs on
gloo.  Slower per handshake, same code path, same results.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gamil_dycore_b200 as gmd  # noqa: E402
from gamil_dycore_b200 import parallel  # noqa: E402
from oracle.oracle import Oracle, OracleConfig  # noqa: E402


def rel(a, b, scale=None):
    return float(np.linalg.norm(a - b) / max(scale if scale is not None else np.linalg.norm(b), 1e-300))


CASES = [
    ("mountain_zonal_flow", dict(num_lon=120, num_lat=61, time_step_size=600.0, subcycles=4, split_scheme="csp2",
                                 uv_adv_scheme="upwind", uv_adv_upwind_lat_beta=0.1,
                                 zonal_tend_filter_cutoff_wavenumber=[4, 4, 4]), 4),
    ("jet_zonal_flow", dict(num_lon=144, num_lat=73, time_step_size=450.0, subcycles=6, split_scheme="csp2",
                            zonal_tend_filter_cutoff_wavenumber=[4] * 4, use_diffusion=True, diffusion_coef=1.0e5), 3),
    ("steady_geostrophic_flow", dict(num_lon=96, num_lat=49, time_step_size=600.0, subcycles=4, split_scheme="isp",
                                     zonal_tend_filter_cutoff_wavenumber=[3, 3]), 2),
    ("mountain_zonal_flow", dict(num_lon=96, num_lat=50, time_step_size=600.0, split_scheme="none",
                                 use_diffusion=True, diffusion_order=4, diffusion_coef=1.0e14,
                                 zonal_tend_filter_cutoff_wavenumber=[4, 4]), 3),
    ("mountain_zonal_flow", dict(num_lon=96, num_lat=49, time_step_size=600.0, subcycles=4, split_scheme="csp2",
                                 uv_adv_scheme="weno", zonal_tend_filter_cutoff_wavenumber=[4, 4]), 3),
    ("rossby_haurwitz_wave", dict(num_lon=360, num_lat=181, time_step_size=240.0, subcycles=6, split_scheme="csp2",
                                  zonal_tend_filter_cutoff_wavenumber=[4] * 5), 12),
    # the specified extensions: runge_kutta and the moving reduced tendency across bands
    ("mountain_zonal_flow", dict(num_lon=96, num_lat=49, time_step_size=600.0, subcycles=4, split_scheme="csp2",
                                 time_scheme="runge_kutta", time_order=3, zonal_tend_filter_cutoff_wavenumber=[4, 4]), 3),
    ("jet_zonal_flow", dict(num_lon=144, num_lat=73, time_step_size=450.0, subcycles=6, split_scheme="csp2",
                            use_zonal_reduce=True, reduce_adv_lon=True, use_reduce_tend_smooth=True,
                            zonal_reduce_factors=[8, 4, 2, 2], use_diffusion=True, diffusion_coef=1.0e5), 3),
]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    ndev = torch.cuda.device_count()
    shared = ndev < world
    dev = local % max(ndev, 1)
    torch.cuda.set_device(dev)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    red_dev = "cpu" if shared else "cuda"
    mode = sys.argv[1] if len(sys.argv) > 1 else "peer"   # "peer" (NVLink peer memory) or "nccl"
    out_path = sys.argv[2] if len(sys.argv) > 2 else None
    if shared and mode == "nccl":
        raise SystemExit("the NCCL path needs one GPU per rank")
    ok, report = True, []
    only = os.environ.get("GMD_CASES")   # e.g. "3,5": run only these cases (debugging aid)
    cases = [c for k, c in enumerate(CASES) if only is None or str(k) in only.split(",")]
    override = json.loads(os.environ.get("GMD_KW", "{}"))   # debugging aid: config keys to override, "nsteps" too
    for tc, kw, nsteps in cases:
        kw = dict(kw)
        kw.update({k: v for k, v in override.items() if k != "nsteps"})
        nsteps = override.get("nsteps", nsteps)
        o = Oracle(OracleConfig(**kw))
        o.set_initial_condition(tc)
        u, v, gd = o.state()
        ghs = o.ghs()
        o.run_init()
        m0 = o.diag()
        o.step(nsteps)
        ref = o.state()
        # noise floor of the case: oracle from a 1e-16-perturbed gd
        o2 = Oracle(OracleConfig(**kw))
        o2.set_state(u, v, gd * (1 + 1e-16 * np.random.default_rng(0).standard_normal(gd.shape)), ghs)
        o2.run_init()
        o2.step(nsteps)
        uscale = np.linalg.norm(ref[0])
        # v is (near) zero in the steady zonal flows: its error is measured against the wind speed, not itself
        scales = [None, max(np.linalg.norm(ref[1]), uscale), None]
        floor = [rel(a, b, s) for a, b, s in zip(o2.state(), ref, scales)]
        # >= 3 ranks: shorter first and last bands (gmd_config.polar_band_rows), as bench.py uses them
        pbr = max(kw["num_lat"] // world - 3, 8) if world >= 3 else 0
        d = gmd.Dycore(gmd.Config(rank=rank, nranks=world, device=dev, polar_band_rows=pbr, **kw))
        assert d.band() == parallel.band(rank, world, kw["num_lat"], pbr)
        used = parallel.connect(d, mode=mode, fallback=False)
        d.set_state(u, v, gd, ghs)
        d.run_init()
        md0 = d.diag()
        d.step(nsteps)
        nlat = kw["num_lat"]
        got = [parallel.gather_field(a, nlat, polar_band_rows=pbr) for a in d.state()]
        m, e, beta = d.diag()
        d.close()
        mo, eo, bo = o.diag()
        errs = [rel(a, b, s) for a, b, s in zip(got, ref, scales)]
        good = (abs(md0[0] / m0[0] - 1) < 1e-13 and abs(md0[1] / m0[1] - 1) < 1e-13 and abs(m / mo - 1) < 1e-13 and
                abs(e / eo - 1) < 1e-13 and all(er <= max(1e-12, 20 * fl) for er, fl in zip(errs, floor)))
        inv = None
        if rank == 0:
            s = gmd.Dycore(gmd.Config(device=dev, **kw))
            s.set_state(u, v, gd, ghs)
            s.run_init()
            s.step(nsteps)
            single = s.state()
            s.close()
            umax = np.abs(single[0]).max()
            inv = [float(np.abs(a - b).max() / max(np.abs(b).max(), umax if k == 1 else 0.0, 1e-300))
                   for k, (a, b) in enumerate(zip(got, single))]
            good = good and all(x <= max(1e-13, 5 * fl) for x, fl in zip(inv, floor))
            report.append({"case": tc, "grid": [kw["num_lon"], nlat], "split": kw["split_scheme"],
                           "adv": kw.get("uv_adv_scheme", "center_diff"), "steps": nsteps, "ranks": world, "comm": used,
                           "shared_gpu": shared, "rel_l2_vs_oracle_u_v_gd": errs, "noise_floor_u_v_gd": floor,
                           "max_rel_vs_one_band_u_v_gd": inv, "mass_rel": abs(m / mo - 1), "energy_rel": abs(e / eo - 1),
                           "beta_abs": abs(beta - bo), "ok": bool(good)})
            print(f"[{world} ranks, {used}{', one shared GPU' if shared else ''}] {tc} {kw['num_lon']}x{nlat} {kw['split_scheme']} "
                  f"{kw.get('uv_adv_scheme', 'center_diff')}: vs oracle rel-L2 u,v,gd = {errs} (floor {floor}), vs one band max-rel = {inv}, "
                  f"mass {abs(m / mo - 1):.1e} energy {abs(e / eo - 1):.1e} beta {abs(beta - bo):.1e} -> {'ok' if good else 'FAIL'}",
                  flush=True)
        ok = ok and good
    flag = torch.tensor([0 if ok else 1], device=red_dev)
    dist.all_reduce(flag)
    if rank == 0 and out_path:
        with open(out_path, "w") as f:
            json.dump(report, f, indent=1)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
