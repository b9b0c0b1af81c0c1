#!/usr/bin/env python
"""bench.py -- throughput of the barotropic shallow-water time step (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libgmd.so, C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's serial CPU algorithm (oracle port)

One "step" = one model step (time_integrate + diag_run, src/dycore_mod.F90:131-140) of the whole globe.
metric = grid-point updates/s = num_lon * num_lat * K / t, whole job.  Workload (BASELINE.json configs[3], the
configuration the metric is quoted on): steady geostrophic flow on the 0.1 degree grid 3600x1801, predict_correct +
beta, csp2 with S subcycles, centred advection, zonal filter on 20 rows per pole.  The reference ships no namelist
for this grid; dt, S and the cutoff vector are builder-chosen and reported in `config`.  BOTH arms run
this workload at its full size and print the same `config`.

  value     inputs resident in HBM when the timed region starts; K steps between CUDA events; max over ranks
  e2e       through the C ABI with HOST buffers: gmd_set_state (H2D of u,v,gd,ghs) + K x {gmd_step(1) +
            gmd_get_diag (D2H)} + gmd_get_state (D2H of u,v,gd), wall clock around the calls; the two state
            transfers are paid once per K steps (that is how dycore_run uses the path: state in, K steps between two
            output alerts, state out) and are reported per step as total / K
  roofline  the fused stage kernel timed alone with CUDA events: the S2 variant (13 words per column) as the headline
            kernel, and `family` = the three variants one predict_correct runs (deferred-update S1 + S2 + S3a,
            36 words), time-weighted
  cpu_baseline  the CPU oracle (-O3 -ffast-math, as the reference's -Ofast), 1 thread (the reference is serial: no
            OpenMP directive, no MPI call in src/), on a bounded sample: 3 steps of the SAME workload and grid
Multi-GPU (torchrun, one rank per GPU): latitude bands; halo rows are stored into the neighbour's ghost rows over
NVLink peer memory and the two-scalar all-reduces are one-shot peer exchanges, all inside libgmd (--comm nccl
selects the ncclSend/Recv + ncclAllReduce path instead); fixed global grid => "strong" scaling.  At N > 1 the line
also carries `decomposition_check`: a 360x181 Rossby-Haurwitz run in N bands against the same run in one band.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (test case, kwargs)
    "sg_0.1deg": ("steady_geostrophic_flow",
                  dict(num_lon=3600, num_lat=1801, time_step_size=10.0, subcycles=10, split_scheme="csp2",
                       zonal_tend_filter_cutoff_wavenumber=[4] * 20)),
    # one eighth of the 0.1 degree rows: with --gpus 2 each rank holds what an edge rank of the 8-GPU run holds
    # (development aid for the small-band regime; not a BASELINE configuration)
    "sg_band8": ("steady_geostrophic_flow",
                 dict(num_lon=3600, num_lat=451, time_step_size=10.0, subcycles=10, split_scheme="csp2",
                      zonal_tend_filter_cutoff_wavenumber=[4] * 20)),
    "rh_0.05deg": ("rossby_haurwitz_wave",
                   dict(num_lon=7200, num_lat=3601, time_step_size=2.0, subcycles=10, split_scheme="csp2",
                        zonal_tend_filter_cutoff_wavenumber=[4] * 20)),
    "jz_0.25deg": ("jet_zonal_flow",
                   dict(num_lon=1440, num_lat=721, time_step_size=30.0, subcycles=6, split_scheme="csp2",
                        zonal_tend_filter_cutoff_wavenumber=[4] * 20, use_diffusion=True, diffusion_coef=6.0e3)),
    "rh_1deg": ("rossby_haurwitz_wave",
                dict(num_lon=360, num_lat=181, time_step_size=240.0, subcycles=6, split_scheme="csp2",
                     zonal_tend_filter_cutoff_wavenumber=[4] * 5)),
}
CPU_BASELINE_STEPS = 3   # + 1 warm-up step: ~20 s of one core at 3600x1801


def initial_condition(test_case, kw):
    """synthetic input: the product's own test-case plugin (gamil_dycore_b200/host/test_cases.cpp behind
    include/gmd_host.h), on the host"""
    import gamil_dycore_b200 as gmd
    return gmd.initial_condition(test_case, kw["num_lon"], kw["num_lat"])


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU during the timed region (pynvml; nvidia-smi fallback)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        while not self.stop_flag:
            try:
                if self.nv:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                        self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
                else:
                    import subprocess
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    a, b = out.strip().split(",")
                    self.samples.append(int(a))
                    self.max_mhz = int(b)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_oracle_rate(test_case, kw, nsteps, warm=1):
    """grid-point updates/s of the serial CPU oracle on the workload's own grid (bounded by the step count)"""
    from oracle.oracle import Oracle, OracleConfig
    o = Oracle(OracleConfig(**kw), kind="fast")
    o.set_initial_condition(test_case)
    o.run_init()
    if warm > 0:
        o.step(warm)
    t0 = time.perf_counter()
    o.step(nsteps)
    t = time.perf_counter() - t0
    o.close()
    return kw["num_lon"] * kw["num_lat"] * nsteps / t, t / nsteps


def config_dict(name, kw):
    """the workload, identical in both arms (how the job is spread over GPUs is reported under `parallel`)"""
    cut = kw["zonal_tend_filter_cutoff_wavenumber"]
    return {"workload": f"{name}: {WORKLOADS[name][0]} {kw['num_lon']}x{kw['num_lat']}", "dt_s": kw["time_step_size"],
            "time_scheme": "predict_correct", "split_scheme": kw["split_scheme"], "subcycles": kw["subcycles"],
            "uv_adv_scheme": kw.get("uv_adv_scheme", "center_diff"),
            "filter_rows_per_pole": sum(1 for c in cut if c), "filter_cutoff": max(cut),
            "use_diffusion": bool(kw.get("use_diffusion", False)),
            "l2_policy": "per-step working set (>= 16 fields x 8 B x columns) exceeds the 126 MB L2 at 0.1 deg on one GPU; "
                         "nothing is flushed between steps"}


def run_reference(args):
    """the reference's own (serial, CPU) implementation of the path: not buildable here (Fortran), so the oracle port,
    on the SAME workload and grid as the GPU arm; K timed steps after W warm-up steps"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    test_case, kw = WORKLOADS[args.workload]
    from oracle import oracle as orc
    orc.build()
    rate, spp = cpu_oracle_rate(test_case, kw, args.steps, warm=max(args.warmup, 0))
    cpu = {"value": rate, "unit": "grid-point-updates/s", "cores": 1, "kind": "port",
           "sample": f"the workload itself ({kw['num_lon']}x{kw['num_lat']}), {args.steps} steps after {args.warmup} warm-up, "
                     "1 thread (the reference is serial: no OpenMP directive / MPI call in src/), gcc -O3 -ffast-math"}
    line = {"impl": "reference", "metric": "grid-point-updates/s", "value": rate, "unit": "grid-point-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": spp * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, kw), "cpu_baseline": cpu,
            "e2e": {"value": rate, "unit": "grid-point-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def decomposition_check(gmd, parallel, dist, rank, world, local, comm):
    """N bands against one band on the SAME GPU code: Rossby-Haurwitz 360x181 (BASELINE config C1), 30 steps.
    Returns max over u, v, gd of max|banded - single| / max|single| (rank 0; None elsewhere)."""
    import numpy as np
    kw = dict(WORKLOADS["rh_1deg"][1])
    u, v, gd, ghs = initial_condition("rossby_haurwitz_wave", kw)
    nlat, nsteps = kw["num_lat"], 30
    pbr = 0   # even bands: 181 rows over N ranks
    d = gmd.Dycore(gmd.Config(rank=rank, nranks=world, device=local, polar_band_rows=pbr, **kw))
    used = parallel.connect(d, mode=comm)
    d.set_state(u, v, gd, ghs)
    d.run_init()
    d.step(nsteps)
    got = [parallel.gather_field(a, nlat, polar_band_rows=pbr) for a in d.state()]
    mb, eb, _ = d.diag()
    d.close()
    if rank != 0:
        return None
    s = gmd.Dycore(gmd.Config(device=local, **kw))
    s.set_state(u, v, gd, ghs)
    s.run_init()
    s.step(nsteps)
    ref = s.state()
    ms, es, _ = s.diag()
    s.close()
    errs = [float(np.abs(a - b).max() / np.abs(b).max()) for a, b in zip(got, ref)]
    return {"case": f"rossby_haurwitz_wave 360x181 dt 240 csp2 x6, {nsteps} steps, {world} bands vs 1 band, comm {used}",
            "max_rel_u_v_gd": errs, "decomposition_max_rel": max(errs),
            "mass_rel": abs(mb / ms - 1), "energy_rel": abs(eb / es - 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sg_0.1deg", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-decomposition-check", action="store_true")
    ap.add_argument("--polar-band-rows", type=int, default=-1,
                    help="rows of the first and last latitude band for >= 3 GPUs (0: even bands; default: balanced "
                         "against the polar-row work, gamil_dycore_b200.parallel.polar_band_rows_for)")
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU data path: NVLink peer memory (default) or NCCL send/recv + all-reduce")
    ap.add_argument("--lib", default=None, help="path of an alternative build of libgmd (experiments; default: the product)")
    ap.add_argument("--trace", default=None,
                    help="write a per-launch device timeline of one model step (libgmd_trace.so build) to this file "
                         "prefix (one JSON per rank) instead of benchmarking")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import gamil_dycore_b200 as gmd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    test_case, kw = WORKLOADS[args.workload]
    # at least 5 untimed steps: a model step is replayed from a CUDA graph captured once per buffer set, and the buffer sets
    # cycle with a period of up to 4 steps -- with fewer warm-up steps a capture would land in the timed region (measured:
    # 3.44 instead of 3.36 ms per step with --warmup 3 --steps 10).  The JSON line reports the number actually run.
    W = max(args.warmup, 5)
    K = args.steps
    u, v, gd, ghs = initial_condition(test_case, kw)
    ncol = kw["num_lon"] * kw["num_lat"]

    from gamil_dycore_b200 import parallel
    pbr = args.polar_band_rows if args.polar_band_rows >= 0 else parallel.polar_band_rows_for(
        world, kw["num_lon"], kw["num_lat"], any(kw["zonal_tend_filter_cutoff_wavenumber"]))
    kind = os.path.abspath(args.lib) if args.lib else ("trace" if args.trace else "fast")
    d = gmd.Dycore(gmd.Config(rank=rank, nranks=world, device=local, polar_band_rows=pbr, **kw), kind=kind)
    if world > 1:
        args.comm = parallel.connect(d, mode=args.comm)   # "peer" falls back to "nccl" if CUDA IPC is not available
    if args.no_graph:
        d.set_graph_mode(False)
    stream = torch.cuda.Stream()   # a real (non-legacy) stream: libgmd launches on it, torch events time it
    torch.cuda.set_stream(stream)
    d.set_stream(stream.cuda_stream)
    d.set_state(u, v, gd, ghs)
    d.run_init()
    m0, e0, _ = d.diag()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.trace:
        d.step(W)
        barrier()
        d.trace_begin()
        d.step(2)
        recs = d.trace_end()
        barrier()
        recs["workload"] = config_dict(args.workload, kw)["workload"]
        recs["comm"] = args.comm if world > 1 else None
        with open(f"{args.trace}.rank{rank}.json", "w") as f:
            json.dump(recs, f)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- device-resident throughput ----------------------------------------------------------------------
    d.step(W)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = d.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    d.step_async(K)
    ev1.record(stream)
    d.sync()
    barrier()
    ms = ev0.elapsed_time(ev1)
    ms_lib = d.last_step_ms()      # libgmd's own CUDA events around the same span (cross-check)
    if abs(ms_lib - ms) > 0.05 * ms + 0.05:
        raise SystemExit(f"timing mismatch: torch events {ms:.3f} ms vs libgmd events {ms_lib:.3f} ms")
    launches = d.kernel_launches() - l0
    sampler.stop_flag = True
    sampler.join()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = ncol * K / (ms * 1e-3)
    m1, e1, beta = d.diag()

    # ---- roofline of the dominant kernel (this rank's band, timed alone) -----------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    pass_name = "fast" if kw["split_scheme"] in ("csp2", "isp") else "all"
    fam, fam_ms, fam_bytes = {}, 0.0, 0.0
    for mode, nm in ((4, "S1_deferred_update"), (1, "S2"), (2, "S3a")):
        d.time_stage_variant(pass_name, mode, 3)
        vms, vbytes = d.time_stage_variant(pass_name, mode, 20)
        fam[nm] = {"ms_per_launch": vms, "algorithmic_bytes_per_launch": vbytes, "achieved": vbytes / (vms * 1e-3) / 1e9,
                   "frac": vbytes / (vms * 1e-3) / 1e9 / peak}
        fam_ms += vms
        fam_bytes += vbytes
    traffic_tab = {}
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "stage_kernel_traffic.json")))
    except Exception:
        pass
    step_gbps = d.algorithmic_bytes_per_column_step() * value / 1e9 / max(world, 1)
    family = {"what": "the three stage-kernel variants of one fast predict_correct (the path of the rows next to the poles, "
                      "of polar bands and of the strict build), each timed alone over the whole band, time-weighted: sum "
                      "of algorithmic bytes / sum of launch times",
              "achieved": fam_bytes / (fam_ms * 1e-3) / 1e9, "frac": fam_bytes / (fam_ms * 1e-3) / 1e9 / peak, "variants": fam}
    fused = None
    fr = d.fused_rows()
    use_fused = fr[1] > fr[0]
    if world > 1:   # the timing call is collective (its warm-up runs the inner-product all-reduce): all ranks or none
        t = torch.tensor([1.0 if use_fused else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        use_fused = bool(t.item() > 0.5)
    if use_fused:
        d.time_stage_variant(pass_name, 5, 3)
        fused = d.time_stage_variant(pass_name, 5, 20)
    if fused is not None:
        # The dominant kernel of the step is k_pc: one launch = one predict_correct (S1 + S2 + S3a as one wavefront kernel,
        # the previous predict_correct's update folded in).  `achieved` follows the contract: SURVEY 8d's per-unit figure
        # (312 B per column and fast predict_correct, 216 B slow) x the columns the launch covers / its duration.  The
        # kernel itself only has to move 13 of those 39 words (`moved`): the stage states travel through shared memory,
        # so `frac` can exceed 1 and the kernel is bound by fp64 issue, not by HBM (profiles/r2_r_ncu_k_pc_summary.txt).
        kms, kbytes = fused
        per_col = 312.0 if pass_name != "slow" else 216.0
        model_bytes = kbytes / (13 * 8.0) * per_col
        achieved = model_bytes / (kms * 1e-3) / 1e9
        traffic = traffic_tab.get(args.workload + ":k_pc") if world == 1 else None
        roofline = {"bound": "hbm", "kernel": f"k_pc<{pass_name}, deferred update>", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                      "(profiles/stage_kernel_traffic.json)" if traffic else None,
                    "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                    "ms_per_launch": kms, "algorithmic_bytes_per_launch": model_bytes,
                    "algorithmic_bytes_what": f"SURVEY 8d: {per_col:.0f} B per column per predict_correct x the columns of the fused rows "
                                              "(one launch = one predict_correct)",
                    "moved": {"what": "what the fused kernel has to move: old state 3 + ghs 1 + deferred tendency 3 words in, "
                                      "materialised state 3 + new tendency 3 out = 13 words per column",
                              "bytes_per_launch": kbytes, "GBps": kbytes / (kms * 1e-3) / 1e9,
                              "frac_of_peak": kbytes / (kms * 1e-3) / 1e9 / peak},
                    "rows_covered": list(d.fused_rows()),
                    "limiter": "fp64 issue / dependent latency (ncu: issue active 57 %, fp64 pipe 38 %, DRAM 30 % of peak)",
                    "three_sweep_family": family,
                    "step_algorithmic_GBps": step_gbps, "step_frac_per_gpu": step_gbps / peak}
    else:
        kms, kbytes = fam["S2"]["ms_per_launch"], fam["S2"]["algorithmic_bytes_per_launch"]
        achieved = kbytes / (kms * 1e-3) / 1e9
        traffic = traffic_tab.get(args.workload) if world == 1 else None
        roofline = {"bound": "hbm", "kernel": f"k_stage<{pass_name}, S2>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                      "(profiles/stage_kernel_traffic.json)" if traffic else None,
                    "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                    "ms_per_launch": kms, "algorithmic_bytes_per_launch": kbytes, "family": family,
                    "step_algorithmic_GBps": step_gbps, "step_frac_per_gpu": step_gbps / peak}

    # ---- end to end through the C ABI with host buffers ----------------------------------------------------
    barrier()
    t0 = time.perf_counter()
    d.set_state(u, v, gd, ghs)
    for _ in range(K):
        d.step(1)
        d.diag()
    uo, vo, gdo = d.state()
    barrier()
    te = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([te], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t.item())
    r0, r1 = d.band()
    h2d = 4 * (r1 - r0 + 4) * kw["num_lon"] * 8
    d2h = 3 * (r1 - r0) * kw["num_lon"] * 8 + K * 2 * 3 * 8   # + per step: the NaN check and gmd_get_diag, 3 doubles each
    e2e = {"value": ncol * K / te, "unit": "grid-point-updates/s", "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
           "what": f"gmd_set_state + {K} x (gmd_step(1) + gmd_get_diag) + gmd_get_state, host arrays in and out; the state "
                   f"upload and download are paid once and amortised over the {K} steps (bytes per step = total / {K})"}
    d.close()

    dec = None
    if world > 1 and not args.no_decomposition_check:
        dec = decomposition_check(gmd, parallel, dist, rank, world, local, args.comm)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            rate, spp = cpu_oracle_rate(test_case, kw, CPU_BASELINE_STEPS)
            cpu = {"value": rate, "unit": "grid-point-updates/s", "cores": 1, "kind": "port", "ms_per_step": spp * 1e3,
                   "sample": f"the workload itself ({kw['num_lon']}x{kw['num_lat']}), {CPU_BASELINE_STEPS} steps after 1 warm-up, "
                             "1 thread (the reference is serial), oracle built with gcc -O3 -ffast-math"}
        line = {"metric": "grid-point-updates/s", "value": value, "unit": "grid-point-updates/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.workload, kw),
                "parallel": {"decomposition": f"{world} latitude band(s)", "polar_band_rows": pbr,
                             "comm": None if world == 1 else ("NVLink peer memory" if args.comm == "peer" else "NCCL")},
                "sim_days_per_day": kw["time_step_size"] * K / (ms * 1e-3), "clocks": sampler.summary(), "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "conservation": {"mass_rel_drift": abs(m1 / m0 - 1), "energy_rel_drift": abs(e1 / e0 - 1), "beta": beta}}
        if dec is not None:
            line["decomposition_check"] = dec
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
