#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native CLASS++ hot path.

Metric (BASELINE.json): lensed-C_l spectra / s for the Planck-2018 LambdaCDM configuration
(base_2018_plikHM_TTTEEE_lowl_lowE_lensing.ini: 1 massive neutrino species, halofit,
l_max_scalars=2500, P(k) to 1 h/Mpc), measured on the hot path this repository replaces:
PerturbationsModule -> (halofit) -> TransferModule -> SpectraModule -> LensingModule + P(k)
(SURVEY.md section 8 with its two "next" rows).

A "step" = one pass of the hot path over one batch of `--batch` cosmologies per GPU: ONE batched
perturbation launch, then the per-cosmology stages.  The batch is `--batch` DIFFERENT cosmologies: the
seed-0 Latin hypercube of BASELINE configs[4] with the Planck-18 settings (`--identical`: copies of the
best fit); their upstream tables (background, thermodynamics, ncdm grids) are produced once, outside the
timed region, by the reference's own upstream modules inside the compiled drop-in library
(shim/_build/libclass_b200.so).  Steps are scheduled by sweep.SweepPipeline on `--sets` sets of contexts:
the batched launches of consecutive steps overlap and the per-cosmology stages run under them;
everything is drained before the clock stops.  The reference arm runs the unmodified reference
(oracle/_ref) in its throughput-optimal arrangement (one single-threaded process per core) on the same
workload and the same five module constructors.

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU)
  python bench.py --impl reference --steps K --warmup W    the reference's CPU implementation
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "lensed C_l spectra/sec (Planck-18 LCDM hot path: perturbations+halofit+transfer+spectra+lensing)"
UNIT = "spectra/s"

# Algorithmic FP64 work of stage 1 per cosmology (SURVEY.md 8d / BASELINE.md 2): oracle stepstat
# counters x fixed per-RHS / per-solve costs -- NOT the GPU's own step count.
ALGO_FLOP_STAGE1 = {"planck18": 3.6e9, "lcdm": 1.0e9, "lcdm_coarse": 1.0e8}
# stage 2: integrand points x 40 flop
ALGO_FLOP_PER_LOS_POINT = 40.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        def num(s):
            try:
                return float(s)
            except Exception:
                return None
        sm = [num(r[0]) for r in self.rows if r and num(r[0]) is not None]
        mx = [num(r[1]) for r in self.rows if len(r) > 1 and num(r[1]) is not None]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


class _NL:
    def __init__(self, arr):
        self.nl_corr_density_m = arr


_RESULT_LINE = []  # the one JSON line of rank 0 (printed by main() after stdout has been restored)

# DRAM bytes one batched perturbation launch moves per cosmology: measured with ncu (profiles/, see roofline.traffic_source)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_traffic.json")


def sweep_inputs(args, M):
    """The B cosmologies of a step.  Default: B DIFFERENT cosmologies -- the seed-0 Latin hypercube of BASELINE config 5
     with the Planck-2018 settings of configs[1] -- whose upstream tables come from the compiled drop-in library
    (shim/_build/libclass_b200.so: the reference's own input/background/thermodynamics modules, out of scope and unchanged).
    --identical (or no drop-in library on this box): B copies of the Planck-18 best fit from tests/golden/planck18.npz."""
    from classpp_public_b200.configs import CONFIGS
    base = CONFIGS[args.config]
    fixture = M.Inputs.load(os.path.join(ROOT, "tests", "golden", args.config + ".npz"))
    if not args.identical:
        try:
            from classpp_public_b200 import upstream
            if not upstream.available():
                raise ImportError("shim/_build/libclass_b200.so not found")
            from concurrent.futures import ThreadPoolExecutor
            # weak scaling: every rank takes its own shard of one Latin hypercube of batch x world points
            rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
            pars = upstream.latin_hypercube_sweep(args.batch * world, base, seed=0)[rank::world]
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
                inps = list(ex.map(upstream.inputs_for, pars))
            prims = [M.AnalyticPrimordial(p["A_s"], p["n_s"]) for p in pars]
            return inps, prims, {"kind": "lhs", "upstream_host_s_per_cosmology": (time.perf_counter() - t0) / args.batch,
                                 "upstream_threads": os.cpu_count()}
        except Exception as e:  # noqa: BLE001
            print("bench.py: falling back to identical cosmologies (%s)" % e, file=sys.stderr)
    prim = M.AnalyticPrimordial(base.get("A_s", 2.215e-9), base.get("n_s", 0.9619))
    return [fixture] * args.batch, [prim] * args.batch, {"kind": "identical"}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from classpp_public_b200 import modules as M
    from classpp_public_b200.sweep import SweepPipeline

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line of the contract
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B, NSET = args.batch, (1 if args.no_pipeline else args.sets)
    inps, prims, wl = sweep_inputs(args, M)
    use_halofit = int(inps[0].meta["nl.method"]) != 0

    # ---- one context (own stream) per cosmology, NSET sets of B contexts used by consecutive steps: the batched
    # perturbation launches of consecutive steps OVERLAP (the lane kernel's duration is set by its longest k chain, not
    # by the batch: the GPU is mostly idle while the last chains finish), and the per-cosmology stages of a step run
    # under the launches of the next ones.  Everything is drained before the clock stops.
    sets = []
    for s_ in range(NSET):
        cs_, ms_ = [], []
        for b in range(B):
            ctx = M.Context(local)
            ctx.set_option("lean_scratch", 1)
            if args.path != "auto":
                ctx.set_option("lane_path", 1 if args.path == "lane" else 0)
            bg = M.BackgroundModule(inps[b], ctx)
            th = M.ThermodynamicsModule(inps[b], bg)
            cs_.append(ctx)
            ms_.append((bg, th))
        sets.append((cs_, ms_))
    ctxs = [c for cs_, _ in sets for c in cs_]

    results = [None] * B
    kms, kms_lock = {}, threading.Lock()
    KEYS = ("perturb", "k_spline", "bessel", "los", "spectra", "perturb_tail", "halofit", "lensing")

    def kms_reset():
        for k_ in KEYS:
            kms[k_] = 0.0

    kms_reset()

    def kms_add(ctx, keys):
        t = ctx.kernel_ms()
        with kms_lock:
            for k_ in keys:
                kms[k_] += t[k_]

    def front(s_, b, host_inputs):
        cs, ms = sets[s_]
        if host_inputs is not None:  # end to end: host -> device copy of this step's inputs (pinned host tables)
            bg = M.BackgroundModule(host_inputs[b], cs[b])
            ms[b] = (bg, M.ThermodynamicsModule(host_inputs[b], bg))
        return M.PerturbationsModule((host_inputs or inps)[b], ms[b][0], ms[b][1], solve=False)

    def on_solved(s_):
        kms_add(sets[s_][0][0], ("perturb", "perturb_tail"))

    def back(s_, b, pt, host_inputs):
        cs, ms = sets[s_]
        x = (host_inputs or inps)[b]
        nlb = M.NonlinearModule(x, ms[b][0], pt, prims[b]) if use_halofit else None  # halofit on the device
        tr = M.TransferModule(x, ms[b][0], ms[b][1], pt, nlb)
        sp = M.SpectraModule(x, pt, prims[b], nlb, tr)
        le = M.LensingModule(x, sp)  # lensed TT/TE/EE/BB on the device: the metric's "lensed C_l"
        pk_lin = pt.pk_linear(prims[b].pk_at_k(pt.k_[0]))  # linear P(k, z=0) on the perturbation k grid
        out_bytes = sp.cl_[0].nbytes + le.cl_lens_.nbytes + pk_lin.nbytes  # device -> host: the results a user reads
        kms_add(cs[b], ("k_spline", "bessel", "los", "spectra", "halofit", "lensing"))
        results[b] = (pt.info, tr.info, out_bytes, float(le.cl_lens_[le.lt_size_ * 10]))
        return out_bytes

    pipe = SweepPipeline(B, front, back, n_sets=NSET, on_solved=on_solved)

    def step(host_inputs=None):
        pipe.submit(host_inputs)
        if NSET == 1:
            pipe.drain()

    def barrier():
        pipe.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: the first launch runs alone (its duration sets the stagger between overlapping launches)
    t_w = time.perf_counter()
    step()
    pipe.drain()
    solo_s = pipe.solve_seconds[-1]
    if NSET > 1 and args.stagger >= 0:  # default: one batched launch at a time, only the per-cosmology stages overlap it
        pipe.stagger = args.stagger * solo_s
    for _ in range(max(args.warmup - 1, 0)):
        step()
    pipe.drain()
    warm_s = time.perf_counter() - t_w

    launches0 = sum(c.launch_count for c in ctxs)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    kms_reset()
    for _ in range(args.steps):
        step()
    barrier()
    kms_timed = dict(kms)
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    elapsed = max(ev0.elapsed_time(ev1) * 1e-3, 1e-9)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    launches = sum(c.launch_count for c in ctxs) - launches0
    if world > 1:
        t = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = B * args.steps * world / elapsed
    launch_log = [(s_, round(a_ - t0, 2), round(b_ - t0, 2)) for s_, a_, b_ in pipe.launch_log[-args.steps:]]

    # ---- end to end through the C ABI from pinned HOST buffers: every step uploads the upstream tables of its B
    # cosmologies (clpp_set_background / clpp_set_thermo: host spline + H2D) and reads the results back (C_l, lensed C_l, P(k))
    def pinned(x):
        return torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()
    cache = {}

    def pin_inputs(i):
        if id(i) not in cache:
            cache[id(i)] = M.Inputs(i.meta, {k: (pinned(v) if v.dtype == np.float64 and v.size > 64 else v) for k, v in i.arrays.items()})
        return cache[id(i)]
    host = [pin_inputs(i) for i in inps]
    h2d = sum(sum(i.arrays[k].nbytes for k in ("bg.tau_table", "bg.background_table", "th.z_table", "th.thermodynamics_table")) * 2
              for i in host)  # tables + their spline tables
    e2e_steps = max(1, min(args.steps, NSET))
    barrier()
    te0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    for _ in range(e2e_steps):
        step(host)
    barrier()
    d2h = sum(r[2] for r in results)
    ee1.record()
    torch.cuda.synchronize()
    e2e_elapsed = max(ee0.elapsed_time(ee1) * 1e-3, time.perf_counter() - te0)
    if world > 1:
        t = torch.tensor([e2e_elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_elapsed = float(t.item())
    e2e_value = e2e_steps * B * world / e2e_elapsed

    # ---- latency of ONE cosmology (the Class.compute() use case): perturbations alone and the whole hot path
    lat = {}
    if rank == 0 and not args.no_latency:
        c1 = M.Context(local)
        bg1 = M.BackgroundModule(inps[0], c1)
        th1 = M.ThermodynamicsModule(inps[0], bg1)
        for rep in range(2):
            t1 = time.perf_counter()
            pt1 = M.PerturbationsModule(inps[0], bg1, th1)
            t2 = time.perf_counter()
            nl1 = M.NonlinearModule(inps[0], bg1, pt1, prims[0]) if use_halofit else None
            tr1 = M.TransferModule(inps[0], bg1, th1, pt1, nl1)
            sp1 = M.SpectraModule(inps[0], pt1, prims[0], nl1, tr1)
            M.LensingModule(inps[0], sp1)
            t3 = time.perf_counter()
        lat = {"latency_single_cosmology_s": t3 - t1, "latency_single_cosmology_perturb_s": t2 - t1,
               "latency_note": "one cosmology alone on the GPU (warp-per-mode kernels: bounded by the k = 22.6/Mpc chain of 3.6e5 "
                               "step attempts); second of two runs"}
        c1.close()

    # ---- the same cosmology over ALL ranks (north-star split: cost-balanced k lists -> NCCL all-gather of S(k,tau) ->
    # q ranges -> all-reduce of the partial C_l); bounded by the one k = 22.6/Mpc chain, so it does not get faster with N
    if world > 1 and not args.no_latency:
        from classpp_public_b200 import multigpu
        fixture = M.Inputs.load(os.path.join(ROOT, "tests", "golden", args.config + ".npz"))
        from classpp_public_b200.configs import CONFIGS
        prim_f = M.AnalyticPrimordial(CONFIGS[args.config].get("A_s", 2.215e-9), CONFIGS[args.config].get("n_s", 0.9619))
        for rep in range(2):
            torch.cuda.synchronize()
            dist.barrier()
            t1 = time.perf_counter()
            multigpu.compute_cl_distributed(fixture, prim_f, None, rank, world, local)
            torch.cuda.synchronize()
            dist.barrier()
            t_dist = time.perf_counter() - t1
        lat["latency_single_cosmology_distributed_s"] = t_dist
        lat["latency_distributed_note"] = "the Planck-18 best fit split over %d GPUs (second of two runs)" % world

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the batched perturbation launch): FP64 vector pipe
    peaks, peaks_kind = load_peaks()
    fp64_peak = ctxs[0].fp64_peak_tflops()
    perturb_busy = min(kms_timed["perturb"] * 1e-3, elapsed)  # launches of consecutive steps overlap on purpose
    algo = ALGO_FLOP_STAGE1.get(args.config, 1.0e9) * B
    achieved = algo * args.steps / max(perturb_busy, 1e-12) / 1e12
    info_pt, info_tr = results[0][0], results[0][1]
    n_back = B * args.steps
    t_los = max(kms_timed["los"] * 1e-3 / n_back, 1e-12)
    sec_achieved = ALGO_FLOP_PER_LOS_POINT * float(info_tr.n_points) / t_los / 1e12
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    traffic = json.load(open(TRAFFIC_FILE)) if os.path.exists(TRAFFIC_FILE) else {}
    kern = "perturb_lane_kernel" if args.path == "lane" else "perturb_kernel + perturb_tail_kernel"
    roofline = {"kernel": kern, "bound": "fp64-vector (latency-bound in practice; neither hbm nor tensor)",
                "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "peak_source": "DFMA microbenchmark run live in bench.py (MEASURED_PEAKS.json has no FP64 entry; "
                               "its hbm_gbs=%s bf16_tflops=%s are %s)" % (peaks.get("hbm_gbs"), peaks.get("bf16_tflops"), peaks_kind),
                "traffic": (traffic.get(kern, {}).get("dram_bytes_per_cosmology", 0.0) * B) or None,
                "traffic_source": traffic.get(kern, {}).get("source", "no ncu capture of this kernel at this batch size yet"),
                "algorithmic_flop_per_launch": algo,
                "algorithmic_bytes_per_launch": 30.3e6 * B,
                "launch_duration_ms_avg": kms_timed["perturb"] / args.steps,
                "launches_in_flight_avg": kms_timed["perturb"] * 1e-3 / elapsed,
                "note": "achieved = algorithmic flop of the K launches (SURVEY 8d: oracle stepstat x fixed per-RHS / per-solve "
                        "costs, 3.6 GFLOP per Planck-18 cosmology) / device time they occupy (= min(sum of launch durations, "
                        "timed region): launches of consecutive steps overlap on purpose)",
                "kernel_ms_per_step": {k_: v / args.steps for k_, v in kms_timed.items()},
                "kernel_share_of_step": {k_: v * 1e-3 / elapsed for k_, v in kms_timed.items()},
                "secondary": [
                    {"kernel": "los_kernel", "bound": "fp64-vector", "achieved": sec_achieved, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": sec_achieved / fp64_peak,
                     "algorithmic_flop_per_launch": ALGO_FLOP_PER_LOS_POINT * float(info_tr.n_points),
                     "launch_duration_ms_avg": t_los * 1e3, "note": "integrand points counted by the kernel (tr_info.n_points) x 40 flop"},
                    {"kernel": "k_spline_kernel", "bound": "hbm",
                     "achieved": 3 * 8.0 * info_pt.tp_size * info_pt.k_size * info_pt.tau_size /
                                 max(kms_timed["k_spline"] * 1e-3 / n_back, 1e-12) / 1e9,
                     "peak": hbm, "unit": "GB/s", "note": "read S, write S'' and the sweep scratch: 3 x 8 B x tp x k x tau per launch"},
                    {"kernel": "spectra_partial_kernel + spectra_final_kernel", "bound": "hbm",
                     "achieved": 8.0 * info_tr.tt_size * info_tr.l_size * info_tr.q_size /
                                 max(kms_timed["spectra"] * 1e-3 / n_back, 1e-12) / 1e9,
                     "peak": hbm, "unit": "GB/s", "note": "reads Delta_l(q) once: 8 B x tt x l x q per launch"}]}
    for s_ in roofline["secondary"][1:]:
        s_["frac"] = s_["achieved"] / s_["peak"]

    workload = ("BASELINE configs[1]: base_2018_plikHM_TTTEEE_lowl_lowE_lensing.ini settings (1 ncdm species, halofit, "
                "l_max_scalars=2500, P_k_max_h/Mpc=1); batch = %d %s" %
                (B, "DIFFERENT cosmologies: seed-0 Latin hypercube of BASELINE configs[4] (omega_b, omega_cdm, h, ln10^10A_s, n_s, "
                    "tau_reio)" if wl["kind"] == "lhs" else "copies of the Planck-18 best fit (identical cosmologies)"))
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload if args.config == "planck18" else args.config, "batch_per_gpu": B,
                   "cosmologies": wl,
                   "halofit": "on the device, inside the step (clpp_nonlinear_halofit)" if use_halofit else "none",
                   "scope": "per cosmology: perturbations -> halofit -> transfer -> spectra -> lensing (fast mode) -> linear P(k)",
                   "k_modes": int(info_pt.k_size), "tau_samples": int(info_pt.tau_size),
                   "q_values": int(info_tr.q_size), "l_values": int(info_tr.l_size),
                   "pipeline": ("%d context sets of %d cosmologies; %s (solo launch %.2f s); the per-cosmology stages of a step run "
                                "under the launch of the next one; all drained before the clock stops" %
                                (NSET, B, ("batched perturbation launches of consecutive steps overlap, %.2f s apart" % pipe.stagger)
                                 if pipe.stagger is not None else "one batched perturbation launch at a time", solo_s)) if NSET > 1 else "off",
                   "launch_log_set_start_end_s": launch_log,
                   "parallelism": "independent cosmologies per GPU (replicas, no data-path collective); per GPU every k mode of the "
                                  "batch in one batched launch, one warp per mode: long-tail modes on a high-priority stream, the "
                                  "bulk in chunks on low-priority streams, generic kernel -> radiation-streaming tail kernel",
                   "l2_policy": "working set per step = batch x (tables 6 MB + sources 24 MB + per-mode state 35 KB x 617 + transfer "
                                "work buffers 90 MB) >> 126 MB L2; every buffer is rewritten every step",
                   "timing": "torch.cuda.Event around K steps after device-wide synchronize, max over ranks; per-kernel "
                             "times from cudaEvents on the launching stream inside libclpp.so"},
        "k_modes_per_s": value * int(info_pt.k_size),
        "wall_s": wall, "warmup_s": warm_s,
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps,
                "note": "per step and per cosmology of the batch: upstream tables from pinned host memory through "
                        "clpp_set_background/clpp_set_thermo (host spline + H2D), grids, batched perturbation launch, "
                        "halofit, transfer, spectra, lensing, P(k), D2H of cl_, cl_lens_ and P(k) (contexts and device "
                        "buffers are reused across steps; the 24 MB source table of a cosmology stays on the device)"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    out.update(lat)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.config, budget_s=20.0, identical=(wl["kind"] != "lhs"))
    _RESULT_LINE.append(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _ref_hot_path_seconds(params, threads):
    """Wall time of the five module constructors of the hot path in the unmodified reference (oracle/_ref)."""
    from oracle import refprobe
    ref = refprobe.RefCosmology(params, threads=threads)
    ref.compute("lensing")
    t3 = ref.scalar("time.perturb") + ref.scalar("time.transfer") + ref.scalar("time.spectra")
    t = t3 + ref.scalar("time.nonlinear") + ref.scalar("time.lensing")
    up = ref.scalar("time.background") + ref.scalar("time.thermo")
    ref.close()
    return t, t3, up


def _sweep_params(config, n, identical=False):
    from classpp_public_b200.configs import CONFIGS
    from classpp_public_b200.upstream import latin_hypercube_sweep
    return [dict(CONFIGS[config])] * n if identical else latin_hypercube_sweep(n, CONFIGS[config], seed=0)


def ref_worker(args):
    """One process of the throughput arrangement of the reference: one cosmology, thread pool of ONE thread."""
    par = _sweep_params(args.config, args.ref_worker + 1, args.identical)[args.ref_worker]
    t, t3, up = _ref_hot_path_seconds(par, 1)
    print(json.dumps({"hot_path_s": t, "three_module_s": t3, "upstream_s": up}))


def reference_throughput(config, n_proc, identical=False, rounds=1):
    """the throughput-optimal CPU arrangement -- one process per cosmology with one thread each,
    packed onto all host cores.  Returns (cosmologies per second over the hot-path constructors, details)."""
    best = None
    for _ in range(rounds):
        t0 = time.perf_counter()
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", config,
                                   "--ref-worker", str(i)] + (["--identical"] if identical else []),
                                  stdout=subprocess.PIPE, stderr=subprocess.DEVNULL) for i in range(n_proc)]
        outs = [json.loads(p.communicate()[0].decode().strip().splitlines()[-1]) for p in procs]
        wall = time.perf_counter() - t0
        # every process also pays interpreter start-up and the upstream modules: score the hot-path constructors only,
        # as if the processes were perfectly packed (upper bound of the CPU throughput)
        hot = float(np.max([o["hot_path_s"] for o in outs]))
        r = {"value": n_proc / hot, "wall_s": wall, "hot_path_s_max": hot, "hot_path_s_mean": float(np.mean([o["hot_path_s"] for o in outs])),
             "upstream_s_mean": float(np.mean([o["upstream_s"] for o in outs])), "processes": n_proc}
        if best is None or r["value"] > best["value"]:
            best = r
    return best


def cpu_baseline(config, budget_s=20.0, threads=None, identical=False):
    """The reference's own CPU implementation of the hot path (oracle/_ref = unmodified CLASS++ built from /root/reference)
    on this box's host cores, on a bounded sample of the bench workload: (1) latency arrangement, one cosmology on all
    threads of the reference's thread pool; (2) throughput arrangement, one single-threaded process per core."""
    from oracle import refprobe
    if not refprobe.available():
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built on this box"}
    cores = threads or os.cpu_count()
    pars = _sweep_params(config, 4, identical)
    lat = []
    t_start = time.perf_counter()
    for i in range(4):
        t, t3, up = _ref_hot_path_seconds(pars[i % len(pars)], cores)
        if i > 0:  # the first run of a process is cold
            lat.append(t)
        if time.perf_counter() - t_start > 0.4 * budget_s and len(lat) >= 2:
            break
    thr = reference_throughput(config, cores, identical)
    value = max(1.0 / min(lat), thr["value"])
    return {"value": value, "unit": UNIT, "cores": int(cores), "kind": "reference",
            "latency_arrangement": {"spectra_per_s": 1.0 / min(lat), "hot_path_s_best": min(lat), "threads": int(cores)},
            "throughput_arrangement": thr,
            "sample": "%d cosmologies of the bench workload on all %d threads of the reference's thread pool (first discarded) and "
                      "%d cosmologies as %d single-threaded processes (BASELINE.md section 3); value = the better of the two; timed: "
                      "Perturbations+Nonlinear+Transfer+Spectra+Lensing module constructors (the scope of the GPU arm's step; "
                      "background/thermodynamics/primordial excluded)" % (len(lat) + 1, cores, cores, cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if args.ref_worker >= 0:
        return ref_worker(args)
    from oracle import refprobe
    if not refprobe.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (make -C oracle needs /root/reference)"}))
        return
    cores = os.cpu_count()
    # a step = a bounded sample of the bench workload: its first `cores` cosmologies in the throughput arrangement (one
    # single-threaded process each, ~8 s per step).  W warm-up and K timed steps as asked; a wall-clock budget of 5 minutes
    # cuts the run short if the box is slower than that (the line then says how many steps were timed).
    n_warm, n_steps = max(0, args.warmup), max(1, args.steps)
    vals = []
    t_begin = time.perf_counter()
    for i in range(n_warm + n_steps):
        t_step = time.perf_counter()
        r = reference_throughput(args.config, cores, args.identical)
        if i >= n_warm:
            vals.append(r)
        t_now = time.perf_counter()
        if vals and (t_now - t_begin) + (t_now - t_step) > 300.0:
            break
    value = float(np.mean([v["value"] for v in vals]))
    lat, _, _ = _ref_hot_path_seconds(_sweep_params(args.config, 1, args.identical)[0], cores)
    lat, _, _ = _ref_hot_path_seconds(_sweep_params(args.config, 1, args.identical)[0], cores)
    from classpp_public_b200.configs import CONFIGS  # noqa: F401
    workload = ("BASELINE configs[1]: base_2018_plikHM_TTTEEE_lowl_lowE_lensing.ini settings (1 ncdm species, halofit, "
                "l_max_scalars=2500, P_k_max_h/Mpc=1); batch = %d %s" %
                (args.batch, "DIFFERENT cosmologies: seed-0 Latin hypercube of BASELINE configs[4] (omega_b, omega_cdm, h, ln10^10A_s, n_s, "
                             "tau_reio)" if not args.identical else "copies of the Planck-18 best fit (identical cosmologies)"))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
           "steps": len(vals), "warmup": n_warm, "ms_per_step": float(np.mean([v["hot_path_s_max"] for v in vals])) * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload if args.config == "planck18" else args.config, "batch_per_gpu": int(args.batch),
                      "sample_cosmologies_per_step": int(cores),
                      "note": "unmodified CLASS++ (oracle/_ref) on the host cores, throughput arrangement: each step = a bounded "
                              "sample of the workload, its first %d cosmologies as %d single-threaded processes; timed = "
                              "Perturbations+Nonlinear+Transfer+Spectra+Lensing module constructors (the scope of the GPU arm's "
                              "step); value = cosmologies / slowest process" % (cores, cores)},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(cores), "kind": "reference",
                            "latency_arrangement": {"spectra_per_s": 1.0 / lat, "hot_path_s": lat, "threads": int(cores)},
                            "throughput_arrangement": vals[-1],
                            "sample": "%d steps of %d cosmologies each, one single-threaded process per cosmology" % (len(vals), cores)},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="planck18")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("CLPP_BENCH_BATCH", 128)),
                    help="cosmologies per GPU and per step (one batched perturbation launch)")
    ap.add_argument("--sets", type=int, default=int(os.environ.get("CLPP_BENCH_SETS", 2)),
                    help="context sets = batched launches in flight (consecutive steps overlap)")
    ap.add_argument("--identical", action="store_true", help="batch = copies of the Planck-18 best fit instead of the Latin hypercube")
    ap.add_argument("--path", default="auto", choices=["auto", "lane", "warp"], help="perturbation kernels (auto: by batch size)")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-cosmology latency measurement")
    ap.add_argument("--ref-worker", type=int, default=-1, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stagger", type=float, default=-1.0,
                    help="pipeline: >= 0 lets the batched launches of consecutive steps overlap, this fraction of a solo "
                         "launch duration apart (useful with --path lane; default: one launch at a time, only the "
                         "per-cosmology stages overlap it)")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="one set of contexts: the per-cosmology stages of a step finish before the next step starts")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        # the contract is ONE JSON line on stdout: libraries (NCCL's version banner) write to the stdout file descriptor
        # directly, so everything but that line is sent to stderr at the descriptor level
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            run_gpu(args)
        finally:
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
            os.close(real_stdout)
            if _RESULT_LINE:
                print(_RESULT_LINE[0], flush=True)


if __name__ == "__main__":
    main()
