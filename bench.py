#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native CLASS++ hot path.

Metric (BASELINE.json): lensed-C_l spectra / s for the Planck-2018 LambdaCDM configuration
(base_2018_plikHM_TTTEEE_lowl_lowE_lensing.ini: 1 massive neutrino species, halofit,
l_max_scalars=2500, P(k) to 1 h/Mpc), measured on the hot path this repository replaces:
PerturbationsModule -> (halofit) -> TransferModule -> SpectraModule -> LensingModule + P(k)
(SURVEY.md section 8 with its two "next" rows).

A "step" = one pass of the hot path over one batch of `--batch` cosmologies per GPU: ONE batched
perturbation launch, then the per-cosmology stages.  Steps are scheduled by sweep.SweepPipeline on two
sets of contexts (per-cosmology stages of step i under the launch of step i+1; --no-pipeline: strictly
one step after the other); everything is drained before the clock stops.  Upstream inputs
(background/thermodynamics tables, ncdm grids, primordial spectrum) are synthetic-by-construction: they
were generated once from the reference for the named configuration and are stored in
tests/golden/planck18.npz.  The reference arm times the same five module constructors of the unmodified
reference (oracle/_ref) on all host cores.

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


def hot_path(M, inp, ctx, bg, th, pk, nl, fetch_tables=False):
    """One cosmology through the three stages on an existing context."""
    pt = M.PerturbationsModule(inp, bg, th)
    tr = M.TransferModule(inp, bg, th, pt, nl)
    sp = M.SpectraModule(inp, pt, M.TabulatedPrimordial(pk), nl, tr)
    out_bytes = sp.cl_[0].nbytes
    if fetch_tables:  # what the reference-facing drop-in hands back as public members
        out_bytes += sum(s.nbytes for s in pt.sources_[0]) + tr.transfer_[0].nbytes
    return pt, tr, sp, out_bytes


# DRAM bytes one batched perturbation launch moves per cosmology (ncu, see roofline.traffic_source)
TRAFFIC_BYTES_PER_COSMOLOGY = {"planck18": 36.44e6}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from classpp_public_b200 import modules as M

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

    inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", args.config + ".npz"))
    a = inp.arrays
    pk = a["pm.pk_at_transfer_k"]
    nl = _NL(a["nl.nl_corr_density_m"]) if "nl.nl_corr_density_m" in a else None
    # `non linear = halofit` (config 2): by default the halofit step runs on the device between stages 1 and 2
    #  instead of taking the reference's correction table as an input
    from classpp_public_b200.configs import CONFIGS
    halofit_on_device = nl is not None and args.halofit == "device"
    prim_k = M.AnalyticPrimordial(CONFIGS[args.config].get("A_s", 2.215e-9), CONFIGS[args.config].get("n_s", 0.9619))
    B = args.batch
    prim_pt = np.ascontiguousarray(prim_k.pk_at_k(a["ref.k"]))

    # ---- device-resident inputs: one context (own stream) per cosmology of the batch. Two sets of contexts, used by
    # alternate steps: the per-cosmology stages of step i (halofit .. P(k), short kernels and host work) run while the
    # batched perturbation launch of step i+1 already occupies the GPU; everything is drained before the clock stops.
    NSET = 1 if args.no_pipeline else 2
    sets = []
    for s_ in range(NSET):
        cs_, ms_ = [], []
        for b in range(B):
            ctx = M.Context(local)
            bg = M.BackgroundModule(inp, ctx)
            th = M.ThermodynamicsModule(inp, bg)
            cs_.append(ctx)
            ms_.append((bg, th))
        sets.append((cs_, ms_))
    ctxs = [c for cs_, _ in sets for c in cs_]

    results = [None] * B
    import threading
    from classpp_public_b200.sweep import SweepPipeline
    kms = {}
    kms_lock = threading.Lock()

    def kms_reset():
        for k_ in ("perturb", "k_spline", "bessel", "los", "spectra", "perturb_tail", "halofit", "lensing"):
            kms[k_] = 0.0

    kms_reset()

    def kms_add(ctx, keys):
        t = ctx.kernel_ms()
        with kms_lock:
            for k_ in keys:
                kms[k_] += t[k_]

    # One pass of the hot path over a batch ("step"): every k mode of the B cosmologies in ONE perturbation launch
    # (longest modes first across the batch), then halofit, transfer, spectra, lensing and P(k) per cosmology on its own
    # stream. With `inputs` (pinned host arrays) the upstream tables are uploaded first and the public result members
    # (sources_, cl_, cl_lens_, P(k)) are read back: the end-to-end variant.
    def front(s_, b, inputs, p_, n_, fetch):
        cs, ms = sets[s_]
        if inputs is not None:  # host -> device copy of this step's inputs
            bg = M.BackgroundModule(inputs, cs[b])
            ms[b] = (bg, M.ThermodynamicsModule(inputs, bg))
        return M.PerturbationsModule(inputs or inp, ms[b][0], ms[b][1], solve=False)

    def on_solved(s_):
        for c in sets[s_][0]:
            kms_add(c, ("perturb", "perturb_tail"))

    def back(s_, b, pt, inputs, p_, n_, fetch):
        cs, ms = sets[s_]
        x = inputs or inp
        nlb = n_
        if halofit_on_device:  # NonlinearModule (halofit) on the device, from the resident delta_m sources
            nlb = M.NonlinearModule(x, ms[b][0], pt, prim_k)
        tr = M.TransferModule(x, ms[b][0], ms[b][1], pt, nlb)
        sp = M.SpectraModule(x, pt, M.TabulatedPrimordial(p_), nlb, tr)
        le = M.LensingModule(x, sp)  # lensed TT/TE/EE/BB on the device: the metric's "lensed C_l"
        pk_lin = pt.pk_linear(prim_pt)  # linear P(k, z=0) on the perturbation k grid
        out_bytes = sp.cl_[0].nbytes + le.cl_lens_.nbytes + pk_lin.nbytes
        if fetch:  # device -> host: the public members downstream modules read (Nonlinear/Lensing/Output)
            out_bytes += sum(s__.nbytes for s__ in pt.sources_[0])
        kms_add(cs[b], ("k_spline", "bessel", "los", "spectra", "halofit", "lensing"))
        results[b] = (pt, tr, sp, out_bytes)
        return out_bytes

    pipe = SweepPipeline(B, front, back, n_sets=NSET, on_solved=on_solved)

    def step(inputs=None, pk_=None, nl_=None, fetch=False):
        pipe.submit(inputs, pk if pk_ is None else pk_, nl if nl_ is None else nl_, fetch)
        if NSET == 1:
            pipe.drain()

    drain = pipe.drain

    def barrier():
        drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w_ in range(args.warmup):
        step()
        if w_ == 0:  # the first launch runs alone: its duration sets the stagger between overlapping launches
            drain()
            if NSET > 1 and args.stagger >= 0:
                pipe.stagger = args.stagger * pipe.solve_seconds[-1]
    drain()
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
    if os.environ.get("CLPP_VERBOSE"):
        print("launch log (set, start, end):", [(s_, round(a_ - t0, 2), round(b_ - t0, 2)) for s_, a_, b_ in pipe.launch_log[-args.steps:]], file=sys.stderr)
    kms_timed = dict(kms)  # frozen: the end-to-end passes below keep accumulating into the live dict
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
    n_cosmo = B * args.steps * world
    value = n_cosmo / elapsed

    # ---- end to end through the public (reference-facing) API from pinned HOST buffers
    def pinned(x):
        return torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()
    inp_h = M.Inputs(inp.meta, {k: (pinned(v) if v.dtype == np.float64 and v.size > 64 else v) for k, v in a.items()})
    pk_h = pinned(pk)
    nl_h = _NL(pinned(nl.nl_corr_density_m)) if nl is not None else None
    h2d = sum(inp_h.arrays[k].nbytes for k in ("bg.tau_table", "bg.background_table", "th.z_table",
                                               "th.thermodynamics_table")) * 2  # tables + their spline tables
    # the halofit table is an input only with --halofit input; on the device it is computed from the resident sources
    h2d += (nl_h.nl_corr_density_m.nbytes if (nl_h is not None and not halofit_on_device) else 0) + pk_h.nbytes
    e2e_steps = max(1, min(args.steps, 3))
    h2d *= B
    step(inp_h, pk_h, nl_h, fetch=True)  # one untimed pass: first-touch allocations of the result buffers
    barrier()
    te0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    for _ in range(e2e_steps):
        step(inp_h, pk_h, nl_h, fetch=True)
    d2h = sum(r[3] for r in results)
    barrier()
    ee1.record()
    torch.cuda.synchronize()
    e2e_elapsed = max(ee0.elapsed_time(ee1) * 1e-3, time.perf_counter() - te0)
    if world > 1:
        t = torch.tensor([e2e_elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_elapsed = float(t.item())
    e2e_value = e2e_steps * B * world / e2e_elapsed

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (perturb_kernel): FP64 vector pipe
    peaks, peaks_kind = load_peaks()
    fp64_peak = ctxs[0].fp64_peak_tflops()
    n_launch = B * args.steps
    # one perturb_kernel launch integrates the whole batch: B cosmologies of algorithmic work each
    # overlapping launches (pipeline): the device time the perturbation kernels occupy is at most the timed region itself
    perturb_busy = min(kms_timed["perturb"] * 1e-3, elapsed)
    t_perturb = max(perturb_busy / args.steps, 1e-12)
    algo = ALGO_FLOP_STAGE1.get(args.config, 1.0e9) * B
    achieved = algo / t_perturb / 1e12
    tr_info = results[0][1].info
    t_los = max(kms_timed["los"] * 1e-3 / n_launch, 1e-12)
    sec_achieved = ALGO_FLOP_PER_LOS_POINT * 1.5e8 / t_los / 1e12
    roofline = {"kernel": "perturb_kernel", "bound": "fp64-vector (latency-bound in practice; neither hbm nor tensor)",
                "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "peak_source": "DFMA microbenchmark run live in bench.py (MEASURED_PEAKS.json has no FP64 entry; "
                               "its hbm_gbs=%s bf16_tflops=%s are %s)" % (peaks.get("hbm_gbs"), peaks.get("bf16_tflops"), peaks_kind),
                "traffic": TRAFFIC_BYTES_PER_COSMOLOGY.get(args.config, 0.0) * B or None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of perturb_kernel + perturb_tail_kernel from the ncu "
                                  "launch list profiles/r01_launches_v3_bench_b16.csv (bench.py --batch 16: 36.4 MB per cosmology "
                                  "per step, scaled to this batch); algorithmic: 24.3 MB of S(k,tau) written once + 6 MB of tables",
                "algorithmic_flop_per_launch": algo,
                "launch_duration_ms_avg": kms_timed["perturb"] / args.steps,
                "launches_in_flight_avg": kms_timed["perturb"] * 1e-3 / elapsed,
                "note": "achieved = algorithmic flop of the K launches / device time they occupy (= min(sum of launch "
                        "durations, timed region): launches of consecutive steps overlap on purpose)",
                "kernel_ms_per_step": {k_: v / args.steps for k_, v in kms_timed.items()},
                "kernel_share_of_step": {k_: v * 1e-3 / elapsed for k_, v in kms_timed.items()},
                "secondary": {"kernel": "los_kernel", "bound": "fp64-vector", "achieved": sec_achieved,
                              "peak": fp64_peak, "unit": "TFLOP/s", "frac": sec_achieved / fp64_peak}}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("BASELINE configs[1]: base_2018_plikHM_TTTEEE_lowl_lowE_lensing.ini (Planck-2018 best fit, "
                                "1 ncdm species, halofit, l_max_scalars=2500, P_k_max_h/Mpc=1)") if args.config == "planck18"
                               else args.config,
                   "fixture": "tests/golden/%s.npz" % args.config, "batch_per_gpu": B,
                   "halofit": ("on the device, inside the step (clpp_nonlinear_halofit)" if halofit_on_device else
                               "correction table is an input" if nl is not None else "none"),
                   "scope": "per cosmology: perturbations -> halofit -> transfer -> spectra -> lensing (fast mode) -> linear P(k)",
                   "k_modes": int(results[0][0].info.k_size), "tau_samples": int(results[0][0].info.tau_size),
                   "q_values": int(tr_info.q_size), "l_values": int(tr_info.l_size),
                   "pipeline": ("%d context sets: the per-cosmology stages of step i run under the batched launch of step "
                                "i+1 (%s); all drained before the clock stops" %
                                (NSET, "launches overlap, %.2f x solo duration apart" % args.stagger if args.stagger >= 0
                                 else "one batched launch at a time")) if NSET > 1 else "off",
                   "parallelism": "independent cosmologies per GPU (replicas, no data-path collective); per GPU all k modes "
                                  "of the batch in one batched launch: long-tail modes on a high-priority stream, the bulk in "
                                  "chunks on low-priority streams, generic kernel -> radiation-streaming tail kernel",
                   "l2_policy": "working set per step = batch x (tables 6 MB + sources 24 MB + source spline 72 MB + Bessel 25 MB "
                                "+ transfer 12 MB) >> 126 MB L2; every buffer is rewritten every step",
                   "timing": "torch.cuda.Event around K steps after device-wide synchronize, max over ranks; per-kernel "
                             "times from cudaEvents on the launching stream inside libclpp.so"},
        "k_modes_per_s": value * int(results[0][0].info.k_size),
        "wall_s": wall,
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps,
                "note": "per step and per cosmology of the batch: upstream tables from pinned host memory through "
                        "clpp_set_background/clpp_set_thermo (host spline + H2D), grids, batched perturbation launch, "
                        "halofit, transfer, spectra, lensing, P(k), D2H of sources_, cl_, cl_lens_ and P(k) (contexts and "
                        "device buffers are reused across steps)"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.config, budget_s=20.0)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(config, budget_s=20.0, threads=None):
    """The reference's own CPU implementation of the hot path (oracle/_ref = unmodified CLASS++ built
    from /root/reference) on this box's host cores, timed per module constructor by the probe."""
    from oracle import refprobe
    from classpp_public_b200.configs import CONFIGS
    if not refprobe.available():
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                "sample": "oracle/_ref not built on this box"}
    cores = threads or os.cpu_count()
    times, times3 = [], []
    t_start = time.perf_counter()
    n = 0
    while True:
        ref = refprobe.RefCosmology(CONFIGS[config], threads=cores)
        ref.compute("lensing")
        t3 = ref.scalar("time.perturb") + ref.scalar("time.transfer") + ref.scalar("time.spectra")
        t = t3 + ref.scalar("time.nonlinear") + ref.scalar("time.lensing")
        ref.close()
        n += 1
        if n > 1:  # the first run of a process is cold: discard
            times.append(t)
            times3.append(t3)
        if (time.perf_counter() - t_start > budget_s and len(times) >= 2) or len(times) >= 8:
            break
    best = min(times)
    return {"value": 1.0 / best, "unit": UNIT, "cores": int(cores), "kind": "reference",
            "hot_path_s_best": best, "hot_path_s_mean": float(np.mean(times)),
            "three_module_s_best": float(min(times3)),
            "sample": "%d full runs of the reference for this config (first discarded); value = 1 / best wall time of "
                      "the Perturbations+Nonlinear+Transfer+Spectra+Lensing module constructors (same scope as the GPU "
                      "arm; three_module_s_best = Perturbations+Transfer+Spectra alone; background/thermodynamics/"
                      "primordial excluded), thread pool = %d threads" % (n, cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import refprobe
    from classpp_public_b200.configs import CONFIGS
    if not refprobe.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (make -C oracle needs /root/reference)"}))
        return
    cores = os.cpu_count()
    ts = []
    for i in range(args.warmup + args.steps):
        ref = refprobe.RefCosmology(CONFIGS[args.config], threads=cores).compute("lensing")
        t = (ref.scalar("time.perturb") + ref.scalar("time.nonlinear") + ref.scalar("time.transfer") +
             ref.scalar("time.spectra") + ref.scalar("time.lensing"))
        ref.close()
        if i >= args.warmup:
            ts.append(t)
    total = float(np.sum(ts))
    value = len(ts) / total
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(ts) * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": args.config, "note": "unmodified CLASS++ (oracle/_ref) on the host cores; each step = "
                      "one cosmology, timed = Perturbations+Nonlinear+Transfer+Spectra+Lensing module constructors "
                      "(the scope of the GPU arm's step)"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(cores), "kind": "reference",
                            "sample": "%d steps of 1 cosmology each, thread pool = %d" % (len(ts), cores)},
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stagger", type=float, default=-1.0,
                    help="pipeline: >= 0 lets the batched launches of consecutive steps overlap, this fraction of a solo "
                         "launch duration apart (default: one launch at a time, only the per-cosmology stages overlap it)")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="one set of contexts: the per-cosmology stages of a step finish before the next step starts")
    ap.add_argument("--halofit", default="device", choices=["device", "input"],
                    help="config with halofit: run it on the device (default) or take the reference's table as input")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
