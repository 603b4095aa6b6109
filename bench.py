#!/usr/bin/env python
"""Benchmark of the AGQ hot path: fused log-likelihood + score evaluations per second.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one fused evaluation (logLik_mixed value + full score_mixed vector,
reference R/Fit_Funs.R:1-293) of the BASELINE.json config-3 workload -- negative
binomial, random intercepts + slopes (q = 2), nAGQ = 11 (K = 121 nodes), 1e6 clusters x 8
observations -- on a fixed quadrature grid, clusters sharded over the N GPUs ("strong"
scaling: the metric is quoted at 1M clusters for every N).  Prints ONE JSON line with
  value         evaluations/s, data resident in HBM, CUDA events on the launching stream, max over ranks
  e2e           the same through the C ABI with host theta in / host result out per step (+ cold-start figure)
  roofline      dominant kernel vs the FP64 peak measured in the same run (flops by SURVEY.md 8(d) convention
                v1, for the separable algorithm and for the dense grid), HBM sub-object, ncu dram traffic
  cpu_baseline  the numpy port of the R code on one host core (bounded sample, extrapolated linearly)
  mixed_model   wall seconds of a whole fit (EM + BFGS + Hessian, default control) on the same data
--workload C2|C4|C5a|C5b and --clusters select the other BASELINE configurations (profiles/*.jsonl).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "loglik+score evals/s (1M clusters, nAGQ=11)"
UNIT = "evals/s"

# SURVEY.md section 8(d), flop convention v1 (dense grid): per-point family cost incl. score
F_PT = {"negative.binomial": 95, "poisson": 36, "binomial": 78, "zi.poisson": 137,
        "censored.normal": 57, "beta.fam": 460}
F_OBS_SPECIALS = {"negative.binomial": 400, "poisson": 100, "binomial": 0, "zi.poisson": 100,
                  "censored.normal": 0, "beta.fam": 200}


def algorithmic_flops(family, n, N, K, p, q_y, q_zi, p_zi):
    q = q_y + q_zi
    f_node = 2 * q * q + 3 * q + 2 * q * q + 30 + 6
    f_obs = 2 * (p + p_zi) + F_OBS_SPECIALS[family]
    return N * K * (F_PT[family] + 2 * q_y + 2 * q_zi) + n * K * f_node + N * f_obs


def needed_flops(family, n, N, K, nagq, p, q_y, q_zi, p_zi):
    """Flops the engine's ALGORITHM needs under the same convention ('if the builder
    uses the separable-exp shortcut it must report both ... and compute roofline.achieved from the
    latter').  Relative to the dense grid: the per-point exp(eta) (30) becomes a product of q table
    factors (q - 1 multiplies), eta itself is no longer formed per point for the families whose
    log-density is linear in y*eta (2 q_y), and q * nAGQ table exponentials (30 + 2 each) are added
    per observation."""
    q = q_y + q_zi
    dense = algorithmic_flops(family, n, N, K, p, q_y, q_zi, p_zi)
    if family not in ("negative.binomial", "poisson", "binomial", "zi.poisson", "beta.fam"):
        return dense                      # no exp(eta) in the family: nothing to separate
    P = N * K
    lin_eta = family in ("negative.binomial", "poisson", "binomial", "zi.poisson")
    return dense - 30 * P - (2 * q_y * P if lin_eta else 0) + (q - 1) * P + N * q * nagq * 32


def load_traffic(kernel_prefix):
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/traffic.json, written by profiles/summarise.py); None if there is no capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        for rec in t["kernels"]:
            if rec["kernel"].startswith(kernel_prefix):
                return rec
    except Exception:
        pass
    return None


def algorithmic_bytes(n, N, p, q_y, q_zi, p_zi, n_phis, c_y=1):
    q = q_y + q_zi
    return 8 * N * (c_y + p + q_y + p_zi + q_zi) + n * (4 + 8 * q + 8 * q * (q + 1) // 2) + \
        8 * (2 + p + n_phis + p_zi + q * q)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_eval_time(d, th, n_sample, reps=1):
    """Times the CPU oracle (numpy restatement of the reference's N x K vectorised R code):
    logLik_mixed + score_mixed, each recomputing everything as optim's fn/gr do
    (R/Fit_Funs.R:1-42 and :44-293).  Returns seconds per evaluation on n_sample clusters."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import agq
    from helpers import oracle_family, thetas_vec
    rp = np.concatenate(([0], np.flatnonzero(np.diff(d["id"])) + 1, [len(d["id"])]))
    r1 = int(rp[n_sample])
    od = agq.Data(y=d["y"][:r1], X=d["X"][:r1], Z=d["Z"][:r1], id=d["id"][:r1])
    fam = oracle_family(d["family"])
    q = od.nRE
    # grid at the data-generating theta: identity-ish Hessians are enough for timing, but use
    # a real grid so that the arithmetic is representative
    invD = np.linalg.inv(th["D"])
    modes, hess, _ = agq.find_modes_newton(np.zeros((min(od.n, 200), q)), _head(od, 200), fam, th["betas"], invD,
                                           th["phis"], th["gammas"])
    H = np.mean(np.array(hess), axis=0)
    gh = agq.GHfun_from_modes(od, np.zeros((od.n, q)), [H] * od.n, th["nAGQ"], q)
    tv = thetas_vec(th)
    best = np.inf
    for _ in range(reps):
        t0 = time.perf_counter()
        agq.logLik_mixed(tv, od, fam, gh)
        agq.score_mixed(tv, od, fam, gh)
        best = min(best, time.perf_counter() - t0)
    return best


def _head(od, k):
    from oracle import agq
    k = min(k, od.n)
    r1 = int(od.row_ptr[k])
    return agq.Data(y=od.y[:r1], X=od.X[:r1], Z=od.Z[:r1], id=od.id[:r1])


def _oracle_worker(args):
    """One pool process: pinned to its own core, single-threaded numpy, its OWN slice of clusters."""
    seed, n_sample, core = args
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    from glmmadaptive_b200 import synth
    d, th = synth.make("C3", n=n_sample, ni=8, seed=seed)
    return oracle_eval_time(d, th, n_sample)


R_TIMING_SCRIPT = os.path.join(ROOT, "oracle", "time_reference.R")


def find_reference_tree():
    """The unmodified reference sources, if they travelled to this machine (they do not travel to the
    GPU box: /root/reference exists in the build container only)."""
    for cand in (os.environ.get("GLMMA_REFERENCE_DIR"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "R", "Fit_Funs.R")):
            return cand
    return None


def try_real_reference(n_sample):
    """when Rscript (with matrixStats) AND the reference tree are present, time the
    reference's own logLik_mixed + score_mixed (R/Fit_Funs.R, sourced unmodified by
    oracle/time_reference.R) on the C3 workload.  Returns (seconds per evaluation on n_sample clusters,
    description) or None."""
    import shutil
    import tempfile
    rs, ref = shutil.which("Rscript"), find_reference_tree()
    if rs is None or ref is None:
        return None
    try:
        if subprocess.run([rs, "-e", "library(matrixStats)"], capture_output=True, timeout=120).returncode != 0:
            return None
        from glmmadaptive_b200 import synth
        d, th = synth.make("C3", n=n_sample, ni=8, seed=3)
        with tempfile.TemporaryDirectory() as tmp:
            for k in ("y", "X", "Z"):
                np.asarray(d[k], np.float64).ravel(order="F").tofile(os.path.join(tmp, k + ".bin"))
            np.asarray(d["id"] + 1, np.int32).tofile(os.path.join(tmp, "id.bin"))
            theta = np.concatenate([th["betas"], np.asarray(th["D"]).ravel(order="F"), th["phis"]])
            theta.tofile(os.path.join(tmp, "theta.bin"))
            r = subprocess.run([rs, R_TIMING_SCRIPT, ref, tmp, str(len(d["y"])), str(d["X"].shape[1]),
                                str(d["Z"].shape[1]), str(th["nAGQ"])], capture_output=True, text=True, timeout=1800)
            if r.returncode != 0:
                return None
            secs = float([ln for ln in r.stdout.splitlines() if ln.startswith("SECONDS_PER_EVAL")][-1].split()[1])
        return secs, f"unmodified {ref}/R/*.R under {rs}: logLik_mixed + score_mixed on {n_sample} clusters, one core"
    except Exception:
        return None


def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores.

    If Rscript and the reference tree are both on this machine the unmodified R code is timed
    (kind "reference").  Otherwise -- this image has no R, and the reference sources do not travel to
    the GPU box -- the numpy port of the R code is timed on ALL host cores (kind "port"): every pool
    process is pinned to its own core, runs single-threaded numpy on its own slice of distinct clusters,
    and a step takes as long as the slowest process.  The port is timed at three slice sizes (linearity
    check: the figure is extrapolated linearly to 1e6 clusters) and the reported value is the MEDIAN over
    the timed steps at the largest size, with the spread alongside."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    n_total = 1_000_000
    real = try_real_reference(4000)
    try:
        avail = sorted(os.sched_getaffinity(0))
    except Exception:
        avail = list(range(os.cpu_count() or 1))
    cores = max(1, min(len(avail), 16))
    sizes = [1000, 2000, 4000]
    steps = max(args.steps, 5)
    med = {}
    spread = None
    with ProcessPoolExecutor(max_workers=cores) as ex:
        for m in sizes:
            k = steps if m == sizes[-1] else 3
            times = []
            for it in range(args.warmup + k):
                jobs = [(1000 + w, m, avail[w]) for w in range(cores)]        # distinct clusters per process
                dt = max(ex.map(_oracle_worker, jobs))
                if it >= args.warmup:
                    times.append(dt)
            med[m] = float(np.median(times))
            if m == sizes[-1]:
                spread = [float(np.min(times)), float(np.max(times))]
    m = sizes[-1]
    n_sample = m * cores
    t_step = med[m]
    value = 1.0 / (t_step * n_total / n_sample)
    kind = "port"
    sample = (f"{n_sample} clusters x 8 obs per step ({m} distinct clusters per process x {cores} pinned, "
              f"single-threaded processes), logLik_mixed + score_mixed of the numpy port; median of {steps} steps "
              f"(min {spread[0]:.3f} s, max {spread[1]:.3f} s); evals/s extrapolated linearly to 1e6 clusters")
    extra = {"seconds_per_step_by_slice": {str(k): v for k, v in med.items()},
             "linearity": {"t(2000)/t(1000)": med[2000] / med[1000], "t(4000)/t(2000)": med[4000] / med[2000],
                           "note": "2.0 = linear in the number of clusters"},
             "spread_s": spread}
    if real is not None:
        secs, what = real
        value_r = 1.0 / (secs * n_total / 4000)
        extra["port_all_cores"] = {"value": value, "sample": sample}
        value, kind, cores, sample = value_r, "reference", 1, what + "; extrapolated linearly to 1e6 clusters"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": "C3: negative.binomial, q=2, nAGQ=11 (K=121), 1e6 clusters x 8 obs"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, **extra},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "Rscript probed: " + ("found, unmodified reference timed" if real is not None else
                                         ("not found" if __import__("shutil").which("Rscript") is None else
                                          "found, but the reference tree or matrixStats is missing") +
                                         "; the reference is pure R, so its numpy port is timed instead")}
    print(json.dumps(out), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import glmmadaptive_b200 as gb
    from glmmadaptive_b200 import synth, sharded
    from glmmadaptive_b200.engine import measure_fp64_peak

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    n_total = args.clusters
    d, th = synth.make(args.workload, n=n_total, ni=args.ni, seed=None, ragged=args.ragged)
    fam_name = d.pop("family")
    cfg = synth.CONFIGS[args.workload]
    ni = cfg["ni"] if args.ni is None else args.ni
    arrays = {k: d.get(k) for k in ("y", "X", "Z", "X_zi", "Z_zi")}
    loc, (c0, c1), (r0, r1) = sharded.shard_arrays(arrays, d["id"], rank, world)
    stream = torch.cuda.current_stream()

    t0 = time.perf_counter()
    eng = gb.Engine(loc["y"], loc["X"], loc["Z"], loc["id"], fam_name, nAGQ=th["nAGQ"],
                    X_zi=loc.get("X_zi"), Z_zi=loc.get("Z_zi"), device=local_rank)
    eng.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()
    t_create = time.perf_counter() - t0
    sh = sharded.ShardedEngine(eng, dist if world > 1 else None)
    invD = np.linalg.inv(th["D"])
    betas, D, phis, gammas = th["betas"], th["D"], th["phis"], th["gammas"]

    # grid: the engine's own Newton search at the data-generating theta (timed separately)
    t0 = time.perf_counter()
    st = eng.find_modes(betas, invD, phis, gammas, warm=False)
    torch.cuda.synchronize()
    t_modes_cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng.find_modes(betas, invD, phis, gammas, warm=True)
    torch.cuda.synchronize()
    t_modes_warm = time.perf_counter() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        sh.eval_enqueue(betas, D, phis, gammas, True)
    barrier()

    # ---- device-resident throughput: K fused evaluations, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        buf = sh.eval_enqueue(betas, D, phis, gammas, True)
    e1.record(stream)
    barrier()
    launches = eng.launch_count - l0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    res = eng.unpack(buf.cpu().numpy())

    # ---- end to end through the C ABI with host buffers: theta in, result vector out, sync per step
    barrier()
    tv = np.concatenate([betas, gb.chol_transf(D), phis if phis is not None else [], gammas if gammas is not None else []])
    t0 = time.perf_counter()
    for s in range(args.steps):
        r = sh.eval(betas, D, phis, gammas, True)
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = args.steps / float(t_e2e.item())
    h2d = 8 * (len(betas) + D.size + (0 if phis is None else len(phis)) + (0 if gammas is None else len(gammas)))
    d2h = 8 * eng.result_len

    # ---- second half of BASELINE's metric: wall seconds of a whole mixed_model() fit (default
    # control: 30 EM iterations, then optim-BFGS outer iterations, then the cd_vec Hessian) on the
    # same data, clusters sharded over the ranks, starting values from glm.fit run on the device (IRLS passes)
    fit_info = None
    if not args.no_fit:
        barrier()
        t0 = time.perf_counter()
        inits = gb.starting_values_device(sh)
        t_inits = time.perf_counter() - t0
        l0 = eng.launch_count
        fit = gb.mixed_fit(sh, dict(inits), {}, hessian=True)
        barrier()
        t_first = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local_rank}")
        # The first fit of a process also pays CUDA's lazy loading of every kernel it touches for the first time
        # (mode search, E-step / M-step forms, observed information: a few tenths of a second, independent of the
        # data).  `seconds` is a second, identical fit (same starting values, cold grid: modes reset to zero);
        # `first_fit_seconds` is the first one.
        barrier()
        t1 = time.perf_counter()
        l0 = eng.launch_count
        fit = gb.mixed_fit(sh, dict(inits), {}, hessian=True)
        barrier()
        t_fit = torch.tensor([time.perf_counter() - t1 + t_inits], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(t_fit, op=dist.ReduceOp.MAX)
            dist.all_reduce(t_first, op=dist.ReduceOp.MAX)
        fit_info = {"seconds": float(t_fit.item()), "first_fit_seconds": float(t_first.item()),
                    "starting_values_s": t_inits, "phases_s": fit["seconds"],
                    "converged": bool(fit["converged"]), "outer_iterations": fit["n_outer"],
                    "calls": fit["counts"], "gpu_launches": eng.launch_count - l0,
                    "logLik": fit["logLik"], "coefficients": [float(v) for v in fit["coefficients"]],
                    "true_betas": [float(v) for v in betas],
                    "what": "glmmadaptive_b200.mixed_fit (host mirror of R/mixed_fit.R with default control) on "
                            "the resident data set; excludes data generation and glmma_create (create_s)"}

    # ---- kernel-level timing of the dominant kernel (rank-local) and the roofline
    ms_eval, ms_kernel = eng.time_eval(betas, D, phis, gammas, True, iters=max(3, min(args.steps, 10)))
    barrier()

    # ---- the single-process multi-GPU handle (glmma_create_multi: the shape an R session needs) checked
    # where several GPUs are visible: rank 0 spreads a 3000-cluster sample over all of them behind ONE
    # handle and compares its evaluation with a single-device engine on the same sample
    multi_check = None
    if rank == 0 and torch.cuda.device_count() > 1:
        try:
            ndev = torch.cuda.device_count()
            dm, thm = synth.make(args.workload, n=3000, ni=args.ni, seed=17, ragged=True)
            fm = dm.pop("family")
            kw = dict(nAGQ=thm["nAGQ"], X_zi=dm.get("X_zi"), Z_zi=dm.get("Z_zi"))
            one = gb.Engine(dm["y"], dm["X"], dm["Z"], dm["id"], fm, device=local_rank, **kw)
            one.find_modes(thm["betas"], np.linalg.inv(thm["D"]), thm["phis"], thm["gammas"], warm=False)
            m_, c_, _ = one.get_modes()
            r1_ = one.eval_raw(thm["betas"], thm["D"], thm["phis"], thm["gammas"])
            many = gb.Engine(dm["y"], dm["X"], dm["Z"], dm["id"], fm, device=list(range(ndev)), **kw)
            many.set_modes(m_, c_)
            rN = many.eval_raw(thm["betas"], thm["D"], thm["phis"], thm["gammas"])
            stN = many.find_modes(thm["betas"], np.linalg.inv(thm["D"]), thm["phis"], thm["gammas"], warm=False)
            mN, _, _ = many.get_modes()
            err = float(np.max(np.abs(rN - r1_) / np.maximum(1.0, np.abs(r1_))))
            # different shards sum in a different order: the rounding of every sum is bounded by 1e-13 of the LARGEST
            # component (the phi-score is the cancelled difference of two sums of the size of the log-likelihood)
            err_scaled = float(np.max(np.abs(rN - r1_)) / np.max(np.abs(r1_)))
            multi_check = {"devices": ndev, "clusters": 3000, "max_rel_diff_vs_single_device": err,
                           "max_diff_over_largest_component": err_scaled,
                           "modes_max_abs_diff": float(np.max(np.abs(mN - m_))), "ok": bool(err_scaled < 1e-13 and err < 1e-9),
                           "not_converged": int(stN["n_not_converged"]),
                           "what": "glmma_create_multi handle over all visible GPUs in ONE process vs a single-device "
                                   "engine: same grid (set_modes), glmma_eval result vector; then its own mode search"}
            one.close(); many.close()
        except Exception as exc:                      # never let the check break the benchmark line
            multi_check = {"error": repr(exc)}
    barrier()
    out = None
    if rank == 0:
        peak = measure_fp64_peak(local_rank)
        N_loc, n_loc = eng.N, eng.n
        dense = algorithmic_flops(fam_name, n_loc, N_loc, eng.K, eng.p, eng.q_y, eng.q_zi, eng.p_zi)
        flops = needed_flops(fam_name, n_loc, N_loc, eng.K, eng.nAGQ, eng.p, eng.q_y, eng.q_zi, eng.p_zi)
        byts = algorithmic_bytes(n_loc, N_loc, eng.p, eng.q_y, eng.q_zi, eng.p_zi, eng.n_phis,
                                 c_y=2 if fam_name == "censored.normal" else 1)
        achieved = flops / (ms_kernel * 1e-3) / 1e12
        traffic = load_traffic("eval_fast_kernel")
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak,
                    "traffic": None if traffic is None else traffic["dram_bytes_per_launch"] * n_loc / traffic["clusters"],
                    "traffic_source": None if traffic is None else
                    f"{traffic['source']}: dram__bytes_read.sum + dram__bytes_write.sum of one launch at "
                    f"{traffic['clusters']} clusters, scaled linearly to this launch's {n_loc} clusters",
                    "kernel": f"eval_fast_kernel<{fam_name}, q_y={eng.q_y}, q_zi={eng.q_zi}, score, NJ=8> "
                              "(+ eval_kernel for deferred clusters; both between the two events)",
                    "kernel_ms": ms_kernel,
                    "ncu": None if traffic is None else {k: traffic.get(k) for k in
                                                         ("fp64_pipe_busy_pct", "issue_slots_busy_pct",
                                                          "shared_mem_pipe_pct", "registers_per_thread")},
                    "ncu_note": "hardware counters of the committed ncu capture (profiles/): the FP64 pipe is busy "
                                "that share of cycles; `frac` credits the convention's flop weights (log = 30, "
                                "division = 10 ...) to a kernel that spends 12 FP64 instructions on log and 1/t "
                                "together, so it is higher than the pipe utilisation (profiles/README.md)",
                    "flops_per_launch": flops,
                    "flop_convention": "SURVEY.md 8(d) v1 weights applied to the separable-exp algorithm the engine "
                                       "runs (bench.py needed_flops); dense-grid figure alongside",
                    "dense_grid": {"flops_per_launch": dense, "achieved": dense / (ms_kernel * 1e-3) / 1e12,
                                   "frac": dense / (ms_kernel * 1e-3) / 1e12 / peak},
                    "why_fp64": "arithmetic intensity ~190 flop/B vs a ridge of ~5.6 flop/B: the HBM roofline "
                                "(sub-object hbm) is two orders of magnitude away; tensor cores do not apply "
                                "(no contraction with inner dimension > q = 2)",
                    "peak_source": "DFMA microbenchmark measured in this run (glmma_measure_fp64_peak); "
                                   "MEASURED_PEAKS.json has no FP64 entry",
                    "hbm": {"achieved": byts / (ms_kernel * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": byts / (ms_kernel * 1e-3) / 1e9 / hbm_peak, "bytes_per_launch": byts,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
        cpu_fit = None
        if world == 1 and not args.no_cpu_baseline and not args.no_fit:
            # the CPU side of "mixed_model() s" -- the oracle's restatement of
            # R/mixed_fit.R (default control: 30 EM iterations, BFGS, cd_vec Hessian) on a sample the
            # port finishes in seconds, with the engine's fit of the SAME sample beside it
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from oracle import agq as _agq, fit as _ofit
            from helpers import oracle_family as _of
            n_f = 1000
            df_, thf = synth.make(args.workload, n=n_f, ni=args.ni, seed=23)
            ff = df_.pop("family")
            t0 = time.perf_counter()
            try:
                rf = _ofit.mixed_fit(_agq.Data(y=df_["y"], X=df_["X"], Z=df_["Z"], id=df_["id"], X_zi=df_.get("X_zi"),
                                               Z_zi=df_.get("Z_zi")), _of(ff), mode_finder="newton", hessian=True)
                t_cpu = time.perf_counter() - t0
                t0 = time.perf_counter()
                gf = gb.mixed_model(df_["y"], df_["X"], df_["Z"], df_["id"], ff, X_zi=df_.get("X_zi"), Z_zi=df_.get("Z_zi"),
                                    device=local_rank)
                t_gpu = time.perf_counter() - t0
                cpu_fit = {"clusters": n_f, "cpu_port_seconds": t_cpu, "engine_seconds": t_gpu, "cores": 1,
                           "kind": "port", "max_abs_diff_coefficients": float(np.max(np.abs(rf["coefficients"] - gf["coefficients"]))),
                           "what": "whole mixed_model() fit (default control) of the same 1000-cluster sample: numpy "
                                   "restatement of R/mixed_fit.R on one core vs the engine (incl. glmma_create); at "
                                   "this size the engine is launch-latency bound, the CPU time grows linearly with n"}
            except Exception as exc:
                cpu_fit = {"error": repr(exc)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            n_s = min(60000, n_total)           # ~6 s per pass on one core, two passes
            dd = dict(d); dd["family"] = fam_name
            t_s = oracle_eval_time(dd, th, n_s, reps=2)
            cpu = {"value": 1.0 / (t_s * n_total / n_s), "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"first {n_s} clusters of the same workload, logLik_mixed + score_mixed of the numpy port "
                             f"of the R code ({t_s:.2f} s), extrapolated linearly to {n_total} clusters; "
                             f"host has {os.cpu_count()} cores, the R path is single-threaded"}
        out = {"metric": METRIC, "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"{args.workload}: {fam_name}, q={eng.q}, nAGQ={eng.nAGQ} (K={eng.K}), "
                                      f"{n_total} clusters x {ni} obs{' (ragged: 1 + Poisson)' if args.ragged else ''}, "
                                      f"clusters sharded over {world} GPU(s)",
                          "l2": "inputs larger than L2 (0.43 GB data + work arrays per evaluation)",
                          "cluster_evals_per_s": n_total * 1e3 / ms_step,
                          "point_evals_per_s": n_total * ni * eng.K * 1e3 / ms_step},
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "what": ("glmma_eval through the C ABI: host theta in (kernel parameters), result vector "
                                "copied back and synchronised every step" if world == 1 else
                                "glmma_eval_device through the C ABI on every rank (host theta in), one all-reduce of "
                                "the result vector, device-to-host copy and synchronisation every step") +
                               "; the data set is uploaded once by glmma_create, as the R glue does",
                       "create_s": t_create, "upload_bytes": eng.input_bytes,
                       "cold_first_eval_s": t_create + t_modes_cold + 1.0 / e2e_value,
                       "cold_what": "from host arrays to the first result: glmma_create (host->device copy of the "
                                    "whole data set, upload_bytes) + cold mode search + one evaluation"},
               "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
               "find_modes": {"cold_clusters_per_s": n_total / t_modes_cold, "warm_clusters_per_s": n_total / t_modes_warm,
                              "mean_iterations_cold": st["mean_iterations"], "not_converged": st["n_not_converged"]},
               "mixed_model": fit_info, "mixed_model_cpu": cpu_fit, "multi_handle_check": multi_check,
               "ms_per_eval_local": ms_eval,
               "check": {"loglik": res["loglik"], "n_nonfinite": res["n_nonfinite"]}}
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--clusters", type=int, default=1_000_000)
    ap.add_argument("--ni", type=int, default=None, help="observations per cluster (default: the configuration's)")
    ap.add_argument("--ragged", action="store_true", help="cluster sizes 1 + Poisson(ni - 1) instead of fixed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fit", action="store_true", help="skip the whole-fit (mixed_model() seconds) leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
