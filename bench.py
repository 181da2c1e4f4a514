#!/usr/bin/env python
"""bench.py -- node-prune-steps/s of the CHD hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun ... bench.py --gpus N ...      (one rank per GPU, NCCL)

Workload (BASELINE.json configs[3], SURVEY.md 8d "C4"): synthetic 512 nodes x 8192 samples, default
kernel list Linear+Quadratic+Gaussian.  One *step* = one backward-elimination step (activations ->
argmin -> Gram -> regression + gamma search + Z-test, reference helper_functions.py:166-176) for a
batch of target nodes per GPU; `value` = node-prune-steps per second over all GPUs (targets are
sharded over ranks, no data-path collective: weak scaling).  The headline kernel of the default list
on this data is the Gaussian mode (the one the nodes choose); linear / quadratic rates are reported
in `per_kernel`.

Timed region: inputs resident in HBM (`value`), or host buffers with H2D/D2H inside (`e2e`).  Every
step streams fresh n x n FP64 matrices (537 MB each) that exceed L2, so no flush is needed.
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

METRIC = "node_prune_steps_per_sec"
UNIT = "node-prune-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--N", type=int, default=512)
    ap.add_argument("--targets", type=int, default=0, help="targets in flight per GPU (0: two waves)")
    ap.add_argument("--kernel", default="mix", choices=["mix", "linear", "quadratic", "gaussian"],
                    help="mix = equal shares of the three kernels of the default list (the config's kernel list)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-kernel", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- work model
def credited_flops_tridiag(n):
    return (4.0 / 3.0) * n ** 3


def step_work(n, p, kernel):
    w = dict(eig=(4.0 / 3.0) * n ** 3, gram=2.0 * n * n * p, z=200.0 * n * n, solve=4.0 * n * n, grid=8e4 * n)
    w["act"] = {"linear": 2.0 * n * p, "quadratic": 2.0 * n * p * p, "gaussian": 2.0 * n * n * p}[kernel]
    return w


# --------------------------------------------------------------------------- CPU arm (oracle)
def cpu_step_time(kernel, n_s, N_s, seed=0):
    """Times one oracle prune step (reference algorithm, NumPy/LAPACK, all host threads) on an
    (n_s x N_s) sample and returns the per-component seconds."""
    from oracle import kernels as okern, regression as oreg, chd as ochd, jaxprng
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(n_s, N_s, seed)
    X = (X - X.mean(0)) / X.std(0)
    specs = okern.make_specs([("linear", 2.0, None), ("quadratic", 1.0, None), ("gaussian", 1.0, 1.0)])
    spec = {s.name: s for s in specs}[kernel]
    im = np.eye(N_s)
    am = im.copy()
    am[:, 0] = 0
    ga = X[:, 0]
    key = jaxprng.PRNGKey(0)
    t = {}
    # the reference caches exps (n x n x N) once in GaussianMode.setup (kernels.py:307-309): not timed
    exps = okern.make_exps(X, spec.l) if kernel == "gaussian" else None
    t0 = time.perf_counter()
    K = okern.gram(spec, X, X, am.sum(0), exps)
    t["gram"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    lam, V = np.linalg.eigh(K)
    t["eig"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    yb, *_ = oreg.perform_regression_and_find_gamma(K, ga, 1e-9, key)
    t["regress_total"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ochd.activations(spec, X, yb, ga, am, False, exps)
    t["act"] = time.perf_counter() - t0
    t["rest"] = max(t["regress_total"] - t["eig"], 0.0)
    return t


def cpu_extrapolate(t, kernel, n_s, N_s, n, N):
    """Scales the sampled component times to (n, N) with the reference's own cost model:
    eigh ~ n^3; Gram ~ n^2 N; grid/solve/Z ~ n (grid dominates) ; activations: linear ~ nN,
    quadratic ~ n^2 N (helper_functions.py:92-109), gaussian ~ C n^2 N (helper_functions.py:113-132 +
    kernels.py:270-278, C = N clusters)."""
    rn, rN = n / n_s, N / N_s
    act = {"linear": rn * rN, "quadratic": rn * rn * rN, "gaussian": rn * rn * rN * rN}[kernel]
    return t["eig"] * rn ** 3 + t["gram"] * rn * rn * rN + t["rest"] * rn + t["act"] * act


def cpu_sample_sizes(kernel):
    return {"linear": (2048, 32), "quadratic": (2048, 32), "gaussian": (1024, 12)}[kernel]


def kernels_of(args):
    return ["linear", "quadratic", "gaussian"] if args.kernel == "mix" else [args.kernel]


def cpu_mixed_step_seconds(args, seed=0):
    """Seconds the reference algorithm needs for one node-prune-step of every kernel in the mix at the
    config's (n, N), extrapolated from bounded samples.  Returns (mean seconds per node-step, per-kernel)."""
    per = {}
    for k in kernels_of(args):
        n_s, N_s = cpu_sample_sizes(k)
        per[k] = cpu_extrapolate(cpu_step_time(k, n_s, N_s, seed), k, n_s, N_s, args.n, args.N)
    return float(np.mean(list(per.values()))), per


def cpu_sample_text(args):
    sizes = ", ".join(f"{k}: n={cpu_sample_sizes(k)[0]}, N={cpu_sample_sizes(k)[1]}" for k in kernels_of(args))
    return (f"oracle restatement of the reference (NumPy/LAPACK syevd, exps cached as the reference does; JAX is not "
            f"installable offline): one prune step of one target per kernel on a bounded sample ({sizes}), component "
            f"times extrapolated to n={args.n}, N={args.N} with the reference's own cost model (eigh n^3, Gram n^2 N, "
            f"activations: linear nN, quadratic n^2 N, gaussian C n^2 N)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, pers = [], []
    for i in range(args.warmup + args.steps):
        sec_i, per_i = cpu_mixed_step_seconds(args, seed=i)
        if i >= args.warmup:
            vals.append(sec_i)
            pers.append(per_i)
    sec = float(np.mean(vals))
    cores = os.cpu_count()
    sample = cpu_sample_text(args)
    per_kernel = {k: 1.0 / float(np.mean([p[k] for p in pers])) for k in pers[0]}
    line = {"impl": "reference", "metric": METRIC, "value": 1.0 / sec, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "per_kernel": per_kernel},
            "e2e": {"value": 1.0 / sec, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, tb):
    return {"workload": f"C4 synthetic {args.N} nodes x {args.n} samples, kernels=[Linear,Quadratic,Gaussian(l=1)], "
                        f"prune-chain steps (helper_functions.py:166-176), chain kernel: "
                        f"{'equal shares of linear/quadratic/gaussian' if args.kernel == 'mix' else args.kernel}",
            "n_samples": args.n, "n_nodes": args.N, "kernel": args.kernel, "targets_in_flight_per_gpu": tb,
            "cache": "inputs larger than L2 (every step streams fresh n x n FP64 matrices, 537 MB each at n=8192)"}


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from computationalhypergraphdiscovery_b200 import _native
    from computationalhypergraphdiscovery_b200.synthetic import make_data

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n, N = args.n, args.N

    X, _names = make_data(n, N, seed=0)
    X = np.ascontiguousarray((X - X.mean(0)) / X.std(0))
    plan = _native.Plan(n, N, N)
    # two-stage path: every launch is batched over the targets in flight (32 per kernel state: 17 GB of work
    # matrices at n = 8192; the latency-bound kernels -- bulge chasing, panel QR -- amortise over them);
    # one-stage path: one cooperative wave
    Tb = args.targets or (min(32, max(1, N)) if plan.two_stage else plan.matrices_per_wave)
    Xd = torch.from_numpy(X).to(dev)
    cl = torch.arange(N, dtype=torch.int32, device=dev)
    descs = {"linear": _native.make_desc(0, 2.0), "quadratic": _native.make_desc(1, 1.0, alpha=1.0),
             "gaussian": _native.make_desc(2, 1.0, alpha=1.0, quad_scale=1.0, l=1.0)}

    # this rank's targets: nodes rank*Tb .. rank*Tb+Tb-1 (mod N)
    tidx = (np.arange(Tb) + rank * Tb) % N
    w0 = np.ones((Tb, N), dtype=np.uint8)
    w0[np.arange(Tb), tidx] = 0
    ga_h = np.ascontiguousarray(X[:, tidx].T)
    from computationalhypergraphdiscovery_b200 import random as chd_random
    keys_h = chd_random.split(chd_random.PRNGKey(0), N + 1)[1:][tidx].astype(np.uint32)

    s1_times = []

    def setup_state(kernel):
        """Stage-1 regression of the targets with `kernel` (untimed) -> chain state on the device."""
        desc = descs[kernel]
        w0d = torch.from_numpy(w0).to(dev)
        ga = torch.from_numpy(ga_h).to(dev)
        keys = torch.from_numpy(keys_h.view(np.int32).copy()).to(dev)
        gmin = torch.full((Tb,), 1e-9, dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = plan.gram(desc, Xd, w0d)
        r = plan.regress(K, ga, keys, gmin, mode=1, base=1e-9)   # stage 1 (MAIN:448-511): Gram + regression + lambda_min
        e1.record()
        torch.cuda.synchronize()
        s1_times.append(e0.elapsed_time(e1))
        del K
        p_cache = (dict(P=torch.empty((Tb, n, plan.ldk), dtype=torch.float64, device=dev), valid=0)
                   if kernel == "gaussian" else None)
        return dict(desc=desc, w0=w0d, ga=ga, alive=torch.ones((Tb, N), dtype=torch.uint8, device=dev),
                    yb=r["yb"], keys=r["keys"], lmin=r["lambda_min"],
                    gmin=torch.full((Tb,), 1e-9, dtype=torch.float64, device=dev), step=0, p_cache=p_cache)

    def chain_step(st):
        out = plan.prune_chain(st["desc"], Xd, cl, st["w0"], st["alive"], st["ga"], st["yb"], st["keys"], st["lmin"],
                               st["gmin"], st["step"], 1, keep_yb=True, p_cache=st["p_cache"])
        st["step"] += 1
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prof = {}

    def timed(fn, warmup, steps, sample_clocks=False):
        for _ in range(warmup):
            fn()
        sampler = ClockSampler(local) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        l0 = _native.launch_count()
        _native.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        tri_ms, tri_launches = _native.profile_read()
        prof.clear()
        prof.update(_native.profile_read_all())
        _native.profile_enable(False)
        launches = _native.launch_count() - l0
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, tri_ms, tri_launches, clocks

    # ---- device-resident throughput (value)
    knames = kernels_of(args)
    nkm = len(knames)
    sts = [setup_state(k) for k in knames]      # first call per kernel includes one-time set-up: not reported
    s1_times.clear()

    def mixed_step():
        for s_ in sts:
            chain_step(s_)

    ms, launches, tri_ms, tri_launches, clocks = timed(mixed_step, args.warmup, args.steps, True)
    total_steps = nkm * Tb * world * args.steps
    value = total_steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel, timed live with CUDA events around every launch inside the timed region
    n_mats = nkm * Tb * args.steps
    peak_tf = _native.dmma_peak_tflops(20000) if rank == 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = ("FP64 DMMA m8n8k4 issue-rate probe run in this process (chd_dmma_peak); MEASURED_PEAKS.json has no "
                "FP64 figure")
    prof_main = dict(prof)
    if plan.two_stage:
        # Stage 1 (dense -> band, NB = 32) is two DMMA GEMM kernels per panel, batched over the targets in flight:
        #   symm  Y = A22 V          2 m^2 NB flop per matrix and panel  (sum over panels: 2/3 n^3)
        #   syr2k A22 -= VW^T+WV^T   2 m^2 NB flop (lower triangle only)  (sum over panels: 2/3 n^3)
        NBW = _native.BAND
        ms_list = [n - p * NBW for p in range(0, (n - 2) // NBW + 1) if n - p * NBW >= 2]
        ms_list = [n] + [m for m in ms_list[1:]]          # panel 0 is H_{-1} on the full matrix
        gemm_flops = sum(2.0 * m * m * NBW for m in ms_list)   # per kernel, per matrix
        kern_stats = {}
        for kname in ("symm", "syr2k", "panel_qr", "wform", "chase", "band_solve", "backtransform"):
            if kname in prof_main:
                kms, kcnt = prof_main[kname]
                kern_stats[kname] = {"ms": kms, "launches": kcnt, "share_of_step": kms / ms}
        dom = max(("symm", "syr2k"), key=lambda k: prof_main.get(k, (0.0, 0))[0])
        dms, dcnt = prof_main.get(dom, (0.0, 1))
        achieved_tf = gemm_flops * n_mats / (dms * 1e-3) / 1e12 if dms else 0.0
        # algorithmic bytes: symm reads the lower triangle twice (tile + its transpose), syr2k reads + writes it once
        alg_bytes = sum(8.0 * m * m for m in ms_list) * n_mats
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "stage1_traffic.json")))
            per_matrix = tr.get(f"{dom}_n{n}_per_matrix")
            traffic = per_matrix * n_mats / max(dcnt, 1) if per_matrix else None
        except Exception:  # noqa: BLE001
            pass
        stage1_ms = sum(prof_main.get(k, (0.0, 0))[0] for k in ("symm", "syr2k", "panel_qr", "wform"))
        kfull = {"symm": "symm_kernel", "syr2k": "syr2k_row_kernel"}[dom]
        roofline = {"kernel": f"{kfull} (dense->band stage of the two-stage reduction, FP64 DMMA m8n8k4, "
                              f"batched over {Tb} targets per launch)",
                    "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic, "peak_source": peak_src,
                    "flops_per_launch": gemm_flops * n_mats / max(dcnt, 1),
                    "avg_launch_ms": dms / max(dcnt, 1), "launches_timed": dcnt, "share_of_step": dms / ms,
                    "stage1": {"credited_flops_per_matrix": credited_flops_tridiag(n),
                               "achieved_tflops": credited_flops_tridiag(n) * n_mats / (stage1_ms * 1e-3) / 1e12
                               if stage1_ms else None,
                               "frac_of_dmma_peak": (credited_flops_tridiag(n) * n_mats / (stage1_ms * 1e-3) / 1e12 /
                                                     peak_tf) if stage1_ms and peak_tf else None,
                               "ms_per_matrix": stage1_ms / n_mats,
                               "note": "symm + syr2k + panel QR + W formation, credited 4/3 n^3 (SURVEY 8d)"},
                    "kernels": kern_stats,
                    "hbm_view": {"bound": "hbm", "achieved": alg_bytes / (dms * 1e-3) / 1e9 if dms else None,
                                 "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / (dms * 1e-3) / 1e9 / hbm_peak if dms else None,
                                 "note": "8 m^2 B per panel and matrix for either GEMM kernel (symm: lower triangle "
                                         "read as tile and as transposed tile; syr2k: read + write)"}}
    else:
        mats_per_launch = n_mats / max(tri_launches, 1)
        dur = tri_ms * 1e-3 / max(tri_launches, 1)
        achieved_tf = credited_flops_tridiag(n) * n_mats / (tri_ms * 1e-3) / 1e12
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "tridiag_traffic.json")))
            per_matrix = tr.get(f"n{n}_per_matrix")
            traffic = per_matrix * mats_per_launch if per_matrix else None
        except Exception:  # noqa: BLE001
            pass
        # bytes the one-stage reduction must stream: lower triangle of the trailing matrix once per column (symv,
        # 4n^3/3 B) + read/write of it once per panel (8n^3/(3 nb) B)
        alg_bytes = (4.0 * n ** 3 / 3.0 + 8.0 * n ** 3 / (3.0 * plan.panel_width)) * mats_per_launch
        roofline = {"kernel": "tridiag_kernel (batched Householder tridiagonalisation, DMMA trailing update)",
                    "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic, "peak_source": peak_src,
                    "flops_per_launch": credited_flops_tridiag(n) * mats_per_launch,
                    "avg_launch_ms": dur * 1e3, "launches_timed": tri_launches,
                    "share_of_step": tri_ms / ms,
                    "hbm_view": {"bound": "hbm", "achieved": alg_bytes / dur / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / dur / 1e9 / hbm_peak,
                                 "note": "one-stage reduction streams the lower triangle of the trailing matrix once "
                                         "per column (4n^3/3 B) plus one read+write per panel (8n^3/(3 nb) B)"}}

    # ---- end to end through the public call with HOST buffers (pinned), H2D + D2H inside the timed region
    sts2 = [setup_state(k) for k in knames]
    # stage-1 regressions per second (SURVEY 8d asks for it next to the headline metric): warm second set-up pass,
    # one Gram + regression (with the smallest-eigenvalue rule) per target and kernel
    s1_ms = float(sum(s1_times))
    if world > 1:
        t1 = torch.tensor([s1_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t1, op=dist.ReduceOp.MAX)
        s1_ms = float(t1.item())
    stage1_rate = nkm * Tb * world / (s1_ms * 1e-3) if s1_ms > 0 else None
    hosts = [{k: s_[k].cpu().pin_memory() for k in ("w0", "ga", "alive", "yb", "keys", "lmin", "gmin")} for s_ in sts2]
    Xh = torch.from_numpy(X).pin_memory()
    res_h = {}
    counters = {"h2d": 0, "d2h": 0}

    def e2e_step():
        h2d = d2h = 0
        Xs = Xh.to(dev, non_blocking=True)
        h2d += Xh.numel() * 8
        for s_, host in zip(sts2, hosts):
            d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            h2d += sum(v.numel() * v.element_size() for v in host.values())
            out = plan.prune_chain(s_["desc"], Xs, cl, d["w0"], d["alive"], d["ga"], d["yb"], d["keys"], d["lmin"],
                                   d["gmin"], s_["step"], 1, keep_yb=True, p_cache=s_["p_cache"])
            s_["step"] += 1
            for k in ("pruned", "noise", "z_low", "z_high", "gamma", "act", "yb", "status"):
                res_h[k] = out[k].cpu()
                d2h += res_h[k].numel() * res_h[k].element_size()
            for k in ("alive", "yb", "keys", "lmin", "gmin"):
                host[k].copy_(d[k])
                d2h += host[k].numel() * host[k].element_size()
        counters["h2d"], counters["d2h"] = h2d, d2h

    e2e_steps = max(2, args.steps // 2)
    ms_e, *_ = timed(e2e_step, args.warmup, e2e_steps, False)
    e2e_value = nkm * Tb * world * e2e_steps / (ms_e * 1e-3)
    del sts2

    # ---- per-kernel rates of the default list
    per_kernel = {}
    if not args.no_per_kernel:
        for kname, s_ in zip(knames, sts):
            ms_k, *_ = timed(lambda: chain_step(s_), 1, 2, False)
            per_kernel[kname] = Tb * world * 2 / (ms_k * 1e-3)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, Tb),
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": counters["h2d"],
                        "d2h_bytes_per_step": counters["d2h"]},
                "roofline": roofline, "per_kernel": per_kernel, "stage1_regressions_per_sec": stage1_rate,
                "plan": {"two_stage": plan.two_stage, "band_half_width": _native.BAND,
                         "ctas_per_matrix": plan.ctas_per_matrix, "matrices_per_wave": plan.matrices_per_wave,
                         "panel_width": plan.panel_width, "threads_per_cta": plan.threads_per_cta,
                         "bulk_symv": plan.bulk_symv}}
        if world == 1 and not args.no_cpu_baseline:
            sec, per = cpu_mixed_step_seconds(args)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": cpu_sample_text(args),
                                    "per_kernel": {k: 1.0 / v for k, v in per.items()}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
