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

`--impl reference` (and the `cpu_baseline` object of the default arm) TIME the reference's algorithm -- the NumPy/LAPACK
oracle, JAX is not installable offline -- on the host cores at the config's real sizes (n = 8192, N = 512): real
prune steps of the Linear and Quadratic chains, no extrapolation in n or N.  The reference's Gaussian mode cannot run
this config at all (its cached `exps` tensor, Modes/kernels.py:307-309, would be n^2 N 8 B = 275 GB), so the CPU
arm covers two of the three kernels of the mix; the B200 `value` is the full three-kernel mix (the slower figure).
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
    ap.add_argument("--no-fit-wall", action="store_true", help="skip the fit() wall-time figures (C1, C2, C3 shapes)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling figure")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("CHD_BENCH_STREAMS", "1")),
                    help="CUDA streams the independent chains of a step are spread over (1: one stream)")
    ap.add_argument("--ref-budget-s", type=float, default=float(os.environ.get("CHD_REF_BUDGET_S", "240")),
                    help="wall-clock budget of the timed CPU steps of --impl reference")
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
GAUSSIAN_NOTE = ("Gaussian mode: the reference cannot run this config -- GaussianMode.setup caches exps (n x n x N, "
                 "Modes/kernels.py:307-309) = {gb:.0f} GB, and its activations cost C full n x n x N products per step "
                 "(helper_functions.py:113-132); the CPU arm therefore times the Linear and Quadratic chains only")


def kernels_of(args):
    return ["linear", "quadratic", "gaussian"] if args.kernel == "mix" else [args.kernel]


class CpuChain:
    """One target's prune chain run by the oracle (reference algorithm restated in NumPy/LAPACK, all host threads)
    at the config's real (n, N): stage 1 (untimed set-up: Gram + eigvalsh + regression, MAIN:448-511), then real
    prune steps (helper_functions.py:166-176), each timed as a whole."""

    def __init__(self, kernel, n, N, target=0):
        from oracle import kernels as okern, regression as oreg, jaxprng
        from computationalhypergraphdiscovery_b200.synthetic import make_data
        X, _names = make_data(n, N, seed=0)
        self.X = np.ascontiguousarray((X - X.mean(0)) / X.std(0))
        specs = okern.make_specs([("linear", 2.0, None), ("quadratic", 1.0, None), ("gaussian", 1.0, 1.0)])
        self.spec = {s.name: s for s in specs}[kernel]
        self.active = np.eye(N)
        self.active[:, target] = 0.0
        self.ga = self.X[:, target].copy()
        key = jaxprng.split(jaxprng.PRNGKey(0), N + 1)[1 + target]
        t0 = time.perf_counter()
        K = okern.gram(self.spec, self.X, self.X, self.active.sum(0))
        self.yb, _noise, _zl, _zh, _g, self.key = oreg.perform_regression_and_find_gamma(K, self.ga, 1e-9, key)
        self.setup_s = time.perf_counter() - t0
        self.steps = 0

    def step(self):
        from oracle import chd as ochd
        t0 = time.perf_counter()
        self.active, self.yb, _g, self.key, _act, noise, _zl, _zh = ochd.prune_step(
            self.spec, self.X, self.active, self.ga, self.yb, self.key, 1e-9, False)
        self.steps += 1
        return time.perf_counter() - t0, float(noise)


def cpu_measure(args, kernels, max_steps, warmup, budget_s):
    """Times real oracle prune steps, alternating over `kernels`, until `max_steps` are timed or the budget is spent
    (at least one per kernel).  Returns (seconds per step list, per-kernel seconds, set-up seconds)."""
    chains = [CpuChain(k, args.n, args.N) for k in kernels]
    setup = sum(c.setup_s for c in chains)
    t_start = time.perf_counter()
    for i in range(warmup):                       # untimed (threads, page faults): bounded by a third of the budget
        if time.perf_counter() - t_start > budget_s / 3:
            break
        chains[i % len(chains)].step()
    times, per = [], {k: [] for k in kernels}
    t_start = time.perf_counter()
    i = 0
    while len(times) < max_steps:
        sec, _noise = chains[i % len(chains)].step()
        times.append(sec)
        per[kernels[i % len(chains)]].append(sec)
        i += 1
        if time.perf_counter() - t_start > budget_s and i >= len(chains):
            break
    return times, {k: float(np.mean(v)) for k, v in per.items() if v}, setup


def cpu_sample_text(args, kernels, nsteps):
    return (f"oracle restatement of the reference (NumPy + LAPACK syevd on all host threads; JAX is not installable "
            f"offline): {nsteps} REAL prune steps (activations -> argmin -> Gram -> eigh -> gamma grid + BFGS -> solve "
            f"-> Z-test, helper_functions.py:166-176) of one target's chain per kernel ({', '.join(kernels)}) at the "
            f"config's own n={args.n}, N={args.N}; every step timed as a whole, nothing extrapolated.  "
            + GAUSSIAN_NOTE.format(gb=args.n ** 2 * args.N * 8 / 1e9))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kernels = [k for k in kernels_of(args) if k != "gaussian"] or ["quadratic"]
    times, per, setup = cpu_measure(args, kernels, args.steps, min(args.warmup, 1), args.ref_budget_s)
    sec = float(np.mean(times))
    cores = os.cpu_count()
    sample = cpu_sample_text(args, kernels, len(times))
    line = {"impl": "reference", "metric": METRIC, "value": 1.0 / sec, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "steps_requested": args.steps, "warmup": min(args.warmup, 1),
            "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "measured": True, "extrapolated": False, "timed_region_s": float(np.sum(times)), "setup_s": setup,
            "note": f"steps capped by --ref-budget-s {args.ref_budget_s:.0f} s (one CPU step is ~{sec:.0f} s); "
                    f"1 target in flight, kernels timed: {kernels}",
            "cpu_baseline": {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "per_kernel": {k: 1.0 / v for k, v in per.items()},
                             "gaussian": GAUSSIAN_NOTE.format(gb=args.n ** 2 * args.N * 8 / 1e9)},
            "e2e": {"value": 1.0 / sec, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    """Identical for both arms (the driver compares it); arm-specific facts live in `plan` / `cpu_baseline`."""
    label = {(8192, 512): "C4", (16384, 2048): "C5"}.get((args.n, args.N), "custom")
    return {"workload": f"{label} synthetic {args.N} nodes x {args.n} samples, kernels=[Linear,Quadratic,Gaussian(l=1)], "
                        f"prune-chain steps (helper_functions.py:166-176), chain kernel: "
                        f"{'equal shares of linear/quadratic/gaussian' if args.kernel == 'mix' else args.kernel}",
            "n_samples": args.n, "n_nodes": args.N, "kernel": args.kernel,
            "cache": f"inputs larger than L2 (every step streams fresh n x n FP64 matrices, "
                     f"{args.n * args.n * 8 / 1e6:.0f} MB each at n={args.n})"}


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from computationalhypergraphdiscovery_b200 import _native
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    from computationalhypergraphdiscovery_b200 import random as chd_random

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")                 # those levels print NCCL's version banner on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n, N = args.n, args.N

    X, _names = make_data(n, N, seed=0)
    X = np.ascontiguousarray((X - X.mean(0)) / X.std(0))
    plan0 = _native.Plan(n, N, N)
    # two-stage path: every launch is batched over the targets in flight (64 per kernel state: 34 GB of work
    # matrices at n = 8192; the latency-bound kernels -- bulge chasing, panel QR -- and the small last panels of the
    # GEMMs amortise over them: 42.4 ms per regression at 64 in flight, 44.4 at 32); one-stage path: one wave
    Tb = args.targets or (min(64, max(1, N)) if plan0.two_stage else plan0.matrices_per_wave)
    Xd = torch.from_numpy(X).to(dev)
    cl = torch.arange(N, dtype=torch.int32, device=dev)
    descs = {"linear": _native.make_desc(0, 2.0), "quadratic": _native.make_desc(1, 1.0, alpha=1.0),
             "gaussian": _native.make_desc(2, 1.0, alpha=1.0, quad_scale=1.0, l=1.0)}
    keys_all = chd_random.split(chd_random.PRNGKey(0), N + 1)[1:]
    knames = kernels_of(args)
    nkm = len(knames)
    nstreams = max(1, min(args.streams, nkm))
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)] if nstreams > 1 else [None]
    # one plan (= one workspace) per chain when the chains run concurrently
    plans = {k: (plan0 if nstreams == 1 else _native.Plan(n, N, N)) for k in knames}
    s1_times = []

    def host_inputs(T, first):
        """Targets first .. first+T-1 (mod N): initial masks, regression targets, PRNG sub-keys (host arrays)."""
        tidx = (np.arange(T) + first) % N
        w0 = np.ones((T, N), dtype=np.uint8)
        w0[np.arange(T), tidx] = 0
        return w0, np.ascontiguousarray(X[:, tidx].T), keys_all[tidx].astype(np.uint32)

    def setup_state(kernel, T, first, plan=None):
        """Stage-1 regression of the targets with `kernel` (untimed) -> chain state on the device."""
        plan = plan or plans[kernel]
        desc = descs[kernel]
        w0, ga_h, keys_h = host_inputs(T, first)
        w0d = torch.from_numpy(w0).to(dev)
        ga = torch.from_numpy(ga_h).to(dev)
        keys = torch.from_numpy(keys_h.view(np.int32).copy()).to(dev)
        gmin = torch.full((T,), 1e-9, dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = plan.gram(desc, Xd, w0d)
        r = plan.regress(K, ga, keys, gmin, mode=1, base=1e-9)   # stage 1 (MAIN:448-511): Gram + regression + lambda_min
        e1.record()
        torch.cuda.synchronize()
        s1_times.append(e0.elapsed_time(e1))
        del K
        p_cache = (dict(P=torch.empty((T, n, plan.ldk), dtype=torch.float64, device=dev), valid=0)
                   if kernel == "gaussian" else None)
        return dict(plan=plan, desc=desc, w0=w0d, ga=ga, alive=torch.ones((T, N), dtype=torch.uint8, device=dev),
                    yb=r["yb"], keys=r["keys"], lmin=r["lambda_min"], T=T,
                    gmin=torch.full((T,), 1e-9, dtype=torch.float64, device=dev), step=0, p_cache=p_cache)

    def chain_step(st):
        out = st["plan"].prune_chain(st["desc"], Xd, cl, st["w0"], st["alive"], st["ga"], st["yb"], st["keys"],
                                     st["lmin"], st["gmin"], st["step"], 1, keep_yb=True, p_cache=st["p_cache"])
        st["step"] += 1
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the one collective of the path: all-gather of the per-node results of a step over NVLink
    RES = ("noise", "z_low", "z_high", "gamma", "pruned", "act", "yb")
    coll = {"ms": 0.0, "calls": 0, "bytes_per_rank": 0, "events": []}

    def gather_results(outs):
        """Packs this rank's step results (noise, Z bounds, gamma, pruned index, activations, yb of every target) into
        one FP64 buffer and all-gathers it to every rank (NCCL all_gather_into_tensor)."""
        send = torch.cat([o[k].reshape(-1).to(torch.float64) for o in outs for k in RES])
        recv = torch.empty((world, send.numel()), dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_gather_into_tensor(recv.view(-1), send)
        e1.record()
        coll["events"].append((e0, e1))
        coll["calls"] += 1
        coll["bytes_per_rank"] = send.numel() * 8
        return recv

    def mixed_step_on(states, gather):
        main = torch.cuda.current_stream()
        outs = []
        if nstreams > 1:
            for i, s_ in enumerate(states):
                sst = streams[i % nstreams]
                sst.wait_stream(main)
                with torch.cuda.stream(sst):
                    outs.append(chain_step(s_))
            for sst in streams:
                main.wait_stream(sst)
        else:
            for s_ in states:
                outs.append(chain_step(s_))
        if gather and world > 1:
            gather_results(outs)
        return outs

    prof = {}

    def timed(fn, warmup, steps, sample_clocks=False):
        for _ in range(warmup):
            fn()
        coll["events"].clear()
        coll["calls"] = 0
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
        coll["ms"] = float(sum(a.elapsed_time(b) for a, b in coll["events"]))
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, tri_ms, tri_launches, clocks

    # ---- device-resident throughput (value): weak scaling, Tb targets of every kernel per rank
    sts = [setup_state(k, Tb, rank * Tb) for k in knames]   # first call per kernel includes one-time set-up
    s1_times.clear()
    ms, launches, tri_ms, tri_launches, clocks = timed(lambda: mixed_step_on(sts, True), args.warmup, args.steps, True)
    total_steps = nkm * Tb * world * args.steps
    value = total_steps / (ms * 1e-3)
    collective = None
    if world > 1:
        collective = {"op": "all_gather_into_tensor (NCCL) of the step's per-node results, inside the timed region",
                      "calls": coll["calls"], "ms_total": coll["ms"], "ms_per_step": coll["ms"] / max(args.steps, 1),
                      "bytes_per_rank_per_step": coll["bytes_per_rank"],
                      "share_of_step": coll["ms"] / ms if ms else None}

    # ---- roofline of the dominant kernel, timed live with CUDA events around every launch inside the timed region
    n_mats = nkm * Tb * args.steps
    peak_tf = _native.dmma_peak_tflops(20000) if rank == 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = ("FP64 DMMA m8n8k4 issue-rate probe run in this process (chd_dmma_peak: 148 SMs x 32 warps of back-to-"
                "back DMMA.8x8x4, SASS in profiles/); MEASURED_PEAKS.json has no FP64 figure")
    prof_main = dict(prof)
    roofline = build_roofline(args, plan0, prof_main, ms, n_mats, Tb, peak_tf, peak_src, hbm_peak, tri_ms, tri_launches)

    # ---- end to end through the public call with HOST buffers (pinned), H2D + D2H inside the timed region
    sts2 = [setup_state(k, Tb, rank * Tb) for k in knames]
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
            out = s_["plan"].prune_chain(s_["desc"], Xs, cl, d["w0"], d["alive"], d["ga"], d["yb"], d["keys"],
                                         d["lmin"], d["gmin"], s_["step"], 1, keep_yb=True, p_cache=s_["p_cache"])
            s_["step"] += 1
            for k in ("pruned", "noise", "z_low", "z_high", "gamma", "act", "yb", "status"):
                res_h[k] = out[k].cpu()
                d2h += res_h[k].numel() * res_h[k].element_size()
            for k in ("alive", "yb", "keys", "lmin", "gmin"):
                host[k].copy_(d[k])
                d2h += host[k].numel() * host[k].element_size()
        counters["h2d"], counters["d2h"] = h2d, d2h

    e2e_steps = max(2, args.steps // 4)
    ms_e, *_ = timed(e2e_step, min(args.warmup, 3), e2e_steps, False)
    e2e_value = nkm * Tb * world * e2e_steps / (ms_e * 1e-3)
    del sts2, hosts

    # ---- per-kernel rates of the default list
    per_kernel = {}
    if not args.no_per_kernel:
        for kname, s_ in zip(knames, sts):
            ms_k, *_ = timed(lambda: chain_step(s_), 1, 2, False)
            per_kernel[kname] = Tb * world * 2 / (ms_k * 1e-3)
    del sts
    torch.cuda.empty_cache()

    # ---- strong scaling: a FIXED set of targets (64 per kernel) split over the ranks, result all-gather included
    strong = None
    if not args.no_strong:
        tot = 64 if N >= 64 else N
        lo, hi = rank * tot // world, (rank + 1) * tot // world
        if hi > lo:
            sst = [setup_state(k, hi - lo, lo) for k in knames]
        else:
            sst = []
        if world > 1:
            # ranks hold different target counts when 64 % world != 0: pad-free variable-size gather via equal split
            assert tot % world == 0, "strong-scaling figure needs the target count to divide by the number of GPUs"
        ms_s, *_ = timed(lambda: mixed_step_on(sst, True), 1, 2, False)
        strong = {"targets_total_per_kernel": tot, "kernels": knames, "steps": 2, "ms_per_step": ms_s / 2,
                  "node_prune_steps_per_sec": nkm * tot * 2 / (ms_s * 1e-3), "targets_per_rank_per_kernel": hi - lo,
                  "note": "fixed work split over the ranks (scaling = 'strong' for this figure only); compare "
                          "node_prune_steps_per_sec across the N = 1/2/4/8 lines"}
        del sst
        torch.cuda.empty_cache()

    # ---- sharded fit() == unsharded fit(), bit for bit (N > 1), and fit() wall times of the small configs (N = 1)
    sharded_identical = None
    if world > 1:
        sharded_identical = sharded_fit_check(dev)
    fit_wall = None
    if world == 1 and not args.no_fit_wall:
        fit_wall = fit_wall_times(args)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": counters["h2d"],
                        "d2h_bytes_per_step": counters["d2h"]},
                "roofline": roofline, "per_kernel": per_kernel, "stage1_regressions_per_sec": stage1_rate,
                "plan": {"two_stage": plan0.two_stage, "band_half_width": _native.BAND,
                         "targets_in_flight_per_gpu": Tb, "streams": nstreams,
                         "ctas_per_matrix": plan0.ctas_per_matrix, "matrices_per_wave": plan0.matrices_per_wave,
                         "panel_width": plan0.panel_width, "threads_per_cta": plan0.threads_per_cta,
                         "bulk_symv": plan0.bulk_symv}}
        if collective:
            line["collective"] = collective
        if strong:
            line["strong_scaling"] = strong
        if sharded_identical is not None:
            line["sharded_fit_identical"] = sharded_identical
        if fit_wall:
            line["fit_wall_s"] = fit_wall
        try:
            line["parity_notes"] = json.load(open(os.path.join(ROOT, "profiles", "r02_parity_notes.json")))
        except Exception:  # noqa: BLE001
            pass
        if world == 1 and not args.no_cpu_baseline:
            times, per, setup = cpu_measure(args, ["quadratic"], 1, 0, 60.0)
            sec = float(np.mean(times))
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": cpu_sample_text(args, ["quadratic"], len(times)),
                                    "measured": True, "seconds_per_step": sec, "setup_s": setup,
                                    "per_kernel": {k: 1.0 / v for k, v in per.items()},
                                    "gaussian": GAUSSIAN_NOTE.format(gb=n ** 2 * N * 8 / 1e9)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def sharded_fit_check(dev):
    """fit() sharded over the ranks (targets split, ONE packed all-gather at the end) must equal the unsharded fit()
    bit for bit: same edges, signals, node attributes.  Returns {identical, edges, ms_sharded, ms_single}."""
    import torch
    import torch.distributed as dist
    import computationalhypergraphdiscovery_b200 as CHD
    from computationalhypergraphdiscovery_b200 import parallel
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(400, 16, seed=2)

    def run(single):
        gd = CHD.GraphDiscovery(X, names, verbose=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if single:
            with parallel.single_process():
                gd.fit()
        else:
            gd.fit()
        torch.cuda.synchronize()
        return gd, time.perf_counter() - t0

    run(True)                                        # warm-up (context, allocator)
    sh, t_sh = run(False)
    ref, t_ref = run(True)
    same = sorted(sh.G.edges(data="signal")) == sorted(ref.G.edges(data="signal"))
    for nm in names:
        a, b = sh.G.nodes[nm], ref.G.nodes[nm]
        same = same and set(a) == set(b)
        if same and "noise" in a:
            same = (a["noise"] == b["noise"] and a["gamma"] == b["gamma"] and a["kernel_index"] == b["kernel_index"]
                    and bool((a["coeff"] == b["coeff"]).all()) and bool((a["active_modes"] == b["active_modes"]).all()))
    flag = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"identical": bool(flag.item()), "config": "synthetic n=400, N=16, default kernels", "comparison": "bitwise",
            "edges": sh.G.number_of_edges(), "fit_s_sharded": t_sh, "fit_s_single_gpu": t_ref}


def fit_wall_times(args):
    """fit() wall time (first half of BASELINE.json's metric, MAIN:272-421) on the config shapes that fit one call:
    C1, C2, C3 on the GPU through the public API, C1 through the oracle on the host cores (graphs compared)."""
    import torch
    import computationalhypergraphdiscovery_b200 as CHD
    from computationalhypergraphdiscovery_b200 import _native, configs
    out = {}
    for name in ("c1", "c2", "c3"):
        ctor, fitkw, desc, okw = configs.BUILDERS[name]()
        if name == "c1":                                 # warm-up: context, allocator, first-use set-up
            CHD.GraphDiscovery(verbose=False, **ctor).fit(**fitkw)
        gd = CHD.GraphDiscovery(verbose=False, **ctor)
        torch.cuda.synchronize()
        l0, t0 = _native.launch_count(), time.perf_counter()
        gd.fit(**fitkw)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        steps = int(sum(ch["noises"].shape[0] * (ch["noises"].shape[1] - 1) for ch in gd.last_fit["chains"].values()))
        out[name] = {"config": desc, "fit_s": sec, "node_prune_steps": steps, "edges": gd.G.number_of_edges(),
                     "stage1_regressions": len(gd.last_fit["targets"]) * len(ctor["kernels"]),
                     "gpu_launches": _native.launch_count() - l0}
        if okw is not None and not args.no_cpu_baseline:
            from oracle import chd as ochd
            t0 = time.perf_counter()
            res = ochd.fit(ctor["X"], ctor["names"], **okw)
            out[name]["oracle_cpu_fit_s"] = time.perf_counter() - t0
            out[name]["cpu_cores"] = os.cpu_count()
            out[name]["graph_identical_to_oracle"] = bool(all(
                sorted(a for a, _ in gd.G.in_edges(nm)) == sorted(res["nodes"].get(nm, {}).get("ancestors", []))
                for nm in ctor["names"]))
        del gd
        torch.cuda.empty_cache()
    return out


def build_roofline(args, plan, prof_main, ms, n_mats, Tb, peak_tf, peak_src, hbm_peak, tri_ms, tri_launches):
    from computationalhypergraphdiscovery_b200 import _native
    n = args.n
    if plan.two_stage:
        # Stage 1 (dense -> band, NB = 32) is two DMMA GEMM kernels per panel, batched over the targets in flight:
        #   symm  Y = A22 V          2 m^2 NB flop per matrix and panel  (sum over panels: 2/3 n^3)
        #   syr2k A22 -= VW^T+WV^T   2 m^2 NB flop (lower triangle only)  (sum over panels: 2/3 n^3)
        NBW = _native.BAND
        ms_list = [n - p * NBW for p in range(0, (n - 2) // NBW + 1) if n - p * NBW >= 2]
        ms_list = [n] + [m for m in ms_list[1:]]          # panel 0 is H_{-1} on the full matrix
        gemm_flops = sum(2.0 * m * m * NBW for m in ms_list)   # symm, per matrix: every panel slot on its trailing block
        # trailing update: once per group of two slots, on the trailing block of the group's LAST slot, K = 2 NB per slot
        upd_flops = sum(2.0 * ms_list[min(g + 1, len(ms_list) - 1)] ** 2 * NBW * min(2, len(ms_list) - g)
                        for g in range(0, len(ms_list), 2))
        flops_of = {"symm": gemm_flops, "syr2k": upd_flops}
        kern_stats = {}
        for kname in ("symm", "syr2k", "panel_qr", "wform", "chase", "band_solve", "backtransform"):
            if kname in prof_main:
                kms, kcnt = prof_main[kname]
                kern_stats[kname] = {"ms": kms, "launches": kcnt, "share_of_step": kms / ms}
        dom = max(("symm", "syr2k"), key=lambda k: prof_main.get(k, (0.0, 0))[0])
        dms, dcnt = prof_main.get(dom, (0.0, 1))
        achieved_tf = flops_of[dom] * n_mats / (dms * 1e-3) / 1e12 if dms else 0.0
        # algorithmic bytes from HBM: symm reads the lower triangle once per panel (4 m^2 B: the transposed use of every
        # block is the paired CTA's L2 hit); the trailing update reads + writes it once per group of two panels
        # (8 m^2 B per group = 4 m^2 B per panel)
        alg_bytes = sum(4.0 * m * m for m in ms_list) * n_mats
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_stage1_traffic.json")))
            per_matrix = tr.get(f"{dom}_n{n}_per_matrix")
            traffic = per_matrix * n_mats / max(dcnt, 1) if per_matrix else None
        except Exception:  # noqa: BLE001
            pass
        stage1_ms = sum(prof_main.get(k, (0.0, 0))[0] for k in ("symm", "syr2k", "panel_qr", "wform"))
        per_gemm = {}
        for k in ("symm", "syr2k"):
            kms = prof_main.get(k, (0.0, 0))[0]
            if kms:
                tf = flops_of[k] * n_mats / (kms * 1e-3) / 1e12
                per_gemm[k] = {"achieved_tflops": tf, "frac_of_dmma_peak": tf / peak_tf if peak_tf else None}
        kfull = {"symm": "symm_tma_kernel (Y = A22 V, paired block schedule)",
                 "syr2k": "syr2k_tma_kernel (trailing update of a group of two panels, K = 128)"}[dom]
        return {"kernel": f"{kfull}: TMA tensor-map tiles + mbarrier ring + FP64 DMMA m8n8k4, dense->band stage of "
                          f"the two-stage reduction, batched over {Tb} targets per launch",
                "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                "traffic_source": "ncu dram__bytes of one capture per kernel (profiles/r02_stage1_traffic.json), "
                                  "scaled over the panels by the kernels' m^2 law, per launch",
                "peak_source": peak_src,
                "flops_per_launch": flops_of[dom] * n_mats / max(dcnt, 1),
                "avg_launch_ms": dms / max(dcnt, 1), "launches_timed": dcnt, "share_of_step": dms / ms,
                "gemm_kernels": per_gemm,
                "stage1": {"credited_flops_per_matrix": credited_flops_tridiag(n),
                           "achieved_tflops": credited_flops_tridiag(n) * n_mats / (stage1_ms * 1e-3) / 1e12
                           if stage1_ms else None,
                           "frac_of_dmma_peak": (credited_flops_tridiag(n) * n_mats / (stage1_ms * 1e-3) / 1e12 /
                                                 peak_tf) if stage1_ms and peak_tf else None,
                           "ms_per_matrix": stage1_ms / n_mats,
                           "note": "symm + trailing update + panel QR (+ thin update) + partial sums / small matrices / "
                                   "W / Z block, credited 4/3 n^3 (SURVEY 8d)"},
                "kernels": kern_stats,
                "hbm_view": {"bound": "hbm", "achieved": alg_bytes / (dms * 1e-3) / 1e9 if dms else None,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (dms * 1e-3) / 1e9 / hbm_peak if dms else None,
                             "note": "4 m^2 B per panel and matrix for either GEMM kernel (symm: lower triangle read "
                                     "once, the transposed use is an L2 hit; trailing update: read + write once per "
                                     "two panels)"}}
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
    return {"kernel": "tridiag_kernel (batched Householder tridiagonalisation, DMMA trailing update)",
            "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic, "peak_source": peak_src,
            "flops_per_launch": credited_flops_tridiag(n) * mats_per_launch,
            "avg_launch_ms": dur * 1e3, "launches_timed": tri_launches,
            "share_of_step": tri_ms / ms,
            "hbm_view": {"bound": "hbm", "achieved": alg_bytes / dur / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": alg_bytes / dur / 1e9 / hbm_peak,
                         "note": "one-stage reduction streams the lower triangle of the trailing matrix once "
                                 "per column (4n^3/3 B) plus one read+write per panel (8n^3/(3 nb) B)"}}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
